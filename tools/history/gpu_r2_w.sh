#!/bin/bash
mkdir -p gpurun_out
STEPS=1 timeout 300 python tools/ncu_one.py > gpurun_out/r2w_plain.log 2>&1 &&
STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'weight_prep_tile' -c 2 -o gpurun_out/r2w_prep -f python tools/ncu_one.py > gpurun_out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2w_prep.ncu-rep --page raw --csv > gpurun_out/r2w_prep_raw.csv 2>/dev/null
ncu -i gpurun_out/r2w_prep.ncu-rep --page source --csv --launch-skip 1 --launch-count 1 > gpurun_out/r2w_prep_src.csv 2>/dev/null
ls -la gpurun_out | grep r2w
