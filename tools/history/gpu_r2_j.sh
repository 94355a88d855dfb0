#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/r2j_$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/r2j_summary.txt; }
run updown python -m pytest tests/test_updown_gpu.py -q -m gpu
run timeline python tools/step_timeline.py --e2e-steps 50
python __graft_entry__.py smoke > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2j_summary.txt
tail -3 gpurun_out/r2j_updown.log; sed -n 1,5p gpurun_out/r2j_timeline.log; grep -n "conv_win\|traced" gpurun_out/r2j_timeline.log; tail -3 gpurun_out/r2j_smoke.log
