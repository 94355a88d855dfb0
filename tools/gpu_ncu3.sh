#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python tools/conv_bench.py --only enc.1 --iters 2 --what fwd,wgrad > gpurun_out/cb_enc1.txt 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 4 -c 2 -o gpurun_out/prof3_enc1 -f python tools/conv_bench.py --only enc.1 --iters 2 --what fwd,wgrad > gpurun_out/ncu3_enc1.log 2>&1; echo "ncu exit $?"
cat gpurun_out/cb_enc1.txt
