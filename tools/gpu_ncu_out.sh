#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python tools/conv_bench.py --only out --iters 1 > gpurun_out/cb_out_plain.txt 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fold_ -c 9 -o gpurun_out/prof_fold -f python tools/conv_bench.py --only out --iters 1 > gpurun_out/ncu_fold.log 2>&1; echo "ncu exit $?"
cat gpurun_out/cb_out_plain.txt; tail -5 gpurun_out/ncu_fold.log
