#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/r2i_$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/r2i_summary.txt; }
run f2 python -m pytest tests/test_f2_gpu.py -q -m gpu
run updown python -m pytest tests/test_updown_gpu.py -q -m gpu
run timeline python tools/step_timeline.py --e2e-steps 50
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2i_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/r2i_summary.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?" | tee -a gpurun_out/r2i_summary.txt
for f in f2 updown all; do echo "== $f"; tail -8 gpurun_out/r2i_$f.log; done; sed -n 1,5p gpurun_out/r2i_timeline.log; grep -n "conv_win\|traced" gpurun_out/r2i_timeline.log
