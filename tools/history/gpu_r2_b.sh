#!/bin/bash
# round 2, call B: whole-model validation after the kernel-level pass of call A
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/r2b_$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/r2b_summary.txt; }
run updown python -m pytest tests/test_updown_gpu.py -q -m gpu
run outconv python -m pytest tests/test_outconv_gpu.py -q -m gpu
run determinism python -m pytest tests/test_determinism_gpu.py -q -m gpu
run parity python -m pytest tests/test_parity_gpu.py -q -m gpu
run layerwise python -m pytest tests/test_layerwise_gpu.py -q -m gpu -s
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" | tee -a gpurun_out/r2b_summary.txt
for f in gpurun_out/r2b_*.log; do echo "== $f"; tail -4 $f; done
