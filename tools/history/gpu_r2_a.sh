#!/bin/bash
# round 2, call A: kernel-level validation of the deterministic reductions and the x2 / s2 geometries
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_updown_gpu.py -x -q -m gpu -s > gpurun_out/r2a_updown.log 2>&1; echo "updown rc=$?" | tee -a gpurun_out/r2a_summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/r2a_kernels.log 2>&1; echo "kernels rc=$?" | tee -a gpurun_out/r2a_summary.txt
timeout 600 python -m pytest tests/test_outconv_gpu.py -q -m gpu -x > gpurun_out/r2a_outconv.log 2>&1; echo "outconv rc=$?" | tee -a gpurun_out/r2a_summary.txt
timeout 600 python -m pytest tests/test_xrank_gpu.py -q -m gpu -x > gpurun_out/r2a_xrank.log 2>&1; echo "xrank rc=$?" | tee -a gpurun_out/r2a_summary.txt
timeout 900 python -m pytest tests/test_determinism_gpu.py -q -m gpu > gpurun_out/r2a_determinism.log 2>&1; echo "determinism rc=$?" | tee -a gpurun_out/r2a_summary.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu > gpurun_out/r2a_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/r2a_summary.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" | tee -a gpurun_out/r2a_summary.txt
tail -3 gpurun_out/r2a_*.log
