#!/bin/bash
# last check of the round: the kernel test file (incl. the large-mean batch-norm statistics test)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -s -k "large_mean or batched_weight or pointwise" > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r2z_summary.txt
grep -E "bn large-mean|passed|failed" gpurun_out/r2z_tests.log | tail -5
