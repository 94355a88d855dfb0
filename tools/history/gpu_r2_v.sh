#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_determinism_gpu.py tests/test_updown_gpu.py tests/test_elr_gpu.py tests/test_f2_gpu.py -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2v_summary.txt
timeout 300 python tools/step_timeline.py > gpurun_out/r2v_timeline.log 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2v_summary.txt
tail -3 gpurun_out/r2v_tests.log
head -4 gpurun_out/r2v_timeline.log
grep -E "weight_prep" gpurun_out/r2v_timeline.log | head -3
