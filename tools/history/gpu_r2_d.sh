#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_layerwise_gpu.py -q -m gpu -s > gpurun_out/r2d_layerwise.log 2>&1; echo "layerwise rc=$?" | tee -a gpurun_out/r2d_summary.txt
timeout 600 python -m pytest tests/test_determinism_gpu.py tests/test_kernels_gpu.py tests/test_updown_gpu.py -q -m gpu -x > gpurun_out/r2d_kernels.log 2>&1; echo "kernels rc=$?" | tee -a gpurun_out/r2d_summary.txt
timeout 600 python tools/step_timeline.py --e2e-steps 50 > gpurun_out/r2d_timeline.txt 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2d_summary.txt
FV_X2_FUSE_STATS=1 timeout 600 python tools/step_timeline.py --e2e-steps 50 > gpurun_out/r2d_timeline_fuse1.txt 2>&1; echo "timeline fuse1 rc=$?" | tee -a gpurun_out/r2d_summary.txt
tail -5 gpurun_out/r2d_layerwise.log; tail -3 gpurun_out/r2d_kernels.log; head -5 gpurun_out/r2d_timeline.txt; head -5 gpurun_out/r2d_timeline_fuse1.txt
