#!/bin/bash
# validation: dual-issuer ring kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv" > gpurun_out/r2x_conv.log 2>&1; echo "conv tests rc=$?" | tee -a gpurun_out/r2x_summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_layerwise_gpu.py tests/test_determinism_gpu.py tests/test_outconv_gpu.py -x -q -m gpu > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2x_summary.txt
timeout 300 python tools/step_timeline.py > gpurun_out/r2x_timeline.log 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2x_summary.txt
FV_RING_DUAL=0 timeout 300 python tools/step_timeline.py > gpurun_out/r2x_timeline_single.log 2>&1; echo "timeline single rc=$?" | tee -a gpurun_out/r2x_summary.txt
tail -3 gpurun_out/r2x_conv.log; tail -3 gpurun_out/r2x_tests.log
head -1 gpurun_out/r2x_timeline.log; grep -E "conv_ring" gpurun_out/r2x_timeline.log | head -4
head -1 gpurun_out/r2x_timeline_single.log; grep -E "conv_ring" gpurun_out/r2x_timeline_single.log | head -4
