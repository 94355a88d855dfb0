#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_layerwise_gpu.py -q -m gpu -s > gpurun_out/r2c_layerwise.log 2>&1; echo "layerwise rc=$?" | tee -a gpurun_out/r2c_summary.txt
timeout 600 python tools/step_timeline.py > gpurun_out/r2c_timeline.txt 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2c_summary.txt
tail -5 gpurun_out/r2c_layerwise.log
