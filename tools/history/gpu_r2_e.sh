#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 "$@" > gpurun_out/r2e_$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/r2e_summary.txt; }
run layerwise python -m pytest tests/test_layerwise_gpu.py -q -m gpu -s
run parity python -m pytest tests/test_parity_gpu.py -q -m gpu -s
run updown python -m pytest tests/test_updown_gpu.py tests/test_determinism_gpu.py -q -m gpu
run timeline python tools/step_timeline.py --e2e-steps 50
for f in layerwise parity updown; do echo "== $f"; tail -4 gpurun_out/r2e_$f.log; done; head -12 gpurun_out/r2e_timeline.log
