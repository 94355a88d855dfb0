#!/bin/bash
# Runs the kernel-level GPU tests in separate processes (a trapping kernel poisons its CUDA context) under timeouts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for grp in "layout or bn_ or reparam or recon" conv_1x1 conv_3x3_tilings conv_3x3_channels "7x7 or rgb_in or residual" many_tiles dgrad_wgrad; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -s > "gpurun_out/kt_${name}.log" 2>&1
  echo "== $grp -> exit $?"
  grep -E "passed|failed|error" "gpurun_out/kt_${name}.log" | tail -2
done
grep -h -E "^(conv|dgrad|wgrad|bn_|fv:)" gpurun_out/kt_*.log | head -150
