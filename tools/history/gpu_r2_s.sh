#!/bin/bash
# validation: flat-grid weight prep; bn_stats grid threshold experiment
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_determinism_gpu.py tests/test_elr_gpu.py tests/test_f2_gpu.py tests/test_updown_gpu.py -x -q -m gpu > gpurun_out/r2s2_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2s2_summary.txt
for B in 64 256 100000; do
  FV_REDUCE_BIG=$B timeout 300 python tools/step_timeline.py > gpurun_out/r2s2_timeline_big$B.log 2>&1; echo "timeline big=$B rc=$?" | tee -a gpurun_out/r2s2_summary.txt
done
tail -3 gpurun_out/r2s2_tests.log
for B in 64 256 100000; do
  echo "== big rows $B"; head -1 gpurun_out/r2s2_timeline_big$B.log
  grep -E "bn_stats_kernel|weight_prep|bn_act_bwd_reduce_kernel<__nv_bfloat16, __nv_bfloat16, 1" gpurun_out/r2s2_timeline_big$B.log | head -6
done
