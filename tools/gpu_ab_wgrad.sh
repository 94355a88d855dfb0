#!/bin/bash
# A/B of the two-stream train step on one box (profiles/r02_ab_wgrad_side_stream.txt): side stream off / weight gradients only /
# + bias column sums and filter preparation (FACEVAE_WGRAD_STREAM=0|1|2), and the capture-stream priority (FACEVAE_CAP_PRIORITY).
# The schedules are bitwise equal (tests/test_determinism_gpu.py), so only the step time differs.
mkdir -p gpurun_out
P=gpurun_out/ab
timeout 150 python -m pytest tests/test_determinism_gpu.py -x -q -k "side_stream or train_step" > ${P}_det.log 2>&1; echo "determinism rc=$?" | tee ${P}_summary.txt
for CFG in "0 -1" "1 0" "1 -1" "2 -1"; do
  set -- $CFG
  FACEVAE_WGRAD_STREAM=$1 FACEVAE_CAP_PRIORITY=$2 timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-glue-roofline --profile-steps 1 > ${P}_bench.json 2>> ${P}_bench.err
  echo "side stream $1 priority $2 rc=$? $(python tools/show_bench.py ${P}_bench.json 2>/dev/null | head -1 | cut -c1-90)" | tee -a ${P}_summary.txt
done
tail -3 ${P}_det.log
