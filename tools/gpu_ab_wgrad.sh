#!/bin/bash
# A/B of the capture-stream priority with the weight-gradient side stream on.
mkdir -p gpurun_out
P=gpurun_out/ab2
: > ${P}_summary.txt
for PR in 0 -1 0 -1; do
  FACEVAE_WGRAD_STREAM=1 FACEVAE_CAP_PRIORITY=$PR timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-glue-roofline --profile-steps 1 > ${P}_bench.json 2>> ${P}_bench.err
  echo "priority $PR rc=$? $(python tools/show_bench.py ${P}_bench.json 2>/dev/null | head -1 | cut -c1-90)" | tee -a ${P}_summary.txt
done
