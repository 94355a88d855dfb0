#!/bin/bash
# A/B of the weight-gradient side stream (FACEVAE_WGRAD_STREAM): correctness with it on, then the bench line both ways.
mkdir -p gpurun_out
P=gpurun_out/ab
timeout 150 python -m pytest tests/test_determinism_gpu.py -x -q -k "side_stream or train_step" > ${P}_det.log 2>&1; echo "determinism rc=$?" | tee ${P}_summary.txt
for F in 0 1; do
  FACEVAE_WGRAD_STREAM=$F timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-glue-roofline --profile-steps 1 > ${P}_bench_$F.json 2>> ${P}_bench.err
  echo "flag $F rc=$? $(python tools/show_bench.py ${P}_bench_$F.json 2>/dev/null | head -1)" | tee -a ${P}_summary.txt
done
tail -5 ${P}_det.log
