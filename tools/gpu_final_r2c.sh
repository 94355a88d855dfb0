#!/bin/bash
# Validation + A/B of side-stream level 2 (bias column sums and filter preparation on the side stream as well).
mkdir -p gpurun_out
P=gpurun_out/r2r
timeout 200 python -m pytest tests -x -q -m gpu > ${P}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee ${P}_summary.txt
timeout 60 python __graft_entry__.py smoke > ${P}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a ${P}_summary.txt
FACEVAE_WGRAD_STREAM=1 timeout 100 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-glue-roofline --profile-steps 1 > ${P}_bench_l1.json 2> ${P}_bench_l1.err
echo "level 1 rc=$? $(python tools/show_bench.py ${P}_bench_l1.json 2>/dev/null | head -1 | cut -c1-90)" | tee -a ${P}_summary.txt
timeout 200 python bench.py --steps 60 --warmup 5 > ${P}_bench_1gpu.json 2> ${P}_bench_1gpu.err
echo "level 2 (default) rc=$? $(python tools/show_bench.py ${P}_bench_1gpu.json 2>/dev/null | head -1 | cut -c1-90)" | tee -a ${P}_summary.txt
tail -2 ${P}_pytest_gpu.log; tail -2 ${P}_smoke.log
