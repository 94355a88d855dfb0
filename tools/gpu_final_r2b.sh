#!/bin/bash
# Final round-2 validation after the weight-gradient side stream (kernels unchanged since gpu_final_r2.sh): all GPU tests, smoke,
# the bench line, the 512-deep line, the CUPTI timeline of the replayed step.
mkdir -p gpurun_out
P=gpurun_out/r2q
timeout 400 python -m pytest tests -x -q -m gpu > ${P}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee ${P}_summary.txt
timeout 120 python __graft_entry__.py smoke > ${P}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a ${P}_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 > ${P}_bench_1gpu.json 2> ${P}_bench_1gpu.err; echo "bench rc=$?" | tee -a ${P}_summary.txt
timeout 120 python bench.py --deep --size 512 --batch 8 --steps 20 --warmup 5 --no-cpu-baseline --no-glue-roofline > ${P}_bench_512deep_1gpu.json 2> ${P}_bench_512deep_1gpu.err; echo "bench 512deep rc=$?" | tee -a ${P}_summary.txt
timeout 120 python tools/step_timeline.py > ${P}_timeline.log 2>&1; echo "timeline rc=$?" | tee -a ${P}_summary.txt
tail -3 ${P}_pytest_gpu.log; tail -3 ${P}_smoke.log
python tools/show_bench.py ${P}_bench_1gpu.json 2>/dev/null | head -1
python tools/show_bench.py ${P}_bench_512deep_1gpu.json 2>/dev/null | head -1 | cut -c1-100
head -6 ${P}_timeline.log
