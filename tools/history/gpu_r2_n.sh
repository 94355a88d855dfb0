#!/bin/bash
# validation of: producer-emitted statistics for the ResBlock2D chain, wide slab sum, shared zero gradients (1 GPU)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/r2n_summary.txt
timeout 300 python tools/step_timeline.py > gpurun_out/r2n_timeline.log 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2n_summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2n_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2n_bench_1gpu.json 2> gpurun_out/r2n_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2n_summary.txt
tail -4 gpurun_out/r2n_all.log
head -5 gpurun_out/r2n_timeline.log
grep -E "bn_stats|slab_sum|Fill|colsum" gpurun_out/r2n_timeline.log | head
tail -3 gpurun_out/r2n_smoke.log
python tools/show_bench.py gpurun_out/r2n_bench_1gpu.json | head -3
