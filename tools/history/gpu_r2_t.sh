#!/bin/bash
# validation: power-of-two index decomposition in the norm + act kernels, one-wave bn_stats
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/r2t_summary.txt
timeout 300 python tools/step_timeline.py > gpurun_out/r2t_timeline.log 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2t_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2t_bench_1gpu.json 2> gpurun_out/r2t_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2t_summary.txt
tail -3 gpurun_out/r2t_all.log
head -4 gpurun_out/r2t_timeline.log
grep -E "bn_act|bn_stats" gpurun_out/r2t_timeline.log | head -14
python tools/show_bench.py gpurun_out/r2t_bench_1gpu.json 2>/dev/null | head -1
