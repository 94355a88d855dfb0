#!/bin/bash
# validation: pipelined pw_bwd_reduce, chunked ring forward for enc.2 (1 GPU)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2o_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/r2o_summary.txt
timeout 300 python tools/step_timeline.py > gpurun_out/r2o_timeline.log 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2o_summary.txt
FV_CONV_RING_CHUNK=0 timeout 300 python tools/step_timeline.py > gpurun_out/r2o_timeline_nochunk.log 2>&1; echo "timeline nochunk rc=$?" | tee -a gpurun_out/r2o_summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2o_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2o_bench_1gpu.json 2> gpurun_out/r2o_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2o_summary.txt
tail -4 gpurun_out/r2o_all.log
head -4 gpurun_out/r2o_timeline.log
grep -E "pw_bwd|conv_ring|conv_igemm_kernel<64>" gpurun_out/r2o_timeline.log | head
head -1 gpurun_out/r2o_timeline_nochunk.log
tail -2 gpurun_out/r2o_smoke.log
python tools/show_bench.py gpurun_out/r2o_bench_1gpu.json 2>/dev/null | head -2
