#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/r2g_$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/r2g_summary.txt; }
run elr python -m pytest tests/test_elr_gpu.py -q -m gpu
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?" | tee -a gpurun_out/r2g_summary.txt
timeout 200 python bench.py --infer --batch 64 --steps 20 > gpurun_out/r2g_infer64.json 2> gpurun_out/r2g_infer.err; echo "infer rc=$?" | tee -a gpurun_out/r2g_summary.txt
bash tools/gpu_sanitize.sh memcheck 2>&1 | tee -a gpurun_out/r2g_summary.txt
tail -5 gpurun_out/r2g_elr.log; tail -3 gpurun_out/r2g_bench.err
