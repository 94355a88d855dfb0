#!/bin/bash
# validation: enc.2 forward on the ring schedule with a 4-slot ring (N = 128 in one pass)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_layerwise_gpu.py tests/test_determinism_gpu.py -x -q -m gpu > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2q_summary.txt
timeout 300 python tools/step_timeline.py > gpurun_out/r2q_timeline.log 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2q_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2q_bench_1gpu.json 2> gpurun_out/r2q_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2q_summary.txt
timeout 300 python bench.py --deep --size 512 --batch 8 --steps 20 --warmup 5 --no-cpu-baseline --no-glue-roofline > gpurun_out/r2q_bench_512deep.json 2> gpurun_out/r2q_bench_512deep.err; echo "bench 512deep rc=$?" | tee -a gpurun_out/r2q_summary.txt
tail -3 gpurun_out/r2q_tests.log
head -4 gpurun_out/r2q_timeline.log
grep -E "conv_ring|conv_igemm_kernel<64>" gpurun_out/r2q_timeline.log | head
python tools/show_bench.py gpurun_out/r2q_bench_1gpu.json 2>/dev/null | head -2
python tools/show_bench.py gpurun_out/r2q_bench_512deep.json 2>/dev/null | head -1
