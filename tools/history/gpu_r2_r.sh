#!/bin/bash
# validation: filter-resident mode of the implicit-GEMM kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_updown_gpu.py tests/test_parity_gpu.py tests/test_layerwise_gpu.py tests/test_determinism_gpu.py tests/test_elr_gpu.py tests/test_f2_gpu.py -x -q -m gpu > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2r_summary.txt
timeout 300 python tools/step_timeline.py > gpurun_out/r2r_timeline.log 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2r_summary.txt
FV_CONV_BRES=0 timeout 300 python tools/step_timeline.py > gpurun_out/r2r_timeline_nobres.log 2>&1; echo "timeline nobres rc=$?" | tee -a gpurun_out/r2r_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2r_bench_1gpu.json 2> gpurun_out/r2r_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2r_summary.txt
timeout 300 python bench.py --deep --size 512 --batch 8 --steps 20 --warmup 5 --no-cpu-baseline --no-glue-roofline > gpurun_out/r2r_bench_512deep.json 2> gpurun_out/r2r_bench_512deep.err; echo "bench 512deep rc=$?" | tee -a gpurun_out/r2r_summary.txt
tail -3 gpurun_out/r2r_tests.log
head -4 gpurun_out/r2r_timeline.log
grep -E "conv_igemm_kernel<" gpurun_out/r2r_timeline.log | head
head -1 gpurun_out/r2r_timeline_nobres.log
grep -E "conv_igemm_kernel<" gpurun_out/r2r_timeline_nobres.log | tail -2
python tools/show_bench.py gpurun_out/r2r_bench_1gpu.json 2>/dev/null | head -1
python tools/show_bench.py gpurun_out/r2r_bench_512deep.json 2>/dev/null | head -1
