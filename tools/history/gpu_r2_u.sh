#!/bin/bash
# validation of the tiled weight prep + 2-rank tests and bench; run under `gpurun --gpus 2`
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_determinism_gpu.py tests/test_updown_gpu.py tests/test_elr_gpu.py tests/test_f2_gpu.py -x -q -m gpu > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2u_summary.txt
timeout 900 python -m pytest tests/test_ddp_gpu.py tests/test_xrank_gpu.py -x -q -m gpu -s > gpurun_out/r2u_ddp.log 2>&1; echo "ddp+xrank rc=$?" | tee -a gpurun_out/r2u_summary.txt
timeout 300 python tools/step_timeline.py > gpurun_out/r2u_timeline.log 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2u_summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 30 --warmup 5 --no-glue-roofline > gpurun_out/r2u_bench_2gpu.json 2> gpurun_out/r2u_bench_2gpu.err; echo "bench2 rc=$?" | tee -a gpurun_out/r2u_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2u_bench_1gpu.json 2> gpurun_out/r2u_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2u_summary.txt
tail -3 gpurun_out/r2u_tests.log
grep -E "passed|failed|ddp " gpurun_out/r2u_ddp.log | tail -8
head -4 gpurun_out/r2u_timeline.log
grep -E "weight_prep" gpurun_out/r2u_timeline.log | head -3
python tools/show_bench.py gpurun_out/r2u_bench_1gpu.json 2>/dev/null | head -1
python tools/show_bench.py gpurun_out/r2u_bench_2gpu.json 2>/dev/null | head -1
