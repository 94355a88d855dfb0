#!/bin/bash
# round-end style validation: all GPU tests, smoke, bench (both arms), ncu launch list + dram traffic of the conv kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm exit $?"
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python tools/show_bench.py gpurun_out/bench.json > gpurun_out/bench_summary.txt; cat gpurun_out/bench_summary.txt
# eager step under ncu: launch list (time) and DRAM traffic of every kernel of ~2 steps
FACEVAE_CUDA_GRAPH=0 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err &&
FACEVAE_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 900 -c 520 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/ncu_final.log 2>&1; echo "ncu launch list exit $?"
# one full-set capture of the dominant kernel (wide-layer implicit GEMM) inside the eager step: DRAM traffic, pipe utilisation
FACEVAE_CUDA_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 22 -c 3 -o gpurun_out/prof_igemm -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/ncu_igemm.log 2>&1; echo "ncu igemm exit $?"
