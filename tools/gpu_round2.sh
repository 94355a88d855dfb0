#!/bin/bash
# parity tests + smoke + bench + ncu launch list, each step logged under gpurun_out/
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "bn_ or layout" > gpurun_out/kt_glue.log 2>&1; echo "glue tests exit $?"; tail -3 gpurun_out/kt_glue.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -s > gpurun_out/parity.log 2>&1; echo "parity exit $?"; grep -E "passed|failed|FAILED|Error|worst" gpurun_out/parity.log | head -40
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json
