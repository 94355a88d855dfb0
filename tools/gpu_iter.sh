#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; grep -v Warning gpurun_out/bench.err | tail -3
python tools/show_bench.py gpurun_out/bench.json
timeout 250 python tools/step_timeline.py --e2e-steps 30 > gpurun_out/timeline3.txt 2>&1; grep -v "Warn\|warn" gpurun_out/timeline3.txt | head -48
