#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "bn_ or layout or recon or reparam" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; grep -v Warning gpurun_out/bench.err | tail -3
python tools/show_bench.py gpurun_out/bench.json
