#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_outconv_gpu.py -x -q -m gpu > gpurun_out/kt_outconv.log 2>&1; echo "outconv tests exit $?"; tail -3 gpurun_out/kt_outconv.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu 2>&1 | tail -4
timeout 300 python tools/conv_bench.py --only out > gpurun_out/cb_out_fold.txt 2>&1; cat gpurun_out/cb_out_fold.txt
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; grep -v Warning gpurun_out/bench.err | tail -3
python tools/show_bench.py gpurun_out/bench.json
timeout 250 python tools/step_timeline.py --e2e-steps 30 > gpurun_out/timeline2.txt 2>&1; head -40 gpurun_out/timeline2.txt
