#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_outconv_gpu.py -x -q -m gpu > gpurun_out/kt_outconv.log 2>&1; echo "outconv tests exit $?"; tail -40 gpurun_out/kt_outconv.log
