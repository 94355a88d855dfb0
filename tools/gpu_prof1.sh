#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu > gpurun_out/parity.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/parity.log
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.txt 2>&1; echo "conv_bench exit $?"; cat gpurun_out/conv_bench.txt
# launch list of one bench step (shares), then full capture of the conv kernel on two layer shapes
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 300 python tools/conv_bench.py --only enc.1 --iters 2 > gpurun_out/cb_enc1.txt 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 2 -c 2 -o gpurun_out/prof_conv_enc1 -f python tools/conv_bench.py --only enc.1 --iters 2 --what fwd > gpurun_out/ncu_enc1.log 2>&1; echo "ncu enc1 exit $?"
timeout 300 python tools/conv_bench.py --only up.0 --iters 2 > gpurun_out/cb_up0.txt 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 2 -c 4 -o gpurun_out/prof_conv_up0 -f python tools/conv_bench.py --only up.0 --iters 2 --what fwd,wgrad > gpurun_out/ncu_up0.log 2>&1; echo "ncu up0 exit $?"
ls -la gpurun_out/
