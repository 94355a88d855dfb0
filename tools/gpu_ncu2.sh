#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for L in enc.1 out up.0; do
  timeout 300 python tools/conv_bench.py --only $L --iters 2 --what fwd > gpurun_out/cb_$L.txt 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 2 -c 1 -o gpurun_out/prof2_$L -f python tools/conv_bench.py --only $L --iters 2 --what fwd > gpurun_out/ncu2_$L.log 2>&1; echo "ncu $L exit $?"
done
ls -la gpurun_out/*.ncu-rep
