#!/bin/bash
# $1 = kernel regex, $2 = output tag, $3 = launches to skip, $4 = launches to capture
mkdir -p gpurun_out
python tools/ncu_one.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-0} -c ${4:-2} -o gpurun_out/prof_$2 -f python tools/ncu_one.py > gpurun_out/ncu_$2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$2.log
