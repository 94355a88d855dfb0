#!/bin/bash
# bench.py at N GPUs (graph-captured data-parallel step); $1 = N
N=${1:-8}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
if [ "$N" = "1" ]; then
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit $?"
else
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit $?"
fi
grep -E "Error|error|Traceback" -A3 gpurun_out/bench_${N}gpu.err | tail -12
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'graph', d['config'].get('cuda_graph'), 'clocks', d['clocks'])"
