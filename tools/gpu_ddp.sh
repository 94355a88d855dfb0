#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1; echo "ddp_check exit $?"; grep -E "ddp_check|Error|error|assert" gpurun_out/ddp_check_$N.log | tail -8
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit $?"; grep -E "Error|error|Traceback" -A3 gpurun_out/bench_${N}gpu.err | tail -12
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'n', d['n_gpus'])"
FACEVAE_CUDA_GRAPH_DDP=0 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_eager.json 2> gpurun_out/bench_${N}gpu_eager.err; echo "bench eager exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_eager.json')); print('EAGER value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'n', d['n_gpus'], d['config'].get('cuda_graph'))"
