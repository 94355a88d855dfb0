#!/bin/bash
# 2-GPU call: multi-rank tests + scaling point
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/r2f_$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/r2f_summary.txt; }
run elr python -m pytest tests/test_elr_gpu.py -q -m gpu
run ddp python -m pytest tests/test_ddp_gpu.py -q -m gpu -s
run layerwise python -m pytest tests/test_layerwise_gpu.py -q -m gpu -s
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; echo "bench2 rc=$?" | tee -a gpurun_out/r2f_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2f_summary.txt
for f in elr ddp layerwise; do echo "== $f"; tail -6 gpurun_out/r2f_$f.log; done
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open(f'gpurun_out/r2f_bench_{n}gpu.json').read().strip().splitlines()[-1])
        print('N', d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
    except Exception as e:
        print(n, 'bench parse failed', e)
PY
