#!/bin/bash
# peer-memory gradient all-reduce: 2-rank tests + bench comparison; run under `gpurun --gpus 2`
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_gpu.py tests/test_xrank_gpu.py -x -q -m gpu -s > gpurun_out/r2l_ddp.log 2>&1; echo "ddp+xrank rc=$?" | tee -a gpurun_out/r2l_summary.txt
run2() {
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 30 --warmup 5 --no-glue-roofline > gpurun_out/r2l_bench_2gpu_$name.json 2> gpurun_out/r2l_bench_2gpu_$name.err; echo "bench2 $name rc=$?" | tee -a gpurun_out/r2l_summary.txt
}
run2 peer FOO=1
run2 nccl1 FACEVAE_GRAD_XRANK=0 FACEVAE_BUCKET_MB=1000
run2 peer_again FOO=1
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2l_bench_1gpu.json 2> gpurun_out/r2l_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2l_summary.txt
grep -E "passed|failed|error|ddp " gpurun_out/r2l_ddp.log | tail -12
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2l_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'N', d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
    except Exception as e:
        print(f, 'parse failed', e)
PY
tail -5 gpurun_out/r2l_bench_2gpu_peer.err
