#!/bin/bash
# peer-memory gradient all-reduce vs NCCL at 2 GPUs (bench), 1-GPU bench with the NVML sampler; run under `gpurun --gpus 2`
mkdir -p gpurun_out
run2() {
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 30 --warmup 5 --no-glue-roofline > gpurun_out/r2m_bench_2gpu_$name.json 2> gpurun_out/r2m_bench_2gpu_$name.err; echo "bench2 $name rc=$?" | tee -a gpurun_out/r2m_summary.txt
}
run2 peer FOO=1
run2 nccl1 FACEVAE_GRAD_XRANK=0 FACEVAE_BUCKET_MB=1000
run2 nccl2mb FACEVAE_GRAD_XRANK=0
run2 peer_again FOO=1
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2m_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2m_bench_1gpu_b.json 2> gpurun_out/r2m_bench_1gpu_b.err; echo "bench1b rc=$?" | tee -a gpurun_out/r2m_summary.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2m_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'N', d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), d['clocks'])
    except Exception as e:
        print(f, 'parse failed', e)
PY
tail -3 gpurun_out/r2m_bench_2gpu_peer.err
