#!/bin/bash
# N-GPU scaling points of bench.py (the driver runs the same at round end); run under `gpurun --gpus 8`
mkdir -p gpurun_out
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 30 --warmup 5 --no-glue-roofline > gpurun_out/r2s_bench_${N}gpu.json 2> gpurun_out/r2s_bench_${N}gpu.err; echo "bench$N rc=$?" | tee -a gpurun_out/r2s_summary.txt
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 --no-glue-roofline --deep --size 512 --batch 8 > gpurun_out/r2s_bench_8gpu_512deep.json 2> gpurun_out/r2s_bench_8gpu_512deep.err; echo "bench8 512deep rc=$?" | tee -a gpurun_out/r2s_summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 --no-glue-roofline --global-batch 256 > gpurun_out/r2s_bench_8gpu_gb256.json 2> gpurun_out/r2s_bench_8gpu_gb256.err; echo "bench8 gb256 rc=$?" | tee -a gpurun_out/r2s_summary.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2s_bench_1gpu.json 2> gpurun_out/r2s_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2s_summary.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'N', d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), d['clocks'])
    except Exception as e:
        print(f, 'parse failed', e)
PY
