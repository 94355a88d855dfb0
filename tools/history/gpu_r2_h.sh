#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/r2h_$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/r2h_summary.txt; }
run updown python -m pytest tests/test_updown_gpu.py -q -m gpu -x
run elr python -m pytest tests/test_elr_gpu.py -q -m gpu
run kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "recon or colsum or bn_eval"
run determinism python -m pytest tests/test_determinism_gpu.py -q -m gpu
run layerwise python -m pytest tests/test_layerwise_gpu.py -q -m gpu -s
run ddp python -m pytest tests/test_ddp_gpu.py -q -m gpu -s
run timeline python tools/step_timeline.py --e2e-steps 50
FV_CONV_WIN=0 timeout 600 python tools/step_timeline.py --e2e-steps 50 > gpurun_out/r2h_timeline_nowin.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-glue-roofline > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err; echo "bench2 rc=$?" | tee -a gpurun_out/r2h_summary.txt
FACEVAE_FUSE_XRANK=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 --no-glue-roofline > gpurun_out/r2h_bench_2gpu_nofuse.json 2> gpurun_out/r2h_bench_2gpu_nofuse.err; echo "bench2 nofuse rc=$?" | tee -a gpurun_out/r2h_summary.txt
for f in updown elr kernels determinism layerwise ddp; do echo "== $f"; tail -4 gpurun_out/r2h_$f.log; done; head -5 gpurun_out/r2h_timeline.log; head -3 gpurun_out/r2h_timeline_nowin.log
python - <<'PY'
import json
for n in ('2gpu', '2gpu_nofuse'):
    try:
        d = json.loads(open(f'gpurun_out/r2h_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
    except Exception as e:
        print(n, 'bench parse failed', e)
PY
