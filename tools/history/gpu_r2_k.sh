#!/bin/bash
# 2-GPU critical-path experiments + validation of the ring x2 weight gradient; run under `gpurun --gpus 2`
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_updown_gpu.py tests/test_determinism_gpu.py -x -q -m gpu > gpurun_out/r2k_updown.log 2>&1; echo "updown+determinism rc=$?" | tee -a gpurun_out/r2k_summary.txt
timeout 300 python tools/step_timeline.py > gpurun_out/r2k_timeline.log 2>&1; echo "timeline rc=$?" | tee -a gpurun_out/r2k_summary.txt
run2() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --no-glue-roofline > gpurun_out/r2k_bench_2gpu_$name.json 2> gpurun_out/r2k_bench_2gpu_$name.err; echo "bench2 $name rc=$?" | tee -a gpurun_out/r2k_summary.txt
}
run2 default FOO=1
run2 bucket1000 FACEVAE_BUCKET_MB=1000
run2 bucket8 FACEVAE_BUCKET_MB=8
run2 maxctas2 NCCL_MAX_CTAS=2
run2 maxctas2_b4 NCCL_MAX_CTAS=2 FACEVAE_BUCKET_MB=4
run2 noreduce FACEVAE_DIAG_SKIP_GRAD_REDUCE=1
run2 noxrank FACEVAE_XRANK=0
timeout 300 python bench.py --steps 30 --warmup 5 --no-glue-roofline --no-cpu-baseline > gpurun_out/r2k_bench_1gpu.json 2> gpurun_out/r2k_bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/r2k_summary.txt
tail -5 gpurun_out/r2k_updown.log
head -5 gpurun_out/r2k_timeline.log
grep -n "wgrad" gpurun_out/r2k_timeline.log | head
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2k_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'N', d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1))
    except Exception as e:
        print(f, 'parse failed', e)
PY
