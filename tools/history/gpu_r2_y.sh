#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_determinism_gpu.py -x -q -m gpu > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2y_summary.txt
rm -f gpurun_out/r2y_infer.jsonl
for B in 1 8 64 256 1024; do
  timeout 200 python bench.py --infer --batch $B --steps 30 --warmup 5 --no-cpu-baseline --no-glue-roofline >> gpurun_out/r2y_infer.jsonl 2>> gpurun_out/r2y_infer.err; echo "infer $B rc=$?" | tee -a gpurun_out/r2y_summary.txt
done
tail -3 gpurun_out/r2y_tests.log
python - <<'PY'
import json
for l in open('gpurun_out/r2y_infer.jsonl'):
    d = json.loads(l)
    print(d['config']['workload'][-50:], 'graph', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms | eager', round(d['eager']['value'], 1), round(d['eager']['ms_per_step'], 3))
PY
tail -3 gpurun_out/r2y_infer.err
