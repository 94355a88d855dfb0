#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv or wgrad" -s > gpurun_out/kt_conv.log 2>&1; echo "conv tests exit $?"; tail -2 gpurun_out/kt_conv.log; grep -E "fv:|Error|wgrad-ring|bad frac" gpurun_out/kt_conv.log | head -30
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.txt 2>&1; echo "conv_bench exit $?"; cat gpurun_out/conv_bench.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu > gpurun_out/parity.log 2>&1; echo "parity exit $?"; tail -3 gpurun_out/parity.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'roof', d['roofline']['achieved'])
for k,v in d['kernels'].items(): print(f"{k:24s} {v['launches_per_step']:5.0f} {v['ms_per_step']:7.3f} ms  {v.get('tflops','')} {v.get('gbs','')}")
PY
