import json, sys
d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.1f} img/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}  conv TF/s {d['roofline']['achieved']:.1f}  graph {d['config'].get('cuda_graph')}  clocks {d['clocks']}")
for k, v in d['kernels'].items():
    print(f"{k:24s} {v['launches_per_step']:5.0f} {v['ms_per_step']:7.3f} ms  {v.get('tflops', '')} {v.get('gbs', '')}")
