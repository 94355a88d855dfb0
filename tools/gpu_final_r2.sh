#!/bin/bash
# Final round-2 evidence (1 GPU): all GPU tests, smoke, both bench arms, timeline, other configs, ncu launch list, full-set capture of
# the kernels changed since the full-step capture.
mkdir -p gpurun_out
P=gpurun_out/r2p
rm -f ${P}_infer.jsonl
timeout 1200 python -m pytest tests -x -q -m gpu > ${P}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee ${P}_summary.txt
timeout 300 python __graft_entry__.py smoke > ${P}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a ${P}_summary.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > ${P}_bench_reference_cpu.json 2> ${P}_bench_reference_cpu.err; echo "reference arm rc=$?" | tee -a ${P}_summary.txt
timeout 600 python bench.py --steps 30 --warmup 5 > ${P}_bench_1gpu.json 2> ${P}_bench_1gpu.err; echo "bench rc=$?" | tee -a ${P}_summary.txt
timeout 300 python tools/step_timeline.py > ${P}_timeline.log 2>&1; echo "timeline rc=$?" | tee -a ${P}_summary.txt
timeout 300 python bench.py --deep --size 512 --batch 8 --steps 20 --warmup 5 --no-cpu-baseline --no-glue-roofline > ${P}_bench_512deep_1gpu.json 2> ${P}_bench_512deep_1gpu.err; echo "bench 512deep rc=$?" | tee -a ${P}_summary.txt
for B in 1 8 64 256 1024; do
  timeout 200 python bench.py --infer --batch $B --steps 20 --warmup 5 --no-cpu-baseline --no-glue-roofline >> ${P}_infer.jsonl 2>> ${P}_infer.err; echo "infer $B rc=$?" | tee -a ${P}_summary.txt
done
FACEVAE_CUDA_GRAPH=0 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-glue-roofline --profile-steps 1 > ${P}_bench_eager.json 2> ${P}_bench_eager.err &&
FACEVAE_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-glue-roofline --profile-steps 1 > ${P}_ncu_list.log 2>&1; echo "ncu launch list rc=$?" | tee -a ${P}_summary.txt
STEPS=1 timeout 300 python tools/ncu_one.py > ${P}_ncu_plain.log 2>&1 &&
STEPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'pw_bwd_reduce_pipe|weight_prep_flat|slab_sum_wide|conv_igemm|conv_ring' -c 40 -o ${P}_changed -f python tools/ncu_one.py > ${P}_ncu_changed.log 2>&1; echo "ncu changed rc=$?" | tee -a ${P}_summary.txt
ncu -i ${P}_changed.ncu-rep --page raw --csv > ${P}_changed_raw.csv 2> ${P}_changed_raw.err
rm -f ${P}_changed.ncu-rep
tail -3 ${P}_pytest_gpu.log; tail -2 ${P}_smoke.log
python tools/show_bench.py ${P}_bench_1gpu.json 2>/dev/null | head -1
cat ${P}_summary.txt
