#!/bin/bash
# Round-2 evidence run (1 GPU): both bench arms, timeline, other configs, ncu launch list of the eager bench step, and
# `ncu --set full` captures of every kernel family of one eager step (exported to CSV on the box: the .ncu-rep is too big to merge).
mkdir -p gpurun_out
P=gpurun_out/r2p
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > ${P}_bench_reference_cpu.json 2> ${P}_bench_reference_cpu.err; echo "reference arm rc=$?" | tee -a ${P}_summary.txt
timeout 600 python bench.py --steps 30 --warmup 5 > ${P}_bench_1gpu.json 2> ${P}_bench_1gpu.err; echo "bench rc=$?" | tee -a ${P}_summary.txt
timeout 300 python tools/step_timeline.py > ${P}_timeline.log 2>&1; echo "timeline rc=$?" | tee -a ${P}_summary.txt
timeout 300 python bench.py --deep --size 512 --batch 8 --steps 20 --warmup 5 --no-cpu-baseline --no-glue-roofline > ${P}_bench_512deep_1gpu.json 2> ${P}_bench_512deep_1gpu.err; echo "bench 512deep rc=$?" | tee -a ${P}_summary.txt
for B in 1 8 64 256 1024; do
  timeout 200 python bench.py --infer --batch $B --steps 20 --warmup 5 --no-cpu-baseline --no-glue-roofline >> ${P}_infer.jsonl 2>> ${P}_infer.err; echo "infer $B rc=$?" | tee -a ${P}_summary.txt
done
# launch list of the eager bench step (cold-cache, serialised: shares, not absolutes)
FACEVAE_CUDA_GRAPH=0 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-glue-roofline --profile-steps 1 > ${P}_bench_eager.json 2> ${P}_bench_eager.err &&
FACEVAE_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-glue-roofline --profile-steps 1 > ${P}_ncu_list.log 2>&1; echo "ncu launch list rc=$?" | tee -a ${P}_summary.txt
# full-set capture of one eager step, all kernel families
STEPS=1 timeout 300 python tools/ncu_one.py > ${P}_ncu_plain.log 2>&1 &&
STEPS=1 timeout 2400 ncu --set full --clock-control none --import-source on -k regex:'fv::' -c 260 -o ${P}_full -f python tools/ncu_one.py > ${P}_ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a ${P}_summary.txt
ncu -i ${P}_full.ncu-rep --page raw --csv > ${P}_full_raw.csv 2> ${P}_full_raw.err
for K in conv_igemm_kernel conv_win_kernel conv_ring_kernel conv_wgrad_kernel conv_wgrad_ring_kernel fold_conv_kernel fold_wgrad_kernel bn_act_bwd_reduce_kernel bn_stats_kernel pw_bwd_reduce_kernel bn_act_bwd_apply_kernel bn_act_fwd_kernel; do
  ncu -i ${P}_full.ncu-rep --page source --csv --kernel-name regex:$K --launch-count 1 > ${P}_src_$K.csv 2>> ${P}_full_raw.err
done
ls -la gpurun_out | grep r2p
sz=$(stat -c %s ${P}_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 30000000 ]; then rm -f ${P}_full.ncu-rep; echo "removed ncu-rep ($sz bytes)"; fi
tail -3 ${P}_ncu_full.log
python tools/show_bench.py ${P}_bench_1gpu.json | head -30
