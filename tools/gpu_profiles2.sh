#!/bin/bash
# Second half of the round-2 evidence run: `ncu --set full` over one eager step (all kernel families), per-tensor parity table
mkdir -p gpurun_out
P=gpurun_out/r2p
timeout 900 python tools/parity_report.py > ${P}_parity_table.md 2> ${P}_parity_table.err; echo "parity table rc=$?" | tee -a ${P}_summary2.txt
STEPS=1 timeout 300 python tools/ncu_one.py > ${P}_ncu_plain.log 2>&1 &&
STEPS=1 timeout 2400 ncu --set full --clock-control none --import-source on -k regex:'conv_|fold_|bn_|pw_|recon|reparam|colsum|wgrad|slab|adam|weight_prep' -c 200 -o ${P}_full -f python tools/ncu_one.py > ${P}_ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a ${P}_summary2.txt
ncu -i ${P}_full.ncu-rep --page raw --csv > ${P}_full_raw.csv 2> ${P}_full_raw.err
for K in conv_igemm_kernel conv_win_kernel conv_ring_kernel conv_wgrad_kernel conv_wgrad_ring_kernel fold_conv_kernel fold_wgrad_kernel bn_act_bwd_reduce_kernel bn_stats_kernel pw_bwd_reduce_kernel bn_act_bwd_apply_kernel bn_act_fwd_kernel; do
  ncu -i ${P}_full.ncu-rep --page source --csv --kernel-name regex:$K --launch-count 1 > ${P}_src_$K.csv 2>> ${P}_full_raw.err
done
# the enc.1 data-gradient ring kernel is the 2nd conv_ring launch, the widest igemm (up.2 forward) the 16th conv_igemm launch
ncu -i ${P}_full.ncu-rep --page source --csv --kernel-name regex:conv_ring_kernel --launch-skip 1 --launch-count 1 > ${P}_src_conv_ring_kernel_dgrad.csv 2>> ${P}_full_raw.err
ls -la gpurun_out | grep r2p_ | grep -v bench
sz=$(stat -c %s ${P}_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 30000000 ]; then rm -f ${P}_full.ncu-rep; echo "removed ncu-rep ($sz bytes)"; fi
tail -3 ${P}_ncu_full.log
tail -5 ${P}_parity_table.md
