#!/bin/bash
# run under `gpurun --gpus 2`: the data-parallel scenarios that exercise the trainer step (rank skew, side stream) + the 2-GPU bench line
mkdir -p gpurun_out
P=gpurun_out/d2
timeout 200 python -m pytest tests/test_ddp_gpu.py -x -q -m gpu -s -k "skew or wstream" > ${P}_ddp.log 2>&1; echo "ddp rc=$?" | tee ${P}_summary.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 30 --warmup 5 --no-glue-roofline > ${P}_bench_2gpu.json 2> ${P}_bench_2gpu.err; echo "bench2 rc=$?" | tee -a ${P}_summary.txt
grep -E "passed|failed|ddp |Error|error" ${P}_ddp.log | tail -8
python tools/show_bench.py ${P}_bench_2gpu.json 2>/dev/null | head -1 | cut -c1-120
tail -3 ${P}_bench_2gpu.err
