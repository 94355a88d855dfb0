#!/usr/bin/env python
"""bench.py -- train images/sec of the face-vae anchor at 256x256 (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (the reference's CPU train step on the host cores)

A step is one full train step of the hot path over one synthetic batch: zero_grad -> encoder / bottleneck / decoder
forward -> 0.2*KL + 10*MSE -> backward -> Adam (the skeleton of reference logger.py:150-164).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train images/sec at 256x256"
UNIT = "images/sec"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 7:
                    self.samples.append((time.perf_counter(), parts))
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t0: float, t1: float):
        rows = [p for t, p in self.samples if t0 <= t <= t1] or [p for _, p in self.samples[-3:]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(rows)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (port in oracle/cpu_train.py), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle.cpu_train import time_cpu_train
    threads = os.cpu_count() or 1
    n = args.cpu_batch
    ips, med = time_cpu_train(n, args.size, args.size, max(args.steps, 1), max(args.warmup, 1), threads)
    sample = f"{max(args.steps, 1)} train steps of batch {n} at {args.size}x{args.size} after {max(args.warmup, 1)} warm-up, fp32 eager, median step"
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"face-vae anchor train step, batch {n} per step at {args.size}x{args.size} on host CPU", "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (BASELINE.json configs[1]: 32 at 256x256)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--cpu-batch", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--deep", action="store_true", help="512x512-deep variant (BASELINE.json configs[3]; use with --size 512 --batch 8)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # libraries (NCCL's version banner, warnings) must not pollute the single JSON line on stdout
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: face_vae_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", init_method="env://", world_size=world, rank=rank)
    from face_vae_b200 import _lib
    from face_vae_b200.models import FaceVAE
    from face_vae_b200.trainer import VAETrainer
    _lib.call("fv_device_ok")
    peaks = _peaks()
    warmup = max(args.warmup, 3)
    B, S = args.batch, args.size

    torch.manual_seed(0)                               # identical initial weights on every rank
    if args.deep:
        from face_vae_b200.models import face_vae_512
        model = face_vae_512().cuda().train()
    else:
        model = FaceVAE().cuda().train()
    trainer = VAETrainer(model)
    dz = model.latent_dim(S, S)
    g = torch.Generator().manual_seed(1 + rank)        # reference seed rule (distributed.py:10)
    host = [(torch.rand((B, 3, S, S), generator=g).pin_memory(), torch.randn((B, dz), generator=g).pin_memory()) for _ in range(2)]
    dev = [(x.cuda(), e.cuda()) for x, e in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        trainer.step(*dev[i % 2])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.15)

    # ---- device-timed region: inputs resident in HBM ------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = _lib.launch_count
    t_start = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        trainer.step(*dev[i % 2])
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = _lib.launch_count - l0
    if trainer.use_cuda_graph and trainer.launches_per_step:
        # the step is a replayed CUDA graph: the C-ABI launches were counted when it was captured
        launches = trainer.launches_per_step * args.steps
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end: pinned host -> device copy of each batch and device -> host read of the loss inside the region
    from face_vae_b200.data import AsyncScalarLog, DevicePrefetcher
    # set-up outside the timed region: the pinned loss buffer (cudaHostAlloc synchronises the device) and two untimed
    # passes so that the caching allocator already holds the prefetch buffers
    log = AsyncScalarLog(args.steps)
    prefetch = DevicePrefetcher()
    for x, e in prefetch.over(host[i % 2] for i in range(3)):
        trainer.step(x, e)
    barrier()
    t0 = time.perf_counter()
    # every step: its batch comes from pinned host memory (copy overlapped with the previous step on a side stream)
    # and its loss goes back to the host (asynchronous copy into pinned memory, read after the loop)
    for x, e in prefetch.over(host[i % 2] for i in range(args.steps)):
        losses, _ = trainer.step(x, e)
        log.push(torch.stack([v.detach() for v in losses.values()]).sum())
    last = float(log.values()[-1])
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_value = world * B * args.steps / e2e_s
    t_clock_end = time.perf_counter()
    sampler.stop()
    clocks = sampler.summary(t_start, t_clock_end)

    # ---- per-kernel timing pass (CUDA events around every C-ABI launch, on the launching stream) ----------
    kernels, roofline = {}, None
    if rank == 0 or world == 1:
        pass
    graph_mode = trainer.use_cuda_graph
    trainer.use_cuda_graph = False                 # eager replay of the same step: one event pair per launch
    trainer.step(*dev[0])
    _lib.profile_start()
    for i in range(args.profile_steps):
        trainer.step(*dev[i % 2])
    recs = _lib.profile_stop()
    trainer.use_cuda_graph = graph_mode
    agg = {}
    for name, t_ms, meta in recs:
        a = agg.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "flops_exec": 0.0, "bytes": 0.0, "big_bytes": 0.0, "big_ms": 0.0})
        a["launches"] += 1
        a["ms"] += t_ms
        if meta:
            a["flops"] += meta.get("flops", 0.0)
            a["flops_exec"] += meta.get("flops_exec", 0.0)
            a["bytes"] += meta.get("bytes", 0.0)
            if meta.get("bytes", 0.0) >= 64e6:        # launches that move >= 64 MB: bandwidth-bound, not launch / latency-bound
                a["big_bytes"] += meta["bytes"]
                a["big_ms"] += t_ms
    total_ms = sum(a["ms"] for a in agg.values()) or 1.0
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        k = {"launches_per_step": a["launches"] / args.profile_steps, "ms_per_step": a["ms"] / args.profile_steps,
             "share_of_kernel_time": a["ms"] / total_ms}
        if a["flops"]:
            k["tflops"] = a["flops"] / (a["ms"] * 1e-3) / 1e12
            k["tflops_executed"] = a["flops_exec"] / (a["ms"] * 1e-3) / 1e12
        if a["bytes"]:
            k["gbs"] = a["bytes"] / (a["ms"] * 1e-3) / 1e9
            k["frac_hbm"] = k["gbs"] / peaks["hbm"]
            if a["big_ms"]:
                k["gbs_launches_over_64MB"] = a["big_bytes"] / (a["big_ms"] * 1e-3) / 1e9
                k["frac_hbm_launches_over_64MB"] = k["gbs_launches_over_64MB"] / peaks["hbm"]
        kernels[name] = k
    # forward + data-gradient convolutions: the generic implicit-GEMM / ring kernels and the tap-folded out_conv kernels
    conv = {"ms": 0.0, "flops": 0.0, "launches": 0}
    for nm in ("fv_conv2d", "fv_conv2d_stats", "fv_outconv_fwd", "fv_outconv_dgrad"):
        if nm in agg:
            for k in conv:
                conv[k] += agg[nm][k]
    if conv["ms"] > 0:
        ach = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
        roofline = {"kernel": "conv_igemm_kernel / conv_ring_kernel / fold_conv_kernel (all forward + data-gradient convolutions of the step)", "bound": "tensor",
                    "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                    "frac_of_burst_peak": ach / peaks["bf16"], "peak_source": peaks["source"] + " (sustained: timed inside the step)",
                    "avg_launch_ms": conv["ms"] / conv["launches"], "launches_per_step": conv["launches"] / args.profile_steps,
                    "traffic": None}
        # DRAM bytes per launch of these kernels from the committed ncu launch list of this command (profiles/)
        tpath = os.path.join(ROOT, "profiles", "r01_final_conv_traffic.json")
        if os.path.exists(tpath) and B == 32 and S == 256 and not args.deep:
            t = json.load(open(tpath))
            roofline["traffic"] = t["bytes_per_launch"]
            roofline["traffic_source"] = "profiles/r01_final_conv_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, average per launch)"

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_train import time_cpu_train
        threads = os.cpu_count() or 1
        ips, med = time_cpu_train(args.cpu_batch, S, S, 6, 2, threads)
        cpu = {"value": ips, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"6 train steps of batch {args.cpu_batch} at {S}x{S} after 2 warm-up (median {med * 1e3:.0f} ms/step), fp32 eager ATen ops"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"face-vae {'512-deep variant' if args.deep else 'anchor'} (SURVEY.md section 8) train step, batch {B} per GPU at {S}x{S}, "
                                       f"0.2*KL + 10*MSE, Adam(5e-5, betas 0.5/0.999), bf16 storage / fp32 accumulate",
                           "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": bool(trainer.use_cuda_graph),
                           "l2": "working set per step (activations + gradients, >3 GB) far exceeds the 126 MB L2; inputs alternate between two batches"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 3 * S * S * 4 + B * dz * 4, "d2h_bytes_per_step": 4,
                        "last_loss": last},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels,
                "peaks": peaks}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        # captured NCCL collectives must be released before the communicator goes away; leave without running the
        # interpreter's teardown (a CUDA graph holding NCCL kernels can block process-group destruction)
        trainer._graph = None
        trainer._static_out = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
