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
    """Samples SM clocks and throttle reasons while the timed region runs: NVML in-process (a few microseconds per query, no child
    process competing for the driver), `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_s: float = 0.025):
        super().__init__(daemon=True)
        self.index = index
        self.period = period_s
        self.samples = []          # (t, [sm, max_sm, power, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap]) as strings
        self.proc = None
        self._halt = threading.Event()
        self.source = "nvml"

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self._visible_index())
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = (nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown, nv.nvmlClocksEventReasonSwThermalSlowdown,
                nv.nvmlClocksEventReasonSwPowerCap)
        while not self._halt.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
            except Exception:
                pw = 0.0
            self.samples.append((time.perf_counter(), [str(sm), str(mx), f"{pw:.1f}"] + ["Active" if r & b else "Not Active" for b in bits]))
            self._halt.wait(self.period)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            self.source = "nvidia-smi"
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
        self._halt.set()
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
                "samples": len(rows), "source": self.source}


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


CONV_KERNELS = ("conv_igemm_kernel", "conv_ring_kernel", "conv_wgrad_kernel", "conv_wgrad_ring_kernel", "fold_conv_kernel", "fold_wgrad_kernel",
                "wgrad_finish_kernel", "wgrad_finish_up_kernel", "slab_sum_kernel")


def _in_graph_shares(torch, trainer, dev, replays=3):
    """Per-kernel device time inside the replayed CUDA graph (CUPTI activity records through torch.profiler).  Only SHARES are
    used: absolute times under a profiler are never reported as bench values."""
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(replays):
            trainer.step(*dev[i % 2])
        torch.cuda.synchronize()
    per = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            d = per.setdefault(ev.name, [0, 0.0])
            d[0] += 1
            d[1] += float(ev.time_range.end - ev.time_range.start)
    total = sum(v[1] for v in per.values()) or 1.0
    conv = sum(v[1] for k, v in per.items() if any(c in k for c in CONV_KERNELS))
    top = sorted(per.items(), key=lambda kv: -kv[1][1])[:24]
    return {"conv_share": conv / total,
            "kernels": {k[:96]: {"launches_per_step": v[0] / replays, "share_of_kernel_time": v[1] / total} for k, v in top}}


def _glue_roofline(torch, peaks):
    """north_star: "fused loss / KL kernels at >= 70 % of HBM bandwidth" -- the reconstruction-loss (+ sigmoid + gradient) and
    re-parameterisation + KL kernels timed alone at a bandwidth-bound size (the inference sweep's batch 1024 for the loss, 8192
    samples for the latent: inputs far larger than the 126 MB L2), CUDA events, median of 10 after 3 warm-ups."""
    from face_vae_b200 import ops
    out = {}

    def timed(fn, nbytes):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = statistics.median(ts)
        gbs = nbytes / (t * 1e-3) / 1e9
        return {"ms": t, "bytes": nbytes, "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"], "bound": "hbm"}

    n, c, h, w = 1024, 3, 256, 256
    logits = torch.randn((n, c, h, w), device="cuda")
    target = torch.rand((n, c, h, w), device="cuda")
    e = logits.numel()
    # sigmoid + MSE + gradient: reads logits and target, writes the fp32 gradient (12 bytes per element)
    out["fv_recon_loss (sigmoid + MSE + gradient, fp32 in / fp32 grad)"] = timed(
        lambda: ops.recon_loss(logits, target, False, True, 1.0 / e, False, True, False), 12.0 * e)
    out["fv_recon_loss_flat (ReconLoss()((a, b)) + gradient)"] = timed(lambda: ops.recon_loss_flat(logits, target, False, 1.0 / e, True), 12.0 * e)
    del logits, target
    nz, dz = 8192, 4096
    hlat = torch.rand((nz, 2 * dz), device="cuda")
    eps = torch.randn((nz, dz), device="cuda")
    mu, ls = hlat[:, :dz], hlat[:, dz:]
    out["fv_reparam_kl_fwd (z + KL partial sums)"] = timed(lambda: ops.reparam_kl_fwd(mu, ls, eps, True, True), 16.0 * nz * dz)
    dzt = torch.randn((nz, dz), device="cuda")
    buf = torch.empty((nz, 2 * dz), device="cuda")
    # reads mu, logstd, eps, dz; writes dmu, dlogstd
    out["fv_reparam_kl_bwd"] = timed(lambda: ops.reparam_kl_bwd(mu, ls, eps, dzt, None, None, 1e-6, None, buf), 24.0 * nz * dz)
    return out


def run_inference(args, torch, _lib, real_stdout, peaks, warmup):
    """BASELINE.json configs[4]: eval-mode encode -> sample -> decode (running statistics, no backward) at one batch size."""
    from face_vae_b200.models import FaceVAE
    torch.manual_seed(0)
    model = FaceVAE().cuda().eval()
    B, S = args.batch, args.size
    dz = model.latent_dim(S, S)
    xs = [torch.rand((B, 3, S, S), device="cuda") for _ in range(2)]
    eps = torch.randn((B, dz), device="cuda")
    l0 = _lib.launch_count
    with torch.no_grad():
        for i in range(warmup):
            model(xs[i % 2], True, eps)
        torch.cuda.synchronize()
        per_step = (_lib.launch_count - l0) // max(warmup, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            model(xs[i % 2], True, eps)
        e1.record()
        torch.cuda.synchronize()
    ms_eager = e0.elapsed_time(e1) / args.steps
    # the same forward replayed from a CUDA graph (face_vae_b200.trainer.GraphedInference): what a serving loop would call
    from face_vae_b200.trainer import GraphedInference
    engine = GraphedInference(model)
    for i in range(max(warmup, 3)):
        out_g = engine(xs[i % 2], True, eps)
    with torch.no_grad():
        ref = model(xs[(max(warmup, 3) - 1) % 2], True, eps)
    torch.cuda.synchronize()
    if not torch.equal(out_g[2], ref[2]):
        raise SystemExit("bench.py --infer: graph replay and eager forward disagree")
    e0.record()
    for i in range(args.steps):
        engine(xs[i % 2], True, eps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    line = {"metric": "inference images/sec at 256x256 (encode -> sample -> decode)", "value": B / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"face-vae anchor eval-mode encode -> sample -> decode, batch {B} at {S}x{S} (BASELINE.json configs[4])",
                       "l2": "two alternating input batches"},
            "gpu_launches": per_step * args.steps, "latency_ms": ms, "cuda_graph": True,
            "eager": {"value": B / (ms_eager * 1e-3), "ms_per_step": ms_eager}, "peaks": peaks}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


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
    ap.add_argument("--global-batch", type=int, default=0, help="fixed global batch split over the ranks (BASELINE.json configs[2]: 256 at "
                    "2/4/8 GPUs -> strong scaling); overrides --batch")
    ap.add_argument("--infer", action="store_true", help="BASELINE.json configs[4]: eval-mode encode -> sample -> decode at --batch (no backward)")
    ap.add_argument("--no-glue-roofline", action="store_true")
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
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        args.batch = args.global_batch // world
    B, S = args.batch, args.size
    if args.infer:
        return run_inference(args, torch, _lib, real_stdout, peaks, warmup)

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
    prefetch = DevicePrefetcher(depth=4)      # the host may run three batches ahead: rides out scheduling hiccups of a few ms
    for x, e in prefetch.over(host[i % 2] for i in range(prefetch.depth + 1)):
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
    from face_vae_b200 import ops as _ops
    side_mode = _ops.set_wgrad_stream(False)       # ... on ONE stream: an event pair must not span a kernel running beside its own
    trainer.step(*dev[0])
    _lib.profile_start()
    for i in range(args.profile_steps):
        # hold the stream (~30 ms of device spin) while the host enqueues the whole eager step: the kernels and their event
        # records then run back to back, and an event pair measures its kernel rather than the host's launch latency in front of it
        torch.cuda._sleep(int(6e7))
        trainer.step(*dev[i % 2])
    recs = _lib.profile_stop()
    trainer.use_cuda_graph = graph_mode
    _ops.set_wgrad_stream(side_mode)
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
    # ---- roofline of the convolutions: ALL of them -- forward, data gradient and weight gradient, the generic / ring / x2 / s2
    # kernels, the tap-folded out_conv kernels and the passes that finish the weight gradients (their time counts, they add no
    # FLOPs).  Algorithmic FLOPs = 2*N*H*W*Ci*Co*k*k per pass of the reference's formulation; the up-sampling
    # convolutions EXECUTE 2.25x fewer (four 2x2 phases instead of a 3x3 on the 4x larger image) -- both figures are given.
    CONV = ("fv_conv2d", "fv_conv2d_stats", "fv_conv2d_x2", "fv_conv2d_s2", "fv_conv2d_ex", "fv_conv2d_wgrad", "fv_conv2d_wgrad_x2",
            "fv_conv2d_wgrad_s2", "fv_outconv_fwd", "fv_outconv_dgrad", "fv_outconv_wgrad", "fv_wgrad_finish", "fv_wgrad_finish_up", "fv_slab_sum")
    conv = {"ms": 0.0, "flops": 0.0, "flops_exec": 0.0, "launches": 0}
    for nm in CONV:
        if nm in agg:
            for k in conv:
                conv[k] += agg[nm][k]
    # the same split measured INSIDE the replayed CUDA graph (CUPTI activity records of three replays): the eager event-pair
    # pass above inflates every launch by its host gap, so the in-graph SHARE of the convolutions times the device-timed step
    # (measured without any profiler, above) is the better estimate of their time in the real step
    in_graph = None
    one_stream_ms = None
    if trainer.use_cuda_graph:
        serial = bool(side_mode) and world == 1
        try:
            if serial:
                # With the weight-gradient side stream the convolutions run BESIDE the norm / activation passes: per-kernel
                # durations then overlap and a share of their sum is no longer a share of the step.  The split is therefore
                # taken on the one-stream schedule of the same kernels (graph re-captured with the side stream off, its step
                # timed with CUDA events without a profiler); the headline value above is the overlapped step.
                _ops.set_wgrad_stream(False)
                trainer._graph = None
                for i in range(3):
                    trainer.step(*dev[i % 2])
                torch.cuda.synchronize()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                for i in range(10):
                    trainer.step(*dev[i % 2])
                s1.record()
                torch.cuda.synchronize()
                one_stream_ms = s0.elapsed_time(s1) / 10
            in_graph = _in_graph_shares(torch, trainer, dev)
        except Exception as e:                      # profiler unavailable: keep the event-pair numbers only
            in_graph = {"error": f"{type(e).__name__}: {e}"}
        finally:
            if serial:
                _ops.set_wgrad_stream(True)
                trainer._graph = None
    if conv["ms"] > 0:
        per_step_ms = conv["ms"] / args.profile_steps
        flops_step = conv["flops"] / args.profile_steps
        ach = flops_step / (per_step_ms * 1e-3) / 1e12
        at_max = clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"] and "sw_power_cap" not in clocks.get("reasons", [])
        peak = peaks["bf16"] if at_max else peaks["bf16_sustained"]
        roofline = {"kernel": "all convolution kernels of the step: conv_igemm (same / x2 / s2), conv_ring, conv_wgrad, conv_wgrad_ring, fold_conv, "
                              "fold_wgrad + the weight-gradient finish passes (forward, data gradient AND weight gradient)",
                    "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "peak_source": peaks["source"] + (" burst (SM clock at max, no power cap during the timed region)" if at_max else
                                                      " sustained (SM clock below max or power cap active during the timed region)"),
                    "frac_of_burst_peak": ach / peaks["bf16"], "frac_of_sustained_peak": ach / peaks["bf16_sustained"],
                    "timing": "CUDA events around every launch of an eager replay of the step, on the launching stream; the stream is held by a device-side spin while the host enqueues the step, so the launches run back to back",
                    "algorithmic_tflop_per_step": flops_step / 1e12, "executed_tflop_per_step": conv["flops_exec"] / args.profile_steps / 1e12,
                    "achieved_executed": conv["flops_exec"] / args.profile_steps / (per_step_ms * 1e-3) / 1e12,
                    "ms_per_step": per_step_ms, "avg_launch_ms": conv["ms"] / conv["launches"],
                    "launches_per_step": conv["launches"] / args.profile_steps, "traffic": None}
        if in_graph and "conv_share" in in_graph:
            ms_in = in_graph["conv_share"] * (one_stream_ms if one_stream_ms else ms / args.steps)
            roofline["in_graph"] = {"conv_share_of_kernel_time": in_graph["conv_share"], "conv_ms_per_step": ms_in,
                                    "achieved": flops_step / (ms_in * 1e-3) / 1e12, "frac": flops_step / (ms_in * 1e-3) / 1e12 / peak,
                                    "how": "share of the convolution kernels in the CUPTI kernel time of three graph replays x the device-timed step"
                                           + (" of the ONE-STREAM schedule (graph re-captured with the weight-gradient side stream off; "
                                              "the headline step overlaps these kernels with the norm / activation passes)" if one_stream_ms else "")}
            if one_stream_ms:
                roofline["in_graph"]["one_stream_ms_per_step"] = one_stream_ms
        elif in_graph:
            roofline["in_graph"] = in_graph
        # DRAM bytes per launch of these kernels from the committed ncu launch list of this command (profiles/)
        for tname in ("r02_conv_traffic.json", "r01_final_conv_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath) and B == 32 and S == 256 and not args.deep:
                try:
                    roofline["traffic"] = float(json.load(open(tpath))["bytes_per_launch"])
                except (KeyError, ValueError, TypeError):
                    continue
                roofline["traffic_source"] = f"profiles/{tname} (ncu dram__bytes_read.sum + dram__bytes_write.sum, average per launch)"
                break
    glue = None
    if rank == 0 and not args.no_glue_roofline:
        glue = _glue_roofline(torch, peaks)

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
                           "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": bool(trainer.use_cuda_graph), "wgrad_side_stream": bool(side_mode),
                           "l2": "working set per step (activations + gradients, >3 GB) far exceeds the 126 MB L2; inputs alternate between two batches"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 3 * S * S * 4 + B * dz * 4, "d2h_bytes_per_step": 4,
                        "last_loss": last},
                "gpu_launches": launches, "roofline": roofline, "roofline_glue": glue, "cpu_baseline": cpu, "kernels": kernels,
                "in_graph_kernels": (in_graph or {}).get("kernels"), "peaks": peaks}
        if args.global_batch:
            line["scaling"] = "strong"
            line["config"]["workload"] += f" (global batch {args.global_batch} split over {world} GPUs)"
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
