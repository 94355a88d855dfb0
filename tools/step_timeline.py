"""Where the replayed (CUDA graph) train step spends its time: per-kernel durations and idle gaps from the CUPTI
activity records torch.profiler collects (not a bench number: tracing adds overhead), plus an e2e loop breakdown.

  python tools/step_timeline.py [--batch 32] [--size 256] [--steps 4]
"""
import argparse, collections, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from face_vae_b200.models import FaceVAE
from face_vae_b200.trainer import VAETrainer
from face_vae_b200.data import AsyncScalarLog, DevicePrefetcher


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=100)
    a = ap.parse_args()
    torch.manual_seed(0)
    model = FaceVAE().cuda().train()
    tr = VAETrainer(model)
    B, S = a.batch, a.size
    dz = model.latent_dim(S, S)
    host = [(torch.rand((B, 3, S, S)).pin_memory(), torch.randn((B, dz)).pin_memory()) for _ in range(2)]
    dev = [(x.cuda(), e.cuda()) for x, e in host]
    for i in range(5):
        tr.step(*dev[i % 2])
    torch.cuda.synchronize()

    def timed(fn, n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(n); torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    def loop_dev(n):
        for i in range(n):
            tr.step(*dev[i % 2])

    def loop_dev_log(n):
        log = AsyncScalarLog(n)
        for i in range(n):
            losses, _ = tr.step(*dev[i % 2])
            log.push(torch.stack([v.detach() for v in losses.values()]).sum())
        log.values()

    def loop_e2e(n):
        log = AsyncScalarLog(n)
        for x, e in DevicePrefetcher(host[i % 2] for i in range(n)):
            losses, _ = tr.step(x, e)
            log.push(torch.stack([v.detach() for v in losses.values()]).sum())
        log.values()

    def loop_h2d(n):
        for i in range(n):
            host[i % 2][0].to("cuda", non_blocking=True)

    n = a.e2e_steps
    print(f"resident inputs        {timed(loop_dev, n):7.3f} ms/step")
    print(f"  + loss read-back     {timed(loop_dev_log, n):7.3f} ms/step")
    print(f"  + prefetched H2D     {timed(loop_e2e, n):7.3f} ms/step")
    print(f"  30-step e2e loop     {timed(loop_e2e, 30):7.3f} ms/step")
    t = timed(loop_h2d, 20)
    print(f"H2D copy alone         {t:7.3f} ms  ({host[0][0].numel() * 4 / t / 1e6:.1f} GB/s)")

    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(a.steps):
            tr.step(*dev[i % 2])
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print("no CUDA activity records"); return
    agg = collections.OrderedDict()
    busy, gaps, last_end = 0.0, 0.0, None
    span0, span1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    for e in evs:
        d = e.time_range.end - e.time_range.start
        k = agg.setdefault(e.name[:90], [0, 0.0]); k[0] += 1; k[1] += d
        if last_end is not None and e.time_range.start > last_end:
            gaps += e.time_range.start - last_end
        last_end = max(last_end or 0, e.time_range.end)
        busy += d
    print(f"\n{a.steps} traced steps: span {(span1 - span0) / 1e3 / a.steps:.3f} ms/step, kernel time {busy / 1e3 / a.steps:.3f}, "
          f"idle gaps {gaps / 1e3 / a.steps:.3f}, records/step {len(evs) / a.steps:.0f}")
    for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{c / a.steps:6.1f} x {us / c:8.2f} us = {us / a.steps / 1e3:7.3f} ms/step  {name}")
    # per-launch durations (last traced step, launch order) of the multi-launch kernels
    per = len(evs) // a.steps
    last = evs[-per:]
    seq = collections.OrderedDict()
    for e in last:
        seq.setdefault(e.name[:60], []).append(e.time_range.end - e.time_range.start)
    print("\nper-launch us (last step, launch order):")
    for name, ds in seq.items():
        if 3 <= len(ds) <= 30 and "fv::" in name:
            print(f"  {name:60s} " + " ".join(f"{d:.0f}" for d in ds))


if __name__ == "__main__":
    main()
