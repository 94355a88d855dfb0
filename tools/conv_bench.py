"""Per-layer timing of the tcgen05 conv kernels at the benchmark shapes (batch 32, 256x256): forward, data gradient,
weight gradient, each timed alone with CUDA events and an L2 flush (256 MB write) between launches.
usage: python tools/conv_bench.py [--only NAME] [--iters 5] [--batch 32] [--size 256]"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from face_vae_b200 import ops

def layers(size):
    s = size
    L = [("enc.0", 3, 32, 1, s), ("enc.1", 32, 64, 3, s), ("enc.2", 64, 128, 3, s // 2), ("enc.3", 128, 256, 3, s // 4),
         ("enc.4", 256, 32, 3, s // 8), ("mid", 16, 256, 1, s // 16), ("res", 256, 256, 3, s // 16), ("up.0", 256, 256, 3, s // 8),
         ("up.1", 256, 128, 3, s // 4), ("up.2", 128, 64, 3, s // 2), ("up.3", 64, 32, 3, s), ("out", 32, 3, 7, s)]
    return L

def timeit(fn, iters, flush):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--what", default="fwd,dgrad,wgrad")
    a = ap.parse_args()
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    tot = {"fwd": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0]}
    mult = {"res": 4}
    print(f"{'layer':7s} {'ci':>4s} {'co':>4s} k {'hw':>4s} | {'fwd ms':>8s} {'TF/s':>7s} | {'dgrad ms':>8s} {'TF/s':>7s} | {'wgrad ms':>8s} {'TF/s':>7s}")
    for name, ci, co, k, hw in layers(a.size):
        if a.only and name != a.only:
            continue
        n = a.batch
        cip, cop = ops.pad_channels(ci), ops.pad_channels(co)
        x = torch.randn((n, hw, hw, cip), device="cuda").bfloat16()
        dy = torch.randn((n, hw, hw, cop), device="cuda").bfloat16()
        w = torch.randn((co, ci, k, k), device="cuda") * 0.05
        wf, wd = ops.weight_prep(w)
        flops = 2.0 * n * hw * hw * ci * co * k * k
        res = {}
        if "fwd" in a.what:
            res["fwd"] = timeit(lambda: ops.conv2d(x, wf, None, co, k), a.iters, flush)
        if "dgrad" in a.what and name != "enc.0":
            res["dgrad"] = timeit(lambda: ops.conv2d(dy, wd, None, cip, k), a.iters, flush)
        if "wgrad" in a.what:
            res["wgrad"] = timeit(lambda: ops.conv2d_wgrad(x, dy, k), a.iters, flush)
        if name == "out" and ops.outconv_supported(n, hw, hw, cip, co, k):
            # the product path of this layer: tap-folded kernels (forward with the fused sigmoid + loss epilogue)
            wq, wdq = ops.outconv_prep(w)
            tgt = torch.rand((n, co, hw, hw), device="cuda")
            g4 = (torch.randn((n, hw, hw, 4), device="cuda") * 0.01).bfloat16()
            one = torch.ones((1,), device="cuda")
            rf = {"fwd": timeit(lambda: ops.outconv_fwd(xf, wq, None, co, target=tgt, gscale=1e-6), a.iters, flush) if (xf := x) is not None else 0,
                  "dgrad": timeit(lambda: ops.outconv_dgrad(g4, wdq, one, co), a.iters, flush),
                  "wgrad": timeit(lambda: ops.outconv_wgrad(x, g4, one, co), a.iters, flush)}
            print(f"{'out/gen':7s} {ci:4d} {co:4d} {k} {hw:4d} | " + " | ".join(f"{res[key]:8.3f} {flops / res[key] / 1e9:7.1f}" for key in ("fwd", "dgrad", "wgrad") if key in res), flush=True)
            res = rf
        cells = []
        for key in ("fwd", "dgrad", "wgrad"):
            if key in res:
                cells.append(f"{res[key]:8.3f} {flops / res[key] / 1e9:7.1f}")
                tot[key][0] += res[key] * mult.get(name, 1)
                tot[key][1] += flops * mult.get(name, 1)
            else:
                cells.append(f"{'-':>8s} {'-':>7s}")
        print(f"{name:7s} {ci:4d} {co:4d} {k} {hw:4d} | " + " | ".join(cells), flush=True)
    for key, (ms, fl) in tot.items():
        if ms:
            print(f"total {key}: {ms:.3f} ms per step, {fl / ms / 1e9:.1f} TFLOP/s algorithmic")

if __name__ == "__main__":
    main()
