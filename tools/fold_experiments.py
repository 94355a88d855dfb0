"""Timing decomposition of the tap-folded out_conv kernels: FV_FOLD_DEBUG switches parts of the kernel off
(1: no MMAs, 2: no epilogue work, 4: no slab fill).  Results are wrong by construction; only the times matter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from face_vae_b200 import ops
n, hw, co = 32, 256, 3
x = torch.randn((n, hw, hw, 32), device="cuda").bfloat16()
w = torch.randn((co, 32, 7, 7), device="cuda") * 0.05
wq, wdq = ops.outconv_prep(w)
tgt = torch.rand((n, co, hw, hw), device="cuda")
g4 = (torch.randn((n, hw, hw, 4), device="cuda") * 0.01).bfloat16()
one = torch.ones((1,), device="cuda")
flush = torch.empty(64 * 1024 * 1024, device="cuda")
def timeit(fn, iters=5):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3
for dbg in ():
    os.environ["FV_FOLD_DEBUG"] = str(dbg)
    d = timeit(lambda: ops.outconv_wgrad(x, g4, one, co))
    print(f"wgrad dbg {dbg} (noMMA {dbg & 1} noBuild {(dbg >> 2) & 1} noSlabLoad {(dbg >> 3) & 1}): {d:6.1f} us", flush=True)
for dbg in (0,):
    os.environ["FV_FOLD_DEBUG"] = str(dbg)
    a = timeit(lambda: ops.outconv_fwd(x, wq, None, co, target=tgt, gscale=1e-6))
    b = timeit(lambda: ops.outconv_fwd(x, wq, None, co))
    c = timeit(lambda: ops.outconv_dgrad(g4, wdq, one, co))
    print(f"dbg {dbg} (noMMA {dbg & 1} noEpi {(dbg >> 1) & 1} noFill {(dbg >> 2) & 1}): fwd+loss {a:6.1f} us  fwd {b:6.1f} us  dgrad {c:6.1f} us", flush=True)
os.environ["FV_FOLD_DEBUG"] = "0"
