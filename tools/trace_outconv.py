"""Role-loop cycle breakdown of the tap-folded out_conv kernels (needs the FV_TRACE build: tools/lib_trace.so)."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import torch
from face_vae_b200 import _lib
_lib.LIB_PATH = os.path.join(root, "tools", "lib_trace.so")
from face_vae_b200 import ops
cnt = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
_lib.call("fv_debug_trace_set", cnt.data_ptr())
def run(name, fn, names, units):
    fn(); torch.cuda.synchronize(); cnt.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    c = cnt.view(148, 8).double().mean(0).tolist()
    print(f"{name}: {e0.elapsed_time(e1)*1e3:.1f} us; mean cycles per CTA per unit ({units:.1f} units/CTA): " + ", ".join(f"{n}={v/units:.0f}" for n, v in zip(names, c) if n))
n, hw, co = 32, 256, 3
x = torch.randn((n, hw, hw, 32), device="cuda").bfloat16()
w = torch.randn((co, 32, 7, 7), device="cuda") * 0.05
wq, wdq = ops.outconv_prep(w)
tgt = torch.rand((n, co, hw, hw), device="cuda")
g4 = (torch.randn((n, hw, hw, 4), device="cuda") * 0.01).bfloat16()
one = torch.ones((1,), device="cuda")
rows = n * hw / 148
names = ["", "mma_wait_full", "mma_wait_tempty", "mma_issue", "mma_commit", "mma_total", "mma_desc", ""]
run("fold fwd + loss", lambda: ops.outconv_fwd(x, wq, None, co, target=tgt, gscale=1e-6), names, rows)
run("fold fwd plain ", lambda: ops.outconv_fwd(x, wq, None, co), names, rows)
run("fold dgrad     ", lambda: ops.outconv_dgrad(g4, wdq, one, co), names, rows)
run("fold wgrad     ", lambda: ops.outconv_wgrad(x, g4, one, co),
    ["", "mma_wait_full", "mma_wait_rfull", "mma_issue", "mma_commit", "mma_total", "", ""], 2 * rows)
import os
for dbg in (7, 13):
    os.environ["FV_FOLD_DEBUG"] = str(dbg)
    run(f"dbg {dbg} fold fwd plain ", lambda: ops.outconv_fwd(x, wq, None, co), names, rows)
    run(f"dbg {dbg} fold wgrad     ", lambda: ops.outconv_wgrad(x, g4, one, co), ["", "mma_wait_full", "mma_wait_rfull", "mma_issue", "mma_commit", "mma_total", "", ""], 2 * rows)
