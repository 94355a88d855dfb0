"""Role-loop cycle breakdown of the ring kernels (needs the FV_TRACE build: tools/lib_trace.so copied over the product lib)."""
import os, sys, shutil
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
    print(f"{name}: {e0.elapsed_time(e1)*1e3:.1f} us; mean cycles per CTA per unit ({units} units/CTA): " + ", ".join(f"{n}={v/units:.0f}" for n, v in zip(names, c) if n))
n, hw = 32, 256
for (ci, co, k) in [(32, 64, 3), (64, 32, 3), (32, 3, 7), (16, 32, 7)]:
    cip, cop = ops.pad_channels(ci), ops.pad_channels(co)
    x = torch.randn((n, hw, hw, cip), device="cuda").bfloat16()
    dy = torch.randn((n, hw, hw, cop), device="cuda").bfloat16()
    w = torch.randn((co, ci, k, k), device="cuda") * 0.05
    wf, wd = ops.weight_prep(w)
    tiles = n * hw * (hw // 128) / 148
    run(f"fprop ring {ci}->{co} k{k}", lambda: ops.conv2d(x, wf, None, co, k, out_mode=(2 if co == 3 else 0)),
        ["prod_wait_empty", "mma_commit", "mma_wait_tempty", "mma_wait_full", "mma_issue", "mma_total", "epi_wait_tfull", "epi_work"], tiles)
    blocks = n * hw * (hw // 64) / 148
    run(f"wgrad ring {ci}->{co} k{k}", lambda: ops.conv2d_wgrad(x, dy, k),
        ["prod_wait_bempty", "prod_wait_empty", "mma_wait_full", "mma_wait_b", "mma_issue", "mma_total", "epilogue(total)", ""], blocks)
