import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from face_vae_b200 import _lib
out = torch.zeros(1, dtype=torch.int64, device="cuda")
iters = 4000
print("cycles per tcgen05.mma (M=128, K=16, bf16), operands resident in smem")
for row in (128, 64, 32):
    vals = []
    for n in (16, 32, 64, 128, 256):
        _lib.call("fv_debug_mma_rate", n, row, iters, 16, 0, 1, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        vals.append(out.item() / iters)
    print(f"K -major row {row:3d} B  all SMs, A start shifted by 0/1/2 pixel rows  N=16/32/64/128/256: " + "  ".join(f"{v:6.1f}" for v in vals))
for mn in (0, 1):
    for allsm in (0, 1):
        for row in (128, 64, 32):
            vals = []
            for n in (16, 32, 64, 128, 256):
                _lib.call("fv_debug_mma_rate", n, row, iters, 1, mn, allsm, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                vals.append(out.item() / iters)
            print(f"{'MN' if mn else 'K '}-major row {row:3d} B  {'all SMs' if allsm else 'one SM '}  N=16/32/64/128/256: " + "  ".join(f"{v:6.1f}" for v in vals))
