"""Informational "stronger bar" (SURVEY.md 8d): the anchor composition written with stock torch.nn layers (cuDNN / ATen
kernels, eager autograd, fused Adam) timed on the same B200 -- fp32 with TF32 convolutions, and bf16 autocast +
channels_last.  Not a parity reference and not the product: a yardstick for what the library stack gives for this model.

  python tools/torch_gpu_baseline.py [--batch 32] [--size 256] [--steps 20]
"""
import argparse, json, time
import torch
from torch import nn
import torch.nn.functional as F


def block(ci, co, k=3):
    return nn.Sequential(nn.Conv2d(ci, co, k, 1, (k - 1) // 2), nn.BatchNorm2d(co), nn.ReLU(inplace=True))


class Res(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.n1, self.c1, self.n2, self.c2 = nn.BatchNorm2d(c), nn.Conv2d(c, c, 3, 1, 1), nn.BatchNorm2d(c), nn.Conv2d(c, c, 3, 1, 1)

    def forward(self, x):
        h = self.c1(F.relu(self.n1(x)))
        return x + self.c2(F.relu(self.n2(h)))


class Anchor(nn.Module):
    def __init__(self):
        super().__init__()
        d, u = [3, 32, 64, 128, 256, 32], [256, 256, 128, 64, 32]
        self.enc = nn.ModuleList([block(d[0], d[1], 1)] + [block(d[i], d[i + 1]) for i in range(1, 5)])
        self.mid = nn.Conv2d(16, 256, 1)
        self.res = nn.Sequential(Res(256), Res(256))
        self.up = nn.ModuleList([block(u[i], u[i + 1]) for i in range(4)])
        self.out = nn.Conv2d(32, 3, 7, 1, 3)

    def forward(self, x, eps):
        h = self.enc[0](x)
        for b in self.enc[1:]:
            h = F.avg_pool2d(b(h), 2)
        n = h.shape[0]
        flat = h.float().reshape(n, -1)
        dz = flat.shape[1] // 2
        mu, ls = flat[:, :dz], flat[:, dz:]
        z = (mu + torch.exp(ls) * eps).view(n, 16, h.shape[2], h.shape[3])
        kl = (-0.5 - ls + 0.5 * mu * mu + 0.5 * torch.exp(2 * ls)).mean()
        t = self.res(self.mid(z.to(h.dtype)))
        for b in self.up:
            t = b(F.interpolate(t, scale_factor=2))
        xh = torch.sigmoid(self.out(t).float())
        return 0.2 * kl + 10.0 * F.mse_loss(xh, x.float())


def run(mode, B, S, steps):
    torch.manual_seed(0)
    m = Anchor().cuda().train()
    if mode == "bf16":
        m = m.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(m.parameters(), lr=5e-5, betas=(0.5, 0.999), fused=True)
    x = torch.rand((B, 3, S, S), device="cuda")
    if mode == "bf16":
        x = x.contiguous(memory_format=torch.channels_last)
    eps = torch.randn((B, 16 * (S // 16) ** 2), device="cuda")
    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            loss = m(x, eps)
        loss.backward()
        opt.step()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"mode": mode, "ms_per_step": ms, "images_per_sec": B / ms * 1e3}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    out = {"what": "stock torch.nn anchor on this GPU (cuDNN/ATen eager), informational", "batch": a.batch, "size": a.size,
           "torch": torch.__version__, "runs": [run("fp32_tf32", a.batch, a.size, a.steps), run("bf16", a.batch, a.size, a.steps)]}
    print(json.dumps(out))
