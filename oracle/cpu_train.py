"""CPU train step of the anchor model (TEST / BASELINE INFRASTRUCTURE ONLY).

Used by bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm.  The reference is Python living in
/root/reference, which does not exist on the GPU box, so its CPU implementation of the path is represented by this
port (kind "port"): the same ATen operators the reference's modules dispatch to on CPU (nn.Conv2d -> mkldnn conv,
batch norm, in-place ReLU, AvgPool2d, nearest Upsample, MSELoss), fp32, eager autograd,
Adam(lr=5e-5, betas=(0.5, 0.999)) (reference logger.py:60, 150-164).  tests/test_oracle_golden.py checks it against
the explicit-formula oracle (and so, transitively, against the golden fixtures from the reference).
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F
from torch import nn

from . import facevae_oracle as O


class _CNA(nn.Module):
    def __init__(self, ci, co, k):
        super().__init__()
        self.layers = nn.Sequential(nn.Conv2d(ci, co, k, 1, (k - 1) // 2), nn.BatchNorm2d(co), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.layers(x)


class _NAC(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layers = nn.Sequential(nn.BatchNorm2d(c), nn.ReLU(inplace=True), nn.Conv2d(c, c, 3, 1, 1))

    def forward(self, x):
        return self.layers(x)


class _Same(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.layers = _CNA(ci, co, 1)

    def forward(self, x):
        return self.layers(x)


class _Down(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.layers = nn.Sequential(_CNA(ci, co, 3), nn.AvgPool2d((2, 2)))

    def forward(self, x):
        return self.layers(x)


class _Up(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.layers = nn.Sequential(nn.Upsample(scale_factor=(2, 2)), _CNA(ci, co, 3))

    def forward(self, x):
        return self.layers(x)


class _Res(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layers = nn.Sequential(_NAC(c), _NAC(c))

    def forward(self, x):
        return x + self.layers(x)


class AtenAnchor(nn.Module):
    """Same composition and state_dict keys as the anchor of SURVEY.md section 8."""

    def __init__(self, cfg: O.AnchorConfig = O.CFG_256):
        super().__init__()
        d, u = cfg.down_seq, cfg.up_seq
        self.cfg = cfg
        self.enc = nn.Sequential(*[_Same(d[i], d[i + 1]) if i == 0 else _Down(d[i], d[i + 1]) for i in range(len(d) - 1)])
        self.mid_conv = nn.Conv2d(cfg.zc, u[0], 1, 1, 0)
        self.res = nn.Sequential(*[_Res(u[0]) for _ in range(cfg.n_res)])
        self.up = nn.Sequential(*[_Up(u[i], u[i + 1]) for i in range(len(u) - 1)])
        self.out_conv = nn.Conv2d(u[-1], 3, 7, 1, 3)
        self.mse = nn.MSELoss()

    def forward(self, x, eps):
        h = self.enc(x)
        mu, logstd, z = O.reparameterise(h, eps, True, self.cfg.zc)
        x_hat = torch.sigmoid(self.out_conv(self.up(self.res(self.mid_conv(z)))))
        K = O.kl_divergence(mu, logstd)
        R = self.mse(x, x_hat)
        return K, R, self.cfg.w_kl * K + self.cfg.w_rec * R


class CpuAnchorTrainer:
    def __init__(self, cfg: O.AnchorConfig = O.CFG_256, base: int = 0, lr: float = 5e-5):
        self.model = AtenAnchor(cfg).train()
        sd = self.model.state_dict()
        for k, v in O.det_anchor_params(cfg, base).items():
            assert k in sd, k
            sd[k] = v
        self.model.load_state_dict(sd)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, betas=(0.5, 0.999))

    def step(self, x, eps):
        self.opt.zero_grad()
        K, R, loss = self.model(x, eps)
        loss.backward()
        self.opt.step()
        return loss.item()


def time_cpu_train(n: int, h: int, w: int, steps: int, warmup: int, threads: int, cfg: O.AnchorConfig = O.CFG_256):
    """-> (images/sec from the median step, median seconds per step)."""
    torch.set_num_threads(threads)
    tr = CpuAnchorTrainer(cfg)
    x, eps = O.det_inputs(n, h, w, cfg, 0)
    for _ in range(warmup):
        tr.step(x, eps)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.step(x, eps)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return n / med, med
