"""Loss modules of the hot path with the reference's names and call conventions (reference losses.py:8-13, 385-403)."""
from __future__ import annotations

import torch
from torch import nn

from . import functional as Fn
from . import ops


def _flat2d(t: torch.Tensor) -> torch.Tensor:
    return t if t.dim() == 2 else t.reshape(t.shape[0], -1)


def l1(x, y):
    """Element-wise |x - y| (reference losses.py:8-9); plain tensor algebra, not on the fused path."""
    return torch.abs(x - y)


def l2(x, y):
    """Element-wise (x - y)^2 (reference losses.py:12-13)."""
    return torch.pow((x - y), 2)


class KLDivergenceLoss(nn.Module):
    """loss = mean_n mean_d(-0.5 - logstd + 0.5 mu^2 + 0.5 exp(2 logstd)) for ``kl = (mu, logstd)``
    (reference losses.py:385-393).  One fused kernel forward, one backward."""

    def forward(self, kl):
        mu, logstd = _flat2d(kl[0]), _flat2d(kl[1])
        if mu.dtype != torch.float32 or mu.stride(-1) != 1 or logstd.stride(-1) != 1 or mu.stride(0) != logstd.stride(0) \
                or mu.shape[1] % 4 or mu.data_ptr() % 16 or logstd.data_ptr() % 16 or (mu.stride(0) * 4) % 16:
            mu, logstd = mu.float().contiguous(), logstd.float().contiguous()
        return Fn.KLDivergence.apply(mu, logstd)


class ReconLoss(nn.Module):
    """nn.MSELoss()(Rec[0], Rec[1]) (reference losses.py:396-403); ``l1=True`` gives nn.L1Loss (losses.py:128)."""

    def __init__(self, l1: bool = False) -> None:
        super().__init__()
        self.l1 = l1

    def forward(self, Rec):
        a, b = Rec[0], Rec[1]
        if a.shape != b.shape:
            raise ValueError("ReconLoss: shape mismatch")
        return Fn.ReconLossFlat.apply(a, b, self.l1)
