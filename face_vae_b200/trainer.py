"""Objective and train step for the VAE path, following the reference's contracts:

* ``VAEGeneratorFull.forward`` returns a loss dict whose entries are already weight-multiplied, like
  ``GeneratorFull.forward`` does for its "K" and "R" entries (reference trainer.py:240-252, 300-317); the intended
  weights K: 0.2, R: 10 are the commented values at trainer.py:250-251 (the shipped defaults are 0, which switches the
  VAE terms off).
* ``VAETrainer.step`` is the skeleton of ``Logger.step`` (reference logger.py:150-164): zero_grad -> forward ->
  sum(losses.values()).backward() -> optimizer.step(), with Adam(lr, betas=(0.5, 0.999)) as at logger.py:60.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from . import distributed as fdist
from .models import FaceVAE


class VAEGeneratorFull(nn.Module):
    def __init__(self, vae: FaceVAE, weights: Optional[Dict[str, float]] = None, l1: bool = False):
        super().__init__()
        self.vae = vae
        self.weights = {"K": 0.2, "R": 10.0}
        if weights:
            self.weights.update(weights)
        self.l1 = l1

    def forward(self, d: torch.Tensor, eps: Optional[torch.Tensor] = None, train_vae: bool = True):
        """d: driving frames [N,3,H,W] fp32 in [0,1].  -> (loss_dict, generated_d, mu, logstd)."""
        if not train_vae:
            _, _, x_hat = self.vae(d, False)
            zero = torch.zeros((), device=d.device)
            return {"K": zero, "R": zero}, x_hat, None, None
        out = self.vae.forward_loss(d, eps, self.l1)
        loss = {"K": self.weights["K"] * out["K"], "R": self.weights["R"] * out["R"]}
        return loss, out["x_hat"], out["mu"], out["logstd"]


class VAETrainer:
    """One-model version of the reference Logger's optimisation step (logger.py:52-63, 150-164)."""

    def __init__(self, vae: FaceVAE, lr: float = 5e-5, weights: Optional[Dict[str, float]] = None, bucket_mb: float = 2.0,
                 fused_adam: bool = True):
        self.g_full = VAEGeneratorFull(vae, weights)
        self.vae = vae
        fdist.broadcast_parameters(vae)
        params = list(vae.parameters())
        on_cuda = bool(params) and params[0].is_cuda
        self.optimizer = torch.optim.Adam(params, lr=lr, betas=(0.5, 0.999), **({"fused": True} if (fused_adam and on_cuda) else {}))
        self.reducer = fdist.GradientReducer(params, bucket_mb) if fdist.get_world_size() > 1 else None

    def step(self, d: torch.Tensor, eps: Optional[torch.Tensor] = None):
        self.optimizer.zero_grad(set_to_none=True)
        losses, generated, mu, logstd = self.g_full(d, eps, True)
        total = sum(losses.values())
        total.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        return losses, generated
