"""Storage-precision emulation of the hot path on CPU (TEST INFRASTRUCTURE ONLY -- nothing in face_vae_b200 imports it).

The CUDA path keeps activations, conv-output gradients and filter operands in bf16 and accumulates in fp32.  Compared
with the reference's fp32 run, a bf16 network differs by percents on ReLU-masked gradients -- for ANY bf16 implementation,
the reference's own autocast included -- so an end-to-end comparison can only be loose.  This module
restates the reference's blocks (same formulas as oracle/facevae_oracle.py, which is pinned against the unmodified
reference classes) with bf16 roundings inserted exactly where the CUDA path stores bf16:

    forward  : filter operands, conv outputs, normalised / pooled activations
    backward : the conv-output gradient dy, the data gradient dx handed to the previous block

through two autograd identities (``rf``: round the value, pass the gradient; ``rb``: pass the value, round the gradient).
With the roundings switched off (``Prec(False)``) every function here IS the fp32 oracle -- tests/test_oracle_golden.py
checks that against the golden fixtures, which pins the structure.  With them on, a block fed with the CUDA path's own
(bf16-exact) input and upstream gradient must agree with the CUDA block to accumulation-order noise: that is the tight
(<= 2e-2, typically 1e-3) per-layer check of tests/test_layerwise_gpu.py, at any size including batch 32 at 256x256.

Reference formulas: modules.py:8-135 (blocks), models.py:559-561 (re-parameterisation), losses.py:385-403 (losses).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import facevae_oracle as O


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        s = ctx.scale
        return (g / s).bfloat16().to(g.dtype) * s, None


class Prec:
    """``Prec(True)``: bf16 storage emulation; ``Prec(False)``: every hook is the identity (the fp32 oracle)."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled

    def rf(self, x):
        return _RoundFwd.apply(x) if self.enabled else x

    def rb(self, x, scale: float = 1.0):
        """Identity whose gradient is rounded to bf16 (after division by ``scale``: a gradient that is stored before an
        upstream scalar is applied, as the out_conv kernels do with d(loss)/d(logits))."""
        return _RoundBwd.apply(x, float(scale)) if self.enabled else x


FP32 = Prec(False)
BF16 = Prec(True)


def _up_phase_rows(a, u):
    return {(0, 0): (0, 0), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2, 2)}[(a, u)]


def upsample_conv3x3(x, w, b, prec: Prec):
    """conv3x3(upsample_nearest2(x)) (UpBlock2D, reference modules.py:78-89).  With bf16 emulation: the four 2x2 phase
    convolutions the CUDA path runs (csrc/fv_conv.cu, X2), whose filters are fp32 SUMS of the 3x3 taps rounded to bf16 once."""
    if not prec.enabled:
        return F.conv2d(O.upsample_nearest2(x), w, b, padding=1)
    n, ci, h, wd = x.shape
    co = w.shape[0]
    xp = F.pad(x, (1, 1, 1, 1))
    y = x.new_zeros((n, co, 2 * h, 2 * wd))
    for a in range(2):
        for bb in range(2):
            taps = []
            for u in range(2):
                r0, r1 = _up_phase_rows(a, u)
                row = []
                for v in range(2):
                    s0, s1 = _up_phase_rows(bb, v)
                    row.append(w[:, :, r0:r1 + 1, s0:s1 + 1].sum(dim=(2, 3)))
                taps.append(torch.stack(row, dim=-1))
            wp = prec.rf(torch.stack(taps, dim=-2))                 # [co, ci, 2, 2]
            yp = F.conv2d(xp, wp)[:, :, a:a + h, bb:bb + wd]       # rows i + u - 1 + a of x  <->  rows i + a + u of xp
            y[:, :, a::2, bb::2] = yp
    return y + b[None, :, None, None] if b is not None else y


def cna_block(x, p: Dict[str, torch.Tensor], prefix: str, ksize: int, prec: Prec, post: str = "none", upsample: bool = False,
              nonlinearity: str = "relu", out_fp32: bool = False, training: bool = True, updates: Optional[dict] = None,
              first_layer_pointwise: bool = False):
    """_ConvBlock pattern "CNA" (+ AvgPool2d / preceded by Upsample): ConvBNAct of face_vae_b200/functional.py.
    ``prefix`` ends with "layers." of the ConvBlock2D.  ``first_layer_pointwise``: the RGB fast path keeps everything in
    fp32 up to the stored activation (csrc/fv_pointwise.cu: no bf16 filter, no materialised conv output)."""
    w, b = p[prefix + "0.weight"], p[prefix + "0.bias"]
    if first_layer_pointwise:
        y = F.conv2d(x, w, b)
    else:
        x = prec.rb(x)                                     # dx leaves the data-gradient kernel as bf16
        if upsample:
            y = upsample_conv3x3(x, w, b, prec)
        else:
            y = F.conv2d(x, prec.rf(w), b, padding=(ksize - 1) // 2)
        y = prec.rb(prec.rf(y))                            # y stored bf16; dy (norm+act backward) stored bf16
    g, be = p[prefix + "1.weight"], p[prefix + "1.bias"]
    if training:
        a, rm, rv = O.batch_norm_train(y, g, be, p.get(prefix + "1.running_mean"), p.get(prefix + "1.running_var"))
        if updates is not None and rm is not None:
            updates[prefix + "1.running_mean"] = rm
            updates[prefix + "1.running_var"] = rv
    else:
        a = O.batch_norm_eval(y, g, be, p[prefix + "1.running_mean"], p[prefix + "1.running_var"])
    a = O._act(a, nonlinearity)
    if post == "pool":
        a = O.avg_pool2(a)
    return a if out_fp32 else prec.rf(a)


def nac_block(x, p, prefix: str, prec: Prec, residual=None, training: bool = True, updates: Optional[dict] = None):
    """_ConvBlock pattern "NAC" (ResBlock2D halves): BNActConv of face_vae_b200/functional.py; the residual is added in the
    conv epilogue in fp32, before the bf16 store."""
    x = prec.rb(x)                                         # dx of the norm+act backward: bf16
    g, be = p[prefix + "0.weight"], p[prefix + "0.bias"]
    if training:
        a, rm, rv = O.batch_norm_train(x, g, be, p.get(prefix + "0.running_mean"), p.get(prefix + "0.running_var"))
        if updates is not None and rm is not None:
            updates[prefix + "0.running_mean"] = rm
            updates[prefix + "0.running_var"] = rv
    else:
        a = O.batch_norm_eval(x, g, be, p[prefix + "0.running_mean"], p[prefix + "0.running_var"])
    a = prec.rb(prec.rf(torch.relu(a)))                    # activation stored bf16; its gradient comes from the dgrad kernel
    y = F.conv2d(a, prec.rf(p[prefix + "2.weight"]), p[prefix + "2.bias"], padding=1)
    if residual is not None:
        y = y + residual
    return prec.rf(y)


def res_block(x, p, prefix: str, prec: Prec, **kw):
    """ResBlock2D (reference modules.py:116-130).  The two gradient contributions to x (through the block, and the skip)
    are bf16 tensors added in bf16."""
    x = prec.rb(x)
    h = nac_block(x, p, prefix + "layers.0.layers.", prec, **kw)
    return nac_block(h, p, prefix + "layers.1.layers.", prec, residual=x, **kw)


def mid_conv(z, p, prec: Prec):
    """z (fp32 NCHW latent) -> bf16 NHWC -> 1x1 conv (reference models.py:750 / 1096)."""
    zb = prec.rb(prec.rf(z))
    return prec.rf(F.conv2d(zb, prec.rf(p["mid_conv.weight"]), p["mid_conv.bias"]))


def out_conv_loss(d, x, p, prec: Prec, w_rec: float, folded: bool):
    """7x7 out_conv -> sigmoid -> MSE (reference models.py:1099,1110; losses.py:396-403).  Logits stay fp32; the gradient
    d(R)/d(logits) is stored bf16 BEFORE the upstream weight w_rec is applied (folded kernels: applied in fp32 in the
    dgrad / wgrad epilogues; generic path: applied by a bf16 scale pass, i.e. a second rounding)."""
    d = prec.rb(d)
    logits = F.conv2d(d, prec.rf(p["out_conv.weight"]), p["out_conv.bias"], padding=3)
    lg = prec.rb(logits, w_rec)
    if not folded:
        lg = prec.rb(lg)
    x_hat = torch.sigmoid(lg)
    return x_hat, O.recon_mse(x, x_hat), logits


def anchor_forward(p, x, eps, cfg: O.AnchorConfig = O.CFG_256, prec: Prec = BF16, updates: Optional[dict] = None,
                   taps: Optional[dict] = None, folded_out_conv: Optional[bool] = None):
    """The anchor composition as face_vae_b200.models.FaceVAE.forward_loss runs it."""
    t = taps if taps is not None else {}
    kw = dict(updates=updates)
    n_enc = len(cfg.down_seq) - 1
    h = cna_block(x, p, "enc.0.layers.layers.", 1, prec, first_layer_pointwise=(x.shape[1] <= 4), **kw)
    t["enc.0"] = h
    for i in range(1, n_enc):
        h = cna_block(h, p, f"enc.{i}.layers.0.layers.", 3, prec, post="pool", out_fp32=(i == n_enc - 1), **kw)
        t[f"enc.{i}"] = h
    mu, logstd, z = O.reparameterise(h, eps, True, cfg.zc)
    t["z"] = z
    d = mid_conv(z, p, prec)
    t["mid_conv"] = d
    for r in range(cfg.n_res):
        d = res_block(d, p, f"res.{r}.", prec, **kw)
        t[f"res.{r}"] = d
    for i in range(len(cfg.up_seq) - 1):
        d = cna_block(d, p, f"up.{i}.layers.1.layers.", 3, prec, upsample=True, **kw)
        t[f"up.{i}"] = d
    if folded_out_conv is None:
        folded_out_conv = x.shape[3] in (128, 256) and cfg.up_seq[-1] == 32
    x_hat, R, logits = out_conv_loss(d, x, p, prec, cfg.w_rec, folded_out_conv)
    K = O.kl_divergence(mu, logstd)
    return {"mu": mu, "logstd": logstd, "z": z, "logits": logits, "x_hat": x_hat, "K": K, "R": R, "loss": cfg.w_kl * K + cfg.w_rec * R}


def anchor_train_grads(p, x, eps, cfg: O.AnchorConfig = O.CFG_256, prec: Prec = BF16):
    """Forward + backward of the weighted loss -> (outputs, grads, running-stat updates, block outputs)."""
    leaf = {}
    for k, v in p.items():
        leaf[k] = v.clone() if k.endswith("running_mean") or k.endswith("running_var") else v.clone().requires_grad_(True)
    updates, taps = {}, {}
    out = anchor_forward(leaf, x, eps, cfg, prec, updates, taps)
    out["loss"].backward()
    grads = {k: v.grad for k, v in leaf.items() if v.requires_grad}
    out = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    return out, grads, updates, {k: v.detach() for k, v in taps.items()}
