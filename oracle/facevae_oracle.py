"""CPU restatement of the face-vae training hot path (TEST INFRASTRUCTURE ONLY).

Every function restates, in explicit fp32 (or fp64) tensor arithmetic, what the
reference's own classes compute for the path scoped in SURVEY.md section 8.  The
convolution contraction itself lives in an un-vendored third-party dependency
(PyTorch; the reference pins no version, SURVEY.md 8c -- torch 2.11.0+cu128 is
what is installed here), so ``conv2d`` below calls ``torch.nn.functional.conv2d``
on CPU in fp32/fp64; everything around it (batch norm, pooling, up-sampling,
reparameterisation, KL, reconstruction loss) is written out from the formulas.

Pinned against the reference: ``tests/golden/make_golden.py`` runs the
*unmodified* reference classes from ``/root/reference`` on deterministic inputs
and commits their outputs under ``tests/golden``; ``tests/test_oracle_golden.py``
checks this file against those fixtures and against the closed-form
known-answer table of SURVEY.md section 4.

Nothing in the product package imports this module.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import detgen

BN_EPS = 1e-5        # nn.SyncBatchNorm default (reference modules.py:19)
BN_MOMENTUM = 0.1    # nn.SyncBatchNorm default


# ----------------------------------------------------------------------------
# Blocks (reference modules.py)
# ----------------------------------------------------------------------------
def conv2d(x, weight, bias, stride=1, padding=0):
    """nn.Conv2d as built by _ConvBlock (reference modules.py:15,32)."""
    return F.conv2d(x, weight, bias, stride=stride, padding=padding)


def batch_norm_train(x, gamma, beta, running_mean=None, running_var=None,
                     eps=BN_EPS, momentum=BN_MOMENTUM):
    """Training-mode SyncBatchNorm without a process group == batch norm over
    (N, H, W) (reference modules.py:19; torch/nn/modules/batchnorm.py:818-819).

    Normalises with the *biased* variance; running_var is updated with the
    *unbiased* variance.  Returns (y, new_running_mean, new_running_var).
    """
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
    invstd = 1.0 / torch.sqrt(var + eps)
    y = (x - mean[None, :, None, None]) * (invstd * gamma)[None, :, None, None] + beta[None, :, None, None]
    new_rm = new_rv = None
    if running_mean is not None:
        new_rm = (1 - momentum) * running_mean + momentum * mean.detach()
        new_rv = (1 - momentum) * running_var + momentum * (var.detach() * n / max(n - 1, 1))
    return y, new_rm, new_rv


def batch_norm_eval(x, gamma, beta, running_mean, running_var, eps=BN_EPS):
    invstd = 1.0 / torch.sqrt(running_var + eps)
    return (x - running_mean[None, :, None, None]) * (invstd * gamma)[None, :, None, None] + beta[None, :, None, None]


def avg_pool2(x):
    """nn.AvgPool2d((2, 2)) (reference modules.py:62,70)."""
    n, c, h, w = x.shape
    return x.reshape(n, c, h // 2, 2, w // 2, 2).mean(dim=(3, 5))


def upsample_nearest2(x):
    """nn.Upsample(scale_factor=(2, 2)), default mode nearest (reference modules.py:81,89)."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _act(x, nonlinearity):
    if nonlinearity == "relu":
        return torch.relu(x)
    if nonlinearity == "leakyrelu":
        return torch.where(x > 0, x, 0.2 * x)
    raise ValueError(nonlinearity)


class BNState:
    """Collects running-stat updates (the oracle is functional)."""

    def __init__(self):
        self.updates: Dict[str, torch.Tensor] = {}


def spectral_norm_weight(w, u, v, training=True, eps=1e-12, n_power_iterations=1):
    """torch.nn.utils.spectral_norm as ``_ConvBlock`` applies it with ``use_weight_norm=True`` (reference modules.py:11,14,32;
    torch/nn/utils/spectral_norm.py): in training mode one power iteration v <- normalize(W^T u), u <- normalize(W v) on the
    [Co, Ci*k*k] matrix without gradient, then W / (u^T W v).  Returns (w_sn, u_new, v_new)."""
    mat = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            for _ in range(n_power_iterations):
                v = F.normalize(torch.mv(mat.t(), u), dim=0, eps=eps)
                u = F.normalize(torch.mv(mat, v), dim=0, eps=eps)
    sigma = torch.dot(u, torch.mv(mat, v))
    return w / sigma, u, v


def instance_norm(x, gamma, beta, eps=BN_EPS):
    """nn.InstanceNorm2d(C, affine=True) (reference modules.py:21): per-(sample, channel) statistics over (H, W), biased variance."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * gamma[None, :, None, None] + beta[None, :, None, None]


def conv_block(pattern, x, p, prefix, kernel_size, stride, padding, training=True,
               nonlinearity="relu", bn_state: Optional[BNState] = None):
    """_ConvBlock.forward (reference modules.py:8-42): layers applied in ``pattern`` order.  The variant is read off the
    parameter names, as a state_dict of the reference block carries them: ``weight_orig`` / ``weight_u`` / ``weight_v`` =
    spectral norm (use_weight_norm=True); a norm layer with running statistics = SyncBatchNorm, with affine parameters only =
    InstanceNorm2d, without parameters = Identity (activation_type "none")."""
    for idx, ch in enumerate(pattern):
        key = f"{prefix}layers.{idx}."
        if ch == "C":
            if key + "weight_orig" in p:
                w, u, v = spectral_norm_weight(p[key + "weight_orig"], p[key + "weight_u"], p[key + "weight_v"], training)
                if bn_state is not None and training:
                    bn_state.updates[key + "weight_u"] = u
                    bn_state.updates[key + "weight_v"] = v
            else:
                w = p[key + "weight"]
            x = conv2d(x, w, p[key + "bias"], stride, padding)
        elif ch == "N":
            if key + "weight" not in p:
                continue
            if key + "running_mean" not in p:
                x = instance_norm(x, p[key + "weight"], p[key + "bias"])
                continue
            if training:
                x, rm, rv = batch_norm_train(x, p[key + "weight"], p[key + "bias"],
                                             p.get(key + "running_mean"), p.get(key + "running_var"))
                if bn_state is not None and rm is not None:
                    bn_state.updates[key + "running_mean"] = rm
                    bn_state.updates[key + "running_var"] = rv
            else:
                x = batch_norm_eval(x, p[key + "weight"], p[key + "bias"],
                                    p[key + "running_mean"], p[key + "running_var"])
        elif ch == "A":
            x = _act(x, nonlinearity)
    return x


def down_block(x, p, prefix, **kw):
    """DownBlock2D (reference modules.py:59-70): CNA 3x3 s1 p1 -> AvgPool2d(2)."""
    return avg_pool2(conv_block("CNA", x, p, prefix + "layers.0.", 3, 1, 1, **kw))


def up_block(x, p, prefix, **kw):
    """UpBlock2D (reference modules.py:78-89): Upsample(x2 nearest) -> CNA 3x3 s1 p1."""
    return conv_block("CNA", upsample_nearest2(x), p, prefix + "layers.1.", 3, 1, 1, **kw)


def same_block(x, p, prefix, **kw):
    """SameBlock2D (reference modules.py:97-108): CNA 1x1."""
    return conv_block("CNA", x, p, prefix + "layers.", 1, 1, 0, **kw)


def res_block(x, p, prefix, **kw):
    """ResBlock2D (reference modules.py:116-130): x + NAC(NAC(x))."""
    h = conv_block("NAC", x, p, prefix + "layers.0.", 3, 1, 1, **kw)
    h = conv_block("NAC", h, p, prefix + "layers.1.", 3, 1, 1, **kw)
    return x + h


# ----------------------------------------------------------------------------
# VAE bottleneck and losses (reference models.py / losses.py)
# ----------------------------------------------------------------------------
def reparameterise(h, eps, train_vae=True, zc=None):
    """flatten_vae_nl.forward (reference models.py:550-570) with eps injected.

    mu / logstd are the first / last half of the channels, flattened; the network
    predicts log(sigma), not log(sigma^2).  train_vae False => z = mu and
    (None, None, x_hat) is returned.
    """
    n, c = h.shape[0], h.shape[1]
    zc = c // 2 if zc is None else zc
    mu = h[:, :zc].flatten(start_dim=1)
    logstd = h[:, zc:].flatten(start_dim=1) * (1 if train_vae else 0)
    z = mu + torch.exp(logstd) * eps * (1 if train_vae else 0)
    x_hat = z.view(n, zc, h.shape[2], h.shape[3])
    if train_vae:
        return mu, logstd, x_hat
    return None, None, x_hat


def kl_divergence(mu, logstd):
    """KLDivergenceLoss.forward (reference losses.py:385-393)."""
    return torch.mean(-0.5 - logstd + 0.5 * mu ** 2 + 0.5 * torch.exp(2 * logstd), dim=-1).mean()


def recon_mse(a, b):
    """ReconLoss.forward == nn.MSELoss mean (reference losses.py:396-403)."""
    return ((a - b) ** 2).mean()


def recon_l1(a, b):
    """nn.L1Loss mean (reference losses.py:128)."""
    return (a - b).abs().mean()


def l1(x, y):
    """reference losses.py:8-9"""
    return torch.abs(x - y)


def l2(x, y):
    """reference losses.py:12-13"""
    return (x - y) ** 2


# ----------------------------------------------------------------------------
# ELR layers, flatten_vae6, input pre-scale
# ----------------------------------------------------------------------------
def _elr_act_gain(act):
    """Gain Conv2dELR / LinearELR derive from their activation (reference models_utils.py:137-146, 648-657):
    ``act`` in {None, "relu", "leaky"} (LeakyReLU(0.2))."""
    if act == "relu":
        return float(np.sqrt(2.0))
    if act == "leaky":
        return float(np.sqrt(2.0 / (1.0 + 0.2 ** 2)))
    return 1.0


def _elr_act(x, act):
    if act == "relu":
        return torch.relu(x)
    if act == "leaky":
        return torch.where(x > 0, x, 0.2 * x)
    return x


def conv2d_elr(x, weight, bias, stride, padding, norm=None, act=None, wround=None):
    """Conv2dELR.forward without style modulation (reference models_utils.py:686-744): weight / ||weight||_(ci,r,s) when
    norm == "demod", times weightgain = actgain * (1 if demod else 1/sqrt(fan_in)); conv; + bias; activation.
    ``wround``: optional rounding applied to the effective filter (oracle/emulate.py's bf16 storage hook)."""
    co, ci, k, _ = weight.shape
    gain = _elr_act_gain(act) * (1.0 if norm == "demod" else 1.0 / float(np.sqrt(ci * k * k)))
    w = weight
    if norm == "demod":
        w = w / w.flatten(1).norm(dim=1).clamp_min(1e-12)[:, None, None, None]
    w = w * gain
    if wround is not None:
        w = wround(w)
    return _elr_act(F.conv2d(x, w, None, stride=stride, padding=padding) + bias[None, :, None, None], act)


def linear_elr(x, weight, bias, norm=None, act=None, lrmult=1.0):
    """LinearELR.forward, un-fused (reference models_utils.py:134-203)."""
    gain = _elr_act_gain(act)
    if norm is None:
        gain = gain * lrmult / float(np.sqrt(weight.shape[1]))
    w = weight / weight.norm(dim=1, keepdim=True).clamp_min(1e-12) if norm == "demod" else weight
    return _elr_act(x @ (w * gain).t() + bias[None], act)


def flatten_vae6_forward(x, p, eps, training=True):
    """flatten_vae6.forward (reference models.py:822-833) with eps injected; ``p``: state_dict of the module."""
    shape = x.shape
    h = x.flatten(start_dim=1)
    i = 0
    while f"encoder.{i}.weight" in p:
        h = linear_elr(h, p[f"encoder.{i}.weight"], p[f"encoder.{i}.bias"], "demod", "leaky")
        i += 1
    mu = linear_elr(h, p["mu_fc.weight"], p["mu_fc.bias"]) * 0.1
    logstd = linear_elr(h, p["logstd_fc.weight"], p["logstd_fc.bias"]) * 0.01
    z = mu + torch.exp(logstd) * eps if training else mu
    i = 0
    while f"decoder.{i}.weight" in p:
        z = linear_elr(z, p[f"decoder.{i}.weight"], p[f"decoder.{i}.bias"], "demod", "leaky")
        i += 1
    return mu, logstd, z.view(shape)


def flatten_vae_forward(x, p, eps, train_vae=True):
    """flatten_vae.forward (reference models.py:509-522) with eps injected: LinearELR encoder -> mu_fc * 0.1, logstd_fc * 0.01 ->
    z = mu + exp(logstd) * eps -> view back to the input shape (no decoder).  train_vae False: logstd and the noise are
    multiplied by 0 (z = mu) and (None, None, x_hat) is returned."""
    shape = x.shape
    h = x.flatten(start_dim=1)
    i = 0
    while f"encoder.{i}.weight" in p:
        h = linear_elr(h, p[f"encoder.{i}.weight"], p[f"encoder.{i}.bias"], "demod", "leaky")
        i += 1
    on = 1 if train_vae else 0
    mu = linear_elr(h, p["mu_fc.weight"], p["mu_fc.bias"]) * 0.1
    logstd = linear_elr(h, p["logstd_fc.weight"], p["logstd_fc.bias"]) * 0.01 * on
    z = mu + torch.exp(logstd) * eps * on
    x_hat = z.view(shape)
    return (mu, logstd, x_hat) if train_vae else (None, None, x_hat)


def local_vae_forward(x, p, **kw):
    """local_vae.forward (reference models.py:471-482): DownBlock2D encoder -> flatten -> map_fc1 (LinearELR demod, LeakyReLU) ->
    map_fc2 -> view(b, up_seq[0], 4, 4) -> UpBlock2D decoder.  The sampling lines are commented out in the reference: the
    module is a deterministic bottleneck and returns (None, None, x_hat).  ``kw``: conv_block options (training, ...)."""
    b = x.shape[0]
    h, i = x, 0
    while f"encoder.{i}.layers.0.layers.0.weight" in p:
        h = down_block(h, p, f"encoder.{i}.", **kw)
        i += 1
    h = linear_elr(h.flatten(start_dim=1), p["map_fc1.weight"], p["map_fc1.bias"], "demod", "leaky")
    h = linear_elr(h, p["map_fc2.weight"], p["map_fc2.bias"], "demod", "leaky")
    c0 = p["decoder.0.layers.1.layers.0.weight"].shape[1] if "decoder.0.layers.1.layers.0.weight" in p else h.shape[1] // 16
    h = h.view(b, c0, 4, 4)
    i = 0
    while f"decoder.{i}.layers.1.layers.0.weight" in p:
        h = up_block(h, p, f"decoder.{i}.", **kw)
        i += 1
    return None, None, h


def bilinear_prescale(x, scale_factor=0.25):
    """F.interpolate(x, mode="bilinear", scale_factor=s, align_corners=False, recompute_scale_factor=True) written out
    (reference models.py:764): Ho = floor(H s); source index (o + 0.5) * (H / Ho) - 0.5 clamped at 0; 4-neighbour blend."""
    n, c, h, w = x.shape
    ho, wo = int(np.floor(h * scale_factor)), int(np.floor(w * scale_factor))

    def axis(size, osize):
        src = (torch.arange(osize, dtype=torch.float32) + 0.5) * (np.float32(size) / np.float32(osize)) - 0.5
        src = src.clamp_min(0)
        i0 = src.floor().long()
        i1 = (i0 + (i0 < size - 1).long())
        l1 = src - i0.float()
        return i0, i1, 1 - l1, l1

    h0, h1, lh0, lh1 = axis(h, ho)
    w0, w1, lw0, lw1 = axis(w, wo)
    top = x[:, :, h0][:, :, :, w0] * lw0 + x[:, :, h0][:, :, :, w1] * lw1
    bot = x[:, :, h1][:, :, :, w0] * lw0 + x[:, :, h1][:, :, :, w1] * lw1
    return top * lh0[:, None] + bot * lh1[:, None]


# ----------------------------------------------------------------------------
# Anchor model ("face-vae")
# ----------------------------------------------------------------------------
@dataclass
class AnchorConfig:
    down_seq: Sequence[int] = (3, 32, 64, 128, 256, 32)   # EFE_conv5.down, reference models.py:731,749
    up_seq: Sequence[int] = (256, 256, 128, 64, 32)        # Generator.up pattern, reference models.py:1098
    n_res: int = 2                                         # Generator.res pattern, reference models.py:1097
    w_kl: float = 0.2                                      # trainer.py:250 (commented intent)
    w_rec: float = 10.0                                    # trainer.py:251 (commented intent)

    @property
    def zc(self):
        return self.down_seq[-1] // 2

    @property
    def n_down(self):
        return len(self.down_seq) - 2


CFG_256 = AnchorConfig()
CFG_512 = AnchorConfig(down_seq=(3, 32, 64, 128, 256, 512, 64), up_seq=(512, 512, 256, 128, 64, 32))


def anchor_param_shapes(cfg: AnchorConfig = CFG_256) -> "OrderedDict[str, Tuple[int, ...]]":
    """state_dict names/shapes of the anchor, matching the reference's own key
    layout for each block."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def conv(prefix, ci, co, k):
        s[prefix + "weight"] = (co, ci, k, k)
        s[prefix + "bias"] = (co,)

    def bn(prefix, c):
        s[prefix + "weight"] = (c,)
        s[prefix + "bias"] = (c,)
        s[prefix + "running_mean"] = (c,)
        s[prefix + "running_var"] = (c,)

    d = cfg.down_seq
    conv("enc.0.layers.layers.0.", d[0], d[1], 1)
    bn("enc.0.layers.layers.1.", d[1])
    for i in range(1, len(d) - 1):
        conv(f"enc.{i}.layers.0.layers.0.", d[i], d[i + 1], 3)
        bn(f"enc.{i}.layers.0.layers.1.", d[i + 1])
    u = cfg.up_seq
    conv("mid_conv.", cfg.zc, u[0], 1)
    for r in range(cfg.n_res):
        for j in range(2):
            bn(f"res.{r}.layers.{j}.layers.0.", u[0])
            conv(f"res.{r}.layers.{j}.layers.2.", u[0], u[0], 3)
    for i in range(len(u) - 1):
        conv(f"up.{i}.layers.1.layers.0.", u[i], u[i + 1], 3)
        bn(f"up.{i}.layers.1.layers.1.", u[i + 1])
    conv("out_conv.", u[-1], 3, 7)
    return s


def det_anchor_params(cfg: AnchorConfig = CFG_256, base: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Closed-form deterministic parameters (no RNG dependence): conv weights
    uniform(+-1/sqrt(fan_in)) like nn.Conv2d's default range, BN gamma in
    [0.5, 1.5], beta in [-0.2, 0.2], running stats at their initial values."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in anchor_param_shapes(cfg).items():
        seed = detgen.name_seed(name, base)
        if name.endswith("running_mean"):
            v = np.zeros(shape, np.float32)
        elif name.endswith("running_var"):
            v = np.ones(shape, np.float32)
        elif len(shape) == 4:
            bound = 1.0 / np.sqrt(shape[1] * shape[2] * shape[3])
            v = detgen.det_uniform(shape, seed, -bound, bound)
        elif name.endswith("weight"):
            v = detgen.det_uniform(shape, seed, 0.5, 1.5)      # BN gamma (the only 1-d weights)
        else:
            # biases: conv bias in +-1/sqrt(fan_in) is tiny; use a visible range
            v = detgen.det_uniform(shape, seed, -0.2, 0.2)
        out[name] = torch.from_numpy(v).to(dtype)
    return out


def det_inputs(n: int, h: int, w: int, cfg: AnchorConfig = CFG_256, base: int = 0, dtype=torch.float32):
    """Synthetic face-shaped frames in [0, 1) and the injected eps ~ approx N(0,1)."""
    x = torch.from_numpy(detgen.det_unit((n, 3, h, w), 1000 + base)).to(dtype)
    f = 2 ** cfg.n_down
    eps = torch.from_numpy(detgen.det_normal((n, cfg.zc * (h // f) * (w // f)), 2000 + base)).to(dtype)
    return x, eps


def anchor_forward(p: Dict[str, torch.Tensor], x, eps, cfg: AnchorConfig = CFG_256, training=True,
                   train_vae=True, bn_state: Optional[BNState] = None, taps: Optional[dict] = None):
    """The anchor composition of SURVEY.md section 8.  Returns a dict with mu,
    logstd, z, logits, x_hat, K, R, loss (K and R un-weighted)."""
    kw = dict(training=training, bn_state=bn_state)
    t = taps if taps is not None else {}
    h = same_block(x, p, "enc.0.", **kw)
    t["enc.0"] = h
    for i in range(1, len(cfg.down_seq) - 1):
        h = down_block(h, p, f"enc.{i}.", **kw)
        t[f"enc.{i}"] = h
    mu, logstd, z = reparameterise(h, eps, train_vae, cfg.zc)
    t["z"] = z
    d = conv2d(z, p["mid_conv.weight"], p["mid_conv.bias"], 1, 0)
    t["mid_conv"] = d
    for r in range(cfg.n_res):
        d = res_block(d, p, f"res.{r}.", **kw)
        t[f"res.{r}"] = d
    for i in range(len(cfg.up_seq) - 1):
        d = up_block(d, p, f"up.{i}.", **kw)
        t[f"up.{i}"] = d
    logits = conv2d(d, p["out_conv.weight"], p["out_conv.bias"], 1, 3)
    x_hat = torch.sigmoid(logits)                       # reference models.py:1110
    out = {"mu": mu, "logstd": logstd, "z": z, "logits": logits, "x_hat": x_hat}
    if train_vae:
        out["K"] = kl_divergence(mu, logstd)            # trainer.py:312 (un-weighted here)
    else:
        out["K"] = torch.zeros((), dtype=x.dtype)
    out["R"] = recon_mse(x, x_hat)                      # trainer.py:314: ReconLoss((d, generated_d))
    out["loss"] = cfg.w_kl * out["K"] + cfg.w_rec * out["R"]
    return out


def anchor_train_grads(p: Dict[str, torch.Tensor], x, eps, cfg: AnchorConfig = CFG_256):
    """Forward + backward of the weighted loss; returns (outputs, grads, bn running-stat updates, taps)."""
    leaf = OrderedDict()
    for k, v in p.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            leaf[k] = v.clone()
        else:
            leaf[k] = v.clone().requires_grad_(True)
    st = BNState()
    taps: dict = {}
    out = anchor_forward(leaf, x, eps, cfg, True, True, st, taps)
    out["loss"].backward()
    grads = OrderedDict((k, v.grad) for k, v in leaf.items() if v.requires_grad)
    out = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    taps = {k: v.detach() for k, v in taps.items()}
    return out, grads, st.updates, taps


def adam_step(param, grad, m, v, step, lr=5e-5, b1=0.5, b2=0.999, eps=1e-8):
    """torch.optim.Adam as configured by the reference (logger.py:60: lr, betas=(0.5, 0.999))."""
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    mh = m / (1 - b1 ** step)
    vh = v / (1 - b2 ** step)
    return param - lr * mh / (vh.sqrt() + eps), m, v


def conv_flops_per_image(h: int, w: int, cfg: AnchorConfig = CFG_256, train: bool = True) -> float:
    """2*MAC count of every conv on the anchor for one image (SURVEY.md section 8,
    'Per-image conv FLOPs'); train = fwd + dgrad + wgrad, no dgrad for the first layer."""
    total = 0.0
    first = True

    def add(ci, co, k, hh, ww):
        nonlocal total, first
        f = 2.0 * hh * ww * ci * co * k * k
        total += f * ((2 if first else 3) if train else 1)
        first = False

    d = cfg.down_seq
    add(d[0], d[1], 1, h, w)
    hh, ww = h, w
    for i in range(1, len(d) - 1):
        add(d[i], d[i + 1], 3, hh, ww)
        hh, ww = hh // 2, ww // 2
    u = cfg.up_seq
    add(cfg.zc, u[0], 1, hh, ww)
    for _ in range(cfg.n_res * 2):
        add(u[0], u[0], 3, hh, ww)
    for i in range(len(u) - 1):
        hh, ww = hh * 2, ww * 2
        add(u[i], u[i + 1], 3, hh, ww)
    add(u[-1], 3, 7, hh, ww)
    return total
