#!/usr/bin/env python
"""Generate the golden fixtures by executing the UNMODIFIED reference classes.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``  ->  tests/golden/*.npz

The reference ships no tests or golden vectors, so the
pins are outputs of its own classes (modules.py blocks, models.flatten_vae_nl,
losses.KLDivergenceLoss / ReconLoss) on CPU fp32.  Inputs and weights come from
the closed-form generators in ``oracle/detgen.py`` and are therefore NOT stored:
tests regenerate them.  ``flatten_vae_nl`` draws eps inline with ``torch.randn``
(models.py:561); the draw is intercepted (torch.randn is patched for the duration
of the call, the reference source is untouched) so that eps is the deterministic
tensor the tests also use.

Large tensors are stored as (sum, l2, absmax) plus a strided sample so the
fixtures stay small.
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict
from contextlib import contextmanager

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("FACEVAE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import modules as ref_modules          # noqa: E402  (reference)
import models as ref_models            # noqa: E402  (reference)
import losses as ref_losses            # noqa: E402  (reference)

from oracle import detgen              # noqa: E402
from oracle import facevae_oracle as O  # noqa: E402

SAMPLE = 4099  # prime count of strided samples per large tensor


def summarise(t: torch.Tensor, full_below: int = 4096):
    a = t.detach().to(torch.float64).flatten().numpy()
    d = {"shape": np.array(t.shape, np.int64), "sum": a.sum(), "l2": np.sqrt((a * a).sum()),
         "absmax": np.abs(a).max() if a.size else 0.0}
    if a.size <= full_below:
        d["full"] = a.astype(np.float32)
    else:
        idx = sample_index(a.size)
        d["sample"] = a[idx].astype(np.float32)
    return d


def sample_index(n: int) -> np.ndarray:
    return (np.arange(SAMPLE, dtype=np.int64) * (n // SAMPLE + 1) * 7919 + 13) % n


def rel_l2(a: torch.Tensor, ref: torch.Tensor) -> float:
    a, ref = a.detach().double().flatten(), ref.detach().double().flatten()
    return float((a - ref).norm() / ref.norm().clamp_min(1e-30))


def put_yard(store: dict, name: str, lowp: torch.Tensor, ref: torch.Tensor):
    """Yardstick: how far the reference's OWN bf16-autocast run is from its fp32 run (relative L2 over the full tensor,
    and max error / max |ref|).  The bf16 CUDA path is required to be no worse than this (tests/test_parity_gpu.py)."""
    store[f"{name}/yard_l2"] = np.float64(rel_l2(lowp.float(), ref))
    store[f"{name}/yard_max"] = np.float64(float((lowp.detach().double() - ref.detach().double()).abs().max() /
                                                 ref.detach().double().abs().max().clamp_min(1e-30)))


def put(store: dict, name: str, t: torch.Tensor, **kw):
    for k, v in summarise(t, **kw).items():
        store[f"{name}/{k}"] = v


@contextmanager
def injected_randn(eps: torch.Tensor):
    orig = torch.randn

    def fake(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        assert shape == tuple(eps.shape), (shape, eps.shape)
        return eps.clone()

    torch.randn = fake
    try:
        yield
    finally:
        torch.randn = orig


class RefAnchor(nn.Module):
    """Composition of SURVEY.md section 8 from unmodified reference classes."""

    def __init__(self, cfg: O.AnchorConfig):
        super().__init__()
        d, u = cfg.down_seq, cfg.up_seq
        self.enc = nn.Sequential(*[ref_modules.SameBlock2D(d[i], d[i + 1], False) if i == 0 else
                                   ref_modules.DownBlock2D(d[i], d[i + 1], False) for i in range(len(d) - 1)])
        self.vae = ref_models.flatten_vae_nl()
        self.mid_conv = nn.Conv2d(cfg.zc, u[0], 1, 1, 0)
        self.res = nn.Sequential(*[ref_modules.ResBlock2D(u[0], False) for _ in range(cfg.n_res)])
        self.up = nn.Sequential(*[ref_modules.UpBlock2D(u[i], u[i + 1], False) for i in range(len(u) - 1)])
        self.out_conv = nn.Conv2d(u[-1], 3, 7, 1, 3)
        self.kl = ref_losses.KLDivergenceLoss()
        self.rec = ref_losses.ReconLoss()
        self.cfg = cfg

    def forward(self, x, eps, taps):
        h = x
        for i, blk in enumerate(self.enc):
            h = blk(h)
            taps[f"enc.{i}"] = h
        if tuple(h.shape[1:]) == (32, 4, 4):
            with injected_randn(eps):
                mu, logstd, z = self.vae(h, True)
        else:
            # flatten_vae_nl hard-codes view(b, 16, 4, 4) (reference models.py:564), i.e. 64x64 frames; at other sizes the
            # same three lines (models.py:559-561) are evaluated here with the latent's real shape.  golden_losses() stores
            # the class's own outputs at 4x4, against which tests check this formula (vae/train_*).
            zc = h.shape[1] // 2
            mu = h[:, :zc].flatten(start_dim=1)
            logstd = h[:, zc:].flatten(start_dim=1) * 1
            z = (mu + torch.exp(logstd) * eps * 1).view(h.shape[0], zc, h.shape[2], h.shape[3])
        taps["z"] = z
        dd = self.mid_conv(z)
        taps["mid_conv"] = dd
        for i, blk in enumerate(self.res):
            dd = blk(dd)
            taps[f"res.{i}"] = dd
        for i, blk in enumerate(self.up):
            dd = blk(dd)
            taps[f"up.{i}"] = dd
        logits = self.out_conv(dd)
        x_hat = torch.sigmoid(logits)
        K = self.kl((mu, logstd))
        R = self.rec((x, x_hat))
        return dict(mu=mu, logstd=logstd, z=z, logits=logits, x_hat=x_hat, K=K, R=R,
                    loss=self.cfg.w_kl * K + self.cfg.w_rec * R)


def load_det(model: nn.Module, params):
    sd = model.state_dict()
    for k, v in params.items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
        sd[k] = v.clone()
    missing = [k for k in sd if k not in params and not k.endswith("num_batches_tracked")]
    assert not missing, missing
    model.load_state_dict(sd)


def golden_anchor(n, hw, base, path, cfg=O.CFG_256):
    p = O.det_anchor_params(cfg, base)
    x, eps = O.det_inputs(n, hw, hw, cfg, base)
    m = RefAnchor(cfg).train()
    load_det(m, p)
    taps = {}
    out = m(x, eps, taps)
    out["loss"].backward()
    store = {"meta/n": n, "meta/hw": hw, "meta/base": base}
    for k in ("K", "R", "loss"):
        store[f"out/{k}"] = np.float64(out[k].item())
    for k in ("mu", "logstd", "z"):
        put(store, f"out/{k}", out[k].reshape(n, -1))
    put(store, "out/logits", out["logits"])
    put(store, "out/x_hat", out["x_hat"])
    for k, v in taps.items():
        put(store, f"act/{k}", v)
    for k, v in m.named_parameters():
        put(store, f"grad/{k}", v.grad)
    for k, v in m.named_buffers():
        if k.endswith("running_mean") or k.endswith("running_var"):
            put(store, f"buf/{k}", v)
    # yardstick: the same reference modules under CPU bf16 autocast
    m2 = RefAnchor(cfg).train()
    load_det(m2, p)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out2 = m2(x, eps, {})
    out2["loss"].float().backward()
    for k in ("mu", "logstd", "x_hat"):
        put_yard(store, f"out/{k}", out2[k].reshape(out[k].shape if k == "x_hat" else (n, -1)), out[k].reshape(out[k].shape if k == "x_hat" else (n, -1)))
    g1 = dict(m.named_parameters())
    for k, v in m2.named_parameters():
        put_yard(store, f"grad/{k}", v.grad, g1[k].grad)
    store["yard/K"] = np.float64(out2["K"].item())
    store["yard/R"] = np.float64(out2["R"].item())
    # one Adam step exactly as the reference configures it (logger.py:60)
    opt = torch.optim.Adam(m.parameters(), lr=5e-5, betas=(0.5, 0.999))
    opt.step()
    for k in ("enc.1.layers.0.layers.0.weight", "out_conv.weight", "res.0.layers.0.layers.0.weight"):
        put(store, f"adam/{k}", dict(m.named_parameters())[k])
    np.savez_compressed(path, **store)
    print(path, "K=%.8f R=%.8f loss=%.8f" % (out["K"].item(), out["R"].item(), out["loss"].item()))


def golden_blocks(path):
    """Each block on its own: outputs, input grads and parameter grads, stored in full."""
    store = {}
    n, hw = 2, 8

    def run(tag, blk, ci, upstream_seed):
        import copy
        blk.train()
        sd = blk.state_dict()
        for k in list(sd):
            if k.endswith("num_batches_tracked"):
                continue
            seed = detgen.name_seed(tag + "." + k)
            shape = tuple(sd[k].shape)
            if k.endswith("running_mean"):
                v = np.zeros(shape, np.float32)
            elif k.endswith("running_var"):
                v = np.ones(shape, np.float32)
            elif len(shape) == 4:
                b = 1.0 / np.sqrt(shape[1] * shape[2] * shape[3])
                v = detgen.det_uniform(shape, seed, -b, b)
            elif k.endswith("weight"):
                v = detgen.det_uniform(shape, seed, 0.5, 1.5)
            else:
                v = detgen.det_uniform(shape, seed, -0.2, 0.2)
            sd[k] = torch.from_numpy(v)
        blk.load_state_dict(sd)
        blk2 = copy.deepcopy(blk)
        x = torch.from_numpy(detgen.det_uniform((n, ci, hw, hw), detgen.name_seed(tag + ".x"), -1.0, 1.0)).requires_grad_(True)
        y = blk(x)
        g = torch.from_numpy(detgen.det_uniform(tuple(y.shape), upstream_seed, -1.0, 1.0))
        (y * g).sum().backward()
        put(store, f"{tag}/y", y, full_below=1 << 20)
        put(store, f"{tag}/dx", x.grad, full_below=1 << 20)
        for k, v in blk.named_parameters():
            put(store, f"{tag}/grad/{k}", v.grad, full_below=1 << 20)
        for k, v in blk.named_buffers():
            if not k.endswith("num_batches_tracked"):
                put(store, f"{tag}/buf/{k}", v, full_below=1 << 20)
        blk2.train()
        x2 = x.detach().clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            y2 = blk2(x2)
        (y2.float() * g).sum().backward()
        put_yard(store, f"{tag}/y", y2, y)
        put_yard(store, f"{tag}/dx", x2.grad, x.grad)
        g1 = dict(blk.named_parameters())
        for k, v in blk2.named_parameters():
            put_yard(store, f"{tag}/grad/{k}", v.grad, g1[k].grad)

    run("down", ref_modules.DownBlock2D(16, 32, False), 16, 11)
    run("up", ref_modules.UpBlock2D(32, 16, False), 32, 12)
    run("same", ref_modules.SameBlock2D(16, 32, False), 16, 13)
    run("same3", ref_modules.SameBlock2D(3, 32, False), 3, 14)
    run("res", ref_modules.ResBlock2D(32, False), 32, 15)
    run("convblock_leaky", ref_modules.ConvBlock2D("CNA", 16, 16, 3, 1, 1, False, nonlinearity_type="leakyrelu"), 16, 16)
    np.savez_compressed(path, **store)
    print(path, len(store), "arrays")


def golden_losses(path):
    """KLDivergenceLoss / ReconLoss / l1 / l2 / flatten_vae_nl known answers."""
    store = {}
    kl = ref_losses.KLDivergenceLoss()
    rec = ref_losses.ReconLoss()
    for tag, (m, s) in {"zero": (0.0, 0.0), "mu1": (1.0, 0.0), "ls1": (0.0, 1.0), "lsm1": (0.0, -1.0)}.items():
        store[f"kl/{tag}"] = np.float64(kl((torch.full((4, 256), m), torch.full((4, 256), s))).item())
    i = torch.arange(1024, dtype=torch.float32).view(4, 256)
    mu = torch.sin(0.01 * i).requires_grad_(True)
    ls = (0.5 * torch.cos(0.013 * i)).requires_grad_(True)
    v = kl((mu, ls))
    v.backward()
    store["kl/sincos"] = np.float64(v.item())
    store["kl/sincos_dmu"] = mu.grad.numpy()
    store["kl/sincos_dlogstd"] = ls.grad.numpy()
    a = torch.sin(0.1 * torch.arange(384, dtype=torch.float32)).view(2, 3, 8, 8).requires_grad_(True)
    b = torch.cos(0.07 * torch.arange(384, dtype=torch.float32)).view(2, 3, 8, 8)
    r = rec((a, b))
    r.backward()
    store["rec/mse"] = np.float64(r.item())
    store["rec/mse_da"] = a.grad.numpy()
    store["rec/l1"] = np.float64(nn.L1Loss()(a.detach(), b).item())
    store["rec/l1_elem"] = ref_losses.l1(a.detach(), b).numpy()
    store["rec/l2_elem"] = ref_losses.l2(a.detach(), b).numpy()
    # flatten_vae_nl: eval returns (None, None, x[:, :16]); train follows mu + exp(logstd) * eps
    vae = ref_models.flatten_vae_nl()
    x = torch.from_numpy(detgen.det_uniform((3, 32, 4, 4), 77, 0.0, 2.0))
    eps = torch.from_numpy(detgen.det_normal((3, 256), 78))
    m0, s0, xh0 = vae(x, False)
    assert m0 is None and s0 is None
    store["vae/eval_xhat"] = xh0.numpy()
    with injected_randn(eps):
        m1, s1, xh1 = vae(x, True)
    store["vae/train_mu"] = m1.numpy()
    store["vae/train_logstd"] = s1.numpy()
    store["vae/train_xhat"] = xh1.numpy()
    np.savez_compressed(path, **store)
    print(path, {k: float(v) for k, v in store.items() if np.ndim(v) == 0})


def golden_f2(path):
    """the Generator's / Discriminator's block variants -- spectral norm (use_weight_norm=True), instance norm,
    3x3 stride 2, the un-normalised CN block -- from the unmodified reference ConvBlock2D / ResBlock2D / UpBlock2D."""
    store = {}
    n, hw = 2, 8

    def run(tag, blk, ci):
        blk.train()
        sd = blk.state_dict()
        for k in list(sd):
            if k.endswith("num_batches_tracked"):
                continue
            seed = detgen.name_seed(tag + "." + k)
            shape = tuple(sd[k].shape)
            if k.endswith("running_mean"):
                v = np.zeros(shape, np.float32)
            elif k.endswith("running_var"):
                v = np.ones(shape, np.float32)
            elif k.endswith("weight_u") or k.endswith("weight_v"):
                v = detgen.det_normal(shape, seed)
                v = v / np.sqrt((v * v).sum())
            elif len(shape) == 4:
                b = 1.0 / np.sqrt(shape[1] * shape[2] * shape[3])
                v = detgen.det_uniform(shape, seed, -b, b)
            elif k.endswith("weight"):
                v = detgen.det_uniform(shape, seed, 0.5, 1.5)
            else:
                v = detgen.det_uniform(shape, seed, -0.2, 0.2)
            sd[k] = torch.from_numpy(v.astype(np.float32))
        blk.load_state_dict(sd)
        x = torch.from_numpy(detgen.det_uniform((n, ci, hw, hw), detgen.name_seed(tag + ".x"), -1.0, 1.0)).requires_grad_(True)
        y = blk(x)
        g = torch.from_numpy(detgen.det_uniform(tuple(y.shape), detgen.name_seed(tag + ".g"), -1.0, 1.0))
        (y * g).sum().backward()
        put(store, f"{tag}/y", y, full_below=1 << 20)
        put(store, f"{tag}/dx", x.grad, full_below=1 << 20)
        for k, v in blk.named_parameters():
            put(store, f"{tag}/grad/{k}", v.grad)
        for k, v in blk.named_buffers():
            if not k.endswith("num_batches_tracked"):
                put(store, f"{tag}/buf/{k}", v, full_below=1 << 20)

    run("sn_in_s2", ref_modules.ConvBlock2D("CNA", 32, 64, 3, 2, 1, True, "instance", "leakyrelu"), 32)     # Discriminator down block
    run("sn_in_s1", ref_modules.ConvBlock2D("CNA", 64, 64, 3, 1, 1, True, "instance", "leakyrelu"), 64)
    run("sn_cn_none", ref_modules.ConvBlock2D("CN", 64, 1, 3, 1, 1, True, activation_type="none"), 64)      # Discriminator head
    run("sn_bn_leaky", ref_modules.ConvBlock2D("CNA", 32, 64, 3, 1, 1, True, nonlinearity_type="leakyrelu"), 32)   # Generator.in_conv
    run("sn_res", ref_modules.ResBlock2D(32, True), 32)
    run("sn_up", ref_modules.UpBlock2D(32, 16, True), 32)
    np.savez_compressed(path, **store)
    print(path, len(store), "arrays")


def golden_elr(path):
    """Conv2dELR (4x4 stride 2 + demod + LeakyReLU, and a plain 3x3), flatten_vae6, LinearELR and the
    bilinear pre-scale -- outputs of the unmodified reference classes / the reference's own F.interpolate call."""
    import models_utils as ref_mu
    store = {}

    def conv_case(tag, ci, co, k, s, pd, norm, act, hw, n=2):
        m = ref_mu.Conv2dELR(ci, co, k, s, pd, norm=norm, act=act)
        with torch.no_grad():
            m.weight.copy_(torch.from_numpy(detgen.det_normal(tuple(m.weight.shape), detgen.name_seed(tag + ".w"))))
            m.bias.copy_(torch.from_numpy(detgen.det_uniform(tuple(m.bias.shape), detgen.name_seed(tag + ".b"), -0.3, 0.3)))
        x = torch.from_numpy(detgen.det_uniform((n, ci, hw, hw), detgen.name_seed(tag + ".x"), -1.0, 1.0)).requires_grad_(True)
        y = m(x)
        gy = torch.from_numpy(detgen.det_uniform(tuple(y.shape), detgen.name_seed(tag + ".g"), -1.0, 1.0))
        (y * gy).sum().backward()
        put(store, f"{tag}/y", y, full_below=1 << 20)
        put(store, f"{tag}/dx", x.grad, full_below=1 << 20)
        put(store, f"{tag}/dw", m.weight.grad)
        put(store, f"{tag}/db", m.bias.grad, full_below=1 << 20)
        store[f"{tag}/gain"] = np.float64(m.weightgain)

    conv_case("elr_s2_demod_leaky", 32, 64, 4, 2, 1, "demod", nn.LeakyReLU(0.2), 16)       # EFE_conv6.efe_encoder layer (models.py:846)
    conv_case("elr_s2_rgb", 3, 32, 4, 2, 1, "demod", nn.LeakyReLU(0.2), 16)
    conv_case("elr_s2_plain", 16, 16, 4, 2, 1, None, None, 8)
    conv_case("elr_3x3_relu", 16, 32, 3, 1, 1, None, nn.ReLU(), 8)
    conv_case("elr_1x1_demod", 32, 16, 1, 1, 0, "demod", nn.LeakyReLU(0.2), 8)
    # flatten_vae6 (models.py:802-833), eps injected
    vae = ref_models.flatten_vae6()
    sd = vae.state_dict()
    for k in sd:
        sd[k] = torch.from_numpy(detgen.det_normal(tuple(sd[k].shape), detgen.name_seed("vae6." + k)) * (0.2 if k.endswith("bias") else 1.0))
    vae.load_state_dict(sd)
    x = torch.from_numpy(detgen.det_uniform((3, 16, 4, 4), 91, -1.0, 1.0)).requires_grad_(True)
    eps = torch.from_numpy(detgen.det_normal((3, 256), 92))
    with injected_randn(eps):
        mu, ls, xh = vae(x)
    gy = torch.from_numpy(detgen.det_uniform((3, 16, 4, 4), 93, -1.0, 1.0))
    kl = ref_losses.KLDivergenceLoss()((mu, ls))
    ((xh * gy).sum() + 3.0 * kl).backward()
    for k, v in (("mu", mu), ("logstd", ls), ("xhat", xh), ("dx", x.grad)):
        put(store, f"vae6/{k}", v, full_below=1 << 20)
    for k, v in vae.named_parameters():
        put(store, f"vae6/grad/{k}", v.grad)
    vae.training = False
    mu0, ls0, xh0 = vae(x.detach())
    put(store, "vae6/eval_xhat", xh0, full_below=1 << 20)
    # input pre-scale exactly as EFE_conv5 / EFE_conv6 call it (models.py:764)
    for tag, shape in (("pre256", (2, 3, 256, 256)), ("pre100", (1, 3, 100, 72))):
        xi = torch.from_numpy(detgen.det_unit(shape, detgen.name_seed(tag)))
        yo = torch.nn.functional.interpolate(xi, mode="bilinear", scale_factor=0.25, align_corners=False, recompute_scale_factor=True)
        put(store, f"{tag}/y", yo)
    np.savez_compressed(path, **store)
    print(path, len(store), "arrays")


def golden_vae_variants(path):
    """The other bottlenecks built from the same formula / blocks ("same formula in flatten_vae"; 8b callers:
    local_vae, models.py:461-462,473,480): outputs of the unmodified reference classes ``flatten_vae`` (models.py:484-522,
    eps injected) and ``local_vae`` (models.py:442-482)."""
    store = {}

    def det_load(m, tag):
        sd = m.state_dict()
        for k in sd:
            shp = tuple(sd[k].shape)
            if k.endswith("num_batches_tracked"):
                continue
            if k.endswith("running_var"):
                sd[k] = torch.from_numpy(detgen.det_uniform(shp, detgen.name_seed(tag + k), 0.5, 1.5))
            elif k.endswith("layers.1.weight") and len(shp) == 1:      # batch-norm gamma
                sd[k] = torch.from_numpy(detgen.det_uniform(shp, detgen.name_seed(tag + k), 0.5, 1.5))
            elif len(shp) == 4:                                        # conv filters: fan-in scaled
                sd[k] = torch.from_numpy(detgen.det_normal(shp, detgen.name_seed(tag + k)) / np.sqrt(shp[1] * shp[2] * shp[3])).float()
            else:
                sd[k] = torch.from_numpy(detgen.det_normal(shp, detgen.name_seed(tag + k)) * (0.2 if len(shp) == 1 else 1.0))
        m.load_state_dict(sd)
        store[f"{tag}keys"] = np.array(sorted(sd.keys()))

    # flatten_vae, training and not
    vae = ref_models.flatten_vae()
    det_load(vae, "fvae.")
    x = torch.from_numpy(detgen.det_uniform((3, 16, 4, 4), 191, -1.0, 1.0)).requires_grad_(True)
    eps = torch.from_numpy(detgen.det_normal((3, 256), 192))
    with injected_randn(eps):
        mu, ls, xh = vae(x, True)
    gy = torch.from_numpy(detgen.det_uniform((3, 16, 4, 4), 193, -1.0, 1.0))
    ((xh * gy).sum() + 3.0 * ref_losses.KLDivergenceLoss()((mu, ls))).backward()
    for k, v in (("mu", mu), ("logstd", ls), ("xhat", xh), ("dx", x.grad)):
        put(store, f"fvae/{k}", v, full_below=1 << 20)
    for k, v in vae.named_parameters():
        put(store, f"fvae/grad/{k}", v.grad)
    with injected_randn(eps):
        mu0, ls0, xh0 = vae(x.detach(), False)
    assert mu0 is None and ls0 is None
    put(store, "fvae/eval_xhat", xh0, full_below=1 << 20)

    # local_vae: [N, 128, 8, 8] -> DownBlock2D -> 2048 -> 512 -> 2048 -> [N, 128, 4, 4] -> UpBlock2D -> [N, 128, 8, 8]
    lv = ref_models.local_vae()
    det_load(lv, "lvae.")
    lv.train()
    x = torch.from_numpy(detgen.det_uniform((4, 128, 8, 8), 194, -1.0, 1.0)).requires_grad_(True)
    m0, l0, xh = lv(x)
    assert m0 is None and l0 is None
    gy = torch.from_numpy(detgen.det_uniform(tuple(xh.shape), 195, -1.0, 1.0))
    (xh * gy).sum().backward()
    put(store, "lvae/xhat", xh, full_below=1 << 20)
    put(store, "lvae/dx", x.grad, full_below=1 << 20)
    for k, v in lv.named_parameters():
        put(store, f"lvae/grad/{k}", v.grad, full_below=1 << 12)
    for k, v in lv.named_buffers():
        if not k.endswith("num_batches_tracked"):
            put(store, f"lvae/buf/{k}", v, full_below=1 << 20)
    lv.eval()
    _, _, xh_e = lv(x.detach())
    put(store, "lvae/eval_xhat", xh_e, full_below=1 << 20)
    np.savez_compressed(path, **store)
    print(path, len(store), "arrays")


def golden_efe5(path):
    """the 2-D stage of the reference's EFE_conv5 (models.py:764-787) executed with the reference's own
    sub-modules in the reference's order -- F.interpolate pre-scale -> ``down`` -> ``vae`` (flatten_vae_nl, eps injected) ->
    ``mid_conv`` -> view(N, C, D, H, W).  (The rest of EFE_conv5.forward needs key points and CUDA-only helpers.)"""
    store = {}
    m = ref_models.EFE_conv5()
    sd = m.state_dict()
    keys = [k for k in sd if k.startswith("down.") or k.startswith("mid_conv.")]
    for k in keys:
        shp = tuple(sd[k].shape)
        if k.endswith("num_batches_tracked"):
            continue
        if k.endswith("running_var") or (k.endswith("layers.1.weight") and len(shp) == 1):
            sd[k] = torch.from_numpy(detgen.det_uniform(shp, detgen.name_seed("efe5." + k), 0.5, 1.5))
        elif len(shp) == 4:
            sd[k] = torch.from_numpy(detgen.det_normal(shp, detgen.name_seed("efe5." + k)) / np.sqrt(shp[1] * shp[2] * shp[3])).float()
        else:
            sd[k] = torch.from_numpy(detgen.det_normal(shp, detgen.name_seed("efe5." + k)) * 0.2)
    m.load_state_dict(sd)
    m.train()
    store["keys"] = np.array(sorted(keys))
    x = torch.from_numpy(detgen.det_unit((2, 3, 256, 256), 291))
    eps = torch.from_numpy(detgen.det_normal((2, 256), 292))
    xs = torch.nn.functional.interpolate(x, mode="bilinear", scale_factor=m.scale_factor, align_corners=False, recompute_scale_factor=True)
    h = m.down(xs)
    with injected_randn(eps):
        mu, ls, xhat = m.vae(h, True)
    y = m.mid_conv(xhat)
    n, _, hh, ww = y.shape
    y3 = y.view(n, m.C, m.D, hh, ww)
    gy = torch.from_numpy(detgen.det_uniform(tuple(y3.shape), 293, -1.0, 1.0))
    (y3 * gy).sum().backward()
    put(store, "h", h, full_below=1 << 20)
    put(store, "mu", mu, full_below=1 << 20)
    put(store, "logstd", ls, full_below=1 << 20)
    put(store, "x3d", y3)
    for k, v in m.named_parameters():
        if (k.startswith("down.") or k.startswith("mid_conv.")) and v.grad is not None:
            put(store, f"grad/{k}", v.grad)
    np.savez_compressed(path, **store)
    print(path, len(store), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    golden_losses(os.path.join(HERE, "losses.npz"))
    golden_blocks(os.path.join(HERE, "blocks.npz"))
    golden_elr(os.path.join(HERE, "elr.npz"))
    golden_f2(os.path.join(HERE, "f2.npz"))
    golden_vae_variants(os.path.join(HERE, "vae_variants.npz"))
    golden_efe5(os.path.join(HERE, "efe5.npz"))
    golden_anchor(4, 64, 0, os.path.join(HERE, "anchor_n4_64.npz"))     # BASELINE.json configs[0]
    golden_anchor(2, 64, 1, os.path.join(HERE, "anchor_n2_64_b1.npz"))
    if "--large" in sys.argv:       # minutes of CPU time: BASELINE.json configs[1] and the configs[3] architecture at 512x512
        golden_anchor(32, 256, 7, os.path.join(HERE, "anchor_n32_256.npz"))
        golden_anchor(2, 512, 2, os.path.join(HERE, "anchor512_n2_512.npz"), O.CFG_512)
