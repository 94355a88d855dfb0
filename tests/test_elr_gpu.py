"""SURVEY.md 8f rows 1 and 3 on the GPU, against outputs of the unmodified reference classes (tests/golden/elr.npz):
Conv2dELR (4x4 stride 2 + weight demodulation + LeakyReLU -- the EFE_conv6.efe_encoder layer, reference models.py:845-852,
models_utils.py:632-744 -- and the stride-1 variants), flatten_vae6 / LinearELR (models.py:802-833), the bilinear input
pre-scale (models.py:764) and the 2-D stage of EFE_conv5 up to the 3-D hand-off (models.py:764-787).
bf16 path: outputs rtol 2e-2; fp32 pieces (LinearELR bottleneck, pre-scale): 1e-4."""
import numpy as np
import pytest
import torch
from torch import nn

from oracle import detgen
from oracle import facevae_oracle as O
from tests import goldenlib as G
from tests.test_oracle_golden import ELR_CASES, elr_case_tensors, vae6_params

pytestmark = pytest.mark.gpu
ACTS = {None: None, "relu": nn.ReLU(), "leaky": nn.LeakyReLU(0.2)}


@pytest.fixture(scope="module")
def fv():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import face_vae_b200.models as MO
    import face_vae_b200.modules as M
    from face_vae_b200 import _lib, ops
    _lib.call("fv_device_ok")

    class NS:
        modules, models, o = M, MO, ops
    return NS


@pytest.mark.parametrize("tag", sorted(ELR_CASES))
def test_conv2d_elr_module(fv, tag):
    g = G.load("elr.npz")
    ci, co, k, s, pd, norm, act, hw = ELR_CASES[tag]
    w, b, x, gy = elr_case_tensors(tag)
    m = fv.modules.Conv2dELR(ci, co, k, s, pd, norm=norm, act=ACTS[act])
    assert abs(m.weightgain - float(g[f"{tag}/gain"])) < 1e-12
    with torch.no_grad():
        m.weight.copy_(w)
        m.bias.copy_(b)
    m = m.cuda()
    # bf16-representable input and upstream gradient, and the effective filter rounded to bf16 where the CUDA path rounds it
    # (oracle/emulate.py's storage hook): what is left is accumulation order and the bf16 rounding of the outputs -- with an
    # fp32 filter on the reference side ~0.3 % of the ReLU masks flip and dx cannot be compared per element
    from oracle import emulate as E
    xq = x.bfloat16().float()
    gy = gy.bfloat16().float()
    xr = xq.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = O.conv2d_elr(xr, wr, br, s, pd, norm, act, wround=E.BF16.rf)
    (yr * gy).sum().backward()
    xc = xq.cuda().requires_grad_(True)
    y = m(xc)
    (y.float() * gy.cuda()).sum().backward()
    torch.cuda.synchronize()
    G.check_like(y.float().cpu(), yr.detach(), 2e-2, 8e-3, tag + " y")
    G.check_like(xc.grad.cpu(), xr.grad, 2e-2, 8e-3, tag + " dx")
    for name, got, ref in (("dw", m.weight.grad, wr.grad), ("db", m.bias.grad, br.grad)):
        rel = float((got.cpu().double() - ref.double()).norm() / ref.double().norm())
        assert rel < 2e-2, (tag, name, rel)
    # and against the reference class's own fp32 outputs (un-rounded input): bf16-level agreement
    y2 = m(x.cuda())
    G.check(g, f"{tag}/y", y2.float().contiguous(), 3e-2, 2e-2)
    # fuse(): normalisation and gain baked into the weight, same forward
    m.fuse()
    y3 = m(xq.cuda())
    torch.testing.assert_close(y3.float(), y.detach().float(), rtol=2e-2, atol=2e-2 * float(y.detach().float().abs().max()))


def test_conv2d_elr_refuses_what_is_out_of_scope(fv):
    C = fv.modules.Conv2dELR
    with pytest.raises(NotImplementedError):
        C(3, 32, 1, 1, 1)                       # the reference's first encoder layer: 1x1 kernel with padding 1
    with pytest.raises(NotImplementedError):
        C(16, 16, 4, 2, 1, wsize=8)             # style modulation
    with pytest.raises(NotImplementedError):
        C(16, 16, 3, 1, 1, act=nn.Tanh())


def test_flatten_vae6_module(fv):
    g = G.load("elr.npz")
    vae = fv.models.flatten_vae6()
    vae.load_state_dict(vae6_params())
    vae = vae.cuda()
    x = torch.from_numpy(detgen.det_uniform((3, 16, 4, 4), 91, -1.0, 1.0)).cuda().requires_grad_(True)
    eps = torch.from_numpy(detgen.det_normal((3, 256), 92)).cuda()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        mu, ls, xh = vae(x, eps)
        import face_vae_b200.losses as L
        gy = torch.from_numpy(detgen.det_uniform((3, 16, 4, 4), 93, -1.0, 1.0)).cuda()
        ((xh * gy).sum() + 3.0 * L.KLDivergenceLoss()((mu, ls))).backward()
        torch.cuda.synchronize()
        for k, v in (("mu", mu), ("logstd", ls), ("xhat", xh), ("dx", x.grad)):
            G.check(g, f"vae6/{k}", v, 1e-4, 1e-4)                    # fp32 path: rtol 1e-4
        for k, v in vae.named_parameters():
            G.check(g, f"vae6/grad/{k}", v.grad, 2e-3, atol_frac=2e-3)
        vae.training = False
        _, _, xh0 = vae(x.detach())
        G.check(g, "vae6/eval_xhat", xh0, 1e-4, 1e-4)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_bilinear_prescale_kernel(fv):
    g = G.load("elr.npz")
    for tag, shape in (("pre256", (2, 3, 256, 256)), ("pre100", (1, 3, 100, 72))):
        xi = torch.from_numpy(detgen.det_unit(shape, detgen.name_seed(tag))).cuda()
        y = fv.o.bilinear_resize(xi, 0.25)
        G.check(g, f"{tag}/y", y, 1e-5, 1e-6)
        ref = torch.nn.functional.interpolate(xi, mode="bilinear", scale_factor=0.25, align_corners=False, recompute_scale_factor=True)
        torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-6)


def test_efe_conv5_2d_stage(fv):
    """Pre-scale -> down -> flatten_vae_nl -> mid_conv -> view(N, C, D, h, w): models.py:764-787 against the oracle."""
    torch.manual_seed(0)
    m = fv.models.EFE_conv5().cuda().train()
    x = torch.rand((2, 3, 256, 256), device="cuda")
    eps = torch.randn((2, 256), device="cuda")
    x3d, x_c, x_a_c, (mu, ls), (x_vae, x_hat) = m.forward_2d(x, None, True, eps)
    assert x3d.shape == (2, 256, 16, 4, 4) and x_c is None and x_a_c is None
    assert mu.shape == (2, 256) and x_vae.shape == (2, 32, 4, 4) and x_hat.shape == (2, 16, 4, 4)
    # oracle: the same composition on the CPU
    p = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    xs = O.bilinear_prescale(x.cpu(), 0.25)
    h = O.same_block(xs, {k[len("down.0."):]: v for k, v in p.items() if k.startswith("down.0.")}, "")
    for i in range(1, 5):
        h = O.down_block(h, {k[len(f"down.{i}."):]: v for k, v in p.items() if k.startswith(f"down.{i}.")}, "")
    mu_r, ls_r, z_r = O.reparameterise(h, eps.cpu(), True, 16)
    y_r = O.conv2d(z_r, p["mid_conv.weight"], p["mid_conv.bias"], 1, 0).view(2, 256, 16, 4, 4)
    G.check_like(mu.cpu(), mu_r, 2e-2, 2e-2, "mu")
    G.check_like(x3d.float().cpu(), y_r, 2e-2, 2e-2, "x3d")
    x3d.float().sum().backward()
    assert all(q.grad is not None and bool(torch.isfinite(q.grad).all()) for q in m.parameters())
    with pytest.raises(NotImplementedError):
        m(x)
