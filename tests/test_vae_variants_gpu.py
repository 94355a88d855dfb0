"""The reference's other bottlenecks on the GPU, against outputs of the unmodified reference classes
(tests/golden/vae_variants.npz): ``flatten_vae`` (reference models.py:484-522 -- the formula of SURVEY.md 8a7 behind a LinearELR
encoder) and ``local_vae`` (models.py:442-482 -- DownBlock2D / UpBlock2D around two demodulated fully connected layers, the caller
listed in SURVEY.md 8b).  fp32 pieces rtol 1e-4; the bf16 block path per element within 4e-2 (two convolution blocks and two
fully connected layers in sequence), parameter gradients by relative L2."""
import pytest
import torch

from oracle import detgen
from tests import goldenlib as G
from tests.test_oracle_golden import FVAE_SHAPES, LVAE_SHAPES, variant_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fv():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import face_vae_b200.losses as L
    import face_vae_b200.models as MO
    from face_vae_b200 import _lib
    _lib.call("fv_device_ok")

    class NS:
        models, losses = MO, L
    return NS


def _load(module, params):
    sd = module.state_dict()
    for k, v in params.items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
        sd[k] = v.clone().float()
    module.load_state_dict(sd)
    return module.cuda()


def test_flatten_vae_module(fv):
    g = G.load("vae_variants.npz")
    vae = _load(fv.models.flatten_vae(), variant_params("fvae.", FVAE_SHAPES))
    assert sorted(vae.state_dict().keys()) == [str(k) for k in g["fvae.keys"]]
    x = torch.from_numpy(detgen.det_uniform((3, 16, 4, 4), 191, -1.0, 1.0)).cuda().requires_grad_(True)
    eps = torch.from_numpy(detgen.det_normal((3, 256), 192)).cuda()
    gy = torch.from_numpy(detgen.det_uniform((3, 16, 4, 4), 193, -1.0, 1.0)).cuda()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        mu, ls, xh = vae(x, True, eps)
        ((xh * gy).sum() + 3.0 * fv.losses.KLDivergenceLoss()((mu, ls))).backward()
        torch.cuda.synchronize()
        for k, v in (("mu", mu), ("logstd", ls), ("xhat", xh), ("dx", x.grad)):
            G.check(g, f"fvae/{k}", v, 1e-4, 1e-4)                    # fp32 path: rtol 1e-4
        for k, v in vae.named_parameters():
            G.check(g, f"fvae/grad/{k}", v.grad, 2e-3, atol_frac=2e-3)
        m0, l0, xh0 = vae(x.detach(), False)
        assert m0 is None and l0 is None
        G.check(g, "fvae/eval_xhat", xh0, 1e-4, 1e-4)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def _rel_l2(got, g, name):
    """Relative L2 against a golden entry stored in full."""
    import numpy as np
    ref = g[f"{name}/full"].astype(np.float64) if f"{name}/full" in g else None
    a = got.detach().double().cpu().numpy().flatten()
    if ref is None:
        ref = g[f"{name}/sample"].astype(np.float64)
        a = a[G.sample_index(a.size)]
    return float(np.sqrt(((a - ref) ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-30))


def test_local_vae_module(fv):
    g = G.load("vae_variants.npz")
    lv = _load(fv.models.local_vae(), variant_params("lvae.", LVAE_SHAPES)).train()
    assert sorted(lv.state_dict().keys()) == [str(k) for k in g["lvae.keys"]]
    x = torch.from_numpy(detgen.det_uniform((4, 128, 8, 8), 194, -1.0, 1.0)).cuda().requires_grad_(True)
    gy = torch.from_numpy(detgen.det_uniform((4, 128, 8, 8), 195, -1.0, 1.0)).cuda()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m0, l0, xh = lv(x)
        assert m0 is None and l0 is None and tuple(xh.shape) == (4, 128, 8, 8)
        (xh.float() * gy).sum().backward()
        torch.cuda.synchronize()
        G.check(g, "lvae/xhat", xh.float().contiguous(), 4e-2, 4e-2)
        assert _rel_l2(x.grad, g, "lvae/dx") < 0.10
        scale = max(float(g[f"lvae/grad/{k}/absmax"]) for k, _ in lv.named_parameters())
        for k, p in lv.named_parameters():
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
            if float(g[f"lvae/grad/{k}/absmax"]) < 1e-4 * scale or k.endswith("layers.0.bias"):
                # a conv bias in front of a batch norm: analytically zero, rounding noise on both sides
                assert float(p.grad.abs().max()) <= 5e-3 * scale, (k, float(p.grad.abs().max()))
                continue
            assert _rel_l2(p.grad, g, f"lvae/grad/{k}") < 0.10, (k, _rel_l2(p.grad, g, f"lvae/grad/{k}"))
        for k, b in lv.named_buffers():
            if k.endswith("running_mean") or k.endswith("running_var"):
                G.check(g, f"lvae/buf/{k}", b, 1e-2, 1e-2)
        lv.eval()
        with torch.no_grad():
            _, _, xh_e = lv(x.detach())
        G.check(g, "lvae/eval_xhat", xh_e.float().contiguous(), 4e-2, 4e-2)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
