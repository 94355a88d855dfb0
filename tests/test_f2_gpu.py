"""SURVEY.md 8f row 2 on the GPU: the block variants of the reference's Generator / Discriminator -- spectral norm
(use_weight_norm=True, reference modules.py:11,14), instance norm (modules.py:21), 3x3 stride 2 (models.py:1120-1123), the
un-normalised CN head -- against outputs of the unmodified reference blocks (tests/golden/f2.npz) and, tightly, against the oracle
with the filter rounded to bf16 where the CUDA path rounds it; plus the 2-D stages of Generator and Discriminator end to end."""
import numpy as np
import pytest
import torch

from oracle import emulate as E
from oracle import facevae_oracle as O
from tests import goldenlib as G
from tests.test_oracle_golden import F2_BLOCKS, f2_block_io, f2_block_params

pytestmark = pytest.mark.gpu

CTORS = {
    "sn_in_s2": lambda M: M.ConvBlock2D("CNA", 32, 64, 3, 2, 1, True, "instance", "leakyrelu"),
    "sn_in_s1": lambda M: M.ConvBlock2D("CNA", 64, 64, 3, 1, 1, True, "instance", "leakyrelu"),
    "sn_cn_none": lambda M: M.ConvBlock2D("CN", 64, 1, 3, 1, 1, True, activation_type="none"),
    "sn_bn_leaky": lambda M: M.ConvBlock2D("CNA", 32, 64, 3, 1, 1, True, nonlinearity_type="leakyrelu"),
    "sn_res": lambda M: M.ResBlock2D(32, True),
    "sn_up": lambda M: M.UpBlock2D(32, 16, True),
}


@pytest.fixture(scope="module")
def fv():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import face_vae_b200.models as MO
    import face_vae_b200.modules as M
    from face_vae_b200 import _lib
    _lib.call("fv_device_ok")

    class NS:
        modules, models = M, MO
    return NS


@pytest.mark.parametrize("tag", sorted(CTORS))
def test_block_variants_match_reference(fv, tag):
    g = G.load("f2.npz")
    p = f2_block_params(g, tag)
    blk = CTORS[tag](fv.modules)
    sd = blk.state_dict()
    for k, v in p.items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
        sd[k] = v.clone()
    blk.load_state_dict(sd)
    blk = blk.cuda().train()
    fn, ci = F2_BLOCKS[tag]
    x, gy = f2_block_io(g, tag, ci)
    xc = x.cuda().requires_grad_(True)
    y = blk(xc)
    (y.float() * gy.cuda()).sum().backward()
    torch.cuda.synchronize()
    # against the reference's fp32 outputs: bf16 path, per element 2e-2 (+ floor of the tensor's scale)
    G.check(g, f"{tag}/y", y.float().contiguous(), 2e-2, 2e-2)
    # power-iteration state: identical update rule
    for k, b in blk.named_buffers():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            G.check(g, f"{tag}/buf/{k}", b, 1e-4, 1e-5)
        elif "running" in k:
            G.check(g, f"{tag}/buf/{k}", b, 1e-2, 1e-2)
    wmax = max(float(g[f"{tag}/grad/{k}/absmax"]) for k, q in blk.named_parameters() if q.dim() == 4)
    for k, q in blk.named_parameters():
        ref_max = float(g[f"{tag}/grad/{k}/absmax"])
        if ref_max < 1e-3 * wmax:
            assert float(q.grad.abs().max()) <= 5e-2 * wmax, k
            continue
        got = q.grad.detach().double().cpu().numpy().flatten()
        ref = g[f"{tag}/grad/{k}/full"].astype(np.float64) if f"{tag}/grad/{k}/full" in g else None
        if ref is None:
            got, ref = got[G.sample_index(got.size)], g[f"{tag}/grad/{k}/sample"].astype(np.float64)
        rel = float(np.sqrt(((got - ref) ** 2).sum()) / np.sqrt((ref ** 2).sum()))
        assert rel <= 8e-2, (tag, k, rel)          # one bf16 block against fp32 (ReLU / LeakyReLU mask flips): the yardstick regime


def test_instance_norm_kernels_against_torch(fv):
    """fv_in_* against F.instance_norm + leaky_relu in fp32 on bf16-exact inputs: forward 1e-2, gradients relative L2 2e-2."""
    import torch.nn.functional as F
    from face_vae_b200 import functional as Fn
    from face_vae_b200.ops import ACT_LEAKY
    torch.manual_seed(0)
    n, c, h, w = 3, 128, 12, 20
    y = (torch.randn(n, c, h, w) * 2 + 0.5).bfloat16().float().cuda().requires_grad_(True)
    gamma = (torch.rand(c) + 0.5).cuda().requires_grad_(True)
    beta = (torch.rand(c) - 0.5).cuda().requires_grad_(True)
    ref = F.leaky_relu(F.instance_norm(y, weight=gamma, bias=beta, eps=1e-5), 0.2)
    gy = torch.randn_like(ref).bfloat16().float()
    (ref * gy).sum().backward()
    yn = y.detach().permute(0, 2, 3, 1).contiguous().bfloat16().requires_grad_(True)
    g2, b2 = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    out = Fn.InstanceNormAct.apply(yn, g2, b2, ACT_LEAKY, 1e-5)
    (out.float() * gy.permute(0, 2, 3, 1)).sum().backward()
    torch.cuda.synchronize()
    G.check_like(out.float().permute(0, 3, 1, 2), ref.detach(), 1e-2, 4e-3, "in fwd")
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    assert rel(yn.grad.float().permute(0, 3, 1, 2), y.grad) < 2e-2
    assert rel(g2.grad, gamma.grad) < 2e-2 and rel(b2.grad, beta.grad) < 2e-2


def test_generator_and_discriminator_2d_stages(fv):
    """Generator.forward_2d / Discriminator.forward_features (reference models.py:1100-1110, 1129-1139) against the oracle's
    composition of the same blocks with the same state_dict."""
    torch.manual_seed(0)
    gen = fv.models.Generator(n_res=1, up_seq=[64, 32], D=2, C=16).cuda().train()
    fs = torch.rand((2, 32, 16, 16), device="cuda") * 2 - 1
    occ = torch.rand((2, 1, 16, 16), device="cuda")
    p = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}       # before the forward updates u / v
    out = gen.forward_2d(fs, occ)
    assert out.shape == (2, 3, 32, 32)
    t = O.conv_block("CNA", fs.cpu(), p, "in_conv.", 3, 1, 1, nonlinearity="leakyrelu")
    t = O.conv2d(t, p["mid_conv.weight"], p["mid_conv.bias"], 1, 0) * occ.cpu()
    t = O.res_block(t, p, "res.0.")
    t = O.up_block(t, p, "up.0.")
    ref = torch.sigmoid(O.conv2d(t, p["out_conv.weight"], p["out_conv.bias"], 1, 3))
    G.check_like(out.cpu(), ref, 2e-2, 2e-2, "generator x_hat")
    out.sum().backward()
    assert all(q.grad is not None and bool(torch.isfinite(q.grad).all()) for q in gen.parameters())
    with pytest.raises(NotImplementedError):
        gen(fs, None, occ)

    disc = fv.models.Discriminator(down_seq=[64, 128, 256], K=13).cuda().train()
    x = torch.rand((2, 16, 32, 32), device="cuda") * 2 - 1
    pd = {k: v.detach().cpu().clone() for k, v in disc.state_dict().items()}
    o, feats = disc.forward_features(x)
    assert o.shape == (2, 1, 8, 8) and [tuple(f.shape[1:]) for f in feats] == [(64, 16, 16), (128, 8, 8), (256, 8, 8)]
    t = x.cpu()
    t = O.conv_block("CNA", t, pd, "layers.0.", 3, 2, 1, nonlinearity="leakyrelu")
    t = O.conv_block("CNA", t, pd, "layers.1.", 3, 2, 1, nonlinearity="leakyrelu")
    t = O.conv_block("CNA", t, pd, "layers.2.", 3, 1, 1, nonlinearity="leakyrelu")
    G.check_like(feats[2].float().cpu(), t, 3e-2, 3e-2, "discriminator features")
    t = O.conv_block("CN", t, pd, "layers.3.", 3, 1, 1)
    G.check_like(o.float().cpu(), t, 3e-2, 3e-2, "discriminator output")
    o.float().sum().backward()
    assert all(q.grad is not None and bool(torch.isfinite(q.grad).all()) for q in disc.parameters())
