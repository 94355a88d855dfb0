"""Parity of the CUDA path (through the nn.Module facade and the C ABI) with the reference:
 * against the committed golden fixtures produced by the unmodified reference classes (tests/golden/*.npz),
 * against the CPU oracle on other seeded inputs / sizes,
 * through size-independent properties at BASELINE.json's full size (batch 32 at 256x256).
Tolerances (north_star): bf16 path rtol 2e-2 (with an absolute floor of a fraction of the tensor's max);
KL term / fp32 kernels rtol 1e-4 (tested at kernel level in test_kernels_gpu.py)."""
import numpy as np
import pytest
import torch

from oracle import detgen
from oracle import facevae_oracle as O
from tests import goldenlib as G
from tests.test_oracle_golden import block_io, block_params

pytestmark = pytest.mark.gpu

RTOL = 2e-2          # bf16 path (north_star)
AFRAC = 2e-2         # absolute floor as a fraction of max |ref| of the tensor


@pytest.fixture(scope="module")
def fv():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import face_vae_b200.modules as M
    import face_vae_b200.models as MO
    import face_vae_b200.losses as L
    import face_vae_b200.trainer as T
    from face_vae_b200 import _lib
    _lib.call("fv_device_ok")

    class NS:
        modules, models, losses, trainer = M, MO, L, T
    return NS


def _load(module, params):
    sd = module.state_dict()
    for k, v in params.items():
        assert k in sd, k
        sd[k] = v.clone()
    module.load_state_dict(sd)
    return module.cuda().train()


BLOCK_CTORS = {
    "down": lambda M: M.DownBlock2D(16, 32, False),
    "up": lambda M: M.UpBlock2D(32, 16, False),
    "same": lambda M: M.SameBlock2D(16, 32, False),
    "same3": lambda M: M.SameBlock2D(3, 32, False),
    "res": lambda M: M.ResBlock2D(32, False),
    "convblock_leaky": lambda M: M.ConvBlock2D("CNA", 16, 16, 3, 1, 1, False, nonlinearity_type="leakyrelu"),
}
BLOCK_CI = {"down": 16, "up": 32, "same": 16, "same3": 3, "res": 32, "convblock_leaky": 16}


@pytest.mark.parametrize("tag", sorted(BLOCK_CTORS))
def test_block_matches_reference_golden(fv, tag):
    g = G.load("blocks.npz")
    blk = _load(BLOCK_CTORS[tag](fv.modules), block_params(g, tag))
    x, gy = block_io(g, tag, BLOCK_CI[tag])
    x = x.cuda().requires_grad_(True)
    y = blk(x)
    assert y.shape == tuple(g[f"{tag}/y/shape"]) or tuple(y.shape) == tuple(int(v) for v in g[f"{tag}/y/shape"])
    (y.float() * gy.cuda()).sum().backward()
    torch.cuda.synchronize()
    G.check(g, f"{tag}/y", y.float().contiguous(), RTOL, AFRAC)            # forward: per element
    scale = max(float(g[f"{tag}/grad/{k}/absmax"]) for k, _ in blk.named_parameters())
    _check_grad(g, f"{tag}/dx", x.grad, slack=2.0)
    for k, p in blk.named_parameters():
        _check_grad(g, f"{tag}/grad/{k}", p.grad, slack=2.0, zero_tol=2e-3 * scale)
    for k, b in blk.named_buffers():
        if k.endswith("running_mean") or k.endswith("running_var"):
            G.check(g, f"{tag}/buf/{k}", b, 1e-2, 1e-2)


def _check_grad(g, name, t, slack=1.5, floor=1e-2, zero_tol=1e-3):
    """Gradients: relative L2 against the fp32 golden no worse than slack x the reference's own bf16-autocast deviation
    (+ floor); analytically-zero gradients (bias of a conv feeding a batch norm: fp32 noise in the golden, whose "yardstick"
    is meaningless) must stay ~0."""
    if float(g[f"{name}/absmax"]) < 1e-5 or (name.endswith("layers.0.bias") and (".enc." in "." + name or ".up." in "." + name)):
        assert float(t.detach().abs().max()) <= zero_tol, (name, float(t.detach().abs().max()))   # bf16 rounding noise only
        return 0.0, 0.0
    return G.check_vs_yardstick(g, name, t, slack, floor)


def _anchor(fv, cfg, base):
    m = fv.models.FaceVAE(cfg.down_seq, cfg.up_seq, cfg.n_res)
    return _load(m, O.det_anchor_params(cfg, base))


@pytest.mark.parametrize("fixture,n,hw,base,cfg", [
    ("anchor_n4_64.npz", 4, 64, 0, O.CFG_256),              # BASELINE.json configs[0]: batch 4 at 64x64
    ("anchor_n2_64_b1.npz", 2, 64, 1, O.CFG_256),
    ("anchor_n32_256.npz", 32, 256, 7, O.CFG_256),          # configs[1]: batch 32 at 256x256 (ring / folded / persistent schedules)
    ("anchor512_n2_512.npz", 2, 512, 2, O.CFG_512),         # configs[3] architecture at 512x512
])
def test_anchor_matches_reference_golden(fv, fixture, n, hw, base, cfg):
    """End to end against the unmodified reference classes (fp32): forward activations, losses per element within 2e-2;
    every parameter gradient within 1.25 x the reference's OWN bf16-autocast deviation from its fp32 self (+ 5e-3) -- the
    tight per-layer gradient check is tests/test_layerwise_gpu.py; running statistics."""
    g = G.load(fixture)
    m = _anchor(fv, cfg, base)
    x, eps = O.det_inputs(n, hw, hw, cfg, base)
    x, eps = x.cuda(), eps.cuda()
    out = m.forward_loss(x, eps)
    loss = cfg.w_kl * out["K"] + cfg.w_rec * out["R"]
    loss.backward()
    torch.cuda.synchronize()
    for k in ("K", "R"):
        ref = float(g[f"out/{k}"])
        assert abs(out[k].item() - ref) <= RTOL * abs(ref), (k, out[k].item(), ref)
    assert abs(loss.item() - float(g["out/loss"])) <= RTOL * abs(float(g["out/loss"]))
    G.check(g, "out/mu", out["mu"], RTOL, AFRAC)
    G.check(g, "out/logstd", out["logstd"], RTOL, AFRAC)
    G.check(g, "out/x_hat", out["x_hat"], RTOL, AFRAC)
    worst = {}
    scale = max(float(g[f"grad/{k}/absmax"]) for k, _ in m.named_parameters())
    for k, p in m.named_parameters():
        worst[k] = _check_grad(g, f"grad/{k}", p.grad, slack=1.25, floor=5e-3, zero_tol=2e-3 * scale)
    for k, b in m.named_buffers():
        if k.endswith("running_mean") or k.endswith("running_var"):
            G.check(g, f"buf/{k}", b, 1e-2, 1e-2)
    top = sorted(worst.items(), key=lambda kv: -kv[1][0])[:3]
    print(f"{fixture}: largest grad rel-L2 (ours, reference-autocast yardstick):", [(k, f"{a:.3f}", f"{b:.3f}") for k, (a, b) in top])


def test_anchor_modular_forward_matches_fused(fv):
    """model(x, train_vae, eps) (separate Reparam / KLDivergenceLoss / ReconLoss modules, the reference's call
    pattern trainer.py:312,314) gives the same numbers as the fused forward_loss path."""
    cfg = O.CFG_256
    m = _anchor(fv, cfg, 0)
    x, eps = O.det_inputs(2, 64, 64, cfg, 3)
    x, eps = x.cuda(), eps.cuda()
    out = m.forward_loss(x, eps)
    (cfg.w_kl * out["K"] + cfg.w_rec * out["R"]).backward()
    g_fused = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    for b in m.buffers():            # same starting running stats do not matter for training-mode outputs
        pass
    mu, logstd, x_hat = m(x, True, eps)
    K = fv.losses.KLDivergenceLoss()((mu, logstd))
    R = fv.losses.ReconLoss()((x, x_hat))
    (cfg.w_kl * K + cfg.w_rec * R).backward()
    torch.cuda.synchronize()
    # two runs differ through fp32 atomics order amplified by bf16 rounding / ReLU-mask flips: bf16-level agreement
    assert abs(K.item() - out["K"].item()) <= 5e-3 * abs(K.item())
    assert abs(R.item() - out["R"].item()) <= 5e-3 * abs(R.item())
    torch.testing.assert_close(x_hat, out["x_hat"], rtol=2e-2, atol=2e-2)
    for k, p in m.named_parameters():
        ref = g_fused[k]
        if ref.abs().max().item() < 1e-5:
            continue
        rel = ((p.grad - ref).norm() / ref.norm()).item()
        assert rel <= 0.35, (k, rel)
    # eval / train_vae False: z == mu exactly, (None, None, x_hat)
    m.eval()
    mu0, ls0, xh0 = m(x, False)
    assert mu0 is None and ls0 is None and xh0.shape == x.shape


def test_anchor_matches_oracle_other_size(fv):
    """Same comparison against the CPU oracle at 128x128, batch 2 (oracle finishes in seconds)."""
    cfg = O.CFG_256
    p = O.det_anchor_params(cfg, 5)
    x, eps = O.det_inputs(2, 128, 128, cfg, 5)
    ref_out, ref_grads, ref_bufs, _ = O.anchor_train_grads(p, x, eps, cfg)
    m = _load(fv.models.FaceVAE(), p)
    out = m.forward_loss(x.cuda(), eps.cuda())
    (cfg.w_kl * out["K"] + cfg.w_rec * out["R"]).backward()
    torch.cuda.synchronize()
    assert abs(out["K"].item() - ref_out["K"].item()) <= RTOL * abs(ref_out["K"].item())
    assert abs(out["R"].item() - ref_out["R"].item()) <= RTOL * abs(ref_out["R"].item())
    G.check_like(out["mu"], ref_out["mu"], RTOL, AFRAC, "mu")
    G.check_like(out["x_hat"], ref_out["x_hat"], RTOL, AFRAC, "x_hat")
    for k, pr in m.named_parameters():
        ref = ref_grads[k]
        if ref.abs().max().item() < 1e-5:
            assert pr.grad.abs().max().item() <= 1e-3
            continue
        rel = ((pr.grad.cpu() - ref).norm() / ref.norm()).item()
        # yardstick measured for this model family: the reference's own bf16 autocast sits at 0.05-0.25 here
        assert rel <= 0.30, (k, rel)


def test_train_step_reduces_loss(fv):
    torch.manual_seed(0)
    m = fv.models.FaceVAE().cuda().train()
    tr = fv.trainer.VAETrainer(m, lr=2e-3)
    x, eps = O.det_inputs(4, 64, 64, O.CFG_256, 9)
    x, eps = x.cuda(), eps.cuda()
    vals = []
    for _ in range(8):
        losses, gen = tr.step(x, eps)
        vals.append(sum(v.item() for v in losses.values()))
    assert all(np.isfinite(vals)), vals
    assert vals[-1] < vals[0], vals
    assert gen.shape == x.shape and float(gen.min()) >= 0 and float(gen.max()) <= 1


def test_full_size_properties(fv):
    """Batch 32 at 256x256 (BASELINE.json configs[1]): properties that hold at any size.
    (1) training-mode BN output has per-channel mean beta and variance gamma^2 before the ReLU -> checked through the
        gradient identities sum(dy) = 0 and sum(dy * xhat) = 0 of the conv-output gradient;
    (2) pooling / up-sampling shapes; (3) finite loss and gradients; (4) eval-mode determinism."""
    from face_vae_b200 import ops
    from face_vae_b200.ops import ACT_RELU, MODE_POOL
    torch.manual_seed(1)
    n, h, w, c = 32, 256, 256, 64
    y = torch.randn((n, h, w, c), device="cuda").bfloat16()
    gamma = (torch.rand(c, device="cuda") + 0.5)
    beta = torch.rand(c, device="cuda") - 0.5
    stat = ops.bn_finalize(ops.bn_stats(y), n * h * w, gamma, beta, None, None)
    a = ops.bn_act_fwd(y, stat, MODE_POOL, ACT_RELU)
    assert a.shape == (n, h // 2, w // 2, c)
    g = torch.randn_like(a)
    s = ops.bn_act_bwd_reduce(y, g, stat, MODE_POOL, ACT_RELU)
    _, _, coef = ops.bn_bwd_finalize(s, s, n * h * w, c)
    dy = ops.bn_act_bwd_apply(y, g, stat, coef, MODE_POOL, ACT_RELU)
    cs = ops.colsum(dy)                                     # sum over 2M pixels of dy ~ 0 relative to sum |dy|
    denom = dy.float().abs().sum(dim=(0, 1, 2))
    assert float((cs.abs() / denom).max()) < 2e-3
    xhat = (y.float() - stat[0]) * stat[1]
    ortho = (dy.float() * xhat).sum(dim=(0, 1, 2)).abs() / (dy.float().abs() * xhat.abs()).sum(dim=(0, 1, 2))
    assert float(ortho.max()) < 2e-3
    # whole model at full size: finite, right shapes, deterministic in eval mode
    m = fv.models.FaceVAE().cuda().train()
    x = torch.rand((32, 3, 256, 256), device="cuda")
    out = m.forward_loss(x)
    (0.2 * out["K"] + 10 * out["R"]).backward()
    assert out["x_hat"].shape == x.shape and out["mu"].shape == (32, 4096)
    assert np.isfinite(out["K"].item()) and np.isfinite(out["R"].item())
    for k, p in m.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
    m.eval()
    with torch.no_grad():
        a1 = m(x[:4], False)[2]
        a2 = m(x[:4], False)[2]
    assert torch.equal(a1, a2)


def test_cuda_graph_step_matches_eager(fv):
    """The captured train step (VAETrainer(use_cuda_graph=True)) replays the eager sequence: same losses / weights
    up to the bf16 run-to-run noise of the path, and the capture warm-up leaves the training state untouched."""
    cfg = O.CFG_256
    p = O.det_anchor_params(cfg, 0)
    x, eps = O.det_inputs(4, 64, 64, cfg, 11)
    x, eps = x.cuda(), eps.cuda()
    res = {}
    for mode in (False, True):
        m = _load(fv.models.FaceVAE(), p)
        tr = fv.trainer.VAETrainer(m, lr=1e-3, use_cuda_graph=mode)
        vals = []
        for _ in range(4):
            losses, gen = tr.step(x, eps)
            vals.append(sum(v.item() for v in losses.values()))
        res[mode] = (vals, {k: v.detach().clone() for k, v in m.named_parameters()}, tr)
    assert res[True][2].use_cuda_graph and res[True][2].launches_per_step > 100
    v0, v1 = res[False][0], res[True][0]
    assert abs(v0[0] - v1[0]) <= 5e-3 * abs(v0[0]), (v0, v1)        # first step: identical weights (warm-up was rolled back)
    for a, b in zip(v0, v1):
        assert abs(a - b) <= 2e-2 * abs(a), (v0, v1)
    assert v1[-1] < v1[0]
    w0, w1 = res[False][1]["out_conv.weight"], res[True][1]["out_conv.weight"]
    assert ((w0 - w1).norm() / w0.norm()).item() < 5e-2


def test_deep_512_variant_matches_oracle(fv):
    """BASELINE.json configs[3] architecture (down_seq (3,32,64,128,256,512,64), up to 512 channels, latent 32 channels)
    at a size the CPU oracle finishes in seconds (batch 2, 128x128): losses per element-level tolerance, gradients bounded
    by the bf16 yardstick of this model family."""
    cfg = O.CFG_512
    p = O.det_anchor_params(cfg, 2)
    x, eps = O.det_inputs(2, 128, 128, cfg, 2)
    ref_out, ref_grads, _, _ = O.anchor_train_grads(p, x, eps, cfg)
    m = _load(fv.models.face_vae_512(), p)
    assert sum(q.numel() for q in m.parameters()) == 15258819
    out = m.forward_loss(x.cuda(), eps.cuda())
    (cfg.w_kl * out["K"] + cfg.w_rec * out["R"]).backward()
    torch.cuda.synchronize()
    assert abs(out["K"].item() - ref_out["K"].item()) <= RTOL * abs(ref_out["K"].item())
    assert abs(out["R"].item() - ref_out["R"].item()) <= RTOL * abs(ref_out["R"].item())
    G.check_like(out["x_hat"], ref_out["x_hat"], RTOL, AFRAC, "x_hat")
    for k, pr in m.named_parameters():
        ref = ref_grads[k]
        assert bool(torch.isfinite(pr.grad).all()), k
        if ref.abs().max().item() < 1e-5:
            continue
        rel = ((pr.grad.cpu() - ref).norm() / ref.norm()).item()
        assert rel <= 0.35, (k, rel)


def test_inference_sweep_shapes(fv):
    """BASELINE.json configs[4]: eval-mode encode -> sample -> decode at several batch sizes (running statistics)."""
    m = fv.models.FaceVAE().cuda().eval()
    with torch.no_grad():
        for n in (1, 3, 16):
            x = torch.rand((n, 3, 256, 256), device="cuda")
            mu, ls, xh = m(x, True, torch.zeros((n, 4096), device="cuda"))
            assert xh.shape == x.shape and mu.shape == (n, 4096) and bool(torch.isfinite(xh).all())
            _, _, xh2 = m(x, False)
            torch.testing.assert_close(xh, xh2, rtol=0, atol=0)        # eps = 0  <=>  z = mu


def test_loss_modules_match_reference_golden(fv):
    """The nn.Module classes themselves (KLDivergenceLoss, ReconLoss, flatten_vae_nl: reference losses.py:385-403,
    models.py:525-570) against tests/golden/losses.npz -- outputs of the unmodified reference classes -- at rtol 1e-4."""
    g = G.load("losses.npz")
    L, MO = fv.losses, fv.models
    kl, rec = L.KLDivergenceLoss(), L.ReconLoss()
    for tag, (mv, sv) in {"zero": (0.0, 0.0), "mu1": (1.0, 0.0), "ls1": (0.0, 1.0), "lsm1": (0.0, -1.0)}.items():
        got = kl((torch.full((4, 256), mv, device="cuda"), torch.full((4, 256), sv, device="cuda"))).item()
        assert abs(got - float(g[f"kl/{tag}"])) <= 1e-4 * abs(float(g[f"kl/{tag}"])) + 1e-7, (tag, got)
    i = torch.arange(1024, dtype=torch.float32, device="cuda").view(4, 256)
    mu = torch.sin(0.01 * i).requires_grad_(True)
    ls = (0.5 * torch.cos(0.013 * i)).requires_grad_(True)
    v = kl((mu, ls))
    v.backward()
    assert abs(v.item() - float(g["kl/sincos"])) <= 1e-4 * float(g["kl/sincos"])
    np.testing.assert_allclose(mu.grad.cpu().numpy(), g["kl/sincos_dmu"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(ls.grad.cpu().numpy(), g["kl/sincos_dlogstd"], rtol=1e-4, atol=1e-8)
    a = torch.sin(0.1 * torch.arange(384, dtype=torch.float32, device="cuda")).view(2, 3, 8, 8).requires_grad_(True)
    b = torch.cos(0.07 * torch.arange(384, dtype=torch.float32, device="cuda")).view(2, 3, 8, 8)
    r = rec((a, b))
    r.backward()
    assert abs(r.item() - float(g["rec/mse"])) <= 1e-4 * float(g["rec/mse"])
    np.testing.assert_allclose(a.grad.cpu().numpy(), g["rec/mse_da"], rtol=1e-4, atol=1e-9)
    assert abs(L.ReconLoss(l1=True)((a.detach(), b)).item() - float(g["rec/l1"])) <= 1e-4 * float(g["rec/l1"])
    np.testing.assert_allclose(L.l1(a.detach(), b).cpu().numpy(), g["rec/l1_elem"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(L.l2(a.detach(), b).cpu().numpy(), g["rec/l2_elem"], rtol=1e-6, atol=1e-7)
    vae = MO.flatten_vae_nl()
    x = torch.from_numpy(detgen.det_uniform((3, 32, 4, 4), 77, 0.0, 2.0)).cuda()
    eps = torch.from_numpy(detgen.det_normal((3, 256), 78)).cuda()
    m0, s0, xh0 = vae(x, False)
    assert m0 is None and s0 is None and np.array_equal(xh0.cpu().numpy(), g["vae/eval_xhat"])      # eval: z == mu exactly
    m1, s1, xh1 = vae(x, True, eps)
    np.testing.assert_allclose(m1.cpu().numpy(), g["vae/train_mu"], rtol=0, atol=0)
    np.testing.assert_allclose(s1.cpu().numpy(), g["vae/train_logstd"], rtol=0, atol=0)
    np.testing.assert_allclose(xh1.cpu().numpy(), g["vae/train_xhat"], rtol=1e-5, atol=1e-5)


def test_generator_full_drop_in(fv):
    """GeneratorFull (reference trainer.py:214-317 call contract) on the VAE path: same numbers as the fused forward_loss,
    ten loss keys with the out-of-scope entries zero, 8-tuple, zero K / R when train_vae is falsy; the reference Logger's
    step skeleton (logger.py:150-164) runs on it."""
    cfg = O.CFG_256
    m = _anchor(fv, cfg, 0)
    x, eps = O.det_inputs(2, 64, 64, cfg, 3)
    x, eps = x.cuda(), eps.cuda()
    out = m.forward_loss(x, eps)
    g_full = fv.trainer.GeneratorFull(generator=m)
    g_full.eps = eps
    res = g_full(x, x, None, None, True)
    assert len(res) == 8 and all(r is None for r in res[2:])
    losses, generated = res[0], res[1]
    assert list(losses) == ["P", "G", "F", "E", "L", "H", "D", "C", "K", "R"]
    assert all(float(losses[k]) == 0.0 for k in "PGFELHDC")
    assert torch.equal(generated, out["x_hat"])                                  # deterministic kernels: same bits
    assert float(losses["K"]) == 0.2 * float(out["K"]) or abs(float(losses["K"]) - 0.2 * float(out["K"])) < 1e-7
    assert abs(float(losses["R"]) - 10 * float(out["R"])) < 1e-6
    opt = torch.optim.Adam(m.parameters(), lr=5e-5, betas=(0.5, 0.999))            # logger.py:60
    opt.zero_grad()
    sum(losses.values()).backward()                                              # logger.py:160-161
    opt.step()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
    off = g_full(x, x, None, None, False)
    assert float(off[0]["K"]) == 0.0 and float(off[0]["R"]) == 0.0 and off[1].shape == x.shape
