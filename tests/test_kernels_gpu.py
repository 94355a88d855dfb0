"""Kernel-level parity on a real B200: every C-ABI entry point against a plain PyTorch fp32 reference of the same
op on identical (bf16-representable) inputs.  bf16 outputs: rtol 2e-2 (north_star); fp32 reductions: rtol 1e-4."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from face_vae_b200 import ops as _ops
    from face_vae_b200 import _lib
    _lib.call("fv_device_ok")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _ops


def _rand(shape, seed, lo=-1.0, hi=1.0, bf16_exact=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = (torch.rand(shape, generator=g) * (hi - lo) + lo)
    if bf16_exact:
        t = t.bfloat16().float()
    return t.cuda()


def _report(name, got, ref, rtol, atol_frac):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = rtol * ref.abs() + atol_frac * ref.abs().max()
    bad = err > tol
    msg = f"{name}: max_err {err.max().item():.4e} ref_absmax {ref.abs().max().item():.4e} bad {bad.float().mean().item():.4f}"
    if bad.any():
        idx = bad.nonzero()[:6]
        for i in idx:
            t = tuple(i.tolist())
            msg += f"\n   at {t}: got {got[t].item():.5f} ref {ref[t].item():.5f}"
        if got.dim() == 4:   # NHWC error maps
            msg += "\n   bad frac by c//16: " + str([round(v, 3) for v in bad.float().mean(dim=(0, 1, 2)).view(-1, min(16, bad.shape[3])).mean(1).tolist()])
            msg += "\n   bad frac by w%8 : " + str([round(bad[:, :, i::8].float().mean().item(), 3) for i in range(min(8, bad.shape[2]))])
            msg += "\n   bad frac by h   : " + str([round(v, 3) for v in bad.float().mean(dim=(0, 2, 3)).tolist()][:16])
            msg += "\n   bad frac by n   : " + str([round(v, 3) for v in bad.float().mean(dim=(1, 2, 3)).tolist()][:16])
    print(msg)
    assert not bad.any(), msg


def _conv_case(ops, n, h, w, ci, co, k, bias=True, residual=False, out_mode=0, seed=0):
    from face_vae_b200.ops import pad_channels, OUT_NCHW_F32
    x = _rand((n, ci, h, w), seed)
    wt = _rand((co, ci, k, k), seed + 1, -1.0 / math.sqrt(ci * k * k), 1.0 / math.sqrt(ci * k * k), bf16_exact=True)
    b = _rand((co,), seed + 2, bf16_exact=False) if bias else None
    ref = F.conv2d(x, wt, b, padding=(k - 1) // 2)
    x_nhwc = ops.nchw_to_nhwc(x)
    assert x_nhwc.shape[-1] == pad_channels(ci)
    wf, _ = ops.weight_prep(wt, True, False)
    res = None
    if residual:
        r = _rand((n, co, h, w), seed + 3)
        res = ops.nchw_to_nhwc(r, pad_channels(co))
        ref = ref + r
    y = ops.conv2d(x_nhwc, wf, b, co, k, residual=res, out_mode=out_mode)
    torch.cuda.synchronize()
    if out_mode == OUT_NCHW_F32:
        got = y.permute(0, 2, 3, 1)
    else:
        got = y[..., :co]
        if y.shape[-1] > co:
            assert float(y[..., co:].float().abs().max()) == 0.0, "padded output channels must be zero"
    _report(f"conv n{n} {h}x{w} ci{ci} co{co} k{k} res{int(residual)} out{out_mode}", got, ref.permute(0, 2, 3, 1),
            2e-2 if out_mode == 0 else 1e-4, 4e-3 if out_mode == 0 else 1e-5)


@pytest.mark.parametrize("ci,co", [(64, 64), (16, 16), (32, 64), (128, 128), (256, 32), (64, 256)])
def test_conv_1x1(ops, ci, co):
    _conv_case(ops, 2, 16, 64, ci, co, 1)


@pytest.mark.parametrize("h,w,n", [(8, 128, 2), (16, 64, 2), (16, 16, 4), (8, 8, 3), (4, 4, 5), (4, 256, 1), (32, 32, 2)])
def test_conv_3x3_tilings(ops, h, w, n):
    _conv_case(ops, n, h, w, 64, 64, 3)


@pytest.mark.parametrize("ci,co", [(32, 64), (64, 128), (128, 256), (256, 32), (256, 256), (64, 32), (16, 256)])
def test_conv_3x3_channels(ops, ci, co):
    _conv_case(ops, 2, 16, 32, ci, co, 3)


def test_conv_7x7_rgb_out_nchw(ops):
    from face_vae_b200.ops import OUT_NCHW_F32
    _conv_case(ops, 2, 32, 32, 32, 3, 7, out_mode=OUT_NCHW_F32)


def test_conv_rgb_in_1x1(ops):
    _conv_case(ops, 2, 16, 64, 3, 32, 1)


def test_conv_residual_and_f32_out(ops):
    from face_vae_b200.ops import OUT_NHWC_F32
    _conv_case(ops, 2, 16, 16, 256, 256, 3, residual=True)
    _conv_case(ops, 2, 16, 16, 64, 32, 3, out_mode=OUT_NHWC_F32)


@pytest.mark.parametrize("n,h,w,ci,co,res", [(8, 16, 16, 256, 256, True), (2, 16, 16, 256, 256, False), (4, 16, 32, 128, 128, False),
                                             (3, 8, 8, 256, 192, True)])
def test_conv_output_channel_split(ops, n, h, w, ci, co, res):
    """Layers with fewer pixel tiles than half the SMs split the output channels over 2-4 CTAs per tile."""
    _conv_case(ops, n, h, w, ci, co, 3, residual=res, seed=60)


@pytest.mark.parametrize("n,h,w,ci,co,k", [
    (3, 37, 256, 32, 64, 3),      # ring kernel (staged-store epilogue), per-lane register accumulators
    (2, 20, 128, 64, 32, 3),      # ring kernel, 32 output channels
    (2, 16, 128, 64, 128, 3),     # implicit GEMM, slab schedule
    (4, 16, 16, 256, 256, 3),     # output channels split over CTAs
    (5, 4, 4, 128, 64, 3),        # tiles spanning several images, masked tail rows
    (2, 8, 8, 64, 512, 1),        # more than 256 output channels: two launches into one statistic block
    (2, 16, 32, 3, 32, 1),        # padded input channels
])
def test_conv_fused_bn_statistics(ops, n, h, w, ci, co, k):
    """fv_conv2d_stats: the epilogue's per-channel sum / sum of squares against fv_bn_stats on the stored output (fp32, rtol 1e-4)."""
    from face_vae_b200.ops import pad_channels
    x = _rand((n, ci, h, w), 70)
    wt = _rand((co, ci, k, k), 71, -1.0 / math.sqrt(ci * k * k), 1.0 / math.sqrt(ci * k * k))
    b = _rand((co,), 72, bf16_exact=False)
    wf, _ = ops.weight_prep(wt, True, False)
    xn = ops.nchw_to_nhwc(x)
    y, sums = ops.conv2d(xn, wf, b, co, k, want_stats=True)
    y_ref = ops.conv2d(xn, wf, b, co, k)
    ref = ops.bn_stats(y_ref)
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref)
    cp = pad_channels(co)
    yf = y.float().reshape(-1, cp)
    exact = torch.cat([yf.sum(0), (yf * yf).sum(0)])
    _report(f"fused stats n{n} {h}x{w} ci{ci} co{co}", sums.reshape(1, 1, 1, -1), exact.reshape(1, 1, 1, -1), 1e-4, 1e-5)
    _report("bn_stats kernel", ref.reshape(1, 1, 1, -1), exact.reshape(1, 1, 1, -1), 1e-4, 1e-5)


def test_conv_fused_statistics_with_residual(ops):
    """The second conv of a ResBlock2D emits the sums of (conv + bias + residual) as stored -- what the next block's norm reads."""
    n, h, w, c = 32, 16, 16, 256
    x = _rand((n, c, h, w), 75)
    r = _rand((n, c, h, w), 76)
    wt = _rand((c, c, 3, 3), 77, -1.0 / math.sqrt(c * 9), 1.0 / math.sqrt(c * 9))
    b = _rand((c,), 78, bf16_exact=False)
    wf, _ = ops.weight_prep(wt, True, False)
    xn, rn = ops.nchw_to_nhwc(x), ops.nchw_to_nhwc(r)
    y, sums = ops.conv2d(xn, wf, b, c, 3, rn, want_stats=True)
    y_ref = ops.conv2d(xn, wf, b, c, 3, rn)
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref)
    yf = y.float().reshape(-1, c)
    exact = torch.cat([yf.sum(0), (yf * yf).sum(0)])
    _report("fused stats with residual", sums.reshape(1, 1, 1, -1), exact.reshape(1, 1, 1, -1), 1e-4, 1e-5)


def test_res_block_chain_passes_statistics(ops):
    """ResBlock2D chain with the producer-emitted statistics (ops.attach_stats) against the same chain with the hints removed."""
    from face_vae_b200.modules import ResBlock2D, chain_res_blocks
    torch.manual_seed(5)
    blocks = torch.nn.Sequential(ResBlock2D(64, False), ResBlock2D(64, False)).cuda().train()
    x = ops.nchw_to_nhwc(_rand((4, 64, 16, 16), 79))
    ref = blocks[1].forward_nhwc(blocks[0].forward_nhwc(x)).float()
    chain_res_blocks(blocks)
    assert blocks[0].layers[0].emit_stats and blocks[0].layers[1].emit_stats and not getattr(blocks[1].layers[1], "emit_stats", False)
    h = blocks[0].forward_nhwc(x)
    assert ops.attached_stats(h) is not None
    got = blocks[1].forward_nhwc(h).float()
    torch.cuda.synchronize()
    assert torch.allclose(got, ref, rtol=2e-2, atol=2e-2 * ref.abs().max().item())
    h.add_(0)                                   # an in-place change invalidates the hint
    assert ops.attached_stats(h) is None


def test_conv_wide_row_tiles_with_256_output_channels(ops):
    # 512-deep variant at 512x512: 128 -> 256 channels on 128-pixel row tiles (the slab stage does not fit twice: tap schedule)
    _conv_case(ops, 1, 6, 128, 128, 256, 3, seed=80)
    _conv_case(ops, 1, 4, 256, 64, 256, 3, seed=81)


def test_fused_adam_matches_torch_adam(ops):
    """fv_adam_multi (one launch over a tensor table) against torch.optim.Adam, three steps, odd sizes and a tail."""
    from face_vae_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(64, 32, 3, 3), (257,), (3, 32, 7, 7), (1,), (256, 256, 3, 3), (33, 5)]
    pa = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = FusedAdam(pa, lr=5e-5, betas=(0.5, 0.999))
    ob = torch.optim.Adam(pb, lr=5e-5, betas=(0.5, 0.999))
    for it in range(3):
        for a, b in zip(pa, pb):
            g = torch.randn_like(a) * (10.0 ** (it - 1))
            a.grad, b.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-6, atol=2e-7), (a - b).abs().max().item()
    for a, b in zip(pa, pb):
        assert torch.allclose(oa.state[a]["exp_avg_sq"], ob.state[b]["exp_avg_sq"], rtol=1e-5, atol=1e-12)   # a few ulp: FMA contraction
        assert float(oa.state[a]["step"]) == 3.0


def test_conv_many_tiles_persistent(ops):
    # more tiles than SMs: exercises the persistent loop, TMEM double buffering and mbarrier phase wrap-around
    _conv_case(ops, 8, 64, 128, 32, 64, 3)


@pytest.mark.parametrize("ci,co,k,h,w,n", [(64, 64, 3, 16, 64, 2), (32, 64, 3, 16, 64, 2), (16, 32, 3, 8, 8, 4),
                                           (128, 256, 3, 16, 16, 2), (256, 256, 3, 16, 16, 2), (64, 32, 3, 8, 128, 2),
                                           (16, 256, 1, 16, 16, 2), (32, 3, 7, 32, 32, 2), (64, 128, 3, 4, 4, 6),
                                           (3, 32, 1, 16, 64, 2)])
def test_conv_dgrad_wgrad(ops, ci, co, k, h, w, n):
    from face_vae_b200.ops import pad_channels
    x = _rand((n, ci, h, w), 10).requires_grad_(True)
    wt = _rand((co, ci, k, k), 11, -0.2, 0.2).requires_grad_(True)
    dy = _rand((n, co, h, w), 12)
    y = F.conv2d(x, wt, None, padding=(k - 1) // 2)
    y.backward(dy)
    x_nhwc = ops.nchw_to_nhwc(x.detach())
    dy_nhwc = ops.nchw_to_nhwc(dy, pad_channels(co))
    _, wd = ops.weight_prep(wt.detach(), False, True)
    dx = ops.conv2d(dy_nhwc, wd, None, pad_channels(ci), k)          # data gradient = conv with the rotated filter
    acc = ops.conv2d_wgrad(x_nhwc, dy_nhwc, k)
    dw = ops.wgrad_finish(acc, co, ci, k)
    torch.cuda.synchronize()
    _report(f"dgrad ci{ci} co{co} k{k} {h}x{w}", dx[..., :ci], x.grad.permute(0, 2, 3, 1), 2e-2, 4e-3)
    err = (dw - wt.grad).abs().max().item()
    scale = wt.grad.abs().max().item()
    print(f"wgrad ci{ci} co{co} k{k} {h}x{w}: max_err {err:.4e} absmax {scale:.4e}")
    assert err <= 2e-3 * scale + 1e-5, f"wgrad mismatch {err} vs {scale}"


def test_layout_roundtrip(ops):
    x = _rand((3, 5, 8, 16), 1, bf16_exact=True)
    y = ops.nchw_to_nhwc(x, 16)
    assert torch.equal(y[..., :5].float(), x.permute(0, 2, 3, 1))
    assert float(y[..., 5:].float().abs().max()) == 0
    back = ops.nhwc_to_nchw(y, 5)
    assert torch.equal(back, x)
    back2 = ops.nhwc_to_nchw(y, 5, out=back.clone(), accumulate=True)
    assert torch.equal(back2, 2 * x)
    yf = ops.nchw_to_nhwc(x, 8, dtype=torch.float32)
    assert torch.equal(yf[..., :5], x.permute(0, 2, 3, 1))


@pytest.mark.parametrize("c,dtype", [(32, torch.bfloat16), (64, torch.bfloat16), (256, torch.bfloat16), (16, torch.float32), (32, torch.float32)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_bn_act_fwd_bwd(ops, c, dtype, mode):
    from face_vae_b200.ops import ACT_RELU
    n, h, w = 3, 8, 16
    y = _rand((n, c, h, w), 3, -2, 2).to(dtype).float().requires_grad_(True)
    gamma = _rand((c,), 4, 0.5, 1.5, False).requires_grad_(True)
    beta = _rand((c,), 5, -0.3, 0.3, False).requires_grad_(True)
    rm, rv = torch.zeros(c).cuda(), torch.ones(c).cuda()
    rm_ref, rv_ref = rm.clone(), rv.clone()
    a = F.relu(F.batch_norm(y, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5))
    if mode == 1:
        a = F.avg_pool2d(a, 2)
    elif mode == 2:
        a = F.interpolate(a, scale_factor=2, mode="nearest")
    g = _rand(tuple(a.shape), 6)
    a.backward(g)
    y_nhwc = y.detach().permute(0, 2, 3, 1).contiguous().to(dtype)
    sums = ops.bn_stats(y_nhwc)
    stat = ops.bn_finalize(sums, n * h * w, gamma.detach(), beta.detach(), rm, rv)
    out = ops.bn_act_fwd(y_nhwc, stat, mode, ACT_RELU, torch.bfloat16)
    torch.cuda.synchronize()
    torch.testing.assert_close(rm, rm_ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv, rv_ref, rtol=1e-4, atol=1e-5)
    _report(f"bn_act_fwd c{c} mode{mode}", out, a.detach().permute(0, 2, 3, 1), 1e-2, 4e-3)
    g_nhwc = g.permute(0, 2, 3, 1).contiguous().bfloat16()
    s = ops.bn_act_bwd_reduce(y_nhwc, g_nhwc, stat, mode, ACT_RELU)
    dgamma, dbeta, coef = ops.bn_bwd_finalize(s, s, n * h * w, c)
    dy = ops.bn_act_bwd_apply(y_nhwc, g_nhwc, stat, coef, mode, ACT_RELU)
    torch.cuda.synchronize()
    torch.testing.assert_close(dgamma, gamma.grad, rtol=2e-3, atol=2e-3 * gamma.grad.abs().max().item())
    torch.testing.assert_close(dbeta, beta.grad, rtol=2e-3, atol=2e-3 * beta.grad.abs().max().item())
    _report(f"bn_act_bwd c{c} mode{mode}", dy, y.grad.permute(0, 2, 3, 1), 2e-2, 4e-3)
    # NCHW fp32 output / gradient variants used at the VAE bottleneck
    if mode != 2:
        out_nchw = ops.bn_act_fwd(y_nhwc, stat, mode, ACT_RELU, torch.float32, nchw_out=True)
        torch.testing.assert_close(out_nchw, a.detach(), rtol=1e-4, atol=1e-4)
        s2 = ops.bn_act_bwd_reduce(y_nhwc, g.contiguous(), stat, mode, ACT_RELU, g_nchw=True)
        _, _, coef2 = ops.bn_bwd_finalize(s2, s2, n * h * w, c)
        dy2 = ops.bn_act_bwd_apply(y_nhwc, g.contiguous(), stat, coef2, mode, ACT_RELU, g_nchw=True)
        _report(f"bn_act_bwd nchw-g c{c} mode{mode}", dy2, y.grad.permute(0, 2, 3, 1), 2e-2, 4e-3)


def test_bn_eval_and_colsum(ops):
    from face_vae_b200.ops import ACT_LEAKY
    c = 64
    y = _rand((2, c, 8, 8), 7, -2, 2)
    gamma, beta = _rand((c,), 8, 0.5, 1.5, False), _rand((c,), 9, -0.3, 0.3, False)
    rm, rv = _rand((c,), 10, -0.5, 0.5, False), _rand((c,), 11, 0.5, 2.0, False)
    ref = F.leaky_relu(F.batch_norm(y, rm, rv, gamma, beta, False, 0.1, 1e-5), 0.2)
    stat = ops.bn_eval_affine(gamma, beta, rm, rv)
    y_nhwc = y.permute(0, 2, 3, 1).contiguous().bfloat16()
    out = ops.bn_act_fwd(y_nhwc, stat, 0, ACT_LEAKY, torch.float32)
    _report("bn eval leaky", out, ref.permute(0, 2, 3, 1), 1e-4, 1e-5)
    cs = ops.colsum(y_nhwc)
    torch.testing.assert_close(cs, y.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)


def test_reparam_kl(ops):
    n, dz = 5, 4096
    h = _rand((n, 2 * dz), 20, 0.0, 2.0, False)
    eps = torch.randn((n, dz), generator=torch.Generator().manual_seed(21)).cuda()
    mu, ls = h[:, :dz], h[:, dz:]
    z, kl = ops.reparam_kl_fwd(mu, ls, eps)
    z_ref = mu + torch.exp(ls) * eps
    kl_ref = (-0.5 - ls + 0.5 * mu ** 2 + 0.5 * torch.exp(2 * ls)).double().sum(dim=1)
    torch.testing.assert_close(z, z_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(kl.double().sum(dim=1), kl_ref, rtol=1e-4, atol=0)            # KL: rtol 1e-4 (north_star)
    # known answers of SURVEY.md section 4
    i = torch.arange(1024, dtype=torch.float32).view(4, 256).cuda()
    mu2, ls2 = torch.sin(0.01 * i), 0.5 * torch.cos(0.013 * i)
    _, kl2 = ops.reparam_kl_fwd(mu2, ls2, None, want_z=False)
    assert abs(kl2.sum().item() / 1024 - 0.37965357) < 1e-4 * 0.37965357
    # backward: dz given, KL weight kscale
    dzt = _rand((n, dz), 22, bf16_exact=False)
    dh = ops.reparam_kl_bwd(mu, ls, eps, dzt, None, None, 0.2 / (n * dz), None)
    torch.testing.assert_close(dh[:, :dz], dzt + 0.2 / (n * dz) * mu, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(dh[:, dz:], dzt * eps * torch.exp(ls) + 0.2 / (n * dz) * (torch.exp(2 * ls) - 1), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("l1", [False, True])
def test_recon_loss(ops, l1):
    n, c, h, w = 3, 3, 16, 32
    logits = _rand((n, c, h, w), 30, -3, 3, False).requires_grad_(True)
    target = _rand((n, c, h, w), 31, 0, 1, False)
    pred = torch.sigmoid(logits)
    ref = (pred - target).abs().sum() if l1 else ((pred - target) ** 2).sum()
    ref.backward()
    loss, p, gf, gn = ops.recon_loss(logits.detach(), target, l1=l1, use_sigmoid=True, gscale=1.0, want_grad_f32=True)
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(p, pred.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gf, logits.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(gn[..., :c].float(), logits.grad.permute(0, 2, 3, 1), rtol=1e-2, atol=1e-3)
    assert float(gn[..., c:].float().abs().max()) == 0
    a, b = _rand((7, 333), 32, bf16_exact=False), _rand((7, 333), 33, bf16_exact=False)
    loss2, g2 = ops.recon_loss_flat(a, b, l1=l1, gscale=0.5)
    ref2 = (a - b).abs().sum() if l1 else ((a - b) ** 2).sum()
    torch.testing.assert_close(loss2[0], ref2, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(g2, 0.5 * (torch.sign(a - b) if l1 else 2 * (a - b)), rtol=1e-5, atol=1e-6)
    s = torch.tensor([3.0]).cuda()
    sc = ops.scale(g2.view(-1)[:2328].contiguous(), s, 2.0)
    torch.testing.assert_close(sc, g2.view(-1)[:2328] * 6.0, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("n,h,w,ci,co,k,out_mode", [
    (3, 37, 256, 32, 64, 3, 0),      # column changes in the middle of a CTA's run of tiles
    (8, 64, 256, 64, 32, 3, 0),      # several tiles per CTA, 128-byte slab rows, 64-byte output rows
    (2, 33, 128, 16, 16, 3, 0),      # 32-byte rows in and out
    (3, 20, 128, 32, 3, 7, 2),       # 7x7 out_conv, NCHW fp32 output
    (3, 20, 128, 16, 32, 7, 0),      # its data gradient shape (16 -> 32 channels, 7x7)
    (2, 16, 384, 64, 128, 3, 0),     # Co_pad > 64: direct stores from the ring kernel
    (2, 16, 128, 32, 64, 3, 1),      # fp32 NHWC output
    (75, 4, 128, 32, 3, 7, 2),       # runs of 3 tiles over columns of 4 rows: a run may START on the last row of a column
    (75, 4, 128, 16, 32, 7, 0),      # (the second issuer's first window then lies entirely behind slabs it never reads)
    (50, 6, 128, 64, 32, 3, 0),
])
def test_conv_ring_schedule(ops, n, h, w, ci, co, k, out_mode):
    """Sliding-window schedule (fv_conv_ring.cu): resident filter, one new input-row slab per tile, TMA-store epilogue."""
    _conv_case(ops, n, h, w, ci, co, k, out_mode=out_mode, seed=40)


@pytest.mark.parametrize("n,h,w,ci,co,k", [
    (3, 37, 256, 32, 64, 3),      # enc.1-like: 3 M tiles, column changes inside a CTA's run
    (4, 64, 128, 64, 32, 3),      # up.3-like: two taps per M tile (128-byte slab rows)
    (3, 20, 128, 32, 3, 7),       # out_conv: 14 accumulators of 16 columns
    (2, 33, 64, 16, 32, 3),       # eight taps per M tile (32-byte slab rows)
    (2, 16, 192, 64, 64, 5),      # 5x5
    (2, 24, 128, 64, 128, 3),     # enc.2-like: dY walked in two 64-channel chunks
    (2, 24, 128, 128, 64, 3),     # up.2-like: x walked in two 64-channel chunks
    (1, 10, 64, 32, 192, 3),      # three dY chunks, 32-channel slabs
    (1, 10, 64, 256, 32, 3),      # four x chunks, dY padded to 32 channels
])
def test_wgrad_ring_schedule(ops, n, h, w, ci, co, k):
    """Sliding-window weight gradient (fv_wgrad_ring.cu) against autograd."""
    from face_vae_b200.ops import pad_channels
    x = _rand((n, ci, h, w), 50)
    wt = _rand((co, ci, k, k), 51, -0.2, 0.2).requires_grad_(True)
    dy = _rand((n, co, h, w), 52)
    F.conv2d(x, wt, None, padding=(k - 1) // 2).backward(dy)
    acc = ops.conv2d_wgrad(ops.nchw_to_nhwc(x), ops.nchw_to_nhwc(dy, pad_channels(co)), k)
    dw = ops.wgrad_finish(acc, co, ci, k)
    torch.cuda.synchronize()
    err = (dw - wt.grad).abs().max().item()
    scale = wt.grad.abs().max().item()
    print(f"wgrad-ring ci{ci} co{co} k{k} {h}x{w}: max_err {err:.4e} absmax {scale:.4e}")
    if err > 2e-3 * scale + 1e-5:
        bad = ((dw - wt.grad).abs() > 2e-3 * scale + 1e-5)
        print("   bad frac by tap:", [round(v, 2) for v in bad.float().mean(dim=(0, 1)).flatten().tolist()])
        print("   bad frac by ci :", [round(v, 2) for v in bad.float().mean(dim=(0, 2, 3)).tolist()][:16])
        print("   bad frac by co :", [round(v, 2) for v in bad.float().mean(dim=(1, 2, 3)).tolist()][:16])
    assert err <= 2e-3 * scale + 1e-5, f"wgrad mismatch {err} vs {scale}"


@pytest.mark.parametrize("c,n,h,w", [(3, 3, 24, 40), (1, 3, 24, 40), (4, 3, 24, 40),
                                     (3, 3, 32, 64), (3, 7, 64, 128)])     # the last two: H*W % 256 == 0 -> the cp.async.bulk pipeline
def test_pointwise_first_layer(ops, c, n, h, w):
    """SameBlock2D(c <= 4 -> 32) fast path (fv_pointwise.cu): statistics from input moments, closed-form parameter
    gradients; against F.conv2d + F.batch_norm + relu in fp32."""
    from face_vae_b200.ops import ACT_RELU
    co = 32
    x = _rand((n, c, h, w), 60, 0.0, 1.0, False)
    wt = _rand((co, c, 1, 1), 61, -0.6, 0.6, False).requires_grad_(True)
    b = _rand((co,), 62, -0.2, 0.2, False).requires_grad_(True)
    gamma = _rand((co,), 63, 0.5, 1.5, False).requires_grad_(True)
    beta = _rand((co,), 64, -0.3, 0.3, False).requires_grad_(True)
    rm, rv = torch.zeros(co).cuda(), torch.ones(co).cuda()
    rm_ref, rv_ref = rm.clone(), rv.clone()
    a = F.relu(F.batch_norm(F.conv2d(x, wt, b), rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5))
    g = _rand(tuple(a.shape), 65)
    a.backward(g)
    count = n * h * w
    fs = ops.pw_moments(x)
    coef, stat = ops.pw_prepare(fs, count, wt.detach().reshape(co, c).contiguous(), b.detach(), gamma.detach(), beta.detach(), rm, rv)
    out = ops.pw_fwd(x, coef, ACT_RELU)
    torch.cuda.synchronize()
    torch.testing.assert_close(rm, rm_ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rv, rv_ref, rtol=1e-4, atol=1e-5)
    _report(f"pointwise fwd c{c}", out, a.detach().permute(0, 2, 3, 1), 1e-2, 4e-3)
    g_nhwc = g.permute(0, 2, 3, 1).contiguous().bfloat16()
    bs = ops.pw_bwd_reduce(x, g_nhwc, coef, ACT_RELU)
    dw, dgamma, dbeta = ops.pw_bwd_finalize(fs, bs, count, wt.detach().reshape(co, c).contiguous(), b.detach(), gamma.detach(), stat)
    torch.cuda.synchronize()
    for name, got, ref in (("dw", dw, wt.grad.reshape(co, c)), ("dgamma", dgamma, gamma.grad), ("dbeta", dbeta, beta.grad)):
        err = (got - ref).abs().max().item()
        scale = ref.abs().max().item()
        print(f"pointwise {name} c{c}: max_err {err:.3e} absmax {scale:.3e}")
        assert err <= 1e-2 * scale + 1e-4, (name, err, scale)


@pytest.mark.parametrize("ci,co,k,h,w,n", [(512, 512, 3, 8, 8, 2), (64, 512, 1, 16, 16, 2), (512, 256, 3, 16, 16, 2), (256, 384, 3, 8, 16, 2)])
def test_conv_more_than_256_output_channels(ops, ci, co, k, h, w, n):
    """Co > 256 (the 512x512 'deeper' variant, BASELINE.json configs[3]): output channels are produced in chunks of one UMMA N."""
    from face_vae_b200.ops import pad_channels
    _conv_case(ops, n, h, w, ci, co, k, seed=70)
    x = _rand((n, ci, h, w), 71).requires_grad_(True)
    wt = _rand((co, ci, k, k), 72, -0.1, 0.1).requires_grad_(True)
    dy = _rand((n, co, h, w), 73)
    F.conv2d(x, wt, None, padding=(k - 1) // 2).backward(dy)
    x_nhwc, dy_nhwc = ops.nchw_to_nhwc(x.detach()), ops.nchw_to_nhwc(dy, pad_channels(co))
    _, wd = ops.weight_prep(wt.detach(), False, True)
    dx = ops.conv2d(dy_nhwc, wd, None, pad_channels(ci), k)
    dw = ops.wgrad_finish(ops.conv2d_wgrad(x_nhwc, dy_nhwc, k), co, ci, k)
    torch.cuda.synchronize()
    _report(f"dgrad ci{ci} co{co}", dx[..., :ci], x.grad.permute(0, 2, 3, 1), 2e-2, 4e-3)
    err, scale = (dw - wt.grad).abs().max().item(), wt.grad.abs().max().item()
    assert err <= 2e-3 * scale + 1e-5, (err, scale)


def test_batched_weight_prep_matches_single_layer_entry_points(ops):
    """ops.step_scope prepares the bf16 operands of every conv of a model with ONE flat-grid launch (fv_weight_prep_flat): bitwise
    equal to the per-layer entry points for the three operand kinds (plain, up-sampling 3x3, 4x4 stride 2)."""
    from face_vae_b200 import ops as O
    torch.manual_seed(3)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Conv2d(3, 32, 1)
            self.b = torch.nn.Conv2d(32, 64, 3, padding=1)
            self.c = torch.nn.Conv2d(256, 256, 3, padding=1)
            self.c.prep_kind = O.PREP_UP
            self.d = torch.nn.Conv2d(64, 128, 4, 2, 1)
            self.d.prep_kind = O.PREP_S2
            self.e = torch.nn.Conv2d(16, 256, 1)
            self.f = torch.nn.Conv2d(128, 64, 3, padding=1)
            self.f.prep_kind = O.PREP_UP
            self.g = torch.nn.Conv2d(16, 32, 5, padding=2)          # more than 16 taps: item-wise blocks inside the tiled launch
            self.h = torch.nn.Conv2d(48, 24, 3, padding=1)          # padded to 64 / 32 channels

    m = M().cuda()
    with O.step_scope(m):
        got = {"a": O.weight_prep(m.a.weight, True, True), "b": O.weight_prep(m.b.weight, True, True), "e": O.weight_prep(m.e.weight, True, True),
               "g": O.weight_prep(m.g.weight, True, True), "h": O.weight_prep(m.h.weight, True, True),
               "c": O.weight_prep_up(m.c.weight, True, True), "f": O.weight_prep_up(m.f.weight, True, True),
               "d": O.weight_prep_s2(m.d.weight, True, True)}
        got = {k: tuple(t.clone() for t in v) for k, v in got.items()}
    ref = {"a": O.weight_prep(m.a.weight, True, True), "b": O.weight_prep(m.b.weight, True, True), "e": O.weight_prep(m.e.weight, True, True),
           "g": O.weight_prep(m.g.weight, True, True), "h": O.weight_prep(m.h.weight, True, True),
           "c": O.weight_prep_up(m.c.weight, True, True), "f": O.weight_prep_up(m.f.weight, True, True),
           "d": O.weight_prep_s2(m.d.weight, True, True)}
    torch.cuda.synchronize()
    for k in ref:
        for g, r in zip(got[k], ref[k]):
            assert g.shape == r.shape and torch.equal(g, r), k


@pytest.mark.parametrize("mean_over_std", [8.0, 64.0])
def test_bn_statistics_with_large_mean(ops, mean_over_std):
    """Round-1 ADVICE: the variance is formed as E[y^2] - E[y]^2 from fp32 sums -- prone to cancellation when |mean| >> std, unlike
    torch's Welford.  The sums are tree-reduced (per-thread partials of <= a few hundred values, then ordered block sums) and the
    subtraction is done in double: with a mean 8 / 64 standard deviations away from zero over 512 k values per channel the variance
    must still be within 1e-3 / 3e-2 relative of the fp64 value (bf16 activations cannot carry a larger ratio with a non-trivial spread)."""
    n, h, w, c = 8, 256, 256, 32
    g = torch.Generator(device="cuda").manual_seed(7)
    y32 = torch.randn((n, h, w, c), device="cuda", generator=g) + mean_over_std
    y = y32.to(torch.bfloat16)
    yd = y.double().reshape(-1, c)
    var_ref = yd.var(dim=0, unbiased=False)
    mean_ref = yd.mean(dim=0)
    gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    stat = ops.bn_finalize(ops.bn_stats(y), n * h * w, gamma, beta, None, None)
    torch.cuda.synchronize()
    mean, invstd = stat[0].double(), stat[1].double()
    var = 1.0 / invstd ** 2 - 1e-5
    assert torch.allclose(mean, mean_ref, rtol=1e-5, atol=0)
    rel = ((var - var_ref).abs() / var_ref).max().item()
    print(f"bn large-mean: mean/std {mean_over_std}: variance rel err {rel:.2e}")
    assert rel < (1e-3 if mean_over_std <= 8 else 3e-2), rel
