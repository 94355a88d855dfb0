"""The two resolution-changing convolution geometries of csrc/fv_conv.cu / fv_wgrad.cu against plain PyTorch fp32:

 * x2: nn.Upsample(x2, nearest) + 3x3 conv (UpBlock2D, reference modules.py:78-89) computed as four 2x2 phase convolutions
   on the coarse grid -- forward, data gradient (a 4x4 stride-2 convolution of dY) and weight gradient;
 * s2: the 4x4 stride-2 pad-1 convolution of Conv2dELR / EFE_conv6.efe_encoder (reference models_utils.py:632-744,
   models.py:845-852) -- forward, data gradient (four 2x2 phases) and weight gradient.
Inputs are bf16-representable, so the only differences are accumulation order and the bf16 rounding of the outputs
(and, for x2, of the summed phase filters): rtol 2e-2 on bf16 outputs, 2e-3 of the maximum on fp32 weight gradients."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.test_kernels_gpu import _rand, _report, ops  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def _ref_up(x, w, b):
    return F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)


@pytest.mark.parametrize("n,h,w,ci,co", [
    (2, 8, 16, 64, 64),        # small tiles spanning rows
    (2, 16, 16, 256, 256),     # up.0 shape: wide, output channels possibly split over CTAs
    (3, 32, 32, 256, 128),     # up.1
    (2, 64, 64, 128, 64),      # up.2
    (2, 16, 128, 64, 32),      # up.3 shape: 128-pixel row tiles, thin output
    (5, 4, 4, 32, 16),         # tiles spanning several images, masked tail
    (2, 8, 8, 512, 512),       # 512-deep variant: more than 256 output channels (chunks)
    (1, 6, 256, 16, 32),       # two tiles per row
    (8, 64, 128, 64, 32),      # window schedule (fv_conv_win.cu): several tiles per CTA, slab ring wraps around
    (3, 5, 256, 32, 64),       # window schedule: column changes inside a CTA's run, 64 output channels (4 x 64 TMEM columns x 2)
    (2, 7, 128, 16, 16),       # window schedule: 32-byte slab rows
    (37, 4, 128, 64, 32),      # window schedule: runs that start on the last row of a column
])
def test_upsample_conv_forward_backward(ops, n, h, w, ci, co):
    from face_vae_b200.ops import pad_channels
    x = _rand((n, ci, h, w), 1).requires_grad_(True)
    b3 = 1.0 / math.sqrt(ci * 9)
    wt = _rand((co, ci, 3, 3), 2, -b3, b3).requires_grad_(True)
    bias = _rand((co,), 3, bf16_exact=False)
    dy = _rand((n, co, 2 * h, 2 * w), 4)
    y_ref = _ref_up(x, wt, bias)
    y_ref.backward(dy)
    xn = ops.nchw_to_nhwc(x.detach())
    wx2, ws2 = ops.weight_prep_up(wt.detach())
    y = ops.conv2d_x2(xn, wx2, bias, co)
    cop = pad_channels(co)
    assert y.shape == (n, 2 * h, 2 * w, cop)
    torch.cuda.synchronize()
    # the phase filters are sums of up to four taps rounded to bf16 once: a slightly different (not worse) rounding than
    # rounding each tap -- allow the bf16 floor relative to the output scale
    _report(f"x2 fwd n{n} {h}x{w} ci{ci} co{co}", y[..., :co], y_ref.detach().permute(0, 2, 3, 1), 2e-2, 8e-3)
    if cop > co:
        assert float(y[..., co:].float().abs().max()) == 0.0
    dyn = ops.nchw_to_nhwc(dy, cop)
    dx = ops.conv2d_s2(dyn, ws2, None, pad_channels(ci), alg_taps=36)
    assert dx.shape == (n, h, w, pad_channels(ci))
    part = ops.conv2d_wgrad_x2(xn, dyn)
    dw = ops.wgrad_finish_up(part, co, ci)
    torch.cuda.synchronize()
    _report(f"x2 dgrad n{n} {h}x{w} ci{ci} co{co}", dx[..., :ci], x.grad.permute(0, 2, 3, 1), 2e-2, 8e-3)
    err, scale = (dw - wt.grad).abs().max().item(), wt.grad.abs().max().item()
    print(f"x2 wgrad: max_err {err:.4e} absmax {scale:.4e}")
    assert err <= 2e-3 * scale + 1e-5, (err, scale)


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 16, 16, 256, 256), (2, 16, 128, 64, 32), (3, 8, 8, 64, 128), (2, 4, 4, 128, 512), (9, 33, 128, 64, 32),
                                         (2, 9, 256, 32, 64), (2, 6, 128, 16, 16)])
def test_upsample_conv_fused_statistics(ops, n, h, w, ci, co):
    """fv_conv2d_x2 with the batch-norm sums of y from the epilogue (or the separate pass, per the library's predicate)."""
    from face_vae_b200.ops import pad_channels
    x = _rand((n, ci, h, w), 5)
    wt = _rand((co, ci, 3, 3), 6, -0.1, 0.1)
    xn = ops.nchw_to_nhwc(x)
    wx2, _ = ops.weight_prep_up(wt, True, False)
    y, sums = ops.conv2d_x2(xn, wx2, None, co, want_stats=True)
    y2 = ops.conv2d_x2(xn, wx2, None, co)
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
    yf = y.float().reshape(-1, pad_channels(co))
    exact = torch.cat([yf.sum(0), (yf * yf).sum(0)])
    _report("x2 fused stats", sums.reshape(1, 1, 1, -1), exact.reshape(1, 1, 1, -1), 1e-4, 1e-5)


@pytest.mark.parametrize("n,h,w,ci,co", [
    (2, 8, 16, 64, 64),
    (2, 32, 32, 32, 64),       # EFE_conv6-like first stages (3 -> 32 -> 64 ...), coarse 32x32
    (2, 16, 16, 128, 256),
    (3, 4, 4, 256, 128),
    (1, 4, 128, 16, 32),       # row tiles
    (2, 8, 8, 3, 32),          # RGB input padded to 16 channels
])
def test_stride2_conv_forward_backward(ops, n, h, w, ci, co):
    """4x4 stride-2 pad-1 convolution: (h, w) is the OUTPUT (coarse) size, the input is 2h x 2w."""
    from face_vae_b200.ops import pad_channels
    x = _rand((n, ci, 2 * h, 2 * w), 11).requires_grad_(True)
    b4 = 1.0 / math.sqrt(ci * 16)
    wt = _rand((co, ci, 4, 4), 12, -b4, b4).requires_grad_(True)
    bias = _rand((co,), 13, bf16_exact=False)
    dy = _rand((n, co, h, w), 14)
    y_ref = F.conv2d(x, wt, bias, stride=2, padding=1)
    y_ref.backward(dy)
    xn = ops.nchw_to_nhwc(x.detach())
    wf, wx2 = ops.weight_prep_s2(wt.detach())
    y = ops.conv2d_s2(xn, wf, bias, co)
    cop, cip = pad_channels(co), pad_channels(ci)
    assert y.shape == (n, h, w, cop)
    torch.cuda.synchronize()
    _report(f"s2 fwd n{n} {h}x{w} ci{ci} co{co}", y[..., :co], y_ref.detach().permute(0, 2, 3, 1), 2e-2, 4e-3)
    dyn = ops.nchw_to_nhwc(dy, cop)
    dx = ops.conv2d_x2(dyn, wx2, None, cip, alg_taps=16)
    assert dx.shape == (n, 2 * h, 2 * w, cip)
    dw = ops.wgrad_finish(ops.conv2d_wgrad_s2(xn, dyn), co, ci, 4)
    torch.cuda.synchronize()
    _report(f"s2 dgrad n{n} {h}x{w} ci{ci} co{co}", dx[..., :ci], x.grad.permute(0, 2, 3, 1), 2e-2, 4e-3)
    err, scale = (dw - wt.grad).abs().max().item(), wt.grad.abs().max().item()
    print(f"s2 wgrad: max_err {err:.4e} absmax {scale:.4e}")
    assert err <= 2e-3 * scale + 1e-5, (err, scale)


def test_upblock_module_never_materialises_the_upsampled_tensor(ops):
    """UpBlock2D through the nn.Module facade: same numbers as the explicit up-sample + block, and the step issues no
    up-sampling norm+act launch (MODE_UP) any more."""
    from face_vae_b200 import _lib
    import face_vae_b200.modules as M
    torch.manual_seed(0)
    blk = M.UpBlock2D(64, 32, False).cuda().train()
    x = _rand((2, 64, 16, 16), 21).requires_grad_(True)
    calls = []
    orig = _lib.call

    def spy(name, *a, **k):
        calls.append((name, a))
        return orig(name, *a, **k)

    _lib.call = spy
    import face_vae_b200.ops as O
    O.call = spy
    try:
        y = blk(x)
        gy = _rand(tuple(y.shape), 22)
        (y.float() * gy).sum().backward()
    finally:
        _lib.call = orig
        O.call = orig
    torch.cuda.synchronize()
    names = [c[0] for c in calls]
    assert "fv_conv2d_x2" in names and "fv_conv2d_s2" in names and "fv_conv2d_wgrad_x2" in names
    assert y.shape == (2, 32, 32, 32)
    ref_blk = torch.nn.Sequential(torch.nn.Upsample(scale_factor=2), torch.nn.Conv2d(64, 32, 3, 1, 1), torch.nn.BatchNorm2d(32),
                                  torch.nn.ReLU()).cuda().train()
    with torch.no_grad():
        ref_blk[1].weight.copy_(blk.layers[1].conv.weight)
        ref_blk[1].bias.copy_(blk.layers[1].conv.bias)
    x2 = x.detach().clone().requires_grad_(True)
    yr = ref_blk(x2)
    (yr * gy).sum().backward()
    _report("UpBlock2D y", y.float().permute(0, 2, 3, 1), yr.detach().permute(0, 2, 3, 1), 2e-2, 2e-2)
    rel = ((blk.layers[1].conv.weight.grad - ref_blk[1].weight.grad).norm() / ref_blk[1].weight.grad.norm()).item()
    assert rel < 5e-2, rel          # one bf16 block against fp32: ReLU-mask flips at |z| ~ 0 (the tight check is test_layerwise_gpu.py)
