"""One small invocation of every kernel family, for `compute-sanitizer --tool {memcheck,racecheck,synccheck}` (SURVEY.md section 5;
round-1 VERDICT item 9).  Shapes are the smallest that still exercise the multi-stage mbarrier pipelines, the TMEM double
buffering, the ring schedules, the shared-memory exchanges of the folded out_conv kernels and the cross-block reductions.

    compute-sanitizer --tool racecheck python tools/sanitize_subset.py        (one tool per gpurun call)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from face_vae_b200 import _lib, ops
from face_vae_b200.ops import ACT_LEAKY, ACT_RELU, MODE_NONE, MODE_POOL, pad_channels


def r(shape, lo=-1.0, hi=1.0, dtype=torch.float32):
    return (torch.rand(shape, device="cuda") * (hi - lo) + lo).to(dtype)


def main():
    _lib.call("fv_device_ok")
    torch.manual_seed(0)
    done = []
    # implicit GEMM: tap schedule, slab schedule, output-channel split, fused statistics, residual
    for (n, h, w, ci, co, k) in ((2, 8, 8, 64, 64, 3), (1, 4, 128, 64, 128, 3), (2, 16, 16, 256, 256, 3), (2, 8, 16, 16, 32, 1)):
        x = ops.nchw_to_nhwc(r((n, ci, h, w)))
        wt = r((co, ci, k, k), -0.1, 0.1)
        wf, wd = ops.weight_prep(wt)
        y, sums = ops.conv2d(x, wf, r((co,)), co, k, want_stats=True)
        y2 = ops.conv2d(x, wf, None, co, k, residual=y)
        dx = ops.conv2d(y, wd, None, pad_channels(ci), k)
        dw = ops.wgrad_finish(ops.conv2d_wgrad(x, y, k), co, ci, k)
        done.append("igemm %dx%d %d->%d" % (h, w, ci, co))
    # ring schedules (thin full-resolution layers)
    for (n, h, w, ci, co, k) in ((1, 6, 128, 32, 64, 3), (1, 5, 128, 64, 32, 3)):
        x = ops.nchw_to_nhwc(r((n, ci, h, w)))
        wt = r((co, ci, k, k), -0.1, 0.1)
        wf, wd = ops.weight_prep(wt)
        y, sums = ops.conv2d(x, wf, r((co,)), co, k, want_stats=True)
        dw = ops.wgrad_finish(ops.conv2d_wgrad(x, y, k), co, ci, k)
        done.append("ring %d->%d" % (ci, co))
    # x2 / s2 geometries
    for (n, h, w, ci, co) in ((2, 8, 8, 64, 32), (1, 4, 128, 64, 32)):
        x = ops.nchw_to_nhwc(r((n, ci, h, w)))
        wt = r((co, ci, 3, 3), -0.1, 0.1)
        wx2, ws2 = ops.weight_prep_up(wt)
        y, sums = ops.conv2d_x2(x, wx2, r((co,)), co, want_stats=True)
        dx = ops.conv2d_s2(y, ws2, None, pad_channels(ci))
        dw = ops.wgrad_finish_up(ops.conv2d_wgrad_x2(x, y), co, ci)
        w4 = r((co, ci, 4, 4), -0.1, 0.1)
        weff, inv = ops.demod_fwd(w4, 1.3, True)
        wf4, wx24 = ops.weight_prep_s2(weff)
        xf = ops.nchw_to_nhwc(r((n, ci, 2 * h, 2 * w)))
        y4 = ops.conv2d_ex(2, xf, wf4, r((co,)), co, 4, ACT_LEAKY)
        dy4 = ops.act_bwd(y4, y4, ACT_LEAKY)
        dw4 = ops.demod_bwd(w4, inv, ops.wgrad_finish(ops.conv2d_wgrad_s2(xf, dy4), co, ci, 4), 1.3, True)
        done.append("x2/s2 %dx%d" % (h, w))
    # tap-folded out_conv (7x7, 32 -> 3) + fused loss
    x = ops.nchw_to_nhwc(r((1, 32, 8, 128)))
    wt = r((3, 32, 7, 7), -0.05, 0.05)
    wq, wdq = ops.outconv_prep(wt)
    tgt = r((1, 3, 8, 128), 0, 1)
    out = ops.outconv_fwd(x, wq, r((3,)), 3, target=tgt, gscale=1e-3)
    one = torch.ones(1, device="cuda")
    dxo = ops.outconv_dgrad(out["g4"], wdq, one, 3)
    dwo = ops.outconv_wgrad(x, out["g4"], one, 3)
    done.append("outconv")
    # glue: statistics, norm + act forward / backward, column sums, losses, re-parameterisation, first layer, optimiser
    y = r((2, 16, 32, 64)).bfloat16()
    gamma, beta = r((64,), 0.5, 1.5), r((64,), -0.3, 0.3)
    rm, rv = torch.zeros(64, device="cuda"), torch.ones(64, device="cuda")
    s = ops.bn_stats(y)
    a, stat = ops.bn_act_fwd_fin(y, s, 2 * 16 * 32, gamma, beta, rm, rv, MODE_POOL, ACT_RELU)
    g = r(tuple(a.shape)).bfloat16()
    sb = ops.bn_act_bwd_reduce(y, g, stat, MODE_POOL, ACT_RELU)
    dy, dg, db = ops.bn_act_bwd_apply_fin(y, g, stat, sb, 2 * 16 * 32, MODE_POOL, ACT_RELU)
    cs = ops.colsum(y)
    big = r((8, 64, 128, 64)).bfloat16()          # several hundred blocks: two reduction levels
    s2 = ops.bn_stats(big)
    lg, tg = r((2, 3, 16, 16)), r((2, 3, 16, 16), 0, 1)
    ops.recon_loss(lg, tg, False, True, 1e-3, True, True, True)
    ops.recon_loss_flat(lg, tg)
    h = r((3, 512), 0, 1)
    z, kl = ops.reparam_kl_fwd(h[:, :256], h[:, 256:], r((3, 256)))
    ops.reparam_kl_bwd(h[:, :256], h[:, 256:], r((3, 256)), z, None, None, 1e-3, None)
    xf = r((2, 3, 16, 24), 0, 1)
    fs = ops.pw_moments(xf)
    coef, st = ops.pw_prepare(fs, 2 * 16 * 24, r((32, 3)), r((32,)), r((32,), 0.5, 1.5), r((32,)), None, None)
    o = ops.pw_fwd(xf, coef, ACT_RELU)
    ops.pw_bwd_reduce(xf, o, coef, ACT_RELU)
    ops.bilinear_resize(r((1, 3, 32, 32), 0, 1), 0.25)
    done.append("glue")
    torch.cuda.synchronize()
    print("sanitize subset ok:", ", ".join(done), "launches", _lib.launch_count)


if __name__ == "__main__":
    main()
