"""autograd.Functions that wire the C-ABI kernels into PyTorch's autograd.

Internal activation format: explicit NHWC bf16 tensors ``[N, H, W, Cp]`` (Cp = ops.pad_channels(C)); fp32 NCHW only at
the edges of the path (input frames, the VAE latent, the reconstruction).  Batch-norm statistics are exchanged across
data-parallel ranks here (sum all-reduce of ``[sum, sum_sq]`` forward and ``[sum_dz, sum_dz_xhat]`` backward), which
is what ``nn.SyncBatchNorm`` does under DDP in the reference (modules.py:19, logger.py:55).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops
from . import xrank
from .ops import (ACT_LEAKY, ACT_NONE, ACT_RELU, MODE_NONE, MODE_POOL, MODE_UP, OUT_NCHW_F32, OUT_NHWC_BF16,
                  OUT_NHWC_F32, pad_channels)

_SYNC_BN = True
import os as _os
_FUSE_STATS = _os.environ.get("FACEVAE_FUSE_BN_STATS", "1") != "0"
_FUSE_FIN = _os.environ.get("FACEVAE_FUSE_BN_FINALIZE", "1") != "0"
_FUSE_XRANK = _os.environ.get("FACEVAE_FUSE_XRANK", "1") != "0"


def _fin_fused(training: bool) -> bool:
    """Single-process training: the statistic finalize steps run inside the norm+act kernels (no cross-rank exchange needed)."""
    return training and _FUSE_FIN and _world() == 1


def set_sync_bn(enabled: bool) -> None:
    """Cross-rank batch statistics on/off (on by default whenever a process group with world_size > 1 exists)."""
    global _SYNC_BN
    _SYNC_BN = bool(enabled)


def _world() -> int:
    if _SYNC_BN and dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def _bn_backward_sums(y, g, stat, post_mode, act, g_nchw, count, training):
    """Backward pass 1 of a norm + act layer -> (dgamma, dbeta, coef) with the cross-rank exchange folded into the reduction
    kernel when data parallel, else None (the caller runs the separate reduce / finalize)."""
    if not training or _world() == 1 or not _FUSE_XRANK:
        return None
    xc = xrank.get()
    if xc is None:
        return None
    return xc.reduce_finalize_bwd(y, g, stat, post_mode, act, g_nchw, count)


def _bn_backward_finalize(s_local, count, c, training):
    """dgamma, dbeta (local sums) and the coupling coefficients (global sums / count)."""
    if not training:
        return ops.bn_bwd_finalize(s_local, torch.zeros_like(s_local), count, c)
    xc = xrank.get() if _world() > 1 else None
    if xc is not None:
        return xc.finalize_bwd(s_local, count, c)
    return ops.bn_bwd_finalize(s_local, _allreduce_sum(s_local), count, c)


def _allreduce_sum(t: torch.Tensor) -> torch.Tensor:
    """Sum across ranks; returns a new tensor (the local sums are still needed for dgamma / dbeta)."""
    if _world() == 1:
        return t
    g = t.clone()
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    return g


# ---------------------------------------------------------------------------------------------------- layout
class ToNHWC(torch.autograd.Function):
    """NCHW fp32 -> NHWC bf16 (channels zero-padded to pad_channels(C))."""

    @staticmethod
    def forward(ctx, x):
        ctx.c = x.shape[1]
        return ops.nchw_to_nhwc(x.contiguous().float())

    @staticmethod
    def backward(ctx, g):
        return ops.nhwc_to_nchw(g.contiguous(), ctx.c)


class ToNCHW(torch.autograd.Function):
    """NHWC (bf16 / fp32) -> NCHW fp32, first ``c`` channels."""

    @staticmethod
    def forward(ctx, x, c):
        ctx.cp = x.shape[3]
        return ops.nhwc_to_nchw(x, c)

    @staticmethod
    def backward(ctx, g):
        return ops.nchw_to_nhwc(g.contiguous().float(), ctx.cp), None


class Upsample2x(torch.autograd.Function):
    """nn.Upsample(scale_factor=2, nearest) on NHWC bf16 (reference modules.py:81); backward sums the 2x2 replicas."""

    @staticmethod
    def forward(ctx, x):
        c = x.shape[3]
        stat = torch.zeros((4, c), device=x.device, dtype=torch.float32)
        stat[1:3].fill_(1.0)                     # mean 0, invstd 1, scale 1, shift 0: identity affine
        ctx.save_for_backward(x, stat)
        return ops.bn_act_fwd(x, stat, MODE_UP, ACT_NONE, torch.bfloat16)

    @staticmethod
    def backward(ctx, g):
        x, stat = ctx.saved_tensors
        coef = torch.zeros((2, x.shape[3]), device=x.device, dtype=torch.float32)
        return ops.bn_act_bwd_apply(x, g.contiguous(), stat, coef, MODE_UP, ACT_NONE)


# ---------------------------------------------------------------------------------------------------- conv blocks
def _bn_forward(y, gamma, beta, running_mean, running_var, training, momentum, eps, sums=None):
    """Returns the [4, C] stat block (mean, invstd, scale, shift) and the global element count per channel.
    ``sums``: local [sum | sum of squares] already produced by the conv epilogue (otherwise one fv_bn_stats pass over y)."""
    n, h, w, c = y.shape
    if not training:
        return ops.bn_eval_affine(gamma, beta, running_mean, running_var, eps), n * h * w
    count = n * h * w * _world()
    xc = xrank.get() if _world() > 1 else None
    if xc is not None and sums is None and _FUSE_XRANK:
        # statistic pass, NVLink exchange and finalize in one launch: the reduction's last block pushes this rank's sums to all
        # peers, gathers theirs and writes the stat block
        return xc.stats_finalize_fwd(y, count, gamma, beta, running_mean, running_var, momentum, eps), count
    if sums is None:
        sums = ops.bn_stats(y)
    if xc is not None:   # one kernel: push partial sums to all peers over NVLink, reduce, finalize
        return xc.finalize_fwd(sums, count, gamma, beta, running_mean, running_var, momentum, eps), count
    sums = _allreduce_sum(sums)
    return ops.bn_finalize(sums, count, gamma, beta, running_mean, running_var, momentum, eps), count


GEOM_SAME, GEOM_UP = 0, 1


def _conv_forward(geom, x, weight, bias, ksize, want_stats, residual=None):
    """-> (y, sums | None, operand for the data gradient).  GEOM_UP: nearest 2x up-sampling + 3x3 conv as four 2x2 phase
    convolutions on the coarse input (csrc/fv_conv.cu, X2) -- the up-sampled tensor never exists."""
    co, ci = weight.shape[0], weight.shape[1]
    if geom == GEOM_UP:
        wx2, ws2 = ops.weight_prep_up(weight, True, x.requires_grad)
        if want_stats:
            y, sums = ops.conv2d_x2(x, wx2, bias, co, OUT_NHWC_BF16, (ci, co), want_stats=True)
        else:
            y, sums = ops.conv2d_x2(x, wx2, bias, co, OUT_NHWC_BF16, (ci, co)), None
        return y, sums, ws2
    wf, wd = ops.weight_prep(weight, True, x.requires_grad)
    if want_stats:      # the conv epilogue also produces the batch-norm sums of y where that is hidden (no second pass over y)
        y, sums = ops.conv2d(x, wf, bias, co, ksize, residual, OUT_NHWC_BF16, (ci, co), want_stats=True)
    else:
        y, sums = ops.conv2d(x, wf, bias, co, ksize, residual, OUT_NHWC_BF16, (ci, co)), None
    return y, sums, wd


def _conv_backward(geom, x, dy, weight, wd, ksize, need_dx):
    """-> (dx | None, dw) for y = conv(x, weight) in geometry ``geom``; dy: NHWC bf16 gradient of y."""
    co, ci = weight.shape[0], weight.shape[1]
    if geom == GEOM_UP:
        with ops.wgrad_stream(x, dy):
            dw = ops.wgrad_finish_up(ops.conv2d_wgrad_x2(x, dy, (ci, co)), co, ci)
        dx = None
        if need_dx:
            if wd is None:
                _, wd = ops.weight_prep_up(weight, False, True)
            # 4x4 stride-2 convolution of dY with the summed, mirrored taps: the 2x2 sum of the up-sampling backward is inside
            dx = ops.conv2d_s2(dy, wd, None, x.shape[3], OUT_NHWC_BF16, (co, ci), alg_taps=36)
        return dx, dw
    with ops.wgrad_stream(x, dy):
        dw = ops.wgrad_finish(ops.conv2d_wgrad(x, dy, ksize, (ci, co)), co, ci, ksize)
    dx = None
    if need_dx:
        if wd is None:
            _, wd = ops.weight_prep(weight, False, True)
        dx = ops.conv2d(dy, wd, None, x.shape[3], ksize, None, OUT_NHWC_BF16, (co, ci))
    return dx, dw


class ConvBNAct(torch.autograd.Function):
    """Pattern "CNA" of _ConvBlock (reference modules.py:8-42) plus the block's own pooling (DownBlock2D,
    modules.py:59-70) folded into the norm+act pass, or (``geom`` = GEOM_UP) UpBlock2D's up-sampling folded into the
    convolution itself (modules.py:78-89).

    x: NHWC bf16 [N,H,W,Ci_pad].  Returns NHWC (bf16 / fp32) or NCHW fp32 per ``out_nchw_f32``.
    """

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, ksize, post_mode, act, training,
                out_nchw_f32, momentum, eps, geom=GEOM_SAME):
        co, ci = weight.shape[0], weight.shape[1]
        # both filter operands come from one pass over the fp32 master weights (the dgrad operand is kept for backward)
        y, sums, wd = _conv_forward(geom, x, weight, bias, ksize, training and _FUSE_STATS)
        out_dtype = torch.float32 if out_nchw_f32 else torch.bfloat16
        if _fin_fused(training):
            count = y.shape[0] * y.shape[1] * y.shape[2]
            out, stat = ops.bn_act_fwd_fin(y, sums if sums is not None else ops.bn_stats(y), count, gamma, beta, running_mean,
                                           running_var, post_mode, act, out_dtype, out_nchw_f32, momentum, eps)
        else:
            stat, count = _bn_forward(y, gamma, beta, running_mean, running_var, training, momentum, eps, sums)
            out = ops.bn_act_fwd(y, stat, post_mode, act, out_dtype, out_nchw_f32)
        ctx.save_for_backward(x, y, stat, weight, wd)
        ctx.cfg = (ksize, post_mode, act, training, out_nchw_f32, count, co, ci, geom)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, y, stat, weight, wd = ctx.saved_tensors
        ksize, post_mode, act, training, out_nchw_f32, count, co, ci, geom = ctx.cfg
        g = g.contiguous()
        c = y.shape[3]
        # eval mode: running statistics are constants, dy = scale * dz, no coupling terms
        fused = _bn_backward_sums(y, g, stat, post_mode, act, out_nchw_f32, count, training)
        if fused is not None:
            dgamma, dbeta, coef = fused
            dy = ops.bn_act_bwd_apply(y, g, stat, coef, post_mode, act, None, out_nchw_f32)
        else:
            s_local = ops.bn_act_bwd_reduce(y, g, stat, post_mode, act, out_nchw_f32)
            if _fin_fused(training):
                dy, dgamma, dbeta = ops.bn_act_bwd_apply_fin(y, g, stat, s_local, count, post_mode, act, None, out_nchw_f32)
            else:
                dgamma, dbeta, coef = _bn_backward_finalize(s_local, count, c, training)
                dy = ops.bn_act_bwd_apply(y, g, stat, coef, post_mode, act, None, out_nchw_f32)
        dx, dw = _conv_backward(geom, x, dy, weight, wd, ksize, ctx.needs_input_grad[0])
        db = None
        if ctx.has_bias:
            # a bias in front of a (training-mode) batch norm has zero gradient analytically; eval mode: real sum
            db = ops.zero_grad_const(co, x.device) if training else ops.colsum(dy)[:co].clone()
        return dx, dw, db, dgamma, dbeta, None, None, None, None, None, None, None, None, None, None


class ConvELRAct(torch.autograd.Function):
    """Conv2dELR.forward (reference models_utils.py:706-744) without style modulation: weight -> F.normalize over (ci, r, s)
    ("demod") -> * weightgain -> conv (4x4 stride 2 pad 1, or stride-1 "same") -> + bias -> activation, the last three in
    ONE tensor-core kernel.  x: NHWC bf16.  Backward: act' from the stored output, bias gradient = column sums, weight
    gradient through the strided / plain wgrad kernels and the normalisation, data gradient = four-phase x2 convolution
    (stride 2) or the rotated-filter convolution (stride 1)."""

    @staticmethod
    def forward(ctx, x, weight, bias, ksize, stride, gain, demod, act):
        co, ci = weight.shape[0], weight.shape[1]
        weff, inv = ops.demod_fwd(weight, gain, demod)
        if stride == 2:
            wf, wx2 = ops.weight_prep_s2(weff, True, x.requires_grad)
            out = ops.conv2d_ex(2, x, wf, bias, co, 4, act, OUT_NHWC_BF16, (ci, co))
            wback = wx2
        else:
            wf, wd = ops.weight_prep(weff, True, x.requires_grad)
            out = ops.conv2d_ex(0, x, wf, bias, co, ksize, act, OUT_NHWC_BF16, (ci, co))
            wback = wd
        ctx.save_for_backward(x, out, weight, inv, wback)
        ctx.cfg = (ksize, stride, gain, demod, act, co, ci)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, out, weight, inv, wback = ctx.saved_tensors
        ksize, stride, gain, demod, act, co, ci = ctx.cfg
        g = g.contiguous()
        dy = ops.act_bwd(out, g, act) if act != ACT_NONE else g
        db = ops.colsum(dy)[:co].clone() if ctx.has_bias else None
        if stride == 2:
            dweff = ops.wgrad_finish(ops.conv2d_wgrad_s2(x, dy, (ci, co)), co, ci, 4)
        else:
            dweff = ops.wgrad_finish(ops.conv2d_wgrad(x, dy, ksize, (ci, co)), co, ci, ksize)
        dw = ops.demod_bwd(weight, inv, dweff, gain, demod)
        dx = None
        if ctx.needs_input_grad[0]:
            if stride == 2:
                if wback is None:
                    _, wback = ops.weight_prep_s2(ops.demod_fwd(weight, gain, demod)[0], False, True)
                dx = ops.conv2d_x2(dy, wback, None, x.shape[3], OUT_NHWC_BF16, (co, ci), alg_taps=16)
            else:
                if wback is None:
                    _, wback = ops.weight_prep(ops.demod_fwd(weight, gain, demod)[0], False, True)
                dx = ops.conv2d(dy, wback, None, x.shape[3], ksize, None, OUT_NHWC_BF16, (co, ci))
        return dx, dw, db, None, None, None, None, None


class InstanceNormAct(torch.autograd.Function):
    """nn.InstanceNorm2d(C, affine=True) + ReLU / LeakyReLU(0.2) on NHWC bf16 (reference modules.py:21,27,29; the Discriminator's
    blocks, models.py:1120-1127): per-(image, channel) statistics over H*W, no running statistics."""

    @staticmethod
    def forward(ctx, y, gamma, beta, act, eps):
        stat = ops.in_stats(y, eps)
        out = ops.in_act_fwd(y, stat, gamma, beta, act)
        ctx.save_for_backward(y, stat, gamma, beta)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, g):
        y, stat, gamma, beta = ctx.saved_tensors
        dy, dgamma, dbeta = ops.in_bwd(y, g.contiguous(), stat, gamma, beta, ctx.act)
        return dy, dgamma, dbeta, None, None


class ActOnly(torch.autograd.Function):
    """ReLU / LeakyReLU(0.2) on an NHWC bf16 tensor (a block without a norm layer)."""

    @staticmethod
    def forward(ctx, y, act):
        c = y.shape[3]
        stat = torch.zeros((4, c), device=y.device, dtype=torch.float32)
        stat[1:3].fill_(1.0)                     # identity affine
        out = ops.bn_act_fwd(y, stat, MODE_NONE, act, torch.bfloat16)
        ctx.save_for_backward(out)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        return ops.act_bwd(out, g.contiguous(), ctx.act), None


class PointwiseBNAct(torch.autograd.Function):
    """SameBlock2D with <= 4 input channels in training mode, straight from the NCHW fp32 frames (the first encoder layer,
    reference modules.py:97-108 via models.py:749): a 1x1 conv followed by batch norm is a per-pixel affine map whose
    batch statistics follow from the input moments, so forward is `moments(x)` + one pass, backward ONE pass over (g, x)
    with dW / dgamma / dbeta in closed form (csrc/fv_pointwise.cu).  The frames get no gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, act, momentum, eps):
        n, c, h, w = x.shape
        co = weight.shape[0]
        w2d = weight.reshape(co, c).contiguous()
        fsums = ops.pw_moments(x)
        if _world() > 1:
            dist.all_reduce(fsums, op=dist.ReduceOp.SUM)
        count = n * h * w * _world()
        coef, stat = ops.pw_prepare(fsums, count, w2d, bias, gamma, beta, running_mean, running_var, momentum, eps)
        out = ops.pw_fwd(x, coef, act)
        ctx.save_for_backward(x, w2d, bias if bias is not None else gamma.new_zeros(0), gamma, coef, stat, fsums)
        ctx.cfg = (act, count, bias is not None, tuple(weight.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        x, w2d, bias, gamma, coef, stat, fsums = ctx.saved_tensors
        act, count, has_bias, wshape = ctx.cfg
        bsums_local = ops.pw_bwd_reduce(x, g.contiguous(), coef, act)
        bsums = bsums_local
        if _world() > 1:
            bsums = bsums_local.clone()
            dist.all_reduce(bsums, op=dist.ReduceOp.SUM)
        b = bias if has_bias else None
        dw, dgamma, dbeta = ops.pw_bwd_finalize(fsums, bsums, count, w2d, b, gamma, stat)
        if _world() > 1:
            # the closed form yields the SUM over ranks of the per-rank gradients (it is built from all-reduced sums);
            # every rank reports 1/R of it so that the gradient all-reduce (mean over ranks) delivers exactly that sum / R,
            # the gradient of the global-mean loss -- the same value per-rank SyncBN gradients average to
            inv = 1.0 / _world()
            dw, dgamma, dbeta = dw * inv, dgamma * inv, dbeta * inv
        db = ops.zero_grad_const(wshape[0], x.device) if has_bias else None
        return None, dw.reshape(wshape), db, dgamma, dbeta, None, None, None, None, None


class BNActConv(torch.autograd.Function):
    """Pattern "NAC" of _ConvBlock as used by ResBlock2D (reference modules.py:116-130): norm + act on the block
    input, then the conv; ``residual`` (NHWC bf16) is added in the conv epilogue (the ``x +`` of modules.py:125).
    ``sums_in``: the batch-norm sums of x when the kernel that produced x already emitted them (replaces the fv_bn_stats pass);
    ``emit``: also return the sums of the output (for the next norm layer) -- second output, not differentiable."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, gamma, beta, running_mean, running_var, ksize, act, training, momentum, eps,
                sums_in=None, emit=False):
        co, ci = weight.shape[0], weight.shape[1]
        if _fin_fused(training):
            count = x.shape[0] * x.shape[1] * x.shape[2]
            a, stat = ops.bn_act_fwd_fin(x, ops.bn_stats(x) if sums_in is None else sums_in, count, gamma, beta, running_mean, running_var,
                                         MODE_NONE, act, torch.bfloat16, False, momentum, eps)
        else:
            stat, count = _bn_forward(x, gamma, beta, running_mean, running_var, training, momentum, eps, sums=sums_in if training else None)
            a = ops.bn_act_fwd(x, stat, MODE_NONE, act, torch.bfloat16)
        wf, wd = ops.weight_prep(weight, True, True)
        sums_out = None
        if emit:
            y, sums_out = ops.conv2d(a, wf, bias, co, ksize, residual, OUT_NHWC_BF16, (ci, co), want_stats=True)
            ctx.mark_non_differentiable(sums_out)
        else:
            y = ops.conv2d(a, wf, bias, co, ksize, residual, OUT_NHWC_BF16, (ci, co))
        ctx.save_for_backward(x, a, stat, weight, wd)
        ctx.cfg = (ksize, act, training, count, co, ci)
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        return y, sums_out

    @staticmethod
    def backward(ctx, g, _gsums=None):
        x, a, stat, weight, wd = ctx.saved_tensors
        ksize, act, training, count, co, ci = ctx.cfg
        g = g.contiguous()
        with ops.wgrad_stream(g, extra=True):       # the bias gradient is off the dependent chain as well
            db = ops.colsum(g)[:co].clone() if ctx.has_bias else None
        da, dw = _conv_backward(GEOM_SAME, a, g, weight, wd, ksize, True)
        c = x.shape[3]
        fused = _bn_backward_sums(x, da, stat, MODE_NONE, act, False, count, training)
        if fused is not None:
            dgamma, dbeta, coef = fused
            dx = ops.bn_act_bwd_apply(x, da, stat, coef, MODE_NONE, act) if ctx.needs_input_grad[0] else None
        else:
            s_local = ops.bn_act_bwd_reduce(x, da, stat, MODE_NONE, act)
            if _fin_fused(training):
                dx, dgamma, dbeta = ops.bn_act_bwd_apply_fin(x, da, stat, s_local, count, MODE_NONE, act)
            else:
                dgamma, dbeta, coef = _bn_backward_finalize(s_local, count, c, training)
                dx = ops.bn_act_bwd_apply(x, da, stat, coef, MODE_NONE, act) if ctx.needs_input_grad[0] else None
        dres = g if ctx.has_res else None
        return dx, dres, dw, db, dgamma, dbeta, None, None, None, None, None, None, None, None, None


class ConvOnly(torch.autograd.Function):
    """Plain nn.Conv2d (mid_conv, reference models.py:750 / 1096; out_conv, models.py:1099).  ``emit``: second output = the
    batch-norm sums of the output for a following "NAC" block (not differentiable)."""

    @staticmethod
    def forward(ctx, x, weight, bias, ksize, out_mode, emit=False):
        co, ci = weight.shape[0], weight.shape[1]
        sums_out = None
        if out_mode == OUT_NCHW_F32 and _outconv_fold_ok(x, weight):
            # full-resolution 7x7 output convolution (inference / eval path): tap-folded forward kernel; the backward below
            # prepares its own operands for the generic kernels when it is ever needed
            wq, _ = ops.outconv_prep(weight, True, False)
            y = ops.outconv_fwd(x, wq, bias, co)["logits"]
            wd = None
        else:
            wf, wd = ops.weight_prep(weight, True, x.requires_grad)
            if emit and out_mode == OUT_NHWC_BF16:
                y, sums_out = ops.conv2d(x, wf, bias, co, ksize, None, out_mode, (ci, co), want_stats=True)
                ctx.mark_non_differentiable(sums_out)
            else:
                y = ops.conv2d(x, wf, bias, co, ksize, None, out_mode, (ci, co))
        ctx.save_for_backward(x, weight, wd)
        ctx.cfg = (ksize, out_mode, co, ci)
        ctx.has_bias = bias is not None
        return y, sums_out

    @staticmethod
    def backward(ctx, g, _gsums=None):
        x, weight, wd = ctx.saved_tensors
        ksize, out_mode, co, ci = ctx.cfg
        if out_mode == OUT_NCHW_F32:
            g = ops.nchw_to_nhwc(g.contiguous().float(), pad_channels(co))
        elif g.dtype != torch.bfloat16:
            g = g.to(torch.bfloat16)
        g = g.contiguous()
        with ops.wgrad_stream(g, extra=True):
            db = ops.colsum(g)[:co].clone() if ctx.has_bias else None
        dx, dw = _conv_backward(GEOM_SAME, x, g, weight, wd, ksize, ctx.needs_input_grad[0])
        return dx, dw, db, None, None, None


# ---------------------------------------------------------------------------------------------------- VAE bottleneck / losses
class Reparam(torch.autograd.Function):
    """z = mu + exp(logstd) * eps (reference models.py:561).  mu/logstd: fp32 [N, Dz] row views."""

    @staticmethod
    def forward(ctx, mu, logstd, eps):
        z, _ = ops.reparam_kl_fwd(mu, logstd, eps, True, False)
        ctx.save_for_backward(mu, logstd, eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        mu, logstd, eps = ctx.saved_tensors
        dh = ops.reparam_kl_bwd(mu, logstd, eps, dz.contiguous(), None, None, 0.0, None)
        d = mu.shape[1]
        return dh[:, :d], dh[:, d:], None


class KLDivergence(torch.autograd.Function):
    """KLDivergenceLoss (reference losses.py:385-393): mean_n mean_d(-0.5 - ls + 0.5 mu^2 + 0.5 exp(2 ls))."""

    @staticmethod
    def forward(ctx, mu, logstd):
        _, rows = ops.reparam_kl_fwd(mu, logstd, None, False, True)
        ctx.save_for_backward(mu, logstd)
        n, d = mu.shape
        return rows.sum() / (n * d)

    @staticmethod
    def backward(ctx, gk):
        mu, logstd = ctx.saved_tensors
        n, d = mu.shape
        dh = ops.reparam_kl_bwd(mu, logstd, None, None, None, None, 1.0 / (n * d), gk.reshape(1).float().contiguous())
        return dh[:, :d], dh[:, d:]


class ReparamKL(torch.autograd.Function):
    """Fused bottleneck of the training path: h [N, 2*Dz] fp32 (mu | logstd) and eps -> (z [N, Dz], KL mean)."""

    @staticmethod
    def forward(ctx, h, eps):
        n, d2 = h.shape
        d = d2 // 2
        mu, ls = h[:, :d], h[:, d:]
        z, rows = ops.reparam_kl_fwd(mu, ls, eps, True, True)
        ctx.save_for_backward(h, eps)
        return z, rows.sum() / (n * d)

    @staticmethod
    def backward(ctx, dz, gk):
        h, eps = ctx.saved_tensors
        n, d2 = h.shape
        d = d2 // 2
        dh = ops.reparam_kl_bwd(h[:, :d], h[:, d:], eps, dz.contiguous(), None, None, 1.0 / (n * d),
                                gk.reshape(1).float().contiguous())
        return dh, None


class ReconLossFlat(torch.autograd.Function):
    """nn.MSELoss / nn.L1Loss mean over two same-shape fp32 tensors (reference losses.py:396-403, 128); the gradient
    is produced by the same pass that computes the loss."""

    @staticmethod
    def forward(ctx, a, b, l1):
        a, b = a.contiguous().float(), b.contiguous().float()
        e = a.numel()
        loss, grad = ops.recon_loss_flat(a, b, l1, 1.0 / e, True)
        ctx.save_for_backward(grad)
        ctx.shape = a.shape
        return loss[0] / e

    @staticmethod
    def backward(ctx, gl):
        (grad,) = ctx.saved_tensors
        n = grad.numel()
        flat = grad.view(-1)
        n8 = n // 8 * 8
        gptr = gl.reshape(1).float().contiguous()
        out = torch.empty_like(flat)
        if n8:
            ops.scale(flat[:n8], gptr, 1.0, out[:n8])
        if n8 < n:
            out[n8:] = flat[n8:] * gptr
        out = out.view(ctx.shape)
        return (out if ctx.needs_input_grad[0] else None), (-out if ctx.needs_input_grad[1] else None), None


def _outconv_fold_ok(d: torch.Tensor, weight: torch.Tensor) -> bool:
    import os
    if os.environ.get("FACEVAE_OUTCONV_FOLD", "1") == "0":
        return False
    n, h, w, cp = d.shape
    co, ci, k = weight.shape[0], weight.shape[1], weight.shape[2]
    return ci == cp and weight.shape[2] == weight.shape[3] and ops.outconv_supported(n, h, w, cp, co, k)


class ConvSigmoidRecon(torch.autograd.Function):
    """out_conv (7x7, reference models.py:1099) -> sigmoid (models.py:1110) -> ReconLoss against the target frame
    (losses.py:396-403, trainer.py:314), fused.  Full-resolution shapes (W = 128 / 256) run the tap-folded kernels of
    csrc/fv_outconv.cu: ONE forward kernel produces x_hat, the loss sum, the compact gradient d(loss)/d(logits) and the
    bias gradient; dgrad / wgrad consume the compact gradient and apply the upstream scalar in their epilogues.  Other
    shapes: generic conv + loss kernel (the loss kernel emits the NHWC bf16 gradient the generic dgrad / wgrad consume).
    Returns (x_hat NCHW fp32, mean loss)."""

    @staticmethod
    def forward(ctx, d, weight, bias, target, l1):
        co, ci, ksize = weight.shape[0], weight.shape[1], weight.shape[2]
        ctx.cfg = (ksize, co, ci)
        ctx.has_bias = bias is not None
        ctx.fold = _outconv_fold_ok(d, weight)
        if ctx.fold:
            wq, wdq = ops.outconv_prep(weight, True, d.requires_grad)
            e = d.shape[0] * co * d.shape[1] * d.shape[2]
            out = ops.outconv_fwd(d, wq, bias, co, target=target.contiguous().float(), l1=l1, gscale=1.0 / e)
            ctx.save_for_backward(d, weight, out["g4"], wdq, out["gsum"])
            ctx.mark_non_differentiable(out["pred"])
            return out["pred"], out["loss_sum"][0] / e
        wf, wd = ops.weight_prep(weight, True, d.requires_grad)
        logits = ops.conv2d(d, wf, bias, co, ksize, None, OUT_NCHW_F32, (ci, co))
        e = logits.numel()
        loss, pred, _, gn = ops.recon_loss(logits, target.contiguous().float(), l1, True, 1.0 / e, True, False, True)
        ctx.save_for_backward(d, weight, gn, wd)
        ctx.mark_non_differentiable(pred)
        return pred, loss[0] / e

    @staticmethod
    def backward(ctx, _gpred, gl):
        ksize, co, ci = ctx.cfg
        gl = gl.reshape(1).float().contiguous()
        if ctx.fold:
            d, weight, g4, wdq, gsum = ctx.saved_tensors
            db = gsum * gl if ctx.has_bias else None
            with ops.wgrad_stream(d, g4, gl):
                dw = ops.outconv_wgrad(d, g4, gl, co)
            dx = None
            if ctx.needs_input_grad[0]:
                if wdq is None:
                    _, wdq = ops.outconv_prep(weight, False, True)
                dx = ops.outconv_dgrad(g4, wdq, gl, co)
            return dx, dw, db, None, None
        d, weight, gn, wd = ctx.saved_tensors
        g = ops.scale(gn, gl, 1.0)
        db = ops.colsum(g)[:co].clone() if ctx.has_bias else None
        dx, dw = _conv_backward(GEOM_SAME, d, g, weight, wd, ksize, ctx.needs_input_grad[0])
        return dx, dw, db, None, None
