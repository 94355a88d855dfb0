"""Drop-in 2-D blocks with the reference's class names, constructor signatures and state_dict keys
(reference modules.py:8-135; SURVEY.md 3.4 for the key layout), backed by the sm_100a kernels.

Tensors cross the module boundary as logical NCHW: fp32 contiguous frames/latents are converted once by a CUDA
kernel; between blocks the tensors stay bf16 channels-last (a zero-copy view of the internal NHWC buffer), which is
what the blocks hand to each other.  In scope: the 2-D blocks of SURVEY.md section 8 -- batch-normalised stride-1 blocks
(rows a1-a6), their spectral-norm / instance-norm / 3x3 stride-2 variants (row f2) and the ELR layers (row f1); 3-D blocks
and any other geometry raise ``NotImplementedError``.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import functional as Fn
from . import ops
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU, MODE_NONE, MODE_POOL, MODE_UP, OUT_NHWC_BF16, pad_channels

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ---------------------------------------------------------------------------------------------------- tensor plumbing
def as_nhwc(x: torch.Tensor) -> torch.Tensor:
    """Logical NCHW tensor -> internal NHWC bf16 [N,H,W,pad_channels(C)] (zero-copy when it already is one)."""
    if x.dim() != 4:
        raise ValueError(f"expected a 4-d NCHW tensor, got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("face_vae_b200 has no CPU path: move the input to a CUDA device")
    c = x.shape[1]
    if x.dtype == torch.bfloat16 and c == pad_channels(c):
        v = x.permute(0, 2, 3, 1)
        if v.is_contiguous():
            return v
    return Fn.ToNHWC.apply(x)


def as_nchw(y: torch.Tensor, c: int) -> torch.Tensor:
    """Internal NHWC -> logical NCHW view (bf16, channels-last memory)."""
    v = y.permute(0, 3, 1, 2)
    return v if c == y.shape[3] else v[:, :c]


def to_float_nchw(x: torch.Tensor) -> torch.Tensor:
    """Logical NCHW (any layout/dtype produced by these blocks) -> contiguous NCHW fp32 via the layout kernel."""
    if x.dtype == torch.float32 and x.is_contiguous():
        return x
    return Fn.ToNCHW.apply(as_nhwc(x), x.shape[1])


# ---------------------------------------------------------------------------------------------------- parameter holders
class _BatchNormParams(nn.Module):
    """Parameter/buffer holder with nn.SyncBatchNorm's state_dict keys (weight, bias, running_mean, running_var,
    num_batches_tracked)."""

    def __init__(self, c: int):
        super().__init__()
        self.num_features = c
        self.eps = BN_EPS
        self.momentum = BN_MOMENTUM
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _Conv2dParams(nn.Module):
    """Parameter holder with nn.Conv2d's keys and default initialisation (kaiming-uniform a=sqrt(5), bias +-1/sqrt(fan_in))."""

    def __init__(self, ci: int, co: int, k: int, stride: int, padding: int):
        super().__init__()
        _check_conv_geometry(k, stride, padding)
        self.in_channels, self.out_channels, self.kernel_size, self.stride = ci, co, k, stride
        self.weight = nn.Parameter(torch.empty(co, ci, k, k))
        self.bias = nn.Parameter(torch.empty(co))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(ci * k * k)
        nn.init.uniform_(self.bias, -bound, bound)

    def effective_weight(self) -> torch.Tensor:
        return self.weight

    def forward(self, x):                      # plain conv on a logical NCHW tensor
        y = _conv_nhwc(as_nhwc(x), self.weight, self.bias, self.kernel_size, self.stride)
        return as_nchw(y, self.out_channels)


def _check_conv_geometry(k: int, stride: int, padding: int) -> None:
    same = stride == 1 and k % 2 == 1 and padding == (k - 1) // 2
    strided = stride == 2 and k == 3 and padding == 1          # the Discriminator's down-sampling blocks (reference models.py:1120-1123)
    if not (same or strided):
        raise NotImplementedError(f"conv kernel {k} stride {stride} padding {padding}: supported are stride-1 'same' convolutions with odd "
                                  "kernels and 3x3 stride-2 padding-1 (4x4 stride 2 lives in Conv2dELR)")


def _conv_nhwc(x, weight, bias, k: int, stride: int):
    """Plain convolution on an NHWC bf16 tensor.  3x3 stride 2 pad 1 is the 4x4 stride-2 pad-1 kernel with a zero fourth row and
    column of taps (output pixel i reads input rows 2i-1 .. 2i+1); the zero padding is a torch op on the (tiny) weight, so its
    gradient is sliced back by autograd."""
    if stride == 2:
        w4 = torch.nn.functional.pad(weight, (0, 1, 0, 1))
        return Fn.ConvELRAct.apply(x, w4.contiguous(), bias, 4, 2, 1.0, False, ACT_NONE)
    return Fn.ConvOnly.apply(x, weight, bias, k, OUT_NHWC_BF16)[0]


class _SpectralConv2dParams(nn.Module):
    """``torch.nn.utils.spectral_norm(nn.Conv2d(...))`` as the reference's ``_ConvBlock`` builds it with ``use_weight_norm=True``
    (reference modules.py:11,14,32): state_dict keys ``weight_orig``, ``bias``, ``weight_u``, ``weight_v``; one power iteration per
    training-mode forward (u, v updated in place, no gradient through them), W = W_orig / (u^T W_orig v).  The power iteration and
    the division are a handful of matrix-vector products on the [Co, Ci*k*k] filter -- library calls on the weight, whose result
    feeds the same convolution kernels; autograd carries the gradient of W back to ``weight_orig``."""

    def __init__(self, ci: int, co: int, k: int, stride: int, padding: int, n_power_iterations: int = 1, eps: float = 1e-12):
        super().__init__()
        _check_conv_geometry(k, stride, padding)
        self.in_channels, self.out_channels, self.kernel_size, self.stride = ci, co, k, stride
        self.n_power_iterations, self.eps = n_power_iterations, eps
        self.weight_orig = nn.Parameter(torch.empty(co, ci, k, k))
        self.bias = nn.Parameter(torch.empty(co))
        nn.init.kaiming_uniform_(self.weight_orig, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(ci * k * k)
        nn.init.uniform_(self.bias, -bound, bound)
        self.register_buffer("weight_u", torch.nn.functional.normalize(torch.randn(co), dim=0, eps=eps))
        self.register_buffer("weight_v", torch.nn.functional.normalize(torch.randn(ci * k * k), dim=0, eps=eps))
        self.prep_kind = -1            # the filter operand depends on sigma: prepared per call

    def effective_weight(self) -> torch.Tensor:
        w = self.weight_orig
        mat = w.reshape(w.shape[0], -1)
        u, v = self.weight_u, self.weight_v
        if self.training:
            with torch.no_grad():
                for _ in range(self.n_power_iterations):
                    v = torch.nn.functional.normalize(torch.mv(mat.t(), u), dim=0, eps=self.eps, out=v)
                    u = torch.nn.functional.normalize(torch.mv(mat, v), dim=0, eps=self.eps, out=u)
                u, v = u.clone(), v.clone()
        sigma = torch.dot(u, torch.mv(mat, v))
        return w / sigma

    def forward(self, x):
        y = _conv_nhwc(as_nhwc(x), self.effective_weight().contiguous(), self.bias, self.kernel_size, self.stride)
        return as_nchw(y, self.out_channels)


class _InstanceNormParams(nn.Module):
    """Parameter holder with nn.InstanceNorm2d(C, affine=True)'s state_dict keys (weight, bias; no running statistics)."""

    def __init__(self, c: int):
        super().__init__()
        self.num_features = c
        self.eps = BN_EPS
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class _Act(nn.Module):
    def __init__(self, code):
        super().__init__()
        self.code = code


class Conv2d(_Conv2dParams):
    """nn.Conv2d stand-in for the un-normalised convs of the path (mid_conv, reference models.py:750,1096)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding)


# ---------------------------------------------------------------------------------------------------- blocks
class ConvBlock2D(nn.Module):
    """_ConvBlock / ConvBlock2D (reference modules.py:8-49): ``pattern`` in {"CNA", "NAC", "CN"}; ``activation_type`` "batch"
    (SyncBatchNorm), "instance" (InstanceNorm2d, affine) or "none"; ``use_weight_norm`` wraps the conv in spectral norm
    (modules.py:11,14).  The hot path of SURVEY.md section 8 is batch norm without weight norm at stride 1; the spectral-norm /
    instance-norm / 3x3 stride-2 variants are the Generator's and Discriminator's blocks (section 8f rank 2)."""

    def __init__(self, pattern, in_channels, out_channels, kernel_size, stride, padding, use_weight_norm,
                 activation_type="batch", nonlinearity_type="relu"):
        super().__init__()
        if activation_type not in ("batch", "instance", "none"):
            raise NotImplementedError(f"activation_type {activation_type!r}: the reference uses batch, instance and none")
        if pattern not in ("CNA", "NAC", "CN"):
            raise NotImplementedError(f"pattern {pattern!r}: the reference uses CNA, NAC and CN only")
        if pattern == "NAC" and (activation_type != "batch" or stride != 1):
            raise NotImplementedError("NAC blocks (ResBlock2D) are batch-normalised stride-1 blocks in the reference")
        if activation_type == "batch" and stride != 1:
            raise NotImplementedError("strided blocks are instance-normalised in the reference (Discriminator)")
        norm_channels = out_channels if pattern.find("C") < pattern.find("N") else in_channels
        if activation_type != "none" and norm_channels != pad_channels(norm_channels):
            raise NotImplementedError("normalised channel counts must be 16, 32 or a multiple of 64")
        self.pattern = pattern
        self.activation_type = activation_type
        self.act = ACT_NONE if "A" not in pattern else (ACT_RELU if nonlinearity_type == "relu" else ACT_LEAKY)
        conv_cls = _SpectralConv2dParams if use_weight_norm else _Conv2dParams
        norm = {"batch": _BatchNormParams, "instance": _InstanceNormParams}.get(activation_type)
        mods = {"C": conv_cls(in_channels, out_channels, kernel_size, stride, padding),
                "N": norm(norm_channels) if norm is not None else nn.Identity(), "A": _Act(self.act)}
        self.layers = nn.Sequential(*[mods[ch] for ch in pattern])      # same indices as the reference => same keys
        self.in_channels, self.out_channels, self.kernel_size, self.stride = in_channels, out_channels, kernel_size, stride

    @property
    def conv(self):
        return self.layers[self.pattern.index("C")]

    @property
    def norm(self):
        return self.layers[self.pattern.index("N")]

    def forward_nhwc(self, x, post_mode=MODE_NONE, residual=None, out_nchw_f32=False, upsample_input=False):
        """``upsample_input``: the conv reads the nearest-2x up-sampled x (UpBlock2D) -- computed on the coarse grid by the
        phase-decomposed kernel, the up-sampled tensor is never written."""
        conv, bn = self.conv, self.norm
        weight = conv.effective_weight()
        if self.activation_type != "batch":
            if post_mode != MODE_NONE or out_nchw_f32 or upsample_input or residual is not None:
                raise NotImplementedError("instance-normalised / un-normalised blocks have no fused pool, up-sampling or residual")
            y = _conv_nhwc(x, weight.contiguous(), conv.bias, self.kernel_size, self.stride)
            if self.activation_type == "instance":
                return Fn.InstanceNormAct.apply(y, bn.weight, bn.bias, self.act, bn.eps)
            return Fn.ActOnly.apply(y, self.act) if self.act != ACT_NONE else y
        if self.training:
            ops.bump_counter(bn.num_batches_tracked)
        if self.pattern in ("CNA", "CN"):
            geom = Fn.GEOM_UP if upsample_input else Fn.GEOM_SAME
            if upsample_input and self.kernel_size != 3:
                raise NotImplementedError("the fused up-sampling convolution is 3x3 (UpBlock2D)")
            return Fn.ConvBNAct.apply(x, weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                      self.kernel_size, post_mode, self.act, self.training, out_nchw_f32, bn.momentum, bn.eps, geom)
        if upsample_input:
            raise NotImplementedError("NAC blocks have no fused up-sampling")
        if post_mode != MODE_NONE or out_nchw_f32:
            raise NotImplementedError("NAC blocks have no fused pool/upsample")
        # statistics of the input come from the kernel that produced it when that kernel emitted them (ops.attach_stats); this
        # block emits those of its output when the parent wired a norm layer behind it (ResBlock2D: emit_stats)
        sums_in = ops.attached_stats(x) if self.training else None
        y, sums = Fn.BNActConv.apply(x, residual, weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                     self.kernel_size, self.act, self.training, bn.momentum, bn.eps, sums_in,
                                     bool(getattr(self, "emit_stats", False)) and self.training)
        return ops.attach_stats(y, sums)

    def forward(self, x):
        return as_nchw(self.forward_nhwc(as_nhwc(x)), self.out_channels)


class _Pool(nn.Module):
    """Marker for nn.AvgPool2d((2, 2)) -- executed inside the block's fused norm+act pass."""


class _Up(nn.Module):
    """Marker for nn.Upsample(scale_factor=(2, 2)) (nearest)."""


class DownBlock2D(nn.Module):
    """reference modules.py:59-70: CNA 3x3 s1 p1 -> AvgPool2d(2); the pool is fused into the norm+ReLU kernel."""

    def __init__(self, in_channels, out_channels, use_weight_norm):
        super().__init__()
        self.layers = nn.Sequential(ConvBlock2D("CNA", in_channels, out_channels, 3, 1, 1, use_weight_norm), _Pool())
        self.out_channels = out_channels

    def forward_nhwc(self, x, out_nchw_f32=False):
        return self.layers[0].forward_nhwc(x, MODE_POOL, None, out_nchw_f32)

    def forward(self, x):
        return as_nchw(self.forward_nhwc(as_nhwc(x)), self.out_channels)


class UpBlock2D(nn.Module):
    """reference modules.py:78-89: Upsample(x2, nearest) -> CNA 3x3 s1 p1.  The up-sampling is folded into the convolution:
    every output pixel (2i + a, 2j + b) sees a 2x2 neighbourhood of the coarse input, so the layer runs as four 2x2 phase
    convolutions on the coarse grid (csrc/fv_conv.cu, X2 geometry) and the 4x larger tensor is never written or re-read;
    backward likewise (data gradient = a 4x4 stride-2 convolution of dY, which contains the 2x2 sum of the up-sampling
    backward).  ``pre_upsampled=True`` takes an input that already is at the output resolution (plain CNA block)."""

    def __init__(self, in_channels, out_channels, use_weight_norm):
        super().__init__()
        self.layers = nn.Sequential(_Up(), ConvBlock2D("CNA", in_channels, out_channels, 3, 1, 1, use_weight_norm))
        if not use_weight_norm:
            self.layers[1].conv.prep_kind = ops.PREP_UP    # ops.step_scope prepares the phase filters for this weight
        self.out_channels = out_channels

    def forward_nhwc(self, x, pre_upsampled=False, post_mode=MODE_NONE):
        return self.layers[1].forward_nhwc(x, post_mode, upsample_input=not pre_upsampled)

    def forward(self, x):
        return as_nchw(self.forward_nhwc(as_nhwc(x)), self.out_channels)


class SameBlock2D(nn.Module):
    """reference modules.py:97-108: CNA 1x1."""

    def __init__(self, in_channels, out_channels, use_weight_norm):
        super().__init__()
        self.layers = ConvBlock2D("CNA", in_channels, out_channels, 1, 1, 0, use_weight_norm)
        self.out_channels = out_channels

    def forward_nhwc(self, x, post_mode=MODE_NONE):
        return self.layers.forward_nhwc(x, post_mode)

    def pointwise_ok(self, x) -> bool:
        """Raw NCHW fp32 frames with <= 4 channels into 32, training mode, no gradient wanted for the frames: the block is
        a per-pixel affine map with statistics from the input moments (functional.PointwiseBNAct)."""
        blk = self.layers
        return (self.training and x.dim() == 4 and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
                and not x.requires_grad and x.shape[1] <= 4 and blk.in_channels == x.shape[1] and blk.out_channels == 32
                and blk.pattern == "CNA" and blk.activation_type == "batch" and isinstance(blk.conv, _Conv2dParams))

    def forward_from_frames(self, x):
        """x: NCHW fp32 frames -> NHWC bf16 [N,H,W,32]."""
        blk = self.layers
        conv, bn = blk.conv, blk.norm
        ops.bump_counter(bn.num_batches_tracked)
        return Fn.PointwiseBNAct.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, blk.act,
                                       bn.momentum, bn.eps)

    def forward(self, x):
        if self.pointwise_ok(x):
            return as_nchw(self.forward_from_frames(x), self.out_channels)
        return as_nchw(self.forward_nhwc(as_nhwc(x)), self.out_channels)


class ResBlock2D(nn.Module):
    """reference modules.py:116-130: x + NAC(NAC(x)); the residual add lives in the second conv's epilogue."""

    def __init__(self, in_channels, use_weight_norm):
        super().__init__()
        self.layers = nn.Sequential(
            ConvBlock2D("NAC", in_channels, in_channels, 3, 1, 1, use_weight_norm),
            ConvBlock2D("NAC", in_channels, in_channels, 3, 1, 1, use_weight_norm),
        )
        self.out_channels = in_channels
        self.layers[0].emit_stats = True       # its output is normalised by layers[1]: the conv epilogue emits the batch sums
        # layers[1].emit_stats is set by a parent that puts another ResBlock2D behind this one (chain_res_blocks)

    def forward_nhwc(self, x):
        h = self.layers[0].forward_nhwc(x)
        return self.layers[1].forward_nhwc(h, residual=x)

    def forward(self, x):
        return as_nchw(self.forward_nhwc(as_nhwc(x)), self.out_channels)


def chain_res_blocks(blocks) -> None:
    """Tell every ResBlock2D that is followed by another one to emit the batch-norm sums of its output (the next block normalises
    its input): saves one statistics pass per block."""
    blocks = list(blocks)
    for a, b in zip(blocks[:-1], blocks[1:]):
        if isinstance(a, ResBlock2D) and isinstance(b, ResBlock2D):
            a.layers[1].emit_stats = True


# ---------------------------------------------------------------------------------------------------- ELR layers
def _act_gain(act) -> float:
    """The gain Conv2dELR / LinearELR derive from their activation module (reference models_utils.py:137-146, 648-657)."""
    try:
        if isinstance(act, nn.LeakyReLU):
            return nn.init.calculate_gain("leaky_relu", act.negative_slope)
        if isinstance(act, nn.ReLU):
            return nn.init.calculate_gain("relu")
        return nn.init.calculate_gain(act)
    except Exception:
        return 1.0


def _act_code(act) -> int:
    if act is None:
        return ACT_NONE
    if isinstance(act, nn.ReLU):
        return ACT_RELU
    if isinstance(act, nn.LeakyReLU) and abs(act.negative_slope - 0.2) < 1e-12:
        return ACT_LEAKY
    raise NotImplementedError(f"activation {act!r}: the fused epilogue knows ReLU and LeakyReLU(0.2) (what the reference uses)")


class Conv2dELR(nn.Module):
    """reference models_utils.py:632-744 -- equalised-learning-rate convolution with optional weight demodulation, as used by
    ``EFE_conv6.efe_encoder`` (models.py:845-852): ``conv(ci, co, 4, 2, 1, norm="demod", act=nn.LeakyReLU(0.2))``.
    Same constructor, parameter names (``weight`` ~ N(0,1), ``bias`` zeros) and ``weightgain`` as the reference.

    In scope: 4x4 stride-2 pad-1 (even input sizes, output width a power of two or a multiple of 128) and stride-1 "same"
    1x1 / 3x3, ``norm`` in {None, "demod"}, ``act`` in {None, ReLU, LeakyReLU(0.2)}.  Out of scope (raise): style modulation
    (``wsize > 0`` / a ``w`` argument), untied biases (``ub``), and the reference's first encoder layer ``conv(3, 32, 1, 1, 1)``
    -- a 1x1 kernel with padding 1, which grows the image to 66x66 and makes every later size odd."""

    def __init__(self, inch, outch, kernel_size, stride, padding, wsize=0, affinelrmult=1., norm=None, ub=None, act=None):
        super().__init__()
        if wsize > 0 or ub is not None:
            raise NotImplementedError("Conv2dELR: style modulation (wsize) and untied biases (ub) are outside the hot path")
        if norm not in (None, "demod"):
            raise NotImplementedError(f"Conv2dELR: norm={norm!r}")
        if (kernel_size, stride, padding) not in ((4, 2, 1), (1, 1, 0), (3, 1, 1)):
            raise NotImplementedError(f"Conv2dELR: kernel {kernel_size} stride {stride} padding {padding}: supported are 4/2/1, 1/1/0, 3/1/1")
        self.inch, self.outch, self.kernel_size, self.stride, self.padding = inch, outch, kernel_size, stride, padding
        self.wsize, self.norm, self.ub, self.act = wsize, norm, ub, act
        self.act_code = _act_code(act)
        fan_in = inch * (kernel_size ** 2)
        initgain = 1.0 if norm == "demod" else 1.0 / math.sqrt(fan_in)
        self.weightgain = _act_gain(act) * initgain
        self.weight = nn.Parameter(torch.randn(outch, inch, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.zeros(outch))
        self.affine = None
        self.fused = False
        self.prep_kind = -1            # the operands depend on the demodulated weight: prepared per call, not by ops.step_scope

    def extra_repr(self):
        return 'inch={}, outch={}, kernel_size={}, stride={}, padding={}, wsize={}, norm={}, ub={}, act={}'.format(
            self.inch, self.outch, self.kernel_size, self.stride, self.padding, self.wsize, self.norm, self.ub, self.act)

    def fuse(self):
        """Bake normalisation and gain into the weight (reference models_utils.py:700-704)."""
        with torch.no_grad():
            w, _ = ops.demod_fwd(self.weight.data.contiguous(), self.weightgain, self.norm == "demod")
            self.weight.data = w
        self.fused = True

    def forward_nhwc(self, x):
        demod = (self.norm == "demod") and not self.fused
        gain = 1.0 if self.fused else self.weightgain
        return Fn.ConvELRAct.apply(x, self.weight, self.bias, self.kernel_size, self.stride, gain, demod, self.act_code)

    def forward(self, x, w: Optional[torch.Tensor] = None):
        if w is not None:
            raise NotImplementedError("Conv2dELR: style modulation is outside the hot path")
        return as_nchw(self.forward_nhwc(as_nhwc(x)), self.outch)


class LinearELR(nn.Module):
    """reference models_utils.py:134-203.  The bottleneck's fully connected layers are [N, 256] x [256, 256] library GEMMs
    ("leave on torch.addmm"); kept as plain torch ops on the CUDA device with the reference's exact
    formulas, constructor and parameter names."""

    def __init__(self, inch, outch, lrmult=1., norm: Optional[str] = None, act=None):
        super().__init__()
        initgain = 1.0 / math.sqrt(inch)
        self.weight = nn.Parameter(torch.randn(outch, inch) / lrmult)
        self.weightgain = _act_gain(act)
        if norm is None:
            self.weightgain = self.weightgain * initgain * lrmult
        self.bias = nn.Parameter(torch.full([outch], 0.))
        self.norm = norm
        self.act = act
        self.fused = False

    def getweight(self):
        if self.fused or self.norm != "demod":
            return self.weight
        return torch.nn.functional.normalize(self.weight, dim=1)

    def fuse(self):
        if not self.fused:
            with torch.no_grad():
                self.weight.data = self.getweight() * self.weightgain
        self.fused = True

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("face_vae_b200 has no CPU path: move the input to a CUDA device")
        weight = self.getweight()
        if self.fused:
            out = torch.addmm(self.bias[None], x, weight.t())
            return self.act(out) if self.act is not None else out
        if self.act is None:
            return torch.addmm(self.bias[None], x, weight.t(), alpha=self.weightgain)
        return self.act(torch.nn.functional.linear(x, weight * self.weightgain, bias=self.bias))
