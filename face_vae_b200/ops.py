"""Tensor-level wrappers over the C ABI (include/facevae_b200.h).

PyTorch is used for device memory and streams only: every function allocates its outputs with torch, passes raw
device pointers plus the current CUDA stream to libfacevae_b200.so and returns immediately (asynchronous).
Activations are explicit NHWC tensors ``[N, H, W, C]`` (bf16 unless stated), C padded with ``pad_channels``.
"""
from __future__ import annotations

import math
import os as _os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, DT_BF16, DT_F32, MODE_NONE, MODE_POOL, MODE_UP, OUT_NCHW_F32,
                   OUT_NHWC_BF16, OUT_NHWC_F32, call)

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def pad_channels(c: int) -> int:
    """Channel padding accepted by the tcgen05 kernels: 16, 32 or a multiple of 64."""
    if c <= 16:
        return 16
    if c <= 32:
        return 32
    return (c + 63) // 64 * 64


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return DT_BF16
    if t.dtype == torch.float32:
        return DT_F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _chk(t: torch.Tensor, name: str, dtype=None):
    if not t.is_cuda:
        raise _lib.FaceVaeError(f"{name}: expected a CUDA tensor (this package has no CPU path)")
    if not t.is_contiguous():
        raise _lib.FaceVaeError(f"{name}: expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise _lib.FaceVaeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t


# ------------------------------------------------------------------------------------------------ reduction workspace
_red_ws_cache = {}


def _red_ws() -> int:
    """Device pointer of the reduction scratch for the CURRENT stream (include/facevae_b200.h, `red_ws`): one buffer per
    (device, stream), zero-initialised once -- the kernels leave its ticket words zero again.  Kernels on one stream run in
    order, so they can share it; another stream gets its own."""
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    buf = _red_ws_cache.get(key)
    if buf is None:
        buf = torch.zeros((int(_lib.load().fv_reduce_ws_bytes()),), dtype=torch.uint8, device="cuda")
        _red_ws_cache[key] = buf
    return buf.data_ptr()


# ------------------------------------------------------------------------------------------------ weight-gradient side stream
# Inside a step scope the weight-gradient kernels (tensor bound, nothing downstream of them until the optimiser) can be issued
# on a second stream: they then run beside the HBM-bound norm / activation passes of the layers further down the backward
# chain instead of in front of them.  Fork: the side stream waits for an event recorded on the compute stream right after dY
# was produced; join: the compute stream waits for the side stream when the scope closes (and before a gradient reducer
# touches the gradients).  Operands of a side-stream kernel are kept referenced until the join, so the caching allocator
# cannot hand their memory to a later compute-stream kernel while the side stream still reads it.  Under CUDA-graph capture
# fork and join become graph edges.  Results do not depend on the schedule (no atomics): bitwise equal to the serial order.
_WGRAD_STREAM = _os.environ.get("FACEVAE_WGRAD_STREAM", "1") != "0"
# "2" (default): the bias column sums of the un-normalised convolutions and the per-step filter preparation use the side stream as well
_SIDE_EXTRAS = _os.environ.get("FACEVAE_WGRAD_STREAM", "2") == "2"
_wgrad_side = {}           # (device, compute stream) -> side stream
_wgrad_live = []           # tensors a pending side-stream kernel reads
_wgrad_dirty = {}          # (device, compute stream) -> (compute stream, side stream) with un-joined work


def set_wgrad_stream(enabled: bool) -> bool:
    """Switch the weight-gradient side stream on / off; -> the previous setting.  (bench.py times kernels one by one with it off.)"""
    global _WGRAD_STREAM
    prev, _WGRAD_STREAM = _WGRAD_STREAM, bool(enabled)
    return prev


class wgrad_stream:
    """``with ops.wgrad_stream(x, dy):`` -- the kernels issued inside run on the weight-gradient side stream (when enabled and
    inside a step scope; otherwise on the current stream as usual).  ``tensors``: what those kernels read."""

    def __init__(self, *tensors, extra: bool = False):
        self.tensors = [t for t in tensors if t is not None]
        self.ctx = None
        self.extra = extra

    def __enter__(self):
        if not (_WGRAD_STREAM and _scope_active) or (self.extra and not _SIDE_EXTRAS):
            return self
        main = torch.cuda.current_stream()
        key = (torch.cuda.current_device(), main.cuda_stream)
        side = _wgrad_side.get(key)
        if side is None:
            side = _wgrad_side[key] = torch.cuda.Stream()
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        _wgrad_live.extend(self.tensors)
        _wgrad_dirty.setdefault(key, (main, side))
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
            self.ctx = None
        return False


def join_wgrad_stream() -> None:
    """The compute stream(s) wait for the pending weight-gradient kernels; their operands may be released afterwards."""
    while _wgrad_dirty:
        _, (main, side) = _wgrad_dirty.popitem()
        main.wait_stream(side)
    _wgrad_live.clear()


# ------------------------------------------------------------------------------------------------ per-step scope
PREP_PLAIN, PREP_UP, PREP_S2 = 0, 1, 2      # fv_prep_desc.kind


class _PrepCache:
    """bf16 filter operands of every convolution of a model, produced by one batched launch per step."""

    def __init__(self):
        self.key = None
        self.map = {}          # plain convolutions: weight ptr -> (wf, wd)
        self.map_up = {}       # up-sampling 3x3 convolutions: weight ptr -> (wx2, ws2)
        self.map_s2 = {}       # 4x4 stride-2 convolutions: weight ptr -> (wf, wx2)
        self.table = None
        self.max_items = 0
        self.valid = False
        self.joined = True     # False while this step's batched launch may still be running on the side stream

    def build(self, weights):
        import numpy as np
        dev = weights[0][0].device
        rec = np.zeros((len(weights),), dtype=np.dtype([("w", "<u8"), ("o0", "<u8"), ("o1", "<u8"), ("dims", "<i4", (6,)),
                                                        ("kind", "<i4"), ("reserved", "<i4")]))
        self.map, self.map_up, self.map_s2 = {}, {}, {}
        self.max_items = 0
        lib = _lib.load()
        self.total_blocks = 0
        for i, (w, kind) in enumerate(weights):
            co, ci, r, s_ = w.shape
            cop, cip = pad_channels(co), pad_channels(ci)
            if kind == PREP_UP:
                o0 = torch.empty((4, cop, 4, cip), device=dev, dtype=torch.bfloat16)
                o1 = torch.empty((cip, 16, cop), device=dev, dtype=torch.bfloat16)
                self.map_up[w.data_ptr()] = (o0, o1)
                items = 16 * cop * cip
            elif kind == PREP_S2:
                o0 = torch.empty((cop, 16, cip), device=dev, dtype=torch.bfloat16)
                o1 = torch.empty((4, cip, 4, cop), device=dev, dtype=torch.bfloat16)
                self.map_s2[w.data_ptr()] = (o0, o1)
                items = 16 * cop * cip
            else:
                o0 = torch.empty((cop, r * s_, cip), device=dev, dtype=torch.bfloat16)
                o1 = torch.empty((cip, r * s_, cop), device=dev, dtype=torch.bfloat16)
                self.map[w.data_ptr()] = (o0, o1)
                items = cop * cip * r * s_
            # `reserved` = first block of this layer in the grid of fv_weight_prep_tiled
            rec[i] = (w.data_ptr(), o0.data_ptr(), o1.data_ptr(), (co, ci, r, s_, cop, cip), kind, self.total_blocks)
            self.total_blocks += int(lib.fv_weight_prep_tiled_blocks(int(kind), cop, cip, r, s_))
            self.max_items = max(self.max_items, items)
        self.table = torch.from_numpy(rec.view(np.uint8).copy()).to(dev)
        self.key = tuple((w.data_ptr(), kind) for w, kind in weights)

    def run(self, weights):
        key = tuple((w.data_ptr(), kind) for w, kind in weights)
        if key != self.key:
            self.build(weights)
        call("fv_weight_prep_tiled", self.table.data_ptr(), len(weights), self.total_blocks, _stream())
        self.valid = True


_prep = _PrepCache()
_scope_active = False
_pending_counters = []


def bump_counter(t: torch.Tensor) -> None:
    """num_batches_tracked += 1: inside a step scope the increments of all layers are issued as ONE multi-tensor launch."""
    if _scope_active:
        _pending_counters.append(t)
    else:
        t += 1


# ---- batch-norm statistics handed from the convolution that PRODUCES a tensor to the norm layer that consumes it ------------------
# A "NAC" block (ResBlock2D) normalises its INPUT: the statistics pass over that input (fv_bn_stats, a separate launch per norm
# layer) is a sum the producing convolution's epilogue can emit.  The producer's module attaches them to the tensor OBJECT it
# returns (no global table: the hint lives and dies with that object, and an in-place change of the tensor invalidates it).
def attach_stats(y: torch.Tensor, sums: Optional[torch.Tensor]) -> torch.Tensor:
    if sums is not None:
        y._fv_sums = (y._version, tuple(y.shape), sums)
    return y


def attached_stats(x: torch.Tensor) -> Optional[torch.Tensor]:
    hit = getattr(x, "_fv_sums", None)
    if hit is None or hit[0] != x._version or hit[1] != tuple(x.shape):
        return None
    return hit[2]


_zero_consts = {}


def zero_grad_const(n: int, device) -> torch.Tensor:
    """fp32 zeros [n] for a gradient that is analytically zero (the bias of a conv that feeds a batch norm).  Inside a step_scope
    (the trainer's step) this is ONE shared, never-written tensor per (device, n) instead of a fill kernel per layer and step --
    the optimiser only reads gradients; outside the scope the caller owns a fresh tensor."""
    if not _scope_active:
        return torch.zeros((n,), device=device, dtype=torch.float32)
    key = (str(device), int(n))
    z = _zero_consts.get(key)
    if z is None:
        z = _zero_consts[key] = torch.zeros((n,), device=device, dtype=torch.float32)
    return z


class step_scope:
    """``with ops.step_scope(model):`` around forward + backward of one train step (VAETrainer does this): the bf16 filter
    operands of all convolutions come from ONE batched launch (a device table of layer descriptors; modules tag their
    4-d weights with ``prep_kind``: plain, up-sampling 3x3, 4x4 stride-2) and the batch counters are bumped by one
    multi-tensor launch.  Outside the scope every op prepares for itself.  (Round 1 also zeroed an arena of accumulators
    here; every reduction now WRITES its result -- fv_reduce.cuh -- so nothing needs zeroing.)"""

    def __init__(self, module: Optional[torch.nn.Module] = None):
        self.module = module

    def __enter__(self):
        global _scope_active
        weights = []
        if self.module is not None:
            for m in self.module.modules():
                w = getattr(m, "weight", None)
                if isinstance(w, torch.nn.Parameter) and w.dim() == 4 and w.is_cuda and w.dtype == torch.float32 and w.is_contiguous():
                    kind = int(getattr(m, "prep_kind", PREP_PLAIN))
                    if kind < 0:                 # the module prepares its own operands (tap-folded out_conv)
                        continue
                    if kind == PREP_UP and tuple(w.shape[2:]) != (3, 3):
                        kind = PREP_PLAIN
                    if kind == PREP_S2 and tuple(w.shape[2:]) != (4, 4):
                        continue
                    weights.append((w, kind))
        _scope_active = True
        try:
            if weights:
                # the filter preparation only depends on the weights: it runs on the side stream beside the first encoder layer
                # (which reads the fp32 frames and its fp32 1x1 filter directly) and is joined by the first convolution that asks
                # for an operand
                with wgrad_stream(extra=True) as ws:
                    _prep.run(weights)
                    _prep.joined = ws.ctx is None
        except BaseException:           # __exit__ does not run when __enter__ raises: leave no scope behind
            join_wgrad_stream()
            _scope_active = False
            _prep.valid = False
            raise
        return self

    def __exit__(self, *exc):
        global _scope_active
        join_wgrad_stream()
        _scope_active = False
        _prep.valid = False
        if _pending_counters:
            torch._foreach_add_(list(_pending_counters), 1)
            _pending_counters.clear()
        return False


# ------------------------------------------------------------------------------------------------ layout
def nchw_to_nhwc(x: torch.Tensor, cp: Optional[int] = None, dtype=torch.bfloat16) -> torch.Tensor:
    _chk(x, "x", torch.float32)
    n, c, h, w = x.shape
    cp = pad_channels(c) if cp is None else cp
    out = torch.empty((n, h, w, cp), device=x.device, dtype=dtype)
    call("fv_nchw_to_nhwc", x.data_ptr(), out.data_ptr(), _dt(out), n, c, h, w, cp, _stream())
    return out


def nhwc_to_nchw(x: torch.Tensor, c: Optional[int] = None, out: Optional[torch.Tensor] = None,
                 accumulate: bool = False) -> torch.Tensor:
    _chk(x, "x")
    n, h, w, cs = x.shape
    c = cs if c is None else c
    if out is None:
        out = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
        accumulate = False
    call("fv_nhwc_to_nchw", x.data_ptr(), _dt(x), _chk(out, "out", torch.float32).data_ptr(), n, c, h, w, cs,
         int(accumulate), _stream())
    return out


def bilinear_resize(x: torch.Tensor, scale_factor: float) -> torch.Tensor:
    """F.interpolate(x, mode="bilinear", scale_factor=s, align_corners=False, recompute_scale_factor=True) on NCHW fp32 frames
    (reference models.py:764): the input pre-scale in front of the encoder.  No gradient (frames are data)."""
    _chk(x, "x", torch.float32)
    n, c, h, w = x.shape
    ho, wo = int(math.floor(h * scale_factor)), int(math.floor(w * scale_factor))
    if ho < 1 or wo < 1:
        raise _lib.FaceVaeError("bilinear_resize: empty output")
    out = torch.empty((n, c, ho, wo), device=x.device, dtype=torch.float32)
    call("fv_bilinear_resize", x.data_ptr(), out.data_ptr(), n, c, h, w, ho, wo, _stream(), meta=_bytes(x, out))
    return out


# ------------------------------------------------------------------------------------------------ convolution
def _conv_meta(n, h, w, ci, cop, k, real_dims):
    """Profiling record: executed (padded) FLOPs and algorithmic FLOPs (real channel counts) of one launch."""
    rci, rco = real_dims if real_dims is not None else (ci, cop)
    return {"shape": (n, h, w, ci, cop, k), "flops_exec": 2.0 * n * h * w * ci * cop * k * k,
            "flops": 2.0 * n * h * w * rci * rco * k * k}


def _bytes(*tensors) -> dict:
    return {"bytes": float(sum(t.numel() * t.element_size() for t in tensors if t is not None))}


def weight_prep(w: torch.Tensor, want_fwd: bool = True, want_dgrad: bool = True
                ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """nn.Conv2d weight [Co,Ci,R,S] fp32 -> (wf [Co_pad,R*S,Ci_pad], wd [Ci_pad,R*S,Co_pad]) bf16."""
    _chk(w, "weight", torch.float32)
    if _prep.valid:
        hit = _prep.map.get(w.data_ptr())
        if hit is not None and not _prep.joined:
            join_wgrad_stream()
            _prep.joined = True
        if hit is not None:                      # prepared by this step's batched launch (ops.step_scope)
            return (hit[0] if want_fwd else None), (hit[1] if want_dgrad else None)
    co, ci, r, s = w.shape
    cop, cip = pad_channels(co), pad_channels(ci)
    wf = torch.empty((cop, r * s, cip), device=w.device, dtype=torch.bfloat16) if want_fwd else None
    wd = torch.empty((cip, r * s, cop), device=w.device, dtype=torch.bfloat16) if want_dgrad else None
    call("fv_weight_prep", w.data_ptr(), _ptr(wf), _ptr(wd), co, ci, r, s, cop, cip, _stream())
    return wf, wd


def conv2d(x: torch.Tensor, wf: torch.Tensor, bias: Optional[torch.Tensor], co: int, ksize: int,
           residual: Optional[torch.Tensor] = None, out_mode: int = OUT_NHWC_BF16, real_dims=None, want_stats: bool = False):
    """x NHWC bf16 [N,H,W,Ci]; wf [Co_pad, k*k, Ci] bf16 -> y (NHWC [N,H,W,Co_pad] or NCHW fp32 [N,co,H,W]).
    ``want_stats``: also return the batch-norm sums [2*Co_pad] (sum | sum of squares) of y, produced by the conv epilogue."""
    _chk(x, "x", torch.bfloat16)
    _chk(wf, "wf", torch.bfloat16)
    n, h, w, ci = x.shape
    cop, taps, ci2 = wf.shape
    if ci2 != ci or taps != ksize * ksize:
        raise _lib.FaceVaeError(f"conv2d: filter {tuple(wf.shape)} does not match input channels {ci} / k={ksize}")
    if out_mode == OUT_NCHW_F32:
        y = torch.empty((n, co, h, w), device=x.device, dtype=torch.float32)
    else:
        y = torch.empty((n, h, w, cop), device=x.device,
                        dtype=torch.bfloat16 if out_mode == OUT_NHWC_BF16 else torch.float32)
    if bias is not None:
        _chk(bias, "bias", torch.float32)
    if residual is not None:
        _chk(residual, "residual", torch.bfloat16)
        if tuple(residual.shape) != (n, h, w, cop):
            raise _lib.FaceVaeError("conv2d: residual shape mismatch")
    if want_stats and not _lib.load().fv_conv2d_fuses_stats(out_mode, n, h, w, ci, cop, ksize, ksize, int(residual is not None)):
        # the library would run a separate statistic pass for this shape: issue the two calls here (timed individually)
        call("fv_conv2d", x.data_ptr(), wf.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(), out_mode, n, h, w, ci, co,
             cop, ksize, ksize, (ksize - 1) // 2, _stream(), meta=_conv_meta(n, h, w, ci, cop, ksize, real_dims))
        return y, bn_stats(y)
    if want_stats:
        sums = torch.empty((2 * cop,), device=x.device, dtype=torch.float32)
        call("fv_conv2d_stats", x.data_ptr(), wf.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(), out_mode, n, h, w, ci, co,
             cop, ksize, ksize, (ksize - 1) // 2, sums.data_ptr(), _red_ws(), _stream(), meta=_conv_meta(n, h, w, ci, cop, ksize, real_dims))
        return y, sums
    call("fv_conv2d", x.data_ptr(), wf.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(), out_mode, n, h, w, ci, co,
         cop, ksize, ksize, (ksize - 1) // 2, _stream(), meta=_conv_meta(n, h, w, ci, cop, ksize, real_dims))
    return y


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, ksize: int, real_dims=None) -> torch.Tensor:
    """x NHWC bf16 [N,H,W,Ci], dy NHWC bf16 [N,H,W,Co_pad] -> partial slabs fp32 [splits, Co_pad, k*k, Ci] (one per pixel split
    of the kernel; ``wgrad_finish`` adds them in a fixed order)."""
    _chk(x, "x", torch.bfloat16)
    _chk(dy, "dy", torch.bfloat16)
    n, h, w, ci = x.shape
    cop = dy.shape[3]
    if tuple(dy.shape[:3]) != (n, h, w):
        raise _lib.FaceVaeError("conv2d_wgrad: x / dy shape mismatch")
    splits = int(_lib.load().fv_conv2d_wgrad_splits(0, n, h, w, ci, cop, ksize, ksize))
    part = torch.empty((max(splits, 1), cop, ksize * ksize, ci), device=x.device, dtype=torch.float32)
    call("fv_conv2d_wgrad", x.data_ptr(), dy.data_ptr(), part.data_ptr(), splits, n, h, w, ci, cop, ksize, ksize,
         (ksize - 1) // 2, _stream(), meta=_conv_meta(n, h, w, ci, cop, ksize, real_dims))
    return part


def wgrad_finish(part: torch.Tensor, co: int, ci: int, ksize: int, grad: Optional[torch.Tensor] = None) -> torch.Tensor:
    """partial slabs [splits, Co_pad, k*k, Ci_pad] -> nn.Conv2d layout gradient [co, ci, k, k] fp32."""
    accumulate = grad is not None
    if grad is None:
        grad = torch.empty((co, ci, ksize, ksize), device=part.device, dtype=torch.float32)
    call("fv_wgrad_finish", part.data_ptr(), part.shape[0], _chk(grad, "grad", torch.float32).data_ptr(), co, ci, ksize, ksize,
         part.shape[1], part.shape[3], int(accumulate), _stream())
    return grad


# ------------------------------------------------------------------------------------------------ up-sampling / stride-2 convolutions
def _conv_meta_x2(n, h, w, ci, cop, real_dims, exec_taps, alg_taps):
    """(h, w) = coarse grid.  Executed FLOPs: exec_taps MACs per coarse pixel and channel pair; algorithmic FLOPs: what the
    reference's formulation costs (alg_taps per coarse pixel: 36 for up-sample + 3x3, 16 for the 4x4 stride-2 conv)."""
    rci, rco = real_dims if real_dims is not None else (ci, cop)
    return {"shape": (n, h, w, ci, cop, exec_taps), "flops_exec": 2.0 * n * h * w * ci * cop * exec_taps,
            "flops": 2.0 * n * h * w * rci * rco * alg_taps}


def weight_prep_up(w: torch.Tensor, want_fwd: bool = True, want_dgrad: bool = True):
    """nn.Conv2d weight [Co,Ci,3,3] of an UpBlock2D conv -> (wx2 [4, Co_pad, 4, Ci_pad], ws2 [Ci_pad, 16, Co_pad]) bf16."""
    _chk(w, "weight", torch.float32)
    if _prep.valid:
        hit = _prep.map_up.get(w.data_ptr())
        if hit is not None and not _prep.joined:
            join_wgrad_stream()
            _prep.joined = True
        if hit is not None:                      # prepared by this step's batched launch (ops.step_scope)
            return (hit[0] if want_fwd else None), (hit[1] if want_dgrad else None)
    co, ci, r, s = w.shape
    if r != 3 or s != 3:
        raise _lib.FaceVaeError("weight_prep_up: the up-sampling convolution is 3x3")
    cop, cip = pad_channels(co), pad_channels(ci)
    wx2 = torch.empty((4, cop, 4, cip), device=w.device, dtype=torch.bfloat16) if want_fwd else None
    ws2 = torch.empty((cip, 16, cop), device=w.device, dtype=torch.bfloat16) if want_dgrad else None
    call("fv_weight_prep_up", w.data_ptr(), _ptr(wx2), _ptr(ws2), co, ci, cop, cip, _stream())
    return wx2, ws2


def weight_prep_s2(w: torch.Tensor, want_fwd: bool = True, want_dgrad: bool = True):
    """weight [Co,Ci,4,4] of a 4x4 stride-2 conv -> (wf [Co_pad, 16, Ci_pad], wx2 [4, Ci_pad, 4, Co_pad]) bf16."""
    _chk(w, "weight", torch.float32)
    if _prep.valid:
        hit = _prep.map_s2.get(w.data_ptr())
        if hit is not None and not _prep.joined:
            join_wgrad_stream()
            _prep.joined = True
        if hit is not None:
            return (hit[0] if want_fwd else None), (hit[1] if want_dgrad else None)
    co, ci, r, s = w.shape
    if r != 4 or s != 4:
        raise _lib.FaceVaeError("weight_prep_s2: the stride-2 convolution is 4x4")
    cop, cip = pad_channels(co), pad_channels(ci)
    wf = torch.empty((cop, 16, cip), device=w.device, dtype=torch.bfloat16) if want_fwd else None
    wx2 = torch.empty((4, cip, 4, cop), device=w.device, dtype=torch.bfloat16) if want_dgrad else None
    call("fv_weight_prep_s2", w.data_ptr(), _ptr(wf), _ptr(wx2), co, ci, cop, cip, _stream())
    return wf, wx2


def conv2d_x2(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], co: int, out_mode: int = OUT_NHWC_BF16,
              real_dims=None, want_stats: bool = False, alg_taps: int = 36):
    """x NHWC bf16 [N,H,W,Ci] -> y [N,2H,2W,Co_pad]: four 2x2 phase convolutions, wp [4, Co_pad, 4, Ci] (UpBlock2D's
    up-sample + 3x3 conv without the up-sampled tensor; also the data gradient of ``conv2d_s2``)."""
    _chk(x, "x", torch.bfloat16)
    _chk(wp, "wp", torch.bfloat16)
    n, h, w, ci = x.shape
    if wp.dim() != 4 or wp.shape[0] != 4 or wp.shape[2] != 4 or wp.shape[3] != ci:
        raise _lib.FaceVaeError(f"conv2d_x2: filter {tuple(wp.shape)} does not match input channels {ci}")
    cop = wp.shape[1]
    if out_mode == OUT_NCHW_F32:
        y = torch.empty((n, co, 2 * h, 2 * w), device=x.device, dtype=torch.float32)
    else:
        y = torch.empty((n, 2 * h, 2 * w, cop), device=x.device, dtype=torch.bfloat16 if out_mode == OUT_NHWC_BF16 else torch.float32)
    fused = want_stats and bool(_lib.load().fv_conv2d_geom_fuses_stats(1, out_mode, n, h, w, ci, cop))
    sums = torch.empty((2 * cop,), device=x.device, dtype=torch.float32) if fused else None
    call("fv_conv2d_x2", x.data_ptr(), wp.data_ptr(), _ptr(bias), y.data_ptr(), out_mode, n, h, w, ci, co, cop, _ptr(sums),
         _red_ws() if fused else None, _stream(), meta=_conv_meta_x2(n, h, w, ci, cop, real_dims, 16, alg_taps))
    if want_stats and not fused:      # the library would run a separate statistic pass: issue it here (timed on its own)
        sums = bn_stats(y)
    return (y, sums) if want_stats else y


def conv2d_s2(x: torch.Tensor, wf: torch.Tensor, bias: Optional[torch.Tensor], co: int, out_mode: int = OUT_NHWC_BF16,
              real_dims=None, want_stats: bool = False, alg_taps: int = 16):
    """x NHWC bf16 [N,2H,2W,Ci] -> y [N,H,W,Co_pad]: 4x4 stride-2 pad-1 convolution, wf [Co_pad, 16, Ci] (Conv2dELR of
    EFE_conv6; with ``weight_prep_up``'s ws2 the data gradient of ``conv2d_x2``)."""
    _chk(x, "x", torch.bfloat16)
    _chk(wf, "wf", torch.bfloat16)
    n, h2, w2, ci = x.shape
    if h2 % 2 or w2 % 2:
        raise _lib.FaceVaeError("conv2d_s2: even input sizes only")
    h, w = h2 // 2, w2 // 2
    if wf.dim() != 3 or wf.shape[1] != 16 or wf.shape[2] != ci:
        raise _lib.FaceVaeError(f"conv2d_s2: filter {tuple(wf.shape)} does not match input channels {ci}")
    cop = wf.shape[0]
    if out_mode == OUT_NCHW_F32:
        y = torch.empty((n, co, h, w), device=x.device, dtype=torch.float32)
    else:
        y = torch.empty((n, h, w, cop), device=x.device, dtype=torch.bfloat16 if out_mode == OUT_NHWC_BF16 else torch.float32)
    fused = want_stats and bool(_lib.load().fv_conv2d_geom_fuses_stats(2, out_mode, n, h, w, ci, cop))
    sums = torch.empty((2 * cop,), device=x.device, dtype=torch.float32) if fused else None
    call("fv_conv2d_s2", x.data_ptr(), wf.data_ptr(), _ptr(bias), y.data_ptr(), out_mode, n, h, w, ci, co, cop, _ptr(sums),
         _red_ws() if fused else None, _stream(), meta=_conv_meta_x2(n, h, w, ci, cop, real_dims, 16, alg_taps))
    if want_stats and not fused:
        sums = bn_stats(y)
    return (y, sums) if want_stats else y


def conv2d_ex(kind: int, x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], co: int, ksize: int = 1, act: int = ACT_NONE,
              out_mode: int = OUT_NHWC_BF16, real_dims=None) -> torch.Tensor:
    """Convolution + bias + activation in one kernel (Conv2dELR.forward, reference models_utils.py:712-742).  kind 0: stride-1
    "same" conv, w [Co_pad, k*k, Ci]; kind 2: 4x4 stride-2 pad-1 conv of x [N,2H,2W,Ci], w [Co_pad, 16, Ci]; kind 1: the
    four-phase x2 geometry, w [4, Co_pad, 4, Ci]."""
    _chk(x, "x", torch.bfloat16)
    _chk(w, "w", torch.bfloat16)
    n, hx, wx, ci = x.shape
    if kind == 2:
        if hx % 2 or wx % 2:
            raise _lib.FaceVaeError("conv2d_ex: the stride-2 convolution needs even input sizes")
        h, wd, ho, wo, cop, taps_exec, taps_alg = hx // 2, wx // 2, hx // 2, wx // 2, w.shape[0], 16, 16
    elif kind == 1:
        h, wd, ho, wo, cop, taps_exec, taps_alg = hx, wx, 2 * hx, 2 * wx, w.shape[1], 16, 16
    else:
        h, wd, ho, wo, cop, taps_exec, taps_alg = hx, wx, hx, wx, w.shape[0], ksize * ksize, ksize * ksize
    if out_mode == OUT_NCHW_F32:
        y = torch.empty((n, co, ho, wo), device=x.device, dtype=torch.float32)
    else:
        y = torch.empty((n, ho, wo, cop), device=x.device, dtype=torch.bfloat16 if out_mode == OUT_NHWC_BF16 else torch.float32)
    call("fv_conv2d_ex", kind, x.data_ptr(), w.data_ptr(), _ptr(bias), None, y.data_ptr(), out_mode, n, h, wd, ci, co, cop, ksize, ksize,
         (ksize - 1) // 2, act, None, None, _stream(), meta=_conv_meta_x2(n, h, wd, ci, cop, real_dims, taps_exec, taps_alg))
    return y


def demod_fwd(w: torch.Tensor, gain: float, demod: bool):
    """Conv2dELR.getweight (reference models_utils.py:686-704): -> (weff fp32 same shape, inv_norm [Co] | None)."""
    _chk(w, "weight", torch.float32)
    co = w.shape[0]
    k = w.numel() // co
    weff = torch.empty_like(w)
    inv = torch.empty((co,), device=w.device, dtype=torch.float32) if demod else None
    call("fv_demod_fwd", w.data_ptr(), weff.data_ptr(), _ptr(inv), co, k, float(gain), int(demod), _stream())
    return weff, inv


def demod_bwd(w: torch.Tensor, inv: Optional[torch.Tensor], dweff: torch.Tensor, gain: float, demod: bool) -> torch.Tensor:
    _chk(dweff, "dweff", torch.float32)
    co = w.shape[0]
    dw = torch.empty_like(w)
    call("fv_demod_bwd", w.data_ptr(), _ptr(inv), dweff.data_ptr(), dw.data_ptr(), co, w.numel() // co, float(gain), int(demod), _stream())
    return dw


def act_bwd(out: torch.Tensor, g: torch.Tensor, act: int) -> torch.Tensor:
    """dy = g * act'(out) for an activation fused into a conv epilogue (NHWC bf16)."""
    _chk(out, "out", torch.bfloat16)
    _chk(g, "g", torch.bfloat16)
    dy = torch.empty_like(out)
    call("fv_act_bwd", out.data_ptr(), g.data_ptr(), dy.data_ptr(), out.numel(), act, _stream(), meta=_bytes(out, g, dy))
    return dy


def conv2d_wgrad_x2(x: torch.Tensor, dy: torch.Tensor, real_dims=None) -> torch.Tensor:
    """x coarse [N,H,W,Ci], dy fine [N,2H,2W,Co_pad] -> partial slabs [splits, 4, Co_pad, 4, Ci] of the phase filters."""
    _chk(x, "x", torch.bfloat16)
    _chk(dy, "dy", torch.bfloat16)
    n, h, w, ci = x.shape
    cop = dy.shape[3]
    if tuple(dy.shape[:3]) != (n, 2 * h, 2 * w):
        raise _lib.FaceVaeError("conv2d_wgrad_x2: x / dy shape mismatch")
    splits = int(_lib.load().fv_conv2d_wgrad_splits(1, n, h, w, ci, cop, 2, 2))
    part = torch.empty((max(splits, 1), 4, cop, 4, ci), device=x.device, dtype=torch.float32)
    call("fv_conv2d_wgrad_x2", x.data_ptr(), dy.data_ptr(), part.data_ptr(), splits, n, h, w, ci, cop, _stream(),
         meta=_conv_meta_x2(n, h, w, ci, cop, real_dims, 16, 36))
    return part


def wgrad_finish_up(part: torch.Tensor, co: int, ci: int, grad: Optional[torch.Tensor] = None) -> torch.Tensor:
    """phase slabs [splits, 4, Co_pad, 4, Ci_pad] -> gradient of the 3x3 filter [co, ci, 3, 3] fp32."""
    accumulate = grad is not None
    if grad is None:
        grad = torch.empty((co, ci, 3, 3), device=part.device, dtype=torch.float32)
    if part.shape[0] > 1:
        # splits first (one wide, vectorised pass over 16 * Co_pad * Ci_pad elements), then the 16 -> 9 fold: the fold kernel has
        # only Co * Ci threads and would walk the splits serially
        tot = torch.empty(part.shape[1:], device=part.device, dtype=torch.float32)
        n = tot.numel()
        call("fv_slab_sum", part.data_ptr(), part.shape[0], n, tot.data_ptr(), n, 0, _stream())
        part = tot.unsqueeze(0)
    call("fv_wgrad_finish_up", part.data_ptr(), part.shape[0], _chk(grad, "grad", torch.float32).data_ptr(), co, ci, part.shape[2],
         part.shape[4], int(accumulate), _stream())
    return grad


def conv2d_wgrad_s2(x: torch.Tensor, dy: torch.Tensor, real_dims=None) -> torch.Tensor:
    """x fine [N,2H,2W,Ci], dy coarse [N,H,W,Co_pad] -> partial slabs [splits, Co_pad, 16, Ci] (``wgrad_finish(ksize=4)``)."""
    _chk(x, "x", torch.bfloat16)
    _chk(dy, "dy", torch.bfloat16)
    n, h, w, cop = dy.shape
    ci = x.shape[3]
    if tuple(x.shape[:3]) != (n, 2 * h, 2 * w):
        raise _lib.FaceVaeError("conv2d_wgrad_s2: x / dy shape mismatch")
    splits = int(_lib.load().fv_conv2d_wgrad_splits(2, n, h, w, ci, cop, 4, 4))
    part = torch.empty((max(splits, 1), cop, 16, ci), device=x.device, dtype=torch.float32)
    call("fv_conv2d_wgrad_s2", x.data_ptr(), dy.data_ptr(), part.data_ptr(), splits, n, h, w, ci, cop, _stream(),
         meta=_conv_meta_x2(n, h, w, ci, cop, real_dims, 16, 16))
    return part


def colsum(y: torch.Tensor) -> torch.Tensor:
    _chk(y, "y", torch.bfloat16)
    c = y.shape[-1]
    sums = torch.empty((c,), device=y.device, dtype=torch.float32)
    call("fv_colsum", y.data_ptr(), sums.data_ptr(), y.numel() // c, c, _red_ws(), _stream())
    return sums


# ------------------------------------------------------------------------------------------------ out_conv (7x7, 32 -> <= 4)
def outconv_supported(n: int, h: int, w: int, ci: int, co: int, k: int) -> bool:
    """Shapes the tap-folded out_conv kernels (csrc/fv_outconv.cu) take; everything else goes through conv2d."""
    return bool(_lib.load().fv_outconv_supported(n, h, w, ci, co, k, k))


def outconv_prep(w: torch.Tensor, want_fwd: bool = True, want_dgrad: bool = True):
    """nn.Conv2d weight [Co,32,7,7] fp32 -> (wq, wdq) bf16 [7,32,32] folded operands of the forward / data-gradient GEMMs."""
    _chk(w, "weight", torch.float32)
    co, ci = w.shape[0], w.shape[1]
    wq = torch.empty((7, 32, 32), device=w.device, dtype=torch.bfloat16) if want_fwd else None
    wdq = torch.empty((7, 32, 32), device=w.device, dtype=torch.bfloat16) if want_dgrad else None
    call("fv_outconv_prep", w.data_ptr(), _ptr(wq), _ptr(wdq), co, ci, _stream())
    return wq, wdq


def outconv_fwd(x: torch.Tensor, wq: torch.Tensor, bias: Optional[torch.Tensor], co: int, target: Optional[torch.Tensor] = None,
                l1: bool = False, use_sigmoid: bool = True, gscale: float = 1.0, want_logits: Optional[bool] = None,
                want_pred: bool = True):
    """x NHWC bf16 [N,H,W,32] -> dict(logits, pred, g4, loss_sum, gsum); the loss entries need ``target`` (NCHW fp32)."""
    _chk(x, "x", torch.bfloat16)
    _chk(wq, "wq", torch.bfloat16)
    n, h, w, ci = x.shape
    fused = target is not None
    if want_logits is None:
        want_logits = not fused
    dev = x.device
    logits = torch.empty((n, co, h, w), device=dev, dtype=torch.float32) if want_logits else None
    pred = g4 = acc = None
    if fused:
        _chk(target, "target", torch.float32)
        if tuple(target.shape) != (n, co, h, w):
            raise _lib.FaceVaeError("outconv_fwd: target shape mismatch")
        pred = torch.empty((n, co, h, w), device=dev, dtype=torch.float32) if want_pred else None
        g4 = torch.empty((n, h, w, 4), device=dev, dtype=torch.bfloat16)
        acc = torch.empty((8,), device=dev, dtype=torch.float32)         # [0]: loss sum, [4:8]: gradient sums per channel
    if bias is not None:
        _chk(bias, "bias", torch.float32)
    meta = _conv_meta(n, h, w, ci, 32, 7, (ci, co))
    meta.update(_bytes(x, logits, target, pred, g4))
    call("fv_outconv_fwd", x.data_ptr(), wq.data_ptr(), _ptr(bias), _ptr(logits), _ptr(target), _ptr(pred), _ptr(g4),
         _ptr(acc), None if acc is None else acc[4:].data_ptr(), n, h, w, ci, co, int(l1), int(use_sigmoid), float(gscale),
         _red_ws() if fused else None, _stream(), meta=meta)
    return {"logits": logits, "pred": pred, "g4": g4, "loss_sum": None if acc is None else acc[:1],
            "gsum": None if acc is None else acc[4:4 + co]}


def outconv_dgrad(g4: torch.Tensor, wdq: torch.Tensor, scale_ptr: Optional[torch.Tensor], co: int) -> torch.Tensor:
    _chk(g4, "g4", torch.bfloat16)
    n, h, w, c4 = g4.shape
    if c4 != 4:
        raise _lib.FaceVaeError("outconv_dgrad: g4 must be [N,H,W,4]")
    dx = torch.empty((n, h, w, 32), device=g4.device, dtype=torch.bfloat16)
    call("fv_outconv_dgrad", g4.data_ptr(), wdq.data_ptr(), _ptr(scale_ptr), dx.data_ptr(), n, h, w, 32, co, _stream(),
         meta=_conv_meta(n, h, w, 32, 32, 7, (co, 32)))
    return dx


def outconv_wgrad(x: torch.Tensor, g4: torch.Tensor, scale_ptr: Optional[torch.Tensor], co: int) -> torch.Tensor:
    """-> dw fp32 [co, 32, 7, 7] (nn.Conv2d layout): per-CTA slabs from the tensor-core kernel, added in CTA order."""
    _chk(x, "x", torch.bfloat16)
    _chk(g4, "g4", torch.bfloat16)
    n, h, w, ci = x.shape
    splits = int(_lib.load().fv_outconv_wgrad_splits(n, h, w))
    part = torch.empty((max(splits, 1), co, ci, 7, 7), device=x.device, dtype=torch.float32)
    call("fv_outconv_wgrad", x.data_ptr(), g4.data_ptr(), _ptr(scale_ptr), part.data_ptr(), splits, n, h, w, ci, co, _stream(),
         meta=_conv_meta(n, h, w, ci, 32, 7, (ci, co)))
    dw = torch.empty((co, ci, 7, 7), device=x.device, dtype=torch.float32)
    call("fv_slab_sum", part.data_ptr(), splits, co * ci * 49, dw.data_ptr(), co * ci * 49, 0, _stream())
    return dw


# ------------------------------------------------------------------------------------------------ batch norm glue
def bn_stats(y: torch.Tensor) -> torch.Tensor:
    _chk(y, "y")
    c = y.shape[-1]
    sums = torch.empty((2 * c,), device=y.device, dtype=torch.float32)
    call("fv_bn_stats", y.data_ptr(), _dt(y), sums.data_ptr(), y.numel() // c, c, _red_ws(), _stream(), meta=_bytes(y))
    return sums


def bn_finalize(sums: torch.Tensor, count: float, gamma: torch.Tensor, beta: torch.Tensor,
                running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor],
                momentum: float = BN_MOMENTUM, eps: float = BN_EPS) -> torch.Tensor:
    c = gamma.numel()
    stat = torch.empty((4, c), device=sums.device, dtype=torch.float32)
    call("fv_bn_finalize", sums.data_ptr(), float(count), gamma.data_ptr(), beta.data_ptr(), _ptr(running_mean),
         _ptr(running_var), momentum, eps, stat.data_ptr(), c, _stream())
    return stat


def bn_eval_affine(gamma, beta, running_mean, running_var, eps: float = BN_EPS) -> torch.Tensor:
    c = gamma.numel()
    stat = torch.empty((4, c), device=gamma.device, dtype=torch.float32)
    call("fv_bn_eval_affine", gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(), running_var.data_ptr(), eps,
         stat.data_ptr(), c, _stream())
    return stat


def bn_act_fwd(y: torch.Tensor, stat: torch.Tensor, mode: int = MODE_NONE, act: int = ACT_RELU,
               out_dtype=torch.bfloat16, nchw_out: bool = False) -> torch.Tensor:
    _chk(y, "y")
    n, h, w, c = y.shape
    ho, wo = (h // 2, w // 2) if mode == MODE_POOL else ((2 * h, 2 * w) if mode == MODE_UP else (h, w))
    shape = (n, c, ho, wo) if nchw_out else (n, ho, wo, c)
    out = torch.empty(shape, device=y.device, dtype=out_dtype)
    call("fv_bn_act_fwd", y.data_ptr(), _dt(y), stat.data_ptr(), out.data_ptr(), _dt(out), int(nchw_out), n, h, w, c,
         mode, act, _stream(), meta=_bytes(y, out))
    return out


def bn_act_fwd_fin(y: torch.Tensor, sums: torch.Tensor, count: float, gamma, beta, running_mean, running_var, mode: int = MODE_NONE,
                   act: int = ACT_RELU, out_dtype=torch.bfloat16, nchw_out: bool = False, momentum: float = BN_MOMENTUM,
                   eps: float = BN_EPS):
    """bn_finalize + bn_act_fwd in one launch -> (out, stat [4, C])."""
    _chk(y, "y")
    n, h, w, c = y.shape
    ho, wo = (h // 2, w // 2) if mode == MODE_POOL else ((2 * h, 2 * w) if mode == MODE_UP else (h, w))
    shape = (n, c, ho, wo) if nchw_out else (n, ho, wo, c)
    out = torch.empty(shape, device=y.device, dtype=out_dtype)
    stat = torch.empty((4, c), device=y.device, dtype=torch.float32)
    call("fv_bn_act_fwd_fin", y.data_ptr(), _dt(y), sums.data_ptr(), float(count), gamma.data_ptr(), beta.data_ptr(),
         _ptr(running_mean), _ptr(running_var), momentum, eps, stat.data_ptr(), out.data_ptr(), _dt(out), int(nchw_out), n, h, w, c,
         mode, act, _stream(), meta=_bytes(y, out))
    return out, stat


def bn_act_bwd_apply_fin(y: torch.Tensor, g: torch.Tensor, stat: torch.Tensor, sums: torch.Tensor, count: float, mode: int, act: int,
                         add: Optional[torch.Tensor] = None, g_nchw: bool = False):
    """bn_bwd_finalize + bn_act_bwd_apply in one launch -> (dy, dgamma, dbeta)."""
    n, h, w, c = y.shape
    dy = torch.empty((n, h, w, c), device=y.device, dtype=torch.bfloat16)
    dgamma = torch.empty((c,), device=y.device, dtype=torch.float32)
    dbeta = torch.empty((c,), device=y.device, dtype=torch.float32)
    if add is not None:
        _chk(add, "add", torch.bfloat16)
    call("fv_bn_act_bwd_apply_fin", y.data_ptr(), _dt(y), g.data_ptr(), _dt(g), int(g_nchw), stat.data_ptr(), sums.data_ptr(),
         float(count), dgamma.data_ptr(), dbeta.data_ptr(), _ptr(add), dy.data_ptr(), n, h, w, c, mode, act, _stream(),
         meta=_bytes(y, g, add, dy))
    return dy, dgamma, dbeta


def bn_act_bwd_reduce(y: torch.Tensor, g: torch.Tensor, stat: torch.Tensor, mode: int, act: int,
                      g_nchw: bool = False) -> torch.Tensor:
    _chk(y, "y")
    _chk(g, "g")
    n, h, w, c = y.shape
    sums = torch.empty((2 * c,), device=y.device, dtype=torch.float32)
    call("fv_bn_act_bwd_reduce", y.data_ptr(), _dt(y), g.data_ptr(), _dt(g), int(g_nchw), stat.data_ptr(),
         sums.data_ptr(), n, h, w, c, mode, act, _red_ws(), _stream(), meta=_bytes(y, g))
    return sums


def bn_bwd_finalize(sums_local: torch.Tensor, sums_global: torch.Tensor, count: float, c: int,
                    dgamma: Optional[torch.Tensor] = None, dbeta: Optional[torch.Tensor] = None):
    accumulate = dgamma is not None
    if dgamma is None:
        dgamma = torch.empty((c,), device=sums_local.device, dtype=torch.float32)
        dbeta = torch.empty((c,), device=sums_local.device, dtype=torch.float32)
    coef = torch.empty((2, c), device=sums_local.device, dtype=torch.float32)
    call("fv_bn_bwd_finalize", sums_local.data_ptr(), sums_global.data_ptr(), float(count), dgamma.data_ptr(),
         dbeta.data_ptr(), coef.data_ptr(), c, int(accumulate), _stream())
    return dgamma, dbeta, coef


def bn_act_bwd_apply(y: torch.Tensor, g: torch.Tensor, stat: torch.Tensor, coef: torch.Tensor, mode: int, act: int,
                     add: Optional[torch.Tensor] = None, g_nchw: bool = False) -> torch.Tensor:
    n, h, w, c = y.shape
    dy = torch.empty((n, h, w, c), device=y.device, dtype=torch.bfloat16)
    if add is not None:
        _chk(add, "add", torch.bfloat16)
    call("fv_bn_act_bwd_apply", y.data_ptr(), _dt(y), g.data_ptr(), _dt(g), int(g_nchw), stat.data_ptr(),
         coef.data_ptr(), _ptr(add), dy.data_ptr(), n, h, w, c, mode, act, _stream(), meta=_bytes(y, g, add, dy))
    return dy


# ------------------------------------------------------------------------------------------------ instance norm (Discriminator blocks)
def in_stats(y: torch.Tensor, eps: float = BN_EPS) -> torch.Tensor:
    """NHWC bf16 [N,H,W,C] -> stat [N, 2, C] (mean | invstd per image and channel): nn.InstanceNorm2d statistics."""
    _chk(y, "y", torch.bfloat16)
    n, h, w, c = y.shape
    stat = torch.empty((n, 2, c), device=y.device, dtype=torch.float32)
    call("fv_in_stats", y.data_ptr(), stat.data_ptr(), n, h, w, c, float(eps), _stream(), meta=_bytes(y))
    return stat


def in_act_fwd(y, stat, gamma, beta, act: int) -> torch.Tensor:
    n, h, w, c = y.shape
    out = torch.empty_like(y)
    call("fv_in_act_fwd", y.data_ptr(), stat.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), n, h, w, c, act, _stream(),
         meta=_bytes(y, out))
    return out


def in_bwd(y, g, stat, gamma, beta, act: int):
    """-> (dy NHWC bf16, dgamma [C], dbeta [C])."""
    _chk(g, "g", torch.bfloat16)
    n, h, w, c = y.shape
    sums = torch.empty((n, 2, c), device=y.device, dtype=torch.float32)
    call("fv_in_bwd_sums", y.data_ptr(), g.data_ptr(), stat.data_ptr(), gamma.data_ptr(), beta.data_ptr(), sums.data_ptr(), n, h, w, c, act,
         _stream(), meta=_bytes(y, g))
    dy = torch.empty_like(y)
    call("fv_in_bwd_apply", y.data_ptr(), g.data_ptr(), stat.data_ptr(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), dy.data_ptr(), n, h,
         w, c, act, _stream(), meta=_bytes(y, g, dy))
    tot = sums.sum(dim=0)                      # a [N, 2, C] -> [2, C] sum: torch's reduction (deterministic for a fixed shape)
    return dy, tot[1].contiguous(), tot[0].contiguous()


# ------------------------------------------------------------------------------------------------ VAE bottleneck + losses
def reparam_kl_fwd(mu: torch.Tensor, logstd: torch.Tensor, eps: Optional[torch.Tensor], want_z: bool = True,
                   want_kl: bool = True):
    """mu / logstd: fp32 [N, Dz] row views (unit inner stride, common row stride).  Returns (z [N,Dz] | None, kl partial sums
    [N, P] | None -- ``.sum()`` of it is the KL sum over all samples and latent dimensions)."""
    n, dz = mu.shape
    if mu.stride(1) != 1 or logstd.stride(1) != 1 or mu.stride(0) != logstd.stride(0):
        raise _lib.FaceVaeError("reparam_kl_fwd: mu/logstd must be row views with a common row stride")
    z = torch.empty((n, dz), device=mu.device, dtype=torch.float32) if want_z else None
    # per-block partial sums [N, P]; the (few) partials are added by the caller in a fixed order (no atomics)
    kl = torch.empty((n, int(_lib.load().fv_reparam_kl_parts(n, dz))), device=mu.device, dtype=torch.float32) if want_kl else None
    if eps is not None:
        _chk(eps, "eps", torch.float32)
    call("fv_reparam_kl_fwd", mu.data_ptr(), logstd.data_ptr(), mu.stride(0), _ptr(eps), _ptr(z), _ptr(kl), n, dz,
         _stream(), meta=_bytes(mu, logstd, eps, z))
    return z, kl


def reparam_kl_bwd(mu, logstd, eps, dz, dmu_ext, dls_ext, kscale: float, kscale_ptr: Optional[torch.Tensor],
                   out: Optional[torch.Tensor] = None):
    """Returns dh fp32 [N, 2*Dz] = (dmu | dlogstd)."""
    n, d = mu.shape
    if out is None:
        out = torch.empty((n, 2 * d), device=mu.device, dtype=torch.float32)
    dmu, dls = out[:, :d], out[:, d:]
    call("fv_reparam_kl_bwd", mu.data_ptr(), logstd.data_ptr(), mu.stride(0), _ptr(eps), _ptr(dz), _ptr(dmu_ext),
         _ptr(dls_ext), float(kscale), _ptr(kscale_ptr), dmu.data_ptr(), dls.data_ptr(), out.stride(0), n, d, _stream())
    return out


def recon_loss(logits: torch.Tensor, target: torch.Tensor, l1: bool = False, use_sigmoid: bool = True,
               gscale: float = 1.0, want_pred: bool = True, want_grad_f32: bool = False, want_grad_nhwc: bool = True):
    """Returns (loss_sum [1], pred | None, grad_f32 | None, grad_nhwc [N,H,W,16] bf16 | None)."""
    _chk(logits, "logits", torch.float32)
    _chk(target, "target", torch.float32)
    n, c, h, w = logits.shape
    cp = pad_channels(c)
    loss = torch.empty((1,), device=logits.device, dtype=torch.float32)
    pred = torch.empty_like(logits) if want_pred else None
    gf = torch.empty_like(logits) if want_grad_f32 else None
    gn = torch.empty((n, h, w, cp), device=logits.device, dtype=torch.bfloat16) if want_grad_nhwc else None
    call("fv_recon_loss", logits.data_ptr(), target.data_ptr(), _ptr(pred), _ptr(gf), _ptr(gn), loss.data_ptr(), n, c, h,
         w, cp, int(l1), int(use_sigmoid), float(gscale), _red_ws(), _stream(), meta=_bytes(logits, target, pred, gf, gn))
    return loss, pred, gf, gn


def recon_loss_flat(a: torch.Tensor, b: torch.Tensor, l1: bool = False, gscale: float = 1.0, want_grad: bool = True):
    _chk(a, "a", torch.float32)
    _chk(b, "b", torch.float32)
    loss = torch.empty((1,), device=a.device, dtype=torch.float32)
    grad = torch.empty_like(a) if want_grad else None
    call("fv_recon_loss_flat", a.data_ptr(), b.data_ptr(), _ptr(grad), loss.data_ptr(), a.numel(), int(l1), float(gscale),
         _red_ws(), _stream(), meta=_bytes(a, b, grad))
    return loss, grad


def scale(t: torch.Tensor, scale_ptr: Optional[torch.Tensor], s: float = 1.0, out: Optional[torch.Tensor] = None):
    _chk(t, "t")
    out = torch.empty_like(t) if out is None else out
    call("fv_scale", t.data_ptr(), out.data_ptr(), _dt(t), t.numel(), _ptr(scale_ptr), float(s), _stream())
    return out


# ------------------------------------------------------------------------------------------------ first encoder layer (C <= 4)
def pw_moments(x: torch.Tensor) -> torch.Tensor:
    """x NCHW fp32 [N,C,H,W], C <= 4 -> double [C + C*C]: sum x_c | sum x_c x_d."""
    _chk(x, "x", torch.float32)
    n, c, h, w = x.shape
    sums = torch.empty((c + c * c,), device=x.device, dtype=torch.float64)
    call("fv_pw_moments", x.data_ptr(), sums.data_ptr(), n, c, h * w, _red_ws(), _stream(), meta=_bytes(x))
    return sums


def pw_prepare(sums, count, w2d, bias, gamma, beta, running_mean, running_var, momentum=BN_MOMENTUM, eps=BN_EPS):
    """-> (coef [Co, C+1] fp32: A | c, stat [2, Co]: mean_y | invstd)."""
    co, c = w2d.shape
    coef = torch.empty((co, c + 1), device=w2d.device, dtype=torch.float32)
    stat = torch.empty((2, co), device=w2d.device, dtype=torch.float32)
    call("fv_pw_prepare", sums.data_ptr(), float(count), w2d.data_ptr(), _ptr(bias), gamma.data_ptr(), beta.data_ptr(),
         _ptr(running_mean), _ptr(running_var), momentum, eps, coef.data_ptr(), stat.data_ptr(), co, c, _stream())
    return coef, stat


def pw_fwd(x: torch.Tensor, coef: torch.Tensor, act: int = ACT_RELU) -> torch.Tensor:
    n, c, h, w = x.shape
    co = coef.shape[0]
    out = torch.empty((n, h, w, co), device=x.device, dtype=torch.bfloat16)
    call("fv_pw_fwd", x.data_ptr(), coef.data_ptr(), out.data_ptr(), n, c, h * w, co, act, _stream(), meta=_bytes(x, out))
    return out


def pw_bwd_reduce(x: torch.Tensor, g: torch.Tensor, coef: torch.Tensor, act: int = ACT_RELU) -> torch.Tensor:
    _chk(g, "g", torch.bfloat16)
    n, c, h, w = x.shape
    co = coef.shape[0]
    sums = torch.empty((co + co * c,), device=x.device, dtype=torch.float64)
    call("fv_pw_bwd_reduce", x.data_ptr(), g.data_ptr(), coef.data_ptr(), sums.data_ptr(), n, c, h * w, co, act, _red_ws(), _stream(),
         meta=_bytes(x, g))
    return sums


def pw_bwd_finalize(fsums, bsums, count, w2d, bias, gamma, stat):
    co, c = w2d.shape
    dw = torch.empty((co, c), device=w2d.device, dtype=torch.float32)
    dgamma = torch.empty((co,), device=w2d.device, dtype=torch.float32)
    dbeta = torch.empty((co,), device=w2d.device, dtype=torch.float32)
    call("fv_pw_bwd_finalize", fsums.data_ptr(), bsums.data_ptr(), float(count), w2d.data_ptr(), _ptr(bias), gamma.data_ptr(),
         stat.data_ptr(), dw.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), co, c, _stream())
    return dw, dgamma, dbeta
