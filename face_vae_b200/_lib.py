"""ctypes binding of libfacevae_b200.so (the C ABI declared in include/facevae_b200.h).

There is no fallback: if the library cannot be loaded (or built) the import of any compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfacevae_b200.so")

# enums of include/facevae_b200.h
DT_BF16, DT_F32 = 0, 1
OUT_NHWC_BF16, OUT_NHWC_F32, OUT_NCHW_F32 = 0, 1, 2
MODE_NONE, MODE_POOL, MODE_UP = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2

_p, _i, _ll, _f, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double

# name -> argument ctypes (all return int unless listed in _STR)
SIGNATURES = {
    "fv_device_ok": [],
    "fv_nchw_to_nhwc": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fv_nhwc_to_nchw": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p],
    "fv_bilinear_resize": [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fv_weight_prep": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fv_weight_prep_batched": [_p, _i, _ll, _p],
    "fv_weight_prep_flat": [_p, _i, _i, _p],
    "fv_weight_prep_tiled": [_p, _i, _i, _p],
    "fv_weight_prep_up": [_p, _p, _p, _i, _i, _i, _i, _p],
    "fv_weight_prep_s2": [_p, _p, _p, _i, _i, _i, _i, _p],
    "fv_conv2d": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "fv_conv2d_stats": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "fv_conv2d_ex": [_i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "fv_demod_fwd": [_p, _p, _p, _i, _i, _f, _i, _p],
    "fv_demod_bwd": [_p, _p, _p, _p, _i, _i, _f, _i, _p],
    "fv_act_bwd": [_p, _p, _p, _ll, _i, _p],
    "fv_conv2d_x2": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "fv_conv2d_s2": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "fv_conv2d_wgrad": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "fv_conv2d_wgrad_x2": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fv_conv2d_wgrad_s2": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fv_wgrad_finish": [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "fv_wgrad_finish_up": [_p, _i, _p, _i, _i, _i, _i, _i, _p],
    "fv_slab_sum": [_p, _i, _ll, _p, _ll, _i, _p],
    "fv_colsum": [_p, _p, _ll, _i, _p, _p],
    "fv_outconv_prep": [_p, _p, _p, _i, _i, _p],
    "fv_outconv_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p],
    "fv_outconv_dgrad": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fv_outconv_wgrad": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fv_bn_stats": [_p, _i, _p, _ll, _i, _p, _p],
    "fv_bn_finalize": [_p, _d, _p, _p, _p, _p, _f, _f, _p, _i, _p],
    "fv_bn_eval_affine": [_p, _p, _p, _p, _f, _p, _i, _p],
    "fv_bn_act_fwd": [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "fv_bn_act_fwd_fin": [_p, _i, _p, _d, _p, _p, _p, _p, _f, _f, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "fv_bn_act_bwd_apply_fin": [_p, _i, _p, _i, _i, _p, _p, _d, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fv_bn_act_bwd_reduce": [_p, _i, _p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p],
    "fv_bn_bwd_finalize": [_p, _p, _d, _p, _p, _p, _i, _i, _p],
    "fv_bn_act_bwd_apply": [_p, _i, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fv_in_stats": [_p, _p, _i, _i, _i, _i, _f, _p],
    "fv_in_act_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fv_in_bwd_sums": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fv_in_bwd_apply": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fv_reparam_kl_fwd": [_p, _p, _ll, _p, _p, _p, _i, _i, _p],
    "fv_reparam_kl_bwd": [_p, _p, _ll, _p, _p, _p, _p, _f, _p, _p, _p, _ll, _i, _i, _p],
    "fv_recon_loss": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p],
    "fv_recon_loss_flat": [_p, _p, _p, _p, _ll, _i, _f, _p, _p],
    "fv_adam_multi": [_p, _i, _ll, _f, _d, _d, _f, _p, _p],
    "fv_scale": [_p, _p, _i, _ll, _p, _f, _p],
    "fv_pw_moments": [_p, _p, _i, _i, _i, _p, _p],
    "fv_pw_prepare": [_p, _d, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _i, _i, _p],
    "fv_pw_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fv_pw_bwd_reduce": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p],
    "fv_pw_bwd_finalize": [_p, _p, _d, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "fv_bn_finalize_xrank": [_p, _p, _i, _i, _p, _i, _d, _p, _p, _p, _p, _f, _f, _p, _p, _p, _i, _i, _p],
    "fv_bn_stats_xrank": [_p, _i, _p, _ll, _i, _p, _p, _i, _i, _p, _d, _p, _p, _p, _p, _f, _f, _p, _p],
    "fv_bn_act_bwd_reduce_xrank": [_p, _i, _p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _d, _p, _p, _p, _p],
    "fv_bn_finalize_xrank_emulate": [_p, _p, _i, _p, _i, _d, _p, _p, _p, _p, _f, _f, _p, _p, _p, _i, _i, _p],
    "fv_grad_allreduce": [_p, _p, _i, _i, _ll, _p, _p, _f, _p],
    "fv_debug_mma_rate": [_i, _i, _i, _i, _i, _i, _p, _p],
    "fv_debug_trace_set": [_p],
}
_STR = ("fv_last_error", "fv_version")
_LL = ("fv_xrank_buffer_floats", "fv_reduce_ws_bytes", "fv_grad_allreduce_flag_words")
# predicates / planning queries: the return value is the answer
_PLAIN_INT = {"fv_outconv_supported": [_i, _i, _i, _i, _i, _i, _i], "fv_conv2d_fuses_stats": [_i, _i, _i, _i, _i, _i, _i, _i, _i],
              "fv_conv2d_wgrad_splits": [_i, _i, _i, _i, _i, _i, _i, _i],
              "fv_conv2d_geom_fuses_stats": [_i, _i, _i, _i, _i, _i, _i], "fv_outconv_wgrad_splits": [_i, _i, _i],
              "fv_reparam_kl_parts": [_i, _i], "fv_abi_version": [], "fv_weight_prep_block_items": [], "fv_weight_prep_tiled_blocks": [_i, _i, _i, _i, _i]}
ABI_VERSION = 2            # include/facevae_b200.h FV_ABI_VERSION

_lock = threading.Lock()
_lib = None


class FaceVaeError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """Load (once) and return the ctypes library handle."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise FaceVaeError(f"{LIB_PATH} is missing; run `python -m face_vae_b200.build` (needs nvcc)")
            _build.build()
        elif build_if_missing and _build.stale() and _build.have_nvcc():
            _build.build()            # sources changed since the library was built: never bind new signatures to an old binary
        lib = C.CDLL(LIB_PATH)
        lib.fv_abi_version.restype = C.c_int
        lib.fv_abi_version.argtypes = []
        if lib.fv_abi_version() != ABI_VERSION:
            raise FaceVaeError(f"{LIB_PATH} has ABI version {lib.fv_abi_version()}, this package binds version {ABI_VERSION}: "
                               "rebuild with `python -m face_vae_b200.build --force`")
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        for name in _STR:
            getattr(lib, name).restype = C.c_char_p
            getattr(lib, name).argtypes = []
        for name in _LL:
            getattr(lib, name).restype = C.c_longlong
            getattr(lib, name).argtypes = []
        for name, args in _PLAIN_INT.items():
            getattr(lib, name).restype = C.c_int
            getattr(lib, name).argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().fv_last_error().decode("utf-8", "replace")


launch_count = 0          # every int-returning entry point launches exactly one kernel
_profile = None


def profile_start() -> None:
    """Record a CUDA-event pair around every subsequent call (bench.py's per-kernel timing pass)."""
    global _profile
    _profile = []


def profile_stop():
    """-> list of (entry point, milliseconds, meta) for the calls since profile_start(); synchronises."""
    global _profile
    import torch
    torch.cuda.synchronize()
    out = [(n, e0.elapsed_time(e1), meta) for n, e0, e1, meta in (_profile or [])]
    _profile = None
    return out


def call(name: str, *args, meta=None):
    """Call an int-returning entry point; non-zero -> FaceVaeError with the library's message."""
    global launch_count
    fn = getattr(load(), name)
    if _profile is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _profile.append((name, e0, e1, meta))
    else:
        rc = fn(*args)
    launch_count += 1
    if rc != 0:
        raise FaceVaeError(f"{name} failed ({rc}): {last_error()}")
