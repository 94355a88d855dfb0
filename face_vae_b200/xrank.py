"""Symmetric (NVLink peer-mapped) buffers for the fused cross-rank batch-norm statistic exchange.

PyTorch's symmetric-memory allocator is used only to obtain a buffer that every rank of the node can address
(cuMem allocation + handle exchange through the process group's store); the exchange itself is the single-block
kernel ``fv_bn_finalize_xrank`` (csrc/fv_xrank.cu), not a library collective.  When the buffers cannot be set up
(world size 1, non-CUDA backend, allocator unavailable) ``get()`` returns None and the callers use a NCCL all-reduce.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib

_state = {"tried": False, "xchg": None}


class StatExchange:
    def __init__(self):
        import torch.distributed._symmetric_memory as symm_mem
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        n = int(_lib.load().fv_xrank_buffer_floats())
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
        self.buf.zero_()
        torch.cuda.synchronize()
        self.handle = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.peer_ptrs_dev = int(self.handle.buffer_ptrs_dev)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        self.handle.barrier()
        torch.cuda.synchronize()

    def finalize_fwd(self, sums, count, gamma, beta, running_mean, running_var, momentum, eps):
        c = gamma.numel()
        stat = torch.empty((4, c), device=sums.device, dtype=torch.float32)
        _lib.call("fv_bn_finalize_xrank", sums.data_ptr(), self.peer_ptrs_dev, self.rank, self.world, self.epoch.data_ptr(), 0,
                  float(count), gamma.data_ptr(), beta.data_ptr(), None if running_mean is None else running_mean.data_ptr(),
                  None if running_var is None else running_var.data_ptr(), momentum, eps, stat.data_ptr(), None, None, 0, c,
                  torch.cuda.current_stream().cuda_stream)
        return stat

    def stats_finalize_fwd(self, y, count, gamma, beta, running_mean, running_var, momentum, eps):
        """Statistics of y, exchange and finalize in ONE launch (the reduction's last block does the exchange) -> stat [4, C]."""
        from . import ops
        c = y.shape[-1]
        sums = torch.empty((2 * c,), device=y.device, dtype=torch.float32)
        stat = torch.empty((4, c), device=y.device, dtype=torch.float32)
        _lib.call("fv_bn_stats_xrank", y.data_ptr(), ops._dt(y), sums.data_ptr(), y.numel() // c, c, ops._red_ws(), self.peer_ptrs_dev, self.rank,
                  self.world, self.epoch.data_ptr(), float(count), gamma.data_ptr(), beta.data_ptr(),
                  None if running_mean is None else running_mean.data_ptr(), None if running_var is None else running_var.data_ptr(), momentum, eps,
                  stat.data_ptr(), torch.cuda.current_stream().cuda_stream, meta=ops._bytes(y))
        return stat

    def reduce_finalize_bwd(self, y, g, stat, mode, act, g_nchw, count):
        """Backward sums of a norm + act layer, exchange and finalize in ONE launch -> (dgamma, dbeta, coef [2, C])."""
        from . import ops
        n, h, w, c = y.shape
        sums = torch.empty((2 * c,), device=y.device, dtype=torch.float32)
        dgamma = torch.empty((c,), device=y.device, dtype=torch.float32)
        dbeta = torch.empty((c,), device=y.device, dtype=torch.float32)
        coef = torch.empty((2, c), device=y.device, dtype=torch.float32)
        _lib.call("fv_bn_act_bwd_reduce_xrank", y.data_ptr(), ops._dt(y), g.data_ptr(), ops._dt(g), int(g_nchw), stat.data_ptr(), sums.data_ptr(), n, h, w, c,
                  mode, act, ops._red_ws(), self.peer_ptrs_dev, self.rank, self.world, self.epoch.data_ptr(), float(count), coef.data_ptr(),
                  dgamma.data_ptr(), dbeta.data_ptr(), torch.cuda.current_stream().cuda_stream, meta=ops._bytes(y, g))
        return dgamma, dbeta, coef

    def finalize_bwd(self, sums_local, count, c):
        dgamma = torch.empty((c,), device=sums_local.device, dtype=torch.float32)
        dbeta = torch.empty((c,), device=sums_local.device, dtype=torch.float32)
        coef = torch.empty((2, c), device=sums_local.device, dtype=torch.float32)
        _lib.call("fv_bn_finalize_xrank", sums_local.data_ptr(), self.peer_ptrs_dev, self.rank, self.world, self.epoch.data_ptr(), 1,
                  float(count), None, None, None, None, 0.0, 0.0, coef.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), 0, c,
                  torch.cuda.current_stream().cuda_stream)
        return dgamma, dbeta, coef


def get() -> Optional[StatExchange]:
    """The process-wide exchange object, or None when the NCCL fallback should be used."""
    if _state["tried"]:
        return _state["xchg"]
    _state["tried"] = True
    if os.environ.get("FACEVAE_XRANK", "1") == "0":
        return None
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and torch.cuda.is_available()):
        return None
    if dist.get_backend() != "nccl" or dist.get_world_size() > 16:
        return None
    try:
        _state["xchg"] = StatExchange()
    except Exception as e:   # allocator / rendezvous not available on this system: keep the NCCL path
        if dist.get_rank() == 0:
            print(f"face_vae_b200: peer-memory statistic exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
        _state["xchg"] = None
    # all ranks must agree, otherwise some would wait in the kernel and others in NCCL
    ok = torch.tensor([1 if _state["xchg"] is not None else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        _state["xchg"] = None
    return _state["xchg"]
