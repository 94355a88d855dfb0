"""Data-parallel launch helpers with the reference's names and meanings (reference distributed.py:9-74), plus the
gradient reducer that replaces the DDP wrap of logger.py:55.

One process per GPU.  Two exchanges per step: (a) the gradient mean all-reduce -- by default ONE peer-memory kernel over
NVLink at the end of backward (csrc/fv_xrank.cu, ``fv_grad_allreduce``) on a flat, symmetric gradient buffer; NCCL buckets
launched from grad-ready hooks on a side stream are the fallback -- and (b) the per-layer batch-norm statistic exchange
issued from face_vae_b200.functional (face_vae_b200.xrank, NCCL all-reduce as the fallback).
"""
from __future__ import annotations

import functools
import random
from typing import Iterable, List, Optional

import numpy as np
import torch
import torch.distributed as dist


def init_seeds(cuda_deterministic=True):
    """seed = 1 + rank for python / numpy / torch (reference distributed.py:9-21).  The cudnn switches of the
    reference have no effect on this package's kernels but are kept for callers that mix in torch ops."""
    seed = 1 + get_rank()
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    import torch.backends.cudnn as cudnn
    cudnn.deterministic = bool(cuda_deterministic)
    cudnn.benchmark = not cuda_deterministic


def init_dist(local_rank, world_size, backend="nccl"):
    """Initialise the process group from env:// (reference distributed.py:24-31).  Unlike the reference this does
    not switch on autograd anomaly detection (distributed.py:26), which serialises every backward."""
    if dist.is_available():
        if not dist.is_initialized():
            if backend == "nccl":
                torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, init_method="env://", world_size=world_size, rank=local_rank)
    print("Rank", get_rank(), "initialized.")


def get_rank():
    rank = 0
    if dist.is_available():
        if dist.is_initialized():
            rank = dist.get_rank()
    return rank


def get_world_size():
    world_size = 1
    if dist.is_available():
        if dist.is_initialized():
            world_size = dist.get_world_size()
    return world_size


def master_only(func):
    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        if get_rank() == 0:
            return func(*args, **kwargs)
        else:
            return None

    return wrapper


def is_master():
    return get_rank() == 0


@master_only
def master_only_print(*args, **kwargs):
    print(*args, **kwargs)


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """One-time weight/buffer broadcast from rank ``src`` (what the DDP constructor does, logger.py:55)."""
    if get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


class GradientReducer:
    """Bucketed mean all-reduce of parameter gradients, overlapped with backward.

    All gradients live in ONE flat fp32 buffer laid out in reverse registration order (the order backward produces them);
    a bucket is a contiguous ~``bucket_mb`` slice of it.  A post-accumulate-grad hook counts ready parameters; when a
    bucket completes, its gradients are moved into their slots with one multi-tensor copy, ``p.grad`` is re-pointed at the
    slot views, and the slice is all-reduced IN PLACE on a side stream (NCCL, ``ReduceOp.AVG``) while backward continues.
    ``finish()`` only joins the side stream: there is no flatten (``torch.cat``) and no copy back, and the optimiser reads
    the averaged gradients straight from the flat buffer.
    Equivalent to DDP's reducer (the reference wraps every sub-model in DistributedDataParallel, logger.py:55) with
    smaller buckets: the default 25 MiB cap puts this 15 MB model in essentially one bucket, i.e. no overlap.

    Stream discipline (round-1 ADVICE, high): the collective is issued synchronously *on the side stream*, i.e. the side
    stream itself is ordered after NCCL's completion, and the compute stream joins the side stream in ``finish()`` --
    nothing reads a bucket that NCCL may still be writing, in eager mode as well as under CUDA-graph capture.
    """

    ALIGN = 64          # floats: slots start on 256-byte boundaries (vectorised optimiser loads)

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 2.0, process_group=None):
        self.group = process_group
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.world = get_world_size()
        self.buckets: List[List[torch.nn.Parameter]] = []
        # Peer-memory mode (default on NVLink boxes): the flat buffer is a symmetric allocation and ONE kernel at the end of
        # backward (fv_grad_allreduce: pull my slice from all ranks, add in rank order, push the mean to all ranks) replaces the
        # NCCL buckets -- measured at 2 GPUs NCCL cost 80 us as one exposed call and 105 us as overlapped 2 MB buckets, whose
        # CTAs displace the persistent one-CTA-per-SM convolution kernels they are meant to overlap with.
        self.peer = self._peer_mode_ok(process_group)
        if self.peer:
            bucket_mb = float("inf")
        cap = int(bucket_mb * (1 << 20)) if bucket_mb != float("inf") else (1 << 62)
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= cap:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {}
        self._slot = {}
        self._range = []
        off = 0
        layout = []
        for bi, b in enumerate(self.buckets):
            start = off
            for p in b:
                self._bucket_of[id(p)] = bi
                layout.append((p, off))
                off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            self._range.append((start, off))
        dev = self.params[0].device if self.params else torch.device("cpu")
        off = (max(off, 1) + 3) // 4 * 4
        self._symm = None
        if self.peer:
            try:
                self._setup_peer(off, dev)
            except Exception as e:   # allocator / rendezvous not available: NCCL buckets
                if get_rank() == 0:
                    print(f"face_vae_b200: peer-memory gradient all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
                self.peer = False
            ok = torch.tensor([1 if self.peer else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # all ranks must take the same path
            self.peer = bool(int(ok.item()))
        if not self.peer:
            self.flat = torch.zeros((off,), dtype=torch.float32, device=dev)     # padding stays zero
        for p, o in layout:
            self._slot[id(p)] = self.flat[o:o + p.numel()].view(p.shape)
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._handles = []
        self.stream: Optional[torch.cuda.Stream] = None
        self.launched = 0
        if self.world > 1:
            for p in self.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def _peer_mode_ok(self, process_group) -> bool:
        import os
        if os.environ.get("FACEVAE_GRAD_XRANK", "1") == "0" or process_group is not None:
            return False
        if not (dist.is_available() and dist.is_initialized() and self.world > 1 and torch.cuda.is_available()):
            return False
        if dist.get_backend() != "nccl" or self.world > 16 or not self.params or not self.params[0].is_cuda:
            return False
        return True

    def _setup_peer(self, n: int, dev) -> None:
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        words = int(_lib.load().fv_grad_allreduce_flag_words())
        self.flat = symm_mem.empty(n, dtype=torch.float32, device=dev)
        self.flat.zero_()
        self._flags = symm_mem.empty(words, dtype=torch.int64, device=dev)
        self._flags.zero_()
        torch.cuda.synchronize()
        h_flat = symm_mem.rendezvous(self.flat, dist.group.WORLD)
        h_flags = symm_mem.rendezvous(self._flags, dist.group.WORLD)
        self._symm = (h_flat, h_flags)
        self._bufs_dev = int(h_flat.buffer_ptrs_dev)
        self._flags_dev = int(h_flags.buffer_ptrs_dev)
        self._epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        h_flags.barrier()
        torch.cuda.synchronize()

    def _hook(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        bucket = self.buckets[bi]
        lo, hi = self._range[bi]
        if self.params and self.params[0].is_cuda:
            from . import ops
            ops.join_wgrad_stream()       # weight gradients issued on the side stream are complete before they are gathered
        # gather this bucket's gradients into their slots (one multi-tensor copy on the compute stream, right behind the
        # kernels that produced them) and make the slots the parameters' .grad
        src = [p.grad for p in bucket if p.grad is not None and p.grad.data_ptr() != self._slot[id(p)].data_ptr()]
        dst = [self._slot[id(p)] for p in bucket if p.grad is not None and p.grad.data_ptr() != self._slot[id(p)].data_ptr()]
        if src:
            torch._foreach_copy_(dst, src)
        for p in bucket:
            if p.grad is None:                       # received no gradient this step: contributes zeros
                self._slot[id(p)].zero_()
            p.grad = self._slot[id(p)]
        piece = self.flat[lo:hi]
        if self.peer:
            from . import _lib
            _lib.call("fv_grad_allreduce", self._bufs_dev, self._flags_dev, get_rank(), self.world, self.flat.numel(), self._epoch.data_ptr(),
                      self._ticket.data_ptr(), 1.0 / self.world, torch.cuda.current_stream().cuda_stream)
        elif piece.is_cuda:
            if self.stream is None:
                self.stream = torch.cuda.Stream()
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                # synchronous call = the SIDE stream waits for the collective; the host does not block
                dist.all_reduce(piece, op=dist.ReduceOp.AVG, group=self.group)
        else:   # gloo (CPU tests): no AVG, no streams
            dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group)
            piece /= self.world
        self._launched[bi] = True
        self.launched += 1

    def finish(self) -> None:
        """Call after backward(): launches the buckets whose parameters received no gradient this step and joins the side
        stream, after which every ``p.grad`` (a view of the flat buffer) holds the gradient averaged over the ranks."""
        if self.world == 1:
            return
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:
                self._launch(bi)
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()
