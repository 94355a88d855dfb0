"""Data-parallel launch helpers with the reference's names and meanings (reference distributed.py:9-74), plus the
gradient reducer that replaces the DDP wrap of logger.py:55.

One process per GPU; NCCL over NVLink/NVSwitch carries (a) the bucketed gradient all-reduce, launched from
grad-ready hooks on a side stream so it overlaps the rest of backward, and (b) the per-layer batch-norm statistic
all-reduces issued from face_vae_b200.functional.
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

    Parameters are packed into buckets of ~``bucket_mb`` in reverse registration order (the order backward produces
    them).  A post-accumulate-grad hook counts ready parameters; when a bucket completes, its gradients are flattened
    and all-reduced (AVG) asynchronously on a side stream.  ``finish()`` waits and scatters the results back.
    Equivalent to DDP's reducer (the reference wraps every sub-model in DistributedDataParallel, logger.py:55) with
    smaller buckets: the default 25 MiB cap puts this 15 MB model in essentially one bucket, i.e. no overlap.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 2.0, process_group=None):
        self.group = process_group
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.world = get_world_size()
        self.buckets: List[List[torch.nn.Parameter]] = []
        cap = int(bucket_mb * (1 << 20))
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
        for bi, b in enumerate(self.buckets):
            for p in b:
                self._bucket_of[id(p)] = bi
        self._pending = [len(b) for b in self.buckets]
        self._inflight = []
        self._handles = []
        self.stream: Optional[torch.cuda.Stream] = None
        self.launched = 0
        if self.world > 1:
            for p in self.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def _hook(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        bucket = self.buckets[bi]
        grads = [p.grad for p in bucket]
        if grads[0].is_cuda:
            if self.stream is None:
                self.stream = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                flat = torch.cat([g.reshape(-1) for g in grads])
                if not torch.cuda.is_current_stream_capturing():
                    for g in grads:
                        g.record_stream(self.stream)
                work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:   # gloo (CPU tests): no AVG, no streams
            flat = torch.cat([g.reshape(-1) for g in grads])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight.append((bi, flat, work))
        self.launched += 1

    def finish(self) -> None:
        """Call after backward(): waits for every bucket and writes the averaged gradients back into ``.grad``."""
        if self.world == 1:
            return
        for bi, n in enumerate(self._pending):     # buckets with parameters that received no gradient this step
            if n > 0:
                for p in self.buckets[bi]:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                self._launch(bi)
        for bi, flat, work in self._inflight:
            work.wait()
            bucket = self.buckets[bi]
            ctx = torch.cuda.stream(self.stream) if (flat.is_cuda and self.stream is not None) else _null()
            with ctx:
                if not flat.is_cuda:
                    flat /= self.world
                off = 0
                views = []
                for p in bucket:
                    n = p.numel()
                    views.append(flat[off:off + n].view_as(p.grad))
                    off += n
                torch._foreach_copy_([p.grad for p in bucket], views)
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self._inflight.clear()
        self._pending = [len(b) for b in self.buckets]

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
