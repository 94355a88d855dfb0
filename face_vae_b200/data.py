"""Host -> device input staging for the train loop.

The reference moves every batch with ``x.cuda(non_blocking=True)`` on the compute stream right before the step
(reference logger.py:144-148) and reads the losses back with ``.cpu()`` after it (logger.py:173), which serialises the
PCIe copy, the step and the read-back.  ``DevicePrefetcher`` issues the copy of batch i+1 on a side stream while step i
runs; ``AsyncScalarLog`` turns the per-step loss read into an asynchronous copy into pinned memory that is consumed
after the loop (or one step late), so nothing in the loop waits for the GPU.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence, Tuple

import torch


class DevicePrefetcher:
    """Iterates over (pinned) host batches, yielding device tensors whose H2D copy overlapped the previous step.

    The device side is a fixed ring of ``depth`` staging buffers per tensor (allocated on first use, re-allocated only when
    a shape changes): no allocation happens in the loop -- a ``cudaMalloc`` would synchronise the device -- and the host
    can run at most ``depth - 1`` batches ahead of the GPU: before a slot is overwritten the host waits for the event
    recorded after the step that consumed it.  The yielded tensors are valid until the next batch is requested."""

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]] = (), device=None, depth: int = 3):
        self.batches = batches
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.depth = max(2, int(depth))
        self._slots = [None] * self.depth          # per slot: tuple of device tensors
        self._consumed = [None] * self.depth       # per slot: event recorded on the compute stream after its last use
        self._n = 0

    def over(self, batches: Iterable[Sequence[torch.Tensor]]) -> "DevicePrefetcher":
        """Iterate another stream of batches with the same staging buffers (``for x in pf.over(loader): ...``)."""
        self.batches = batches
        return self

    def _issue(self, batch):
        slot = self._n % self.depth
        self._n += 1
        if self._consumed[slot] is not None:
            self._consumed[slot].synchronize()     # bounds the host's run-ahead; normally long complete
        bufs = self._slots[slot]
        if bufs is None or len(bufs) != len(batch) or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, batch)):
            bufs = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch)
            self._slots[slot] = bufs
        with torch.cuda.stream(self.stream):
            for b, t in zip(bufs, batch):
                b.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return bufs, ev, slot

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        it = iter(self.batches)
        try:
            nxt = self._issue(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dev, ev, slot = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            try:
                nxt = self._issue(next(it))
            except StopIteration:
                nxt = None
            yield dev
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))      # the consumer's work on `dev` has been enqueued
            self._consumed[slot] = done


class AsyncScalarLog:
    """Per-step device scalars -> pinned host buffer without a synchronisation in the loop."""

    def __init__(self, capacity: int):
        self.buf = torch.empty((capacity,), dtype=torch.float32).pin_memory()
        self.n = 0

    def push(self, value: torch.Tensor) -> None:
        self.buf[self.n].copy_(value.detach().reshape(()), non_blocking=True)
        self.n += 1

    def values(self) -> torch.Tensor:
        torch.cuda.current_stream().synchronize()
        return self.buf[:self.n].clone()
