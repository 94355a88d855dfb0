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
    """Iterates over (pinned) host batches, yielding device tensors whose H2D copy overlapped the previous step."""

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device=None):
        self.batches = batches
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(self.device)

    def _issue(self, batch):
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in batch)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        it = iter(self.batches)
        try:
            nxt = self._issue(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dev, ev = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in dev:
                t.record_stream(cur)
            try:
                nxt = self._issue(next(it))
            except StopIteration:
                nxt = None
            yield dev


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
