"""Adam as one multi-tensor launch (the per-model Adam of reference logger.py:60 as a fused optimiser).

``torch.optim.Adam(fused=True)`` walks this model's 52 parameter tensors in three multi-tensor launches of ~26 us each;
``FusedAdam`` issues ONE kernel (``fv_adam_multi``) over a device table of (param, grad, exp_avg, exp_avg_sq, n)
descriptors.  Same update rule and state layout (``exp_avg``, ``exp_avg_sq``, ``step``) as torch.optim.Adam without amsgrad /
weight decay, so optimiser state dicts saved by the reference's Logger (logger.py:93-101) load unchanged.  The step
count lives on the device, which makes the step capturable in a CUDA graph.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        # descriptor tables: one set for eager steps and one for steps recorded into a CUDA graph (the graph's memcpy node reads
        # its pinned source again at every replay, so an eager step in between must not overwrite it)
        self._bufs = {}
        self._check_steps = False

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._check_steps = True

    def _init_state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or \
                        not p.grad.is_contiguous():
                    raise _lib.FaceVaeError("FusedAdam: contiguous fp32 CUDA parameters and gradients only (no CPU path)")
            states = [self._init_state(p) for p in ps]
            for p, st in zip(ps, states):          # state loaded from a torch.optim.Adam checkpoint: python / CPU step counts
                if not torch.is_tensor(st["step"]) or st["step"].device != p.device or st["step"].dtype != torch.float32:
                    st["step"] = torch.tensor(float(st["step"]), dtype=torch.float32, device=p.device)
            # one bias correction per launch: all tensors of a group are assumed to share their step count (they do when
            # every parameter receives a gradient every step, which holds for this model; torch.optim.Adam state that
            # disagrees is refused rather than silently mis-corrected -- checked only when state was loaded from outside)
            steps = [st["step"] for st in states]
            if self._check_steps:
                vals = {float(t) for t in steps}
                if len(vals) > 1:
                    raise _lib.FaceVaeError(f"FusedAdam: parameters of one group have different step counts {sorted(vals)}")
                self._check_steps = False
            torch._foreach_add_(steps, 1.0)
            key = tuple((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr())
                        for p, st in zip(ps, states))
            # descriptor tables are per (parameter group, mode): a graph's memcpy node re-reads its pinned source at every
            # replay, so eager steps and other groups must never write to it
            mode = "capture" if torch.cuda.is_current_stream_capturing() else "eager"
            nbytes = 40 * len(ps)
            buf = self._bufs.get((gi, mode))
            if buf is None or buf["host"].numel() != nbytes:
                buf = {"host": torch.empty(nbytes, dtype=torch.uint8).pin_memory(),
                       "table": torch.empty(nbytes, dtype=torch.uint8, device=ps[0].device), "key": None, "copied": None,
                       "max_n": 1}
                self._bufs[(gi, mode)] = buf
            if key != buf["key"]:
                rec = np.zeros((len(ps),), dtype=np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8")]))
                for i, (p, st) in enumerate(zip(ps, states)):
                    rec[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
                if buf["copied"] is not None:
                    buf["copied"].synchronize()          # the previous step's async H2D copy has read the pinned table
                buf["host"].copy_(torch.from_numpy(rec.view(np.uint8).copy()))
                buf["table"].copy_(buf["host"], non_blocking=True)   # stream-ordered (a memcpy node under graph capture)
                if mode == "eager":
                    buf["copied"] = torch.cuda.Event()
                    buf["copied"].record()
                buf["max_n"] = max(p.numel() for p in ps)
                buf["key"] = key
            b1, b2 = group["betas"]
            _lib.call("fv_adam_multi", buf["table"].data_ptr(), len(ps), buf["max_n"], float(group["lr"]), float(b1), float(b2),
                      float(group["eps"]), steps[0].data_ptr(), torch.cuda.current_stream().cuda_stream)
        return loss
