"""Objective and train step for the VAE path, following the reference's contracts:

* ``VAEGeneratorFull.forward`` returns a loss dict whose entries are already weight-multiplied, like
  ``GeneratorFull.forward`` does for its "K" and "R" entries (reference trainer.py:240-252, 300-317); the intended
  weights K: 0.2, R: 10 are the commented values at trainer.py:250-251 (the shipped defaults are 0, which switches the
  VAE terms off).
* ``VAETrainer.step`` is the skeleton of ``Logger.step`` (reference logger.py:150-164): zero_grad -> forward ->
  sum(losses.values()).backward() -> optimizer.step(), with Adam(lr, betas=(0.5, 0.999)) as at logger.py:60.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from . import distributed as fdist
from .models import FaceVAE


class VAEGeneratorFull(nn.Module):
    def __init__(self, vae: FaceVAE, weights: Optional[Dict[str, float]] = None, l1: bool = False):
        super().__init__()
        self.vae = vae
        self.weights = {"K": 0.2, "R": 10.0}
        if weights:
            self.weights.update(weights)
        self.l1 = l1

    def forward(self, d: torch.Tensor, eps: Optional[torch.Tensor] = None, train_vae: bool = True):
        """d: driving frames [N,3,H,W] fp32 in [0,1].  -> (loss_dict, generated_d, mu, logstd)."""
        if not train_vae:
            _, _, x_hat = self.vae(d, False)
            zero = torch.zeros((), device=d.device)
            return {"K": zero, "R": zero}, x_hat, None, None
        out = self.vae.forward_loss(d, eps, self.l1)
        loss = {"K": self.weights["K"] * out["K"], "R": self.weights["R"] * out["R"]}
        return loss, out["x_hat"], out["mu"], out["logstd"]


class GeneratorFull(nn.Module):
    """Drop-in for the reference's ``GeneratorFull`` (reference trainer.py:214-317) on the VAE path.

    Same constructor parameter names, same ``weights`` dict (keys P, G, F, E, L, H, D, C, K, R), same
    ``forward(s, d, s_a=None, d_a=None, train_vae=None)`` and the same 8-tuple
    ``(loss_dict, generated_d, transformed_d, kp_s, kp_d, transformed_kp, occlusion, mask)``, so the reference's
    ``Logger.step`` (logger.py:150-164: ``losses_g, generated_d, ... = self.g_full(s, d, s_a, d_a, train_vae)``;
    ``sum(losses_g.values()).backward()``) runs unchanged.  What is computed is the hot path of SURVEY.md section 8:
    ``generator`` is the image -> image VAE (``FaceVAE``), "K" is the weighted KL term of the DRIVING frame's (mu, logstd)
    (trainer.py:312) and "R" the weighted reconstruction loss of ``(d, generated_d)`` (trainer.py:314); the keypoint /
    motion / GAN / perceptual networks are outside the path (they need 3-D blocks, downloaded VGG weights and the
    git-ignored hopenet checkpoint), so their sub-modules may be ``None``, their loss entries are zero tensors of the
    reference's shape ``[1]`` and their outputs ``None``.  As in the reference, a falsy ``train_vae`` yields zero K / R
    (the bottleneck then returns ``(None, None, x_hat)``, models.py:567-570).
    The shipped reference weights K: 0, R: 0 switch the VAE terms off (trainer.py:250-251); the defaults here are the
    values commented next to them (0.2 and 10), since this class exists to train that path.  ``eps`` (optional attribute or
    keyword) injects the re-parameterisation noise the reference draws inline with ``torch.randn`` (models.py:561)."""

    OUT_OF_SCOPE = ("P", "G", "F", "E", "L", "H", "D", "C")

    def __init__(self, efe=None, afe=None, ckd=None, hpe_ede=None, mfe=None, generator: Optional[FaceVAE] = None, discriminator=None,
                 pretrained_path=None, n_bins=66):
        super().__init__()
        if generator is None:
            raise ValueError("GeneratorFull: `generator` must be the FaceVAE of the hot path")
        self.efe, self.afe, self.ckd, self.hpe_ede, self.mfe = efe, afe, ckd, hpe_ede, mfe
        self.generator = generator
        self.discriminator = discriminator
        self.weights = {"P": 10, "G": 1, "F": 10, "E": 20, "L": 10, "H": 20, "D": 0.5, "C": 10, "K": 0.2, "R": 10}
        self.eps = None
        self.l1 = False

    def forward(self, s, d, s_a=None, d_a=None, train_vae=None, eps: Optional[torch.Tensor] = None):
        zero = lambda: torch.zeros(1, device=d.device)          # the reference's `torch.Tensor([0.0]).cuda()`
        loss = {k: zero() for k in self.OUT_OF_SCOPE}
        if not train_vae:
            _, _, generated_d = self.generator(d, False)
            loss["K"], loss["R"] = zero(), zero()
        else:
            out = self.generator.forward_loss(d, self.eps if eps is None else eps, self.l1)
            generated_d = out["x_hat"]
            loss["K"] = self.weights["K"] * out["K"]
            loss["R"] = self.weights["R"] * out["R"]
        return loss, generated_d, None, None, None, None, None, None


class VAETrainer:
    """One-model version of the reference Logger's optimisation step (logger.py:52-63, 150-164).

    ``use_cuda_graph`` (default: on for CUDA parameters; FACEVAE_CUDA_GRAPH=0 switches it off, FACEVAE_CUDA_GRAPH_DDP=0 only
    under data parallelism): the whole step -- zero_grad, forward, backward, gradient all-reduce, Adam -- is captured once per
    input shape into a CUDA graph and replayed, which removes the ~160 per-launch host round trips of the eager step.  The
    captured sequence is exactly the eager one: same kernels, same dependencies -- including the fork / join of the
    weight-gradient side stream (ops.wgrad_stream) -- on a high-priority capture stream."""

    def __init__(self, vae: FaceVAE, lr: float = 5e-5, weights: Optional[Dict[str, float]] = None, bucket_mb: float = 2.0,
                 fused_adam: bool = True, use_cuda_graph: Optional[bool] = None):
        import os
        self.g_full = VAEGeneratorFull(vae, weights)
        self.vae = vae
        fdist.broadcast_parameters(vae)
        params = list(vae.parameters())
        on_cuda = bool(params) and params[0].is_cuda
        world = fdist.get_world_size()
        if use_cuda_graph is None:
            use_cuda_graph = on_cuda and os.environ.get("FACEVAE_CUDA_GRAPH", "1") != "0"
            if world > 1:   # capturing the NCCL collectives (SyncBN statistics, gradient buckets) is opt-out as well
                use_cuda_graph = use_cuda_graph and os.environ.get("FACEVAE_CUDA_GRAPH_DDP", "1") != "0"
        self.use_cuda_graph = bool(use_cuda_graph) and on_cuda
        kw = {}
        if fused_adam and on_cuda and os.environ.get("FACEVAE_FUSED_ADAM", "1") != "0":
            from .optim import FusedAdam              # one multi-tensor launch, device-resident step count (graph-capturable)
            self.optimizer = FusedAdam(params, lr=lr, betas=(0.5, 0.999))
        else:
            if fused_adam and on_cuda:
                kw["fused"] = True
                if self.use_cuda_graph:
                    kw["capturable"] = True
            self.optimizer = torch.optim.Adam(params, lr=lr, betas=(0.5, 0.999), **kw)
        bucket_mb = float(os.environ.get("FACEVAE_BUCKET_MB", bucket_mb))
        self.reducer = fdist.GradientReducer(params, bucket_mb) if world > 1 else None
        if os.environ.get("FACEVAE_DIAG_SKIP_GRAD_REDUCE") == "1":      # timing diagnostic only: ranks diverge
            if self.reducer is not None:
                self.reducer.remove()
            self.reducer = None
        self._graph = None
        self._graph_key = None
        self._cap_stream = None
        self.launches_per_step = None            # C-ABI kernel launches inside one captured step

    # -- eager ----------------------------------------------------------------------------------------------------
    def _eager_step(self, d: torch.Tensor, eps: Optional[torch.Tensor]):
        from . import ops
        self.optimizer.zero_grad(set_to_none=True)
        with ops.step_scope(self.vae):          # one memset for all accumulators, one launch for all filter operands
            losses, generated, mu, logstd = self.g_full(d, eps, True)
            total = sum(losses.values())
            total.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        return losses, generated

    # -- CUDA graph -----------------------------------------------------------------------------------------------
    def _capture(self, d: torch.Tensor, eps: torch.Tensor) -> None:
        import copy
        from . import _lib
        self._static_d = d.clone()
        self._static_eps = eps.clone()
        # warm-up on a side stream (lazy initialisation, allocator pools) without changing the training state
        snap_model = copy.deepcopy(self.vae.state_dict())
        # optimizer state must EXIST before capture (tensors created inside the capture would be re-initialised by every
        # replay), so it is restored in place afterwards: copied back if there was state, zeroed if there was none
        snap_opt = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                    for p, st in self.optimizer.state.items()}
        # warm-up AND capture run on one dedicated stream, so that everything keyed by stream (the reduction workspace of
        # ops._red_ws, the NCCL side stream's dependencies) already exists when the capture starts: nothing is allocated
        # or zero-initialised inside the graph
        if self._cap_stream is None:
            import os
            # high priority: the kernels of the dependent chain get free SMs before the weight-gradient kernels queued on the
            # (default-priority) side stream of ops.wgrad_stream -- 3.34 -> 3.23 ms per step
            self._cap_stream = torch.cuda.Stream(priority=int(os.environ.get("FACEVAE_CAP_PRIORITY", "-1")))
        side = self._cap_stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager_step(self._static_d, self._static_eps)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.vae.load_state_dict(snap_model)
        for p, st in self.optimizer.state.items():
            for k, v in st.items():
                if torch.is_tensor(v):
                    if p in snap_opt and k in snap_opt[p]:
                        old = snap_opt[p][k]
                        v.copy_(old) if torch.is_tensor(old) else v.fill_(float(old))   # a loaded checkpoint may hold python numbers
                    else:
                        v.zero_()
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        l0 = _lib.launch_count
        from . import ops
        with torch.cuda.graph(graph, stream=self._cap_stream):
            with ops.step_scope(self.vae):
                losses, generated, _, _ = self.g_full(self._static_d, self._static_eps, True)
                total = sum(losses.values())
                total.backward()
            if self.reducer is not None:
                self.reducer.finish()
            self.optimizer.step()
        self.launches_per_step = _lib.launch_count - l0
        self._graph, self._static_out = graph, (losses, generated)
        self._graph_key = (tuple(d.shape), tuple(eps.shape))

    def _graph_step(self, d: torch.Tensor, eps: Optional[torch.Tensor]):
        if eps is None:
            eps = torch.randn((d.shape[0], self.vae.latent_dim(d.shape[2], d.shape[3])), device=d.device)
        eps = eps.reshape(d.shape[0], -1)
        if self._graph is None or self._graph_key != (tuple(d.shape), tuple(eps.shape)):
            try:
                self._capture(d.float().contiguous(), eps.float().contiguous())
            except Exception as e:      # e.g. an autograd graph from an earlier eager pass is still referenced by the caller
                import warnings
                warnings.warn(f"face_vae_b200: CUDA-graph capture of the train step failed ({type(e).__name__}: {e}); "
                              "continuing with eager launches")
                self.use_cuda_graph = False
                self._graph = None
                torch.cuda.synchronize()
                return self._eager_step(d, eps)
        self._static_d.copy_(d, non_blocking=True)
        self._static_eps.copy_(eps, non_blocking=True)
        self._graph.replay()
        return self._static_out

    # -- checkpoints (reference Logger.save_cpk / load_cpk, logger.py:93-115) ---------------------------------------------
    def checkpoint(self, name: str = "vae", epoch: int = 0) -> dict:
        """The reference's checkpoint dict for one model: {name: state_dict, "optimizer_" + name: ..., "epoch": epoch}."""
        return {name: self.vae.state_dict(), "optimizer_" + name: self.optimizer.state_dict(), "epoch": epoch}

    def save_cpk(self, path: str, name: str = "vae", epoch: int = 0) -> None:
        if fdist.is_master():
            torch.save(self.checkpoint(name, epoch), path)

    def load_cpk(self, path_or_dict, name: str = "vae") -> int:
        """Loads model and optimiser state written by save_cpk or by the reference's Logger; -> the epoch to resume at
        (checkpoint epoch + 1, logger.py:115).  A captured CUDA graph is dropped: the optimiser state tensors are new."""
        ckp = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location=torch.device("cpu"))
        self.vae.load_state_dict(ckp[name])
        self.optimizer.load_state_dict(ckp["optimizer_" + name])
        self._graph = None
        self._graph_key = None
        return int(ckp["epoch"]) + 1

    def step(self, d: torch.Tensor, eps: Optional[torch.Tensor] = None):
        """-> (loss dict {"K", "R"} already weight-multiplied, generated frames).  With CUDA graphs the returned tensors are
        the graph's static outputs: they are overwritten by the next step."""
        if self.use_cuda_graph and d.is_cuda:
            return self._graph_step(d, eps)
        return self._eager_step(d, eps)


class GraphedInference:
    """Eval-mode ``vae(x, train_vae, eps)`` (encode -> sample -> decode, reference models.py:550-570 with running statistics)
    replayed from a CUDA graph per input shape: the eager forward issues ~60 launches of 5-20 us and is bound by the host at small
    batch (1.4 ms at batch 1); a replay costs one launch.  Returns the graph's STATIC output tensors ``(mu, logstd, x_hat)`` --
    valid until the next call with the same shape; clone what must outlive it.  ``eps`` None draws fresh noise per call."""

    def __init__(self, vae: FaceVAE):
        self.vae = vae
        self._graphs = {}
        self._stream = None

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, train_vae: bool = True, eps: Optional[torch.Tensor] = None):
        if self.vae.training:
            raise RuntimeError("GraphedInference replays the eval-mode forward: call vae.eval() first (training-mode batch norm updates "
                               "running statistics, which a replayed graph would apply again on every call)")
        if not x.is_cuda:
            raise RuntimeError("GraphedInference needs CUDA tensors: face_vae_b200 has no CPU path")
        key = (tuple(x.shape), bool(train_vae))
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._graphs[key] = self._capture(x.float().contiguous(), bool(train_vae))
        graph, sx, seps, outs = ent
        sx.copy_(x)
        if seps is not None:
            if eps is None:
                seps.normal_()
            else:
                seps.copy_(eps.reshape(seps.shape))
        graph.replay()
        return outs

    def _capture(self, x: torch.Tensor, train_vae: bool):
        sx = x.clone()
        seps = None
        if train_vae:
            seps = torch.randn((x.shape[0], self.vae.latent_dim(x.shape[2], x.shape[3])), device=x.device)
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        side = self._stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up on the capture stream: per-stream workspaces exist before the capture
            for _ in range(2):
                self.vae(sx, train_vae, seps)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            outs = self.vae(sx, train_vae, seps)
        return graph, sx, seps, outs
