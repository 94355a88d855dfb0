"""Host-side logic that needs no GPU: shape predicates and planning queries of the C ABI, argument checking of the entry
points, the per-step scope, and the optimiser's refusal of CPU tensors (the package has no CPU compute path)."""
import ctypes

import pytest
import torch


@pytest.fixture(scope="module")
def lib():
    from face_vae_b200 import _lib
    return _lib.load()


def test_outconv_shape_predicate(lib):
    ok = lib.fv_outconv_supported
    assert ok(32, 256, 256, 32, 3, 7, 7) == 1          # BASELINE.json configs[1]: batch 32 at 256x256
    assert ok(8, 128, 128, 32, 4, 7, 7) == 1
    assert ok(8, 64, 64, 32, 3, 7, 7) == 0             # W = 64: generic kernels (the CPU-anchor size of the parity tests)
    assert ok(8, 512, 512, 32, 3, 7, 7) == 0           # W = 512: generic kernels
    assert ok(8, 256, 256, 64, 3, 7, 7) == 0           # 64 input channels
    assert ok(8, 256, 256, 32, 5, 7, 7) == 0           # more than 4 output channels
    assert ok(8, 256, 256, 32, 3, 3, 3) == 0           # not a 7x7 filter


def test_conv_stats_fusion_predicate_on_the_anchor_layers(lib):
    """Which layers of the anchor produce their batch-norm sums in the conv epilogue: the 32-channel ring
    layer and the K >= 1152 layers; the 64-channel ring layer and the K = 576 / N = 128 layer keep a separate pass."""
    f = lambda ci, co, hw: lib.fv_conv2d_fuses_stats(0, 32, hw, hw, ci, co, 3, 3, 0)
    assert f(64, 32, 256) == 1 and f(256, 256, 32) == 1 and f(256, 128, 64) == 1 and f(128, 64, 128) == 1 and f(256, 32, 32) == 1
    assert f(32, 64, 256) == 0 and f(64, 128, 128) == 0 and f(128, 256, 64) == 0
    assert lib.fv_conv2d_fuses_stats(2, 32, 64, 64, 128, 256, 3, 3, 0) == 0        # NCHW fp32 output: never


def test_new_entry_points_reject_bad_arguments_without_a_gpu(lib):
    assert lib.fv_outconv_fwd(None, None, None, None, None, None, None, None, None, 1, 8, 128, 32, 3, 0, 1, 1.0, None, None) != 0
    assert b"null pointer" in lib.fv_last_error()
    assert lib.fv_outconv_dgrad(None, None, None, None, 1, 8, 128, 32, 3, None) != 0
    assert lib.fv_outconv_wgrad(None, None, None, None, 1, 1, 8, 128, 32, 3, None) != 0
    assert lib.fv_outconv_prep(None, None, None, 3, 32, None) != 0
    assert lib.fv_conv2d_stats(None, None, None, None, None, 0, 1, 8, 8, 16, 16, 16, 3, 3, 1, None, None, None) != 0
    assert lib.fv_conv2d_x2(None, None, None, None, 0, 1, 8, 8, 16, 16, 16, None, None, None) != 0
    assert lib.fv_conv2d_s2(None, None, None, None, 0, 1, 8, 8, 16, 16, 16, None, None, None) != 0
    assert lib.fv_conv2d_wgrad(None, None, None, 1, 1, 8, 8, 16, 16, 3, 3, 1, None) != 0
    assert lib.fv_conv2d_wgrad_x2(None, None, None, 1, 1, 8, 8, 16, 16, None) != 0
    assert lib.fv_conv2d_wgrad_s2(None, None, None, 1, 1, 8, 8, 16, 16, None) != 0
    assert lib.fv_wgrad_finish(None, 1, None, 16, 16, 3, 3, 16, 16, 0, None) != 0
    assert lib.fv_wgrad_finish_up(None, 1, None, 16, 16, 16, 16, 0, None) != 0
    assert lib.fv_weight_prep_up(None, None, None, 16, 16, 16, 16, None) != 0
    assert lib.fv_weight_prep_s2(None, None, None, 16, 16, 16, 16, None) != 0
    assert lib.fv_slab_sum(None, 1, 10, None, 10, 0, None) != 0
    assert lib.fv_weight_prep_batched(None, 1, 10, None) != 0
    assert lib.fv_adam_multi(None, 1, 10, 1e-3, 0.9, 0.999, 1e-8, None, None) != 0
    assert lib.fv_bn_act_fwd_fin(None, 0, None, 1.0, None, None, None, None, 0.1, 1e-5, None, None, 0, 0, 1, 8, 8, 16, 0, 1, None) != 0
    assert lib.fv_bn_act_bwd_apply_fin(None, 0, None, 0, 0, None, None, 1.0, None, None, None, None, 1, 8, 8, 16, 0, 1, None) != 0
    # reductions refuse to run without their workspace (include/facevae_b200.h, `red_ws`)
    one = ctypes.c_void_p(16)            # any non-null address: the checks fire before anything is dereferenced
    assert lib.fv_bn_stats(one, 0, one, 64, 16, None, None) != 0 and b"workspace" in lib.fv_last_error()
    assert lib.fv_colsum(one, one, 64, 16, None, None) != 0
    assert lib.fv_recon_loss_flat(one, one, None, one, 64, 0, 1.0, None, None) != 0


def test_reduction_workspace_and_split_planning(lib):
    """Sizes the host allocates from: the reduction scratch and the number of partial slabs of the weight-gradient kernels
    (deterministic: one slab per pixel split, added in split order by fv_wgrad_finish)."""
    assert lib.fv_reduce_ws_bytes() >= 512 + 2048 * 1024 * 4
    assert lib.fv_abi_version() == 2
    sp = lib.fv_conv2d_wgrad_splits
    # thin full-resolution layer (enc.1, 32 -> 64 at 256x256): ring schedule, one slab per CTA
    assert 1 <= sp(0, 32, 256, 256, 32, 64, 3, 3) <= 148
    # wide layer: generic schedule, groups x splits <= SM count
    assert 1 <= sp(0, 32, 16, 16, 256, 256, 3, 3) <= 148
    # up-sampling conv (x2 geometry, four phases) and the 4x4 stride-2 conv
    assert 1 <= sp(1, 32, 32, 32, 256, 128, 2, 2) <= 37 and 1 <= sp(2, 32, 32, 32, 64, 128, 4, 4) <= 148
    assert sp(0, 1, 8, 8, 24, 16, 3, 3) == 0            # unsupported channel count: the launch reports why
    assert lib.fv_outconv_wgrad_splits(32, 256, 256) >= 1 and lib.fv_outconv_wgrad_splits(32, 64, 64) == 0
    assert lib.fv_reparam_kl_parts(32, 4096) >= 1
    # statistics of the x2 / s2 geometries: never for NCHW outputs
    assert lib.fv_conv2d_geom_fuses_stats(1, 2, 32, 64, 64, 128, 64) == 0
    assert lib.fv_conv2d_geom_fuses_stats(1, 0, 32, 16, 16, 256, 256) in (0, 1)


def test_step_scope_collects_tagged_weights():
    """ops.step_scope walks the module tree for 4-d weights and their ``prep_kind`` tag (plain / up-sampling / own prep);
    with CPU parameters nothing is launched (the package has no CPU path) and the scope still batches the counters."""
    from face_vae_b200 import ops
    from face_vae_b200.models import FaceVAE
    m = FaceVAE()
    kinds = {name: getattr(mod, "prep_kind", 0) for name, mod in m.named_modules() if getattr(getattr(mod, "weight", None), "dim", lambda: 0)() == 4}
    assert kinds["up.0.layers.1.layers.0"] == ops.PREP_UP and kinds["enc.1.layers.0.layers.0"] == ops.PREP_PLAIN
    assert kinds["out_conv"] == -1
    t = torch.zeros((), dtype=torch.long)
    with ops.step_scope(m):
        ops.bump_counter(t)
        assert int(t) == 0                # deferred to the end of the scope
    assert int(t) == 1
    ops.bump_counter(t)
    assert int(t) == 2


def test_fused_adam_refuses_cpu_tensors():
    from face_vae_b200 import _lib
    from face_vae_b200.optim import FusedAdam
    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    opt = FusedAdam([p], lr=1e-3)
    with pytest.raises(_lib.FaceVaeError):
        opt.step()


def test_ops_refuse_cpu_tensors():
    from face_vae_b200 import _lib, ops
    with pytest.raises(_lib.FaceVaeError):
        ops.bn_stats(torch.zeros((1, 2, 2, 16)))
    with pytest.raises(_lib.FaceVaeError):
        ops.outconv_prep(torch.zeros((3, 32, 7, 7)))


def test_checkpoint_format_and_optimizer_state_interchange(tmp_path):
    """Reference Logger checkpoint layout (logger.py:93-115) and FusedAdam <-> torch.optim.Adam state dicts (host side only)."""
    from face_vae_b200.models import FaceVAE
    from face_vae_b200.optim import FusedAdam
    from face_vae_b200.trainer import VAETrainer
    torch.manual_seed(0)
    tr = VAETrainer(FaceVAE())                              # CPU parameters: construction and (de)serialisation need no GPU
    for p in tr.vae.parameters():                           # give the optimiser some state, as after a few steps
        p.grad = torch.full_like(p, 1e-3)
    tr.optimizer.step()
    path = tmp_path / "00000007-checkpoint.pth.tar"
    tr.save_cpk(str(path), "vae", epoch=7)
    ckp = torch.load(str(path), map_location="cpu")
    assert set(ckp) == {"vae", "optimizer_vae", "epoch"} and ckp["epoch"] == 7
    assert "enc.0.layers.layers.0.weight" in ckp["vae"] or any(k.endswith("layers.0.weight") for k in ckp["vae"])
    tr2 = VAETrainer(FaceVAE())
    assert tr2.load_cpk(str(path), "vae") == 8
    for a, b in zip(tr.vae.state_dict().values(), tr2.vae.state_dict().values()):
        assert torch.equal(a, b)
    # the same optimiser state loads into FusedAdam (and its state dict back into torch.optim.Adam)
    params = list(tr2.vae.parameters())
    fa = FusedAdam(params, lr=5e-5, betas=(0.5, 0.999))
    fa.load_state_dict(ckp["optimizer_vae"])
    st = fa.state[params[0]]
    assert set(st) >= {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 1.0
    ta = torch.optim.Adam(params, lr=5e-5, betas=(0.5, 0.999))
    ta.load_state_dict(fa.state_dict())
    assert torch.equal(ta.state[params[0]]["exp_avg"], tr.optimizer.state[next(iter(tr.vae.parameters()))]["exp_avg"])


def test_generator_full_keeps_the_reference_call_contract():
    """GeneratorFull drop-in (reference trainer.py:214-317): constructor and forward parameter names, the ten loss keys.
    Compared with the reference's own class when /root/reference is present (this container), with the recorded names
    otherwise (the GPU box)."""
    import inspect
    import os
    import sys
    from face_vae_b200.models import FaceVAE
    from face_vae_b200.trainer import GeneratorFull
    ctor = ["efe", "afe", "ckd", "hpe_ede", "mfe", "generator", "discriminator", "pretrained_path", "n_bins"]
    fwd = ["s", "d", "s_a", "d_a", "train_vae"]
    keys = ["P", "G", "F", "E", "L", "H", "D", "C", "K", "R"]
    if os.path.isdir("/root/reference"):
        src = open("/root/reference/trainer.py").read()
        import ast
        cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "GeneratorFull"][0]
        fns = {f.name: f for f in cls.body if isinstance(f, ast.FunctionDef)}
        assert [a.arg for a in fns["__init__"].args.args][1:] == ctor
        assert [a.arg for a in fns["forward"].args.args][1:] == fwd
    assert list(inspect.signature(GeneratorFull.__init__).parameters)[1:] == ctor
    assert list(inspect.signature(GeneratorFull.forward).parameters)[1:6] == fwd
    g = GeneratorFull(generator=FaceVAE())
    assert list(g.weights) == keys
    with pytest.raises(ValueError):
        GeneratorFull()


def test_vae_variants_keep_the_reference_state_dict_keys_and_refuse_cpu_tensors():
    """flatten_vae / local_vae (reference models.py:484-522, 442-482): the key lists stored next to the golden outputs are the
    unmodified reference modules' state_dict keys, so a reference checkpoint loads unchanged; no CPU path."""
    from tests import goldenlib as G
    from face_vae_b200 import models as MO
    g = G.load("vae_variants.npz")
    assert sorted(MO.flatten_vae().state_dict().keys()) == [str(k) for k in g["fvae.keys"]]
    assert sorted(MO.local_vae().state_dict().keys()) == [str(k) for k in g["lvae.keys"]]
    with pytest.raises(RuntimeError):
        MO.flatten_vae()(torch.zeros(2, 16, 4, 4), True)
    with pytest.raises(RuntimeError):
        MO.local_vae()(torch.zeros(2, 128, 8, 8))


def test_side_stream_context_is_inert_outside_a_step_scope():
    """ops.wgrad_stream only forks inside ops.step_scope (and never without CUDA work): outside a scope the body runs on the
    caller's stream and nothing is left to join; set_wgrad_stream returns the previous setting."""
    from face_vae_b200 import ops
    prev = ops.set_wgrad_stream(True)
    try:
        with ops.wgrad_stream(torch.zeros(3)) as ws:
            assert ws.ctx is None
        assert not ops._wgrad_dirty and not ops._wgrad_live
        ops.join_wgrad_stream()                      # nothing pending: a no-op
        assert ops.set_wgrad_stream(False) is True
        with ops.step_scope(torch.nn.Linear(2, 2)):  # a CPU module: no 4-d CUDA weights, nothing prepared, nothing forked
            with ops.wgrad_stream(torch.zeros(3)) as ws:
                assert ws.ctx is None
        assert not ops._scope_active
    finally:
        ops.set_wgrad_stream(prev)


def test_step_scope_leaves_no_scope_behind_when_the_filter_preparation_fails(monkeypatch):
    """__exit__ does not run when __enter__ raises: a failing batched preparation must not leave the process inside a scope
    (shared zero gradients, side stream) for ever."""
    from face_vae_b200 import ops

    class CudaLookingWeight(torch.nn.Parameter):
        @property
        def is_cuda(self):
            return True

    m = torch.nn.Conv2d(3, 4, 3)
    m.weight = CudaLookingWeight(m.weight.data)

    def boom(weights):
        raise RuntimeError("preparation failed")

    monkeypatch.setattr(ops._prep, "run", boom)
    prev = ops.set_wgrad_stream(False)
    try:
        with pytest.raises(RuntimeError):
            with ops.step_scope(m):
                pass
        assert not ops._scope_active and not ops._prep.valid and not ops._wgrad_dirty
    finally:
        ops.set_wgrad_stream(prev)
