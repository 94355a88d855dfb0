"""Worker of tests/test_ddp_gpu.py: one process per GPU (torchrun), NCCL.  Scenarios (argv[1]):

 block   one DownBlock2D / UpBlock2D / ResBlock2D per rank on its shard vs the same block on the concatenated batch in a single
         process (SyncBatchNorm over R ranks == batch norm over the global batch): only the order of the statistic
         sums differs, so this is tight -- forward per element 2e-2, gradients relative L2 <= 2e-2;
 model   the whole data-parallel train step: identical (bitwise) averaged gradients and running statistics on every rank,
         and agreement with the single-process global-batch run within the bf16 yardstick;
 skew    round-1 ADVICE (high): one rank is delayed before backward; gradients must still be bitwise equal across ranks, in
         eager mode and through the CUDA-graph-captured step;
 gradar  the peer-memory gradient all-reduce kernel against NCCL;
 wstream the weight-gradient side stream (ops.wgrad_stream) gives the one-stream bits under data parallelism."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from face_vae_b200 import distributed as fd, functional as Fn
from face_vae_b200.models import FaceVAE
from face_vae_b200.trainer import VAETrainer
from oracle import facevae_oracle as O


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def same_on_all_ranks(name, t, rank):
    t0 = t.clone()
    dist.broadcast(t0, 0)
    assert torch.equal(t0, t), f"rank {rank}: {name} differs from rank 0"


def make_model(p):
    m = FaceVAE()
    sd = m.state_dict()
    for k, v in p.items():
        sd[k] = v.clone()
    m.load_state_dict(sd)
    return m.cuda().train()


def scenario_block(rank, world):
    import face_vae_b200.modules as M
    torch.manual_seed(0)
    per = 4
    for name, ctor, ci, hw in (("down", lambda: M.DownBlock2D(64, 128, False), 64, 32), ("up", lambda: M.UpBlock2D(128, 64, False), 128, 16),
                               ("res", lambda: M.ResBlock2D(64, False), 64, 16)):
        torch.manual_seed(1)
        blk = ctor().cuda().train()
        ref = ctor().cuda().train()
        ref.load_state_dict(blk.state_dict())
        g = torch.Generator().manual_seed(7)
        x = (torch.rand((per * world, ci, hw, hw), generator=g) * 2 - 1).bfloat16().float().cuda()
        xs = x[rank * per:(rank + 1) * per].clone().requires_grad_(True)
        y = blk(xs)
        gy = torch.rand((per * world,) + tuple(y.shape[1:]), generator=g).cuda() * 2 - 1
        (y.float() * gy[rank * per:(rank + 1) * per]).sum().backward()
        torch.cuda.synchronize()
        Fn.set_sync_bn(False)                        # single-process reference on the global batch
        xg = x.clone().requires_grad_(True)
        yg = ref(xg)
        (yg.float() * gy).sum().backward()
        torch.cuda.synchronize()
        Fn.set_sync_bn(True)
        ys, yr = y.float(), yg.float()[rank * per:(rank + 1) * per]
        err = (ys - yr).abs()
        assert bool((err <= 2e-2 * yr.abs() + 4e-3 * yr.abs().max()).all()), (name, "forward", err.max().item())
        assert rel_l2(xs.grad, xg.grad[rank * per:(rank + 1) * per]) <= 2e-2, (name, "dx")
        for (k, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
            gsum = p.grad.clone()
            dist.all_reduce(gsum)                    # per-rank parameter gradients add up to the global-batch gradient
            if float(q.grad.abs().max()) < 1e-6 * float(max(r.grad.abs().max() for r in ref.parameters())):
                continue
            r = rel_l2(gsum, q.grad)
            assert r <= 2e-2, (name, k, r)
        for (k, b), (_, c) in zip(blk.named_buffers(), ref.named_buffers()):
            if "running" in k:
                same_on_all_ranks(name + "." + k, b, rank)
                torch.testing.assert_close(b, c, rtol=1e-4, atol=1e-5)
    if rank == 0:
        print(f"ddp block world={world}: ok")


def scenario_model(rank, world):
    cfg = O.CFG_256
    p = O.det_anchor_params(cfg, 0)
    per, hw = 8, 128
    x, eps = O.det_inputs(per * world, hw, hw, cfg, 2)
    m = make_model(p)
    red = fd.GradientReducer(m.parameters(), bucket_mb=1.0)
    xs, es = x[rank * per:(rank + 1) * per].cuda(), eps[rank * per:(rank + 1) * per].cuda()
    out = m.forward_loss(xs, es)
    (cfg.w_kl * out["K"] + cfg.w_rec * out["R"]).backward()
    red.finish()
    torch.cuda.synchronize()
    g = {k: q.grad.clone() for k, q in m.named_parameters()}
    bufs = {k: b.clone() for k, b in m.named_buffers() if "running" in k}
    for k, t in list(g.items()) + list(bufs.items()):
        same_on_all_ranks(k, t, rank)
    assert all(q.grad.data_ptr() == red._slot[id(q)].data_ptr() for q in m.parameters()), "gradients must live in the flat buffer"
    if rank == 0:
        Fn.set_sync_bn(False)
        m1 = make_model(p)
        o1 = m1.forward_loss(x.cuda(), eps.cuda())
        (cfg.w_kl * o1["K"] + cfg.w_rec * o1["R"]).backward()
        torch.cuda.synchronize()
        worst = 0.0
        for k, q in m1.named_parameters():
            if q.grad.abs().max().item() < 1e-5:
                continue
            worst = max(worst, rel_l2(g[k], q.grad))
        for k, b in m1.named_buffers():
            if "running" in k:
                assert torch.allclose(bufs[k], b, rtol=2e-3, atol=1e-4), k
        print(f"ddp model world={world}: buckets {len(red.buckets)} launched {red.launched}; worst grad rel-L2 vs global batch {worst:.3e}")
        assert worst < 0.25, worst            # two bf16 runs with differently ordered statistic sums: the bf16 yardstick
        Fn.set_sync_bn(True)
    dist.barrier()


def scenario_skew(rank, world):
    cfg = O.CFG_256
    p = O.det_anchor_params(cfg, 0)
    per, hw = 4, 64
    x, eps = O.det_inputs(per * world, hw, hw, cfg, 5)
    xs, es = x[rank * per:(rank + 1) * per].cuda(), eps[rank * per:(rank + 1) * per].cuda()
    for use_graph in (False, True):
        m = make_model(p)
        tr = VAETrainer(m, lr=1e-3, use_cuda_graph=use_graph)
        for it in range(4):
            if rank == (it % world):
                torch.cuda._sleep(int(4e8))              # ~0.2 s of skew on a different rank every step
            tr.step(xs, es)
        torch.cuda.synchronize()
        for k, q in m.named_parameters():
            same_on_all_ranks(f"graph={use_graph} grad {k}", q.grad, rank)
            same_on_all_ranks(f"graph={use_graph} weight {k}", q.data, rank)
        assert tr.use_cuda_graph == use_graph, "graph capture must not silently fall back"
    if rank == 0:
        print(f"ddp skew world={world}: ok")


def scenario_wstream(rank, world):
    """ops.wgrad_stream under data parallelism: with the weight-gradient kernels on the side stream (joined before the gradient
    all-reduce gathers the gradients) three optimiser steps give bitwise the weights of the one-stream order, eager and captured,
    with a different rank delayed every step; identical on all ranks."""
    from face_vae_b200 import ops
    cfg = O.CFG_256
    p = O.det_anchor_params(cfg, 0)
    per, hw = 4, 64
    x, eps = O.det_inputs(per * world, hw, hw, cfg, 6)
    xs, es = x[rank * per:(rank + 1) * per].cuda(), eps[rank * per:(rank + 1) * per].cuda()
    prev = ops.set_wgrad_stream(False)
    try:
        for use_graph in (False, True):
            res = []
            for side in (False, True):
                ops.set_wgrad_stream(side)
                m = make_model(p)
                tr = VAETrainer(m, lr=1e-3, use_cuda_graph=use_graph)
                for it in range(3):
                    if side and rank == (it % world):
                        torch.cuda._sleep(int(2e8))
                    tr.step(xs, es)
                torch.cuda.synchronize()
                assert tr.use_cuda_graph == use_graph, "graph capture must not silently fall back"
                assert not ops._wgrad_dirty and not ops._wgrad_live
                res.append({k: v.detach().clone() for k, v in m.state_dict().items()})
                for k, q in m.named_parameters():
                    same_on_all_ranks(f"graph={use_graph} side={side} weight {k}", q.data, rank)
                del tr
            for k in res[0]:
                assert torch.equal(res[0][k], res[1][k]), (use_graph, k)
    finally:
        ops.set_wgrad_stream(prev)
    if rank == 0:
        print(f"ddp wstream world={world}: ok")


def scenario_gradar(rank, world):
    """fv_grad_allreduce (peer-memory, one kernel) against NCCL on odd-sized tensors, 20 back-to-back epochs."""
    sizes = [5, 1023, 4096, 333, 70001, 1 << 20, 7]
    params = [torch.nn.Parameter(torch.zeros(n, device="cuda")) for n in sizes]
    red = fd.GradientReducer(params)
    assert red.peer, "peer-memory gradient all-reduce must be active on an NVLink box"
    for it in range(20):
        torch.manual_seed(1000 * it + rank)
        gs = [torch.randn(n, device="cuda") * (1 + it) for n in sizes]
        ref = []
        for g in gs:
            t = g.clone()
            dist.all_reduce(t)
            ref.append(t / world)
        if rank == it % world:
            torch.cuda._sleep(int(1e8))
        for q, g in zip(params, gs):
            q.grad = g
        for q in reversed(params):
            red._hook(q)
        red.finish()
        torch.cuda.synchronize()
        for i, (q, r) in enumerate(zip(params, ref)):
            assert q.grad.data_ptr() == red._slot[id(q)].data_ptr()
            assert torch.allclose(q.grad, r, rtol=1e-5, atol=1e-5 * (1 + it)), (it, i, (q.grad - r).abs().max().item())
            same_on_all_ranks(f"gradar it={it} tensor {i}", q.grad, rank)
    if rank == 0:
        print(f"ddp gradar world={world}: ok")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    fd.init_dist(int(os.environ["LOCAL_RANK"]), world, "nccl")
    {"block": scenario_block, "model": scenario_model, "skew": scenario_skew, "gradar": scenario_gradar, "wstream": scenario_wstream}[sys.argv[1]](rank, world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
