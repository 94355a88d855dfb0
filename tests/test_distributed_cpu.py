"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucketed gradient reducer, rank helpers, and the
batch-norm statistic exchange algebra (sum all-reduce of [sum, sum_sq] == statistics of the global batch)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from face_vae_b200 import distributed as fd
    from face_vae_b200 import functional as Fn
    fd.init_dist(rank, world, backend="gloo")
    assert fd.get_rank() == rank and fd.get_world_size() == world and fd.is_master() == (rank == 0)
    fd.init_seeds()
    a = torch.rand(1).item()                      # seed = 1 + rank -> different draws per rank
    # identical model on every rank after broadcast
    torch.manual_seed(100 + rank)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 1))
    fd.broadcast_parameters(model)
    red = fd.GradientReducer(model.parameters(), bucket_mb=0.00005)     # tiny buckets -> several all-reduces
    assert len(red.buckets) >= 3
    g = torch.Generator().manual_seed(7)
    x_all = torch.rand((8, 8), generator=g)
    x = x_all[rank * 4:(rank + 1) * 4]
    for _ in range(2):                              # two steps: reducer state must reset
        model.zero_grad(set_to_none=True)
        model(x).mean().backward()
        red.finish()
    grads = [p.grad.clone() for p in model.parameters()]
    # BN statistic exchange: global mean/var from summed [sum, sum_sq]
    y = x_all[rank * 4:(rank + 1) * 4]
    sums = torch.cat([y.sum(0), (y * y).sum(0)])
    tot = Fn._allreduce_sum(sums)
    mean = tot[:8] / 8
    var = tot[8:] / 8 - mean * mean
    q.put((rank, a, [g.numpy() for g in grads], mean.numpy(), var.numpy(), [p.detach().numpy() for p in model.parameters()],
           red.launched))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_reducer_and_stat_exchange_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, g0, m0, v0, p0, l0), (r1, a1, g1, m1, v1, p1, l1) = res
    assert a0 != a1                                                  # init_seeds: seed = 1 + rank
    for x, y in zip(p0, p1):
        assert (x == y).all()                                        # broadcast made the replicas identical
    for x, y in zip(g0, g1):
        assert abs(x - y).max() < 1e-7                               # both ranks hold the averaged gradient
    assert l0 >= 6                                                   # >= 3 buckets x 2 steps were launched
    # oracle for R ranks: single process on the concatenated global batch
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 1))
    g = torch.Generator().manual_seed(7)
    x_all = torch.rand((8, 8), generator=g)
    model(x_all).mean().backward()
    for got, p in zip(g0, model.parameters()):
        assert abs(got - p.grad.numpy()).max() < 1e-6
    assert abs(m0 - x_all.mean(0).numpy()).max() < 1e-6
    assert abs(v0 - x_all.var(0, unbiased=False).numpy()).max() < 1e-6


def test_rank_helpers_without_process_group():
    from face_vae_b200 import distributed as fd
    assert fd.get_rank() == 0 and fd.get_world_size() == 1 and fd.is_master()
    calls = []

    @fd.master_only
    def f(v):
        calls.append(v)
        return v

    assert f(3) == 3 and calls == [3]
    fd.master_only_print("ok")
    red = fd.GradientReducer(torch.nn.Linear(2, 2).parameters())
    red.finish()                                                     # no-op at world size 1
