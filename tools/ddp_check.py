"""torchrun --nproc-per-node R tools/ddp_check.py : R-rank data-parallel step (SyncBN statistic exchange + bucketed
gradient all-reduce over NCCL) against a single-process run on the concatenated global batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from face_vae_b200 import distributed as fd, functional as Fn
from face_vae_b200.models import FaceVAE
from oracle import facevae_oracle as O

def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    fd.init_dist(int(os.environ["LOCAL_RANK"]), world, "nccl")
    cfg = O.CFG_256
    p = O.det_anchor_params(cfg, 0)
    per, hw = 8, 128
    x, eps = O.det_inputs(per * world, hw, hw, cfg, 2)
    def make():
        m = FaceVAE(); sd = m.state_dict()
        for k, v in p.items(): sd[k] = v.clone()
        m.load_state_dict(sd); return m.cuda().train()
    m = make()
    red = fd.GradientReducer(m.parameters(), bucket_mb=1.0)
    xs, es = x[rank * per:(rank + 1) * per].cuda(), eps[rank * per:(rank + 1) * per].cuda()
    out = m.forward_loss(xs, es)
    (cfg.w_kl * out["K"] + cfg.w_rec * out["R"]).backward()
    red.finish()
    torch.cuda.synchronize()
    g = {k: q.grad.clone() for k, q in m.named_parameters()}
    bufs = {k: b.clone() for k, b in m.named_buffers() if "running" in k}
    # all ranks must hold identical averaged gradients and running stats
    for k, t in list(g.items()) + list(bufs.items()):
        t0 = t.clone(); dist.broadcast(t0, 0)
        assert torch.equal(t0, t), f"rank {rank}: {k} differs from rank 0"
    if rank == 0:
        Fn.set_sync_bn(False)                      # single-process oracle for R ranks: global batch on one GPU
        m1 = make()
        o1 = m1.forward_loss(x.cuda(), eps.cuda())
        (cfg.w_kl * o1["K"] + cfg.w_rec * o1["R"]).backward()
        torch.cuda.synchronize()
        worst = 0.0
        for k, q in m1.named_parameters():
            ref = q.grad
            if ref.abs().max().item() < 1e-5: continue
            rel = ((g[k] - ref).norm() / ref.norm()).item(); worst = max(worst, rel)
        for k, b in m1.named_buffers():
            if "running" in k:
                assert torch.allclose(bufs[k], b, rtol=1e-3, atol=1e-4), k
        print(f"ddp_check world={world}: buckets {len(red.buckets)} launched {red.launched}; worst grad rel-L2 vs single-process global batch {worst:.3e}")
        assert worst < 0.25, worst
        print("ddp_check: ok")
    dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
