"""Per-tensor error table: CUDA path vs fp32 oracle, next to the oracle's own bf16-autocast deviation (yardstick)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import facevae_oracle as O
from face_vae_b200.models import FaceVAE

def metrics(got, ref, rtol=2e-2, afrac=2e-2):
    got, ref = got.double().flatten().cpu(), ref.double().flatten().cpu()
    err = (got - ref).abs()
    am = ref.abs().max().item()
    bad = (err > rtol * ref.abs() + afrac * am).double().mean().item()
    l2 = (err.pow(2).sum().sqrt() / ref.pow(2).sum().sqrt().clamp_min(1e-30)).item()
    return l2, err.max().item() / max(am, 1e-30), bad, am

def run(n, hw, base):
    cfg = O.CFG_256
    p = O.det_anchor_params(cfg, base)
    x, eps = O.det_inputs(n, hw, hw, cfg, base)
    ref_out, ref_g, _, ref_t = O.anchor_train_grads(p, x, eps, cfg)
    # yardstick: the same oracle under CPU bf16 autocast
    leaf = {k: (v.clone() if "running" in k else v.clone().requires_grad_(True)) for k, v in p.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ac = O.anchor_forward(leaf, x, eps, cfg, True, True, None, {})
    ac["loss"].float().backward()
    ac_g = {k: v.grad for k, v in leaf.items() if v.requires_grad}
    m = FaceVAE()
    sd = m.state_dict()
    for k, v in p.items():
        sd[k] = v.clone()
    m.load_state_dict(sd)
    m = m.cuda().train()
    outs = []
    for rep in range(2):
        m.zero_grad(set_to_none=True)
        out = m.forward_loss(x.cuda(), eps.cuda())
        (cfg.w_kl * out["K"] + cfg.w_rec * out["R"]).backward()
        torch.cuda.synchronize()
        outs.append((out, {k: q.grad.clone() for k, q in m.named_parameters()}))
    out, g = outs[0]
    print(f"\n## batch {n}, {hw}x{hw} (deterministic weights / inputs base {base})\n")
    print(f"K: ours {out['K'].item():.6f}, fp32 reference {ref_out['K'].item():.6f}, reference under bf16 autocast {ac['K'].item():.6f}; "
          f"R: ours {out['R'].item():.6f}, fp32 {ref_out['R'].item():.6f}, autocast {ac['R'].item():.6f}; second run of ours: K {outs[1][0]['K'].item():.6f} "
          f"(bitwise {'equal' if outs[1][0]['K'].item() == out['K'].item() else 'DIFFERENT'})\n")
    print("| tensor | ours: rel L2 vs fp32 | max err / max|ref| | reference-autocast: rel L2 vs fp32 | max err / max|ref| | bound 1.25 x yard + 5e-3 | ours run-to-run | max|ref| |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    rows = [("mu", out["mu"], ref_out["mu"], ac["mu"]), ("logstd", out["logstd"], ref_out["logstd"], ac["logstd"]),
            ("x_hat", out["x_hat"], ref_out["x_hat"], ac["x_hat"])]
    for name, a, r, c in rows:
        l2, mx, bad, am = metrics(a, r)
        l2c, mxc, badc, _ = metrics(c.float(), r)
        print(f"| {name} | {l2:.2e} | {mx:.2e} | {l2c:.2e} | {mxc:.2e} | | | {am:.3e} |")
    worst = 0.0
    for k in ref_g:
        l2, mx, bad, am = metrics(g[k], ref_g[k])
        l2c, mxc, badc, _ = metrics(ac_g[k].float(), ref_g[k])
        l2r = metrics(outs[1][1][k], g[k])[0]
        zero = am < 1e-6
        note = " (analytically zero: noise only)" if zero else ""
        if not zero:
            worst = max(worst, l2 / (1.25 * l2c + 5e-3))
        print(f"| grad {k}{note} | {l2:.2e} | {mx:.2e} | {l2c:.2e} | {mxc:.2e} | {1.25 * l2c + 5e-3:.2e} | {l2r:.1e} | {am:.3e} |")
    print(f"\nlargest ours / bound over the non-zero gradients: {worst:.2f}")

if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 16)
    print("# Per-tensor parity: CUDA path vs the fp32 oracle, next to the oracle under CPU bf16 autocast (the yardstick of DESIGN.md section 2)\n")
    print("`python tools/parity_report.py` on a B200 box. The oracle (oracle/facevae_oracle.py) is pinned to the unmodified reference classes by "
          "tests/test_oracle_golden.py; `ours` = face_vae_b200.models.FaceVAE.forward_loss + backward in training mode.")
    run(4, 64, 0)
    run(32, 256, 0)
