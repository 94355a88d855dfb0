"""Helpers to compare tensors with the summarised golden fixtures (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE = 4099


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name)))


def sample_index(n):
    return (np.arange(SAMPLE, dtype=np.int64) * (n // SAMPLE + 1) * 7919 + 13) % n


def check(store, name, t, rtol, atol_frac=None, what="", zero_floor=1e-4):
    """Compare tensor ``t`` with golden entry ``name``.  Tolerance per element is
    rtol * |ref| + atol where atol = atol_frac * absmax(ref) (default atol_frac = rtol)."""
    if isinstance(t, torch.Tensor):
        t = t.detach().to(torch.float64).cpu().numpy()
    a = np.asarray(t, np.float64).flatten()
    shape = tuple(store[f"{name}/shape"])
    assert a.size == int(np.prod(shape)), (name, a.size, shape)
    absmax = float(store[f"{name}/absmax"])
    if absmax < zero_floor:
        # analytically-zero quantity (e.g. the bias of a conv that feeds a batch norm): only noise to compare
        assert np.abs(a).max() <= zero_floor, f"{what}{name}: expected ~0, got absmax {np.abs(a).max():.3e}"
        return float(np.abs(a).max())
    atol = (rtol if atol_frac is None else atol_frac) * max(absmax, 1e-30)
    if f"{name}/full" in store:
        ref = store[f"{name}/full"].astype(np.float64)
        got = a
    else:
        ref = store[f"{name}/sample"].astype(np.float64)
        got = a[sample_index(a.size)]
    err = np.abs(got - ref)
    bad = err > rtol * np.abs(ref) + atol
    assert not bad.any(), f"{what}{name}: {bad.sum()}/{bad.size} off, max err {err.max():.3e} (absmax {absmax:.3e})"
    # global moments guard against errors away from the sampled entries
    l2 = float(store[f"{name}/l2"])
    got_l2 = float(np.sqrt((a * a).sum()))
    assert abs(got_l2 - l2) <= 4 * rtol * l2 + atol, f"{what}{name}: l2 {got_l2} vs {l2}"
    return float(err.max())


def check_vs_yardstick(store, name, t, slack=1.5, floor=5e-3, what=""):
    """bf16-path criterion for gradients: relative L2 error against the fp32 golden no worse than ``slack`` x the
    reference's OWN bf16-autocast deviation from its fp32 self (stored by make_golden.py as <name>/yard_l2) + floor.
    Per-element rtol cannot hold for ReLU-masked gradients in bf16 (mask flips at |z| ~ 0), for either implementation."""
    if isinstance(t, torch.Tensor):
        t = t.detach().to(torch.float64).cpu().numpy()
    a = np.asarray(t, np.float64).flatten()
    absmax = float(store[f"{name}/absmax"])
    if f"{name}/full" in store:
        ref, got = store[f"{name}/full"].astype(np.float64), a
    else:
        ref, got = store[f"{name}/sample"].astype(np.float64), a[sample_index(a.size)]
    l2 = float(np.sqrt(((got - ref) ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-30))
    yard = float(store[f"{name}/yard_l2"])
    assert l2 <= slack * yard + floor, f"{what}{name}: rel-L2 {l2:.3e} vs reference-autocast yardstick {yard:.3e} (absmax {absmax:.3e})"
    return l2, yard


def check_like(got, ref, rtol, atol_frac, name=""):
    got, ref = got.detach().double().cpu().flatten(), ref.detach().double().cpu().flatten()
    err = (got - ref).abs()
    tol = rtol * ref.abs() + atol_frac * ref.abs().max()
    assert bool((err <= tol).all()), f"{name}: max err {err.max().item():.3e} (absmax {ref.abs().max().item():.3e})"
