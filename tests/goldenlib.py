"""Helpers to compare tensors with the summarised golden fixtures (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE = 257


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
