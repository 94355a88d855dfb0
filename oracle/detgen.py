"""Deterministic, platform-independent tensor generators (test infrastructure).

Golden fixtures store only *outputs* of the reference; the inputs and weights
that produced them are regenerated from these closed-form generators, so the
fixtures stay small and do not depend on any RNG implementation (torch's CPU
and CUDA Philox streams differ; the reference draws ``eps`` inline with
``torch.randn``, models.py:561).  Integer arithmetic only up to the final
float conversion: splitmix64 over the flat element index.
"""
from __future__ import annotations

import zlib

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(idx: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64((seed * 0x9E3779B97F4A7C15 + 0x1234567) & 0xFFFFFFFFFFFFFFFF)
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return z


def name_seed(name: str, base: int = 0) -> int:
    """Stable 32-bit seed from a parameter name (crc32, not Python's salted hash)."""
    return (zlib.crc32(name.encode()) + 7919 * base) & 0x7FFFFFFF


def det_unit(shape, seed: int) -> np.ndarray:
    """Uniform in [0, 1) with 24 random bits, float32-exact."""
    n = int(np.prod(shape)) if len(shape) else 1
    bits = _splitmix64(np.arange(n, dtype=np.uint64), seed) >> np.uint64(40)  # top 24 bits
    return (bits.astype(np.float64) / float(1 << 24)).astype(np.float32).reshape(shape)


def det_uniform(shape, seed: int, lo: float, hi: float) -> np.ndarray:
    return (lo + (hi - lo) * det_unit(shape, seed).astype(np.float64)).astype(np.float32)


def det_normal(shape, seed: int) -> np.ndarray:
    """Approximately N(0,1): Irwin-Hall sum of 12 exact uniforms minus 6 (no libm)."""
    n = int(np.prod(shape)) if len(shape) else 1
    acc = np.zeros(n, dtype=np.float64)
    idx = np.arange(n, dtype=np.uint64)
    for j in range(12):
        bits = _splitmix64(idx * np.uint64(12) + np.uint64(j), seed ^ 0x5BD1E995) >> np.uint64(40)
        acc += bits.astype(np.float64) / float(1 << 24)
    return (acc - 6.0).astype(np.float32).reshape(shape)
