"""Small host-side scalar recipes the kernels take as inputs (no device work, no oracle)."""
from __future__ import annotations

import numpy as np


def percentile_plan(n: int, ratio: float):
    """Where `np.percentile(x_f32, ratio*100)` looks (eval.py:257), as numpy >= 2 computes it for a
    float32 array: q and the virtual index are float32 (weak Python scalars), method 'linear'
    (alpha = beta = 1):  vidx = n*q + (1 + q*(1-1-1)) - 1.
    Returns (lo, gamma): the lower order-statistic index and the float32 interpolation weight."""
    q = np.true_divide(ratio * 100, np.float32(100))          # -> float32
    q = np.float32(q)
    vidx = np.float32(n) * q + (np.float32(1.0) + q * np.float32(-1.0)) - np.float32(1.0)
    vidx = np.float32(vidx)
    lo = int(np.floor(vidx))
    lo = max(0, min(lo, n - 1))
    gamma = np.float32(np.float64(vidx) - np.float64(lo))
    return lo, gamma


def lerp_f32(a, b, t):
    """numpy's _lerp in float32: a + (b-a)*t, replaced by b - (b-a)*(1-t) when t >= 0.5."""
    a, b, t = np.float32(a), np.float32(b), np.float32(t)
    d = np.float32(b - a)
    out = np.float32(a + np.float32(d * t))
    if t >= np.float32(0.5):
        out = np.float32(b - np.float32(d * np.float32(np.float32(1.0) - t)))
    return out
