"""CPU oracle for the fixture generators next to the hot path -- TEST INFRASTRUCTURE ONLY.

Restates, in NumPy,
  * ``generate_smooth_synthetic`` (reference ``python/generate_test_frames_natural.py:49-64``):
    the sum-of-sinusoids texture, clipped and truncated to uint8;
  * ``apply_motion`` (``:67-73``): ``scipy.ndimage.shift(frame, (dy, dx), order=1,
    mode="constant", cval=128)`` on a uint8 frame.  SciPy (1.18 here; unpinned in the reference's
    pyproject.toml) evaluates it as: coordinate = index - shift in float64; outside
    [0, n-1] on either axis -> cval; else the 4-tap blend sum((value * wy) * wx) in float64 in
    row-major tap order; uint8 output = clamp(t + 0.5) truncated (ni_interpolation.c,
    CASE_INTERP_OUT_UINT).

Pinned by ``tests/golden/make_golden_motion.py``, which runs the reference's own two functions in
the build container and commits inputs + outputs (``tests/golden/motion.npz``).
"""

from __future__ import annotations

import numpy as np


def generate_smooth_synthetic(width: int, height: int) -> np.ndarray:
    x = np.linspace(0, 4 * np.pi, width)
    y = np.linspace(0, 3 * np.pi, height)
    X, Y = np.meshgrid(x, y)
    pattern = (
        128
        + 50 * np.sin(X) * np.cos(Y)
        + 30 * np.cos(2 * X + 0.5) * np.sin(1.5 * Y)
        + 20 * np.sin(3 * X - 0.3) * np.cos(2.5 * Y + 0.7)
    )
    return np.clip(pattern, 0, 255).astype(np.uint8)


def apply_motion(frame: np.ndarray, dx: float, dy: float, cval: float = 128.0) -> np.ndarray:
    """uint8 [H, W] -> uint8 [H, W], the frame's content moved by (+dx, +dy) pixels."""
    assert frame.dtype == np.uint8 and frame.ndim == 2
    h, w = frame.shape
    yy = np.arange(h, dtype=np.float64) - float(dy)
    xx = np.arange(w, dtype=np.float64) - float(dx)
    Y, X = np.meshgrid(yy, xx, indexing="ij")
    inside = (Y >= 0) & (Y <= h - 1) & (X >= 0) & (X <= w - 1)
    y0, x0 = np.floor(Y), np.floor(X)
    fy, fx = Y - y0, X - x0
    y0i = np.clip(y0.astype(np.int64), 0, h - 1)
    x0i = np.clip(x0.astype(np.int64), 0, w - 1)
    y1i = np.minimum(y0i + 1, h - 1)  # weight exactly 0 where it would leave the frame
    x1i = np.minimum(x0i + 1, w - 1)
    f = frame.astype(np.float64)
    t = np.zeros((h, w), np.float64)
    t += f[y0i, x0i] * (1.0 - fy) * (1.0 - fx)
    t += f[y0i, x1i] * (1.0 - fy) * fx
    t += f[y1i, x0i] * fy * (1.0 - fx)
    t += f[y1i, x1i] * fy * fx
    t = np.where(inside, t, float(cval))
    v = np.where(t > 0, t + 0.5, 0.0)
    return np.clip(v, 0, 255).astype(np.uint8)
