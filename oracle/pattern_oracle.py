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


# --------------------------------------------------------------------------------------
# apply_motion_opencv (reference python/generate_test_suite.py:165-204): cv2.getRotationMatrix2D +
# cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT, 128).  OpenCV (4.13 here; >= 4.8 per the reference's
# pyproject.toml) evaluates the bilinear warp in FIXED POINT (imgwarp.cpp): the inverse map is sampled
# with 10 fractional bits (AB_BITS), rounded to 1/32 pixel (INTER_BITS = 5), the four weights are
# 15-bit integers, and the result is (sum + 2^14) >> 15.  Restated here; pinned against cv2 itself by
# tests/golden/make_golden_motion.py (13 verifier parameter sets + random affine maps: 0 differing
# pixels in the build container).
# --------------------------------------------------------------------------------------
import math


def rotation_matrix(cx: float, cy: float, angle_deg: float, scale: float) -> np.ndarray:
    """cv2.getRotationMatrix2D((cx, cy), angle, scale), bit for bit (libm cos / sin)."""
    a = angle_deg * (math.pi / 180.0)
    alpha = math.cos(a) * scale
    beta = math.sin(a) * scale
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]], np.float64)


def motion_matrix(width: int, height: int, dx: float, dy: float, rotation: float, scale: float) -> np.ndarray:
    """The 2x3 matrix apply_motion_opencv builds (generate_test_suite.py:183-190)."""
    m = rotation_matrix(width / 2.0, height / 2.0, rotation, scale)
    m[0, 2] += dx
    m[1, 2] += dy
    return m


def invert_affine(m: np.ndarray) -> np.ndarray:
    """warpAffine's inversion of the forward matrix (imgwarp.cpp, !WARP_INVERSE_MAP), same operation order."""
    M = np.asarray(m, np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    a11, a22 = M[4] * D, M[0] * D
    M[0] = a11
    M[1] *= -D
    M[3] *= -D
    M[4] = a22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def warp_affine_u8(frame: np.ndarray, m: np.ndarray, cval: int = 128) -> np.ndarray:
    """cv2.warpAffine(frame, m, (W, H), flags=INTER_LINEAR, borderMode=BORDER_CONSTANT, borderValue=cval)."""
    assert frame.dtype == np.uint8 and frame.ndim == 2
    h, w = frame.shape
    M = invert_affine(m)
    xs, ys = np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64)
    adelta = np.rint(M[0] * xs * 1024.0).astype(np.int64)  # saturate_cast<int> = round half to even
    bdelta = np.rint(M[3] * xs * 1024.0).astype(np.int64)
    x0 = np.rint((M[1] * ys + M[2]) * 1024.0).astype(np.int64) + 16  # + AB_SCALE / INTER_TAB_SIZE / 2
    y0 = np.rint((M[4] * ys + M[5]) * 1024.0).astype(np.int64) + 16
    X = (x0[:, None] + adelta[None, :]) >> 5
    Y = (y0[:, None] + bdelta[None, :]) >> 5
    sx, sy, ax, ay = X >> 5, Y >> 5, X & 31, Y & 31
    padded = np.full((h + 2, w + 2), int(cval), np.int64)
    padded[1:-1, 1:-1] = frame

    def px(yy, xx):
        inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        return np.where(inside, padded[np.clip(yy, -1, h) + 1, np.clip(xx, -1, w) + 1], int(cval))

    s = (px(sy, sx) * ((32 - ay) * (32 - ax)) + px(sy, sx + 1) * ((32 - ay) * ax)
         + px(sy + 1, sx) * (ay * (32 - ax)) + px(sy + 1, sx + 1) * (ay * ax))
    return ((s * 32 + (1 << 14)) >> 15).astype(np.uint8)


def apply_motion_opencv(frame: np.ndarray, dx: float = 0.0, dy: float = 0.0, rotation: float = 0.0, scale: float = 1.0):
    h, w = frame.shape
    return warp_affine_u8(frame, motion_matrix(w, h, dx, dy, rotation, scale), 128)
