"""CPU oracle for the float32 Lucas-Kanade path -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference's float pipeline.  It is the
checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline leg may import it.  The product modules in
``optical-flow-fpga_b200/`` never do, and they raise when the CUDA library is
missing instead of falling back to anything in here.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the reference's own
modules (``/root/reference/python``) in the build container and commits their
outputs; ``tests/test_oracle_golden.py`` checks every function below against
those fixtures bit for bit, and against ``python/verification_baseline.json``.

Third-party arithmetic the reference delegates to (not vendored there, versions
unpinned in its ``pyproject.toml:34-40``; installed here: numpy 2.3.5, scipy
1.18.1) is restated explicitly so the operation order is visible:

* ``scipy.signal.convolve2d(mode="same", boundary="symm")`` -> ``_conv3x3_symm``
* ``scipy.ndimage.gaussian_filter(sigma=2)``                -> ``gaussian_blur_sigma``
* ``scipy.ndimage.map_coordinates(order=1, mode="constant")`` -> ``bilinear_sample``
* ``np.sum`` over a contiguous w*w vector (pairwise, 8 lanes) -> ``_numpy_order_sum``

Everything is float32 unless a comment says float64, because that is what the
reference computes in.
"""

from __future__ import annotations

from typing import List, Tuple

import numpy as np

F32 = np.float32

# Sobel taps exactly as the reference builds them (lucas_kanade_core.py:32-33).
SOBEL_X = (np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=F32) / 8.0).astype(F32)
SOBEL_Y = (np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], dtype=F32) / 8.0).astype(F32)

DET_EPS = 1e-4  # lucas_kanade_core.py:131
CONVERGENCE_EPS = 0.01  # lucas_kanade_pyramidal.py:221


# --------------------------------------------------------------------------
# a1  compute_gradients  (lucas_kanade_core.py:15-45)
# --------------------------------------------------------------------------
def _conv3x3_symm(img: np.ndarray, kern: np.ndarray) -> np.ndarray:
    """True 2-D convolution, 'same' size, symmetric boundary, float32.

    Restates scipy.signal.convolve2d as called at lucas_kanade_core.py:39-40:
    the kernel is flipped (convolution, not correlation) and the accumulator is
    a float32 that receives the nine products one after the other in kernel
    order (row j, then column k), zero taps included.  For a 3x3 kernel the
    symmetric boundary equals replicating the edge pixel.
    """
    img = np.asarray(img, dtype=F32)
    pad = np.pad(img, 1, mode="symmetric")
    h, w = img.shape
    acc = np.zeros((h, w), dtype=F32)
    for j in range(3):
        for k in range(3):
            # out[m, n] += in[m + 1 - j, n + 1 - k] * K[j, k]   (padded index +1)
            tap = pad[2 - j : 2 - j + h, 2 - k : 2 - k + w]
            acc = (acc + tap * kern[j, k]).astype(F32)
    return acc


def compute_gradients(frame_prev: np.ndarray, frame_curr: np.ndarray):
    """Ix, Iy on the averaged frame and It = prev - curr (lucas_kanade_core.py:36-43)."""
    frame_prev = np.asarray(frame_prev, dtype=F32)
    frame_curr = np.asarray(frame_curr, dtype=F32)
    frame_avg = ((frame_prev + frame_curr) / F32(2.0)).astype(F32)
    ix = _conv3x3_symm(frame_avg, SOBEL_X)
    iy = _conv3x3_symm(frame_avg, SOBEL_Y)
    it = (frame_prev - frame_curr).astype(F32)
    return ix, iy, it


# --------------------------------------------------------------------------
# a2  lucas_kanade_from_gradients  (lucas_kanade_core.py:73-135)
# --------------------------------------------------------------------------
def _numpy_order_sum(taps: List[np.ndarray]) -> np.ndarray:
    """Sum n = w*w float32 arrays in the order np.sum uses on a contiguous
    n-vector (n < 128): eight running lanes, a fixed tree over the lanes, then
    the n % 8 tail added one by one.  Restates what lucas_kanade_core.py:115-119
    gets from NumPy's pairwise summation for one window.  The reduction result
    starts from the add identity +0.0 (out = 0.0 + pairwise(...)), which only
    shows when every term is -0.0: the sum is then +0.0, not -0.0 (pinned by the
    no_motion pattern, where It == 0 and the sign of u, v depends on it).
    """
    n = len(taps)
    if n < 8:
        res = taps[0].astype(F32)
        for t in taps[1:]:
            res = (res + t).astype(F32)
        return (F32(0.0) + res).astype(F32)
    assert n < 128, "window_size >= 13 recurses in NumPy; outside the reference's presets"
    lanes = [taps[l].astype(F32) for l in range(8)]
    full = n - (n % 8)
    i = 8
    while i < full:
        for l in range(8):
            lanes[l] = (lanes[l] + taps[i + l]).astype(F32)
        i += 8
    res = ((lanes[0] + lanes[1]) + (lanes[2] + lanes[3])) + (
        (lanes[4] + lanes[5]) + (lanes[6] + lanes[7])
    )
    res = res.astype(F32)
    for t in taps[full:]:
        res = (res + t).astype(F32)
    return (F32(0.0) + res).astype(F32)


def window_sums(ix, iy, it, window_size: int = 5):
    """Five structure-tensor sums for every interior pixel, reference order.

    Returns arrays of shape (H - 2*hw, W - 2*hw): sIx2, sIy2, sIxIy, sIxIt, sIyIt.
    """
    ix = np.asarray(ix, dtype=F32)
    iy = np.asarray(iy, dtype=F32)
    it = np.asarray(it, dtype=F32)
    h, w = ix.shape
    hw = window_size // 2
    oh, ow = h - 2 * hw, w - 2 * hw
    pxx = (ix * ix).astype(F32)
    pyy = (iy * iy).astype(F32)
    pxy = (ix * iy).astype(F32)
    pxt = (ix * it).astype(F32)
    pyt = (iy * it).astype(F32)

    def taps_of(p):
        # row-major window order: t = w*i + j  (a fresh contiguous copy of the slice)
        return [p[i : i + oh, j : j + ow] for i in range(window_size) for j in range(window_size)]

    return tuple(_numpy_order_sum(taps_of(p)) for p in (pxx, pyy, pxy, pxt, pyt))


def cramer_solve(sxx, syy, sxy, sxt, syt):
    """2x2 solve exactly as lucas_kanade_core.py:122-133 (float32, no FMA)."""
    a = sxx.astype(F32)
    d = syy.astype(F32)
    b = sxy.astype(F32)
    b0 = (-sxt).astype(F32)
    b1 = (-syt).astype(F32)
    det = ((a * d).astype(F32) - (b * b).astype(F32)).astype(F32)
    ok = np.abs(det) > F32(DET_EPS)
    safe = np.where(ok, det, F32(1.0)).astype(F32)
    nu = ((d * b0).astype(F32) - (b * b1).astype(F32)).astype(F32)
    nv = ((a * b1).astype(F32) - (b * b0).astype(F32)).astype(F32)
    u = np.where(ok, (nu / safe).astype(F32), F32(0.0)).astype(F32)
    v = np.where(ok, (nv / safe).astype(F32), F32(0.0)).astype(F32)
    return u, v


def lucas_kanade_from_gradients(ix, iy, it, window_size: int = 5):
    ix = np.asarray(ix, dtype=F32)
    h, w = ix.shape
    hw = window_size // 2
    u = np.zeros((h, w), dtype=F32)
    v = np.zeros((h, w), dtype=F32)
    if h - 2 * hw <= 0 or w - 2 * hw <= 0:
        return u, v
    sums = window_sums(ix, iy, it, window_size)
    ui, vi = cramer_solve(*sums)
    u[hw : h - hw, hw : w - hw] = ui
    v[hw : h - hw, hw : w - hw] = vi
    return u, v


def lucas_kanade_from_gradients_loop(ix, iy, it, window_size: int = 5):
    """Per-pixel scalar loop, the literal shape of lucas_kanade_core.py:107-133.

    Only for small inputs: used to pin the vectorised version above and to time
    what the reference's own loop costs per pixel.
    """
    ix = np.asarray(ix, dtype=F32)
    iy = np.asarray(iy, dtype=F32)
    it = np.asarray(it, dtype=F32)
    h, w = ix.shape
    hw = window_size // 2
    u = np.zeros((h, w), dtype=F32)
    v = np.zeros((h, w), dtype=F32)
    for y in range(hw, h - hw):
        for x in range(hw, w - hw):
            wx = ix[y - hw : y + hw + 1, x - hw : x + hw + 1]
            wy = iy[y - hw : y + hw + 1, x - hw : x + hw + 1]
            wt = it[y - hw : y + hw + 1, x - hw : x + hw + 1]
            a = np.sum(wx * wx)
            d = np.sum(wy * wy)
            b = np.sum(wx * wy)
            b0 = -np.sum(wx * wt)
            b1 = -np.sum(wy * wt)
            det = a * d - b * b
            if abs(det) > DET_EPS:
                u[y, x] = (d * b0 - b * b1) / det
                v[y, x] = (a * b1 - b * b0) / det
    return u, v


# a3  lucas_kanade_single_scale (lucas_kanade_core.py:48-70)
def lucas_kanade_single_scale(frame_prev, frame_curr, window_size: int = 5):
    ix, iy, it = compute_gradients(frame_prev, frame_curr)
    return lucas_kanade_from_gradients(ix, iy, it, window_size)


# --------------------------------------------------------------------------
# a4  build_gaussian_pyramid  (lucas_kanade_pyramidal.py:23-63)
# --------------------------------------------------------------------------
def gaussian_weights(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """float64 taps of scipy.ndimage's 1-D Gaussian: radius int(truncate*sigma+0.5),
    exp(-0.5/sigma^2 * k^2) normalised by its (NumPy) sum."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x**2)
    return phi / phi.sum()


def _reflect_index(idx: np.ndarray, n: int) -> np.ndarray:
    """scipy 'reflect' (d c b a | a b c d | d c b a) index map, any overhang."""
    period = 2 * n
    m = np.mod(idx, period)
    return np.where(m >= n, period - 1 - m, m)


def _correlate1d_symmetric(x: np.ndarray, wts: np.ndarray, axis: int) -> np.ndarray:
    """One axis of gaussian_filter: float64 accumulate, float32 store.

    Order follows scipy's symmetric-kernel loop: centre tap first, then for
    k = radius .. 1 the pair (x[c-k] + x[c+k]) * w[k].
    """
    x = np.asarray(x, dtype=F32)
    r = (len(wts) - 1) // 2
    n = x.shape[axis]
    xm = np.moveaxis(x, axis, 0).astype(np.float64)
    base = np.arange(n)
    acc = xm[base] * wts[r]
    for ii in range(-r, 0):
        lo = xm[_reflect_index(base + ii, n)]
        hi = xm[_reflect_index(base - ii, n)]
        acc = acc + (lo + hi) * wts[ii + r]
    return np.moveaxis(acc.astype(F32), 0, axis)


def gaussian_blur_sigma(img: np.ndarray, sigma: float) -> np.ndarray:
    """gaussian_filter(img, sigma): axis 0 first, then axis 1, reflect boundary."""
    wts = gaussian_weights(sigma)
    tmp = _correlate1d_symmetric(img, wts, axis=0)
    return _correlate1d_symmetric(tmp, wts, axis=1)


def bilinear_sample(img: np.ndarray, yy: np.ndarray, xx: np.ndarray) -> np.ndarray:
    """map_coordinates(img, [yy, xx], order=1, mode="constant", cval=0) -> float32.

    float64 coordinates; a sample is inside iff 0 <= y <= H-1 and 0 <= x <= W-1,
    anything else is exactly 0.  Blend in float64 over the taps in row-major
    order, each as (value * wy) * wx, summed from 0.0, then stored as float32.
    """
    img = np.asarray(img, dtype=F32)
    h, w = img.shape
    yy = np.asarray(yy, dtype=np.float64)
    xx = np.asarray(xx, dtype=np.float64)
    inside = (yy >= 0) & (yy <= h - 1) & (xx >= 0) & (xx <= w - 1)
    ys = np.where(inside, yy, 0.0)
    xs = np.where(inside, xx, 0.0)
    y0 = np.floor(ys)
    x0 = np.floor(xs)
    fy = ys - y0
    fx = xs - x0
    y0 = y0.astype(np.int64)
    x0 = x0.astype(np.int64)
    # second tap past the last row/column is mirrored by scipy; its weight is 0 there
    y1 = np.where(y0 + 1 > h - 1, max(h - 2, 0), y0 + 1)
    x1 = np.where(x0 + 1 > w - 1, max(w - 2, 0), x0 + 1)
    im = img.astype(np.float64)
    wy0, wy1 = 1.0 - fy, fy
    wx0, wx1 = 1.0 - fx, fx
    t = np.zeros(yy.shape, dtype=np.float64)
    t = t + (im[y0, x0] * wy0) * wx0
    t = t + (im[y0, x1] * wy0) * wx1
    t = t + (im[y1, x0] * wy1) * wx0
    t = t + (im[y1, x1] * wy1) * wx1
    return np.where(inside, t, 0.0).astype(F32)


def build_gaussian_pyramid(image, num_levels: int, scale_factor: float = 0.5):
    """List of levels, coarse -> fine; the last entry is a copy of the input."""
    pyramid: List[np.ndarray] = []
    current = np.array(image, dtype=F32, copy=True)
    for level in range(num_levels):
        if level > 0:
            sigma = 1.0 / scale_factor
            smoothed = gaussian_blur_sigma(current, sigma)
            height, width = smoothed.shape
            new_h = int(height * scale_factor)
            new_w = int(width * scale_factor)
            yc = np.linspace(0, height - 1, new_h)
            xc = np.linspace(0, width - 1, new_w)
            yy, xx = np.meshgrid(yc, xc, indexing="ij")
            current = bilinear_sample(smoothed, yy, xx)
        pyramid.insert(0, current)
    return pyramid


# a5  warp_image (lucas_kanade_pyramidal.py:66-97)
def warp_image(image, flow_u, flow_v):
    image = np.asarray(image, dtype=F32)
    h, w = image.shape
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    xw = xx + np.asarray(flow_u, dtype=F32)  # int64 + float32 -> float64
    yw = yy + np.asarray(flow_v, dtype=F32)
    return bilinear_sample(image, yw, xw)


# a6  upsample_flow (lucas_kanade_pyramidal.py:100-138)
def upsample_flow(flow_u, flow_v, target_shape: Tuple[int, int]):
    flow_u = np.asarray(flow_u, dtype=F32)
    flow_v = np.asarray(flow_v, dtype=F32)
    ch, cw = flow_u.shape
    th, tw = target_shape
    scale_y = th / ch
    scale_x = tw / cw
    yt = np.linspace(0, ch - 1, th)
    xt = np.linspace(0, cw - 1, tw)
    yy, xx = np.meshgrid(yt, xt, indexing="ij")
    uu = bilinear_sample(flow_u, yy, xx)
    vv = bilinear_sample(flow_v, yy, xx)
    return (uu * F32(scale_x)).astype(F32), (vv * F32(scale_y)).astype(F32)


# a7  lucas_kanade_pyramidal (lucas_kanade_pyramidal.py:141-228)
def lucas_kanade_pyramidal(
    frame_prev,
    frame_curr,
    num_levels: int = 3,
    window_size: int = 5,
    num_iterations: int = 3,
    trace: list | None = None,
):
    """Coarse-to-fine loop with the reference's global early exit.

    ``trace`` (optional list) receives (level, iteration, mean|du|, mean|dv|) for
    every executed iteration, so tests can compare the control decisions too.
    No prints and no plotting side effect (lucas_kanade_pyramidal.py:226 is
    visualisation, not arithmetic).
    """
    pyr_prev = build_gaussian_pyramid(frame_prev, num_levels)
    pyr_curr = build_gaussian_pyramid(frame_curr, num_levels)
    h, w = pyr_prev[0].shape
    flow_u = np.zeros((h, w), dtype=F32)
    flow_v = np.zeros((h, w), dtype=F32)
    for level in range(num_levels):
        img_prev = pyr_prev[level]
        img_curr = pyr_curr[level]
        if level > 0:
            flow_u, flow_v = upsample_flow(flow_u, flow_v, img_prev.shape)
        for iteration in range(num_iterations):
            warped = warp_image(img_curr, flow_u, flow_v)
            du, dv = lucas_kanade_single_scale(img_prev, warped, window_size)
            flow_u = (flow_u + du).astype(F32)
            flow_v = (flow_v + dv).astype(F32)
            mean_du = np.mean(np.abs(du))
            mean_dv = np.mean(np.abs(dv))
            if trace is not None:
                trace.append((level, iteration, float(mean_du), float(mean_dv)))
            if mean_du < CONVERGENCE_EPS and mean_dv < CONVERGENCE_EPS:
                break
    return flow_u, flow_v
