"""Arithmetic identities the exact-mode CUDA kernels rely on, checked on the CPU against the oracle.

The kernels themselves only run on the GPU (tests/test_gpu_parity.py); what can be pinned here is that the
REARRANGED operation sequences they use are the reference's bits: each function below restates one kernel
stage in NumPy float32, operation by operation, and is compared bit for bit with the oracle (which the golden
fixtures pin to the reference itself).
"""

import numpy as np
import pytest

from oracle import lk_float_oracle as orc

f32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def frames(kind, shape, rng):
    if kind == "uint8":
        return rng.integers(0, 256, shape).astype(f32), rng.integers(0, 256, shape).astype(f32)
    if kind == "float":
        p = (rng.standard_normal(shape) * 50).astype(f32)
        return p, (p + rng.standard_normal(shape)).astype(f32)
    # denormal range, with signed zeros: products by 0.125 / 0.25 round, x * 0.0 keeps its sign
    p = (rng.standard_normal(shape) * 1e-38).astype(f32)
    c = (rng.standard_normal(shape) * 1e-38).astype(f32)
    p[3, 4], c[3, 4] = 0.0, -0.0
    p[5, 5], c[5, 5] = -0.0, -0.0
    p[7, 1], c[7, 1] = -0.0, 0.0
    return p, c


def sobel_like_lk_tile5(p, c):
    """Stage A + B of lk_tile5_kernel (optical-flow-fpga_b200/csrc/lk_tile5.cu): E = avg * 0.125 and
    D = avg * 0.25 are stored once, a tap is one add / subtract of E or D, a zero tap adds D * 0.0.
    Reference: scipy.signal.convolve2d(avg, sobel / 8, 'same', 'symm') as python/lucas_kanade_core.py:39-40
    calls it (true convolution, float32 accumulator fed in kernel order)."""
    h, w = p.shape
    avg = ((p + c) * f32(0.5)).astype(f32)
    it = (p - c).astype(f32)
    e = np.pad((avg * f32(0.125)).astype(f32), 1, mode="edge")
    d = np.pad((avg * f32(0.25)).astype(f32), 1, mode="edge")
    z = (d * f32(0.0)).astype(f32)

    def s(a, dy, dx):
        return a[dy:dy + h, dx:dx + w]

    ax = np.zeros_like(p)
    ay = np.zeros_like(p)
    # j = 0 (frame row 2): kx = -.125, 0, .125   ky = -.125, -.25, -.125
    ax = ax - s(e, 2, 2); ay = ay - s(e, 2, 2)
    ax = ax + s(z, 2, 1); ay = ay - s(d, 2, 1)
    ax = ax + s(e, 2, 0); ay = ay - s(e, 2, 0)
    # j = 1 (frame row 1): kx = -.25, 0, .25      ky = 0, 0, 0
    ax = ax - s(d, 1, 2); ay = ay + s(z, 1, 2)
    ax = ax + s(z, 1, 1); ay = ay + s(z, 1, 1)
    ax = ax + s(d, 1, 0); ay = ay + s(z, 1, 0)
    # j = 2 (frame row 0): kx = -.125, 0, .125   ky = .125, .25, .125
    ax = ax - s(e, 0, 2); ay = ay + s(e, 0, 2)
    ax = ax + s(z, 0, 1); ay = ay + s(d, 0, 1)
    ax = ax + s(e, 0, 0); ay = ay + s(e, 0, 0)
    assert ax.dtype == f32 and ay.dtype == f32
    return ax, ay, it


@pytest.mark.parametrize("kind", ["uint8", "float", "denormal"])
@pytest.mark.parametrize("shape", [(37, 53), (8, 8)])
def test_scaled_tap_sobel_is_the_reference_sobel(kind, shape):
    rng = np.random.default_rng(hash((kind, shape)) % 2**32)
    p, c = frames(kind, shape, rng)
    ix, iy, it = orc.compute_gradients(p, c)
    ax, ay, at = sobel_like_lk_tile5(p, c)
    assert np.array_equal(bits(ax), bits(ix))
    assert np.array_equal(bits(ay), bits(iy))
    assert np.array_equal(bits(at), bits(it))


def np25_streaming(v):
    """np25_add / np25_finish of lk_tile5.cu (np_add<5> of lk_tile.cu): taps in index order, lane t & 7 for
    t < 24, the fixed tree, the tail, and the + 0.0 identity np.add.reduce starts from."""
    lane = [None] * 8
    for t in range(24):
        lane[t & 7] = v[t] if t < 8 else f32(lane[t & 7] + v[t])
    res = f32(f32(f32(lane[0] + lane[1]) + f32(lane[2] + lane[3])) + f32(f32(lane[4] + lane[5]) + f32(lane[6] + lane[7])))
    res = f32(res + v[24])
    return f32(f32(0.0) + res)


def test_streaming_lane_sum_is_np_sum_on_25_taps():
    rng = np.random.default_rng(5)
    with np.errstate(over="ignore"):
        for trial in range(400):
            scale = 10.0 ** rng.integers(-20, 20)
            v = (rng.standard_normal(25) * scale).astype(f32)
            if trial % 7 == 0:
                v[:] = -0.0  # all products -0.0: np.sum gives +0.0 (identity), so must the kernel
            assert bits(np25_streaming(v)) == bits(np.sum(v)), trial


def test_adjacent_windows_share_three_lane_sums():
    """The compiler folds the lane sums two horizontally adjacent 5 x 5 windows have in common (392 instead of 500
    FADD per thread in lk_tile5_kernel's stage C).  That is legitimate only if those lanes receive the same
    values in the same order: lanes 1, 3, 6 of window x are lanes 0, 2, 5 of window x + 1."""
    def lane_taps(lane):  # (row, column) of the taps lane `lane` accumulates, in order
        return [divmod(t, 5) for t in range(24) if (t & 7) == lane]

    shared = []
    for la in range(8):
        for lb in range(8):
            if [(i, k) for i, k in lane_taps(la)] == [(i, k + 1) for i, k in lane_taps(lb)]:
                shared.append((la, lb))
    assert shared == [(1, 0), (3, 2), (6, 5)]


def test_warp_sample_fraction_float64_exact_float32_not_always():
    """warp_image's sample is taken at the float64 coordinate y + v (lucas_kanade_pyramidal.py:88-92, int64 + float32
    -> float64).  warp_rows_kernel splits it into the integer y + floor(v) and the fraction v - floor(v).
    * <double> (exact mode) takes the difference in float64.  That is the reference's fraction whenever the
      reference's sum y + v is exact: v == 0 or |v| >= 2^-14 for coordinates below 2^16 -- the kernel's test; any
      other warp goes through bilinear_f64, which forms the rounded sum like the reference does.
    * <float> (fast mode) takes it in float32, which is exact for v >= 0 only (documented deviation)."""
    rng = np.random.default_rng(11)
    v = np.concatenate([rng.uniform(-40, 40, 200000), rng.uniform(-1e-3, 1e-3, 20000), rng.uniform(-1e-6, 1e-6, 1000),
                        np.arange(-50, 50), [2.0**-14, -(2.0**-14), np.nextafter(f32(2.0**-14), f32(1))]]).astype(f32)
    y = rng.integers(0, 65536, v.size).astype(np.int64)
    coord = y + v  # the reference's promotion
    assert coord.dtype == np.float64
    ref_floor = np.floor(coord)
    ref_frac = coord - ref_floor
    fl = np.floor(v)  # float32, exact
    frac64 = v.astype(np.float64) - fl.astype(np.float64)
    kernel_takes_split = (v == 0) | (np.abs(v) >= f32(2.0**-14))
    assert np.array_equal(ref_floor[kernel_takes_split], (y + fl.astype(np.int64))[kernel_takes_split])
    assert np.array_equal(ref_frac[kernel_takes_split], frac64[kernel_takes_split])
    # outside that set the reference's own coordinate is rounded: the split would not be its fraction
    assert not np.array_equal(ref_frac[~kernel_takes_split], frac64[~kernel_takes_split])
    frac32 = (v - fl).astype(f32)
    exact32 = frac32.astype(np.float64) == frac64
    assert exact32[v >= 0].all()
    assert not exact32[v < 0].all()  # the documented fast-mode deviation (DESIGN.md K3 fast)
