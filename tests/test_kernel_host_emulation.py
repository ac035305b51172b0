"""Exact-mode kernels run ON THE CPU from their own source files: lk_tile5_kernel (the default for window 5) and
warp_rows_kernel<double> / <float> (warp_image of the split refinement iteration).

tests/host_emul/emul_*.cpp #include optical-flow-fpga_b200/csrc/lk_tile5.cu / warp_rows.cuh and compile them with
g++ on top of tests/host_emul/cuda_on_host.h (every CUDA thread of a block is an OS thread, __syncthreads() a barrier,
__shared__ arrays statics, warp votes / shuffles exchanges through a scratch line, the *_rn / *_rd intrinsics single
IEEE operations).  What the GPU tests establish on the device
-- the kernel's bits are the oracle's -- is established here for the kernel's SOURCE: tile / halo indexing, the
scaled-tap Sobel, the product planes, np.sum's order, the border rule, the flow update and the per-block sums,
on ragged shapes.  It cannot see what only the device has (register allocation, the SASS the compiler emits for
a division); tests/test_gpu_parity.py covers that.
"""

import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import lk_float_oracle as orc

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "optical-flow-fpga_b200" / "csrc"
CUDA_INC = Path("/usr/local/cuda/include")
f32 = np.float32
_vp, _i = C.c_void_p, C.c_int


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers (vector types only; nothing CUDA is linked or run)")
    out = tmp_path_factory.mktemp("host_emul") / "libemul_lk_tile5.so"
    cmd = [gxx, "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-DOF_HOST_EMULATION",
           "-I", str(CSRC), "-I", str(CUDA_INC), str(ROOT / "tests" / "host_emul" / "emul_lk_tile5.cpp"), "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(str(out))
    lib.emul_lk_tile5_frames.argtypes = [_vp] * 4 + [_i] * 3
    lib.emul_lk_tile5_warped.argtypes = [_vp] * 6 + [_vp, _i, _vp, _vp] + [_i] * 7
    return lib


@pytest.fixture(scope="module")
def emul_warp(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers (vector types only; nothing CUDA is linked or run)")
    out = tmp_path_factory.mktemp("host_emul_warp") / "libemul_warp_rows.so"
    cmd = [gxx, "-O1", "-ffp-contract=off", "-frounding-math", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w",
           "-DOF_HOST_EMULATION", "-I", str(CSRC), "-I", str(CUDA_INC),
           str(ROOT / "tests" / "host_emul" / "emul_warp_rows.cpp"), "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(str(out))
    lib.emul_warp_rows.argtypes = [_vp] * 4 + [_i] * 6
    return lib


def bits(a):
    return np.ascontiguousarray(a, dtype=f32).view(np.uint32)


def ptr(a):
    return a.ctypes.data if a is not None else None


def run_frames(lib, prev, curr):
    p = np.ascontiguousarray(prev, f32)
    c = np.ascontiguousarray(curr, f32)
    u = np.full_like(p, np.nan)
    v = np.full_like(p, np.nan)
    lib.emul_lk_tile5_frames(ptr(p), ptr(c), ptr(u), ptr(v), p.shape[0], p.shape[1], p.shape[2])
    return u, v


@pytest.mark.parametrize("shape", [(1, 1), (5, 5), (4, 7), (16, 64), (17, 65), (37, 61), (40, 130), (70, 203)])
def test_single_scale_source_on_cpu_equals_oracle(emul, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a = (rng.standard_normal((2,) + shape) * 50).astype(f32)
    b = (a + rng.standard_normal((2,) + shape) * 3).astype(f32)
    a[1] = rng.integers(0, 256, shape)  # a uint8-valued pair as well (the verifier's kind of frame)
    b[1] = rng.integers(0, 256, shape)
    u, v = run_frames(emul, a, b)
    for k in range(2):
        uo, vo = orc.lucas_kanade_single_scale(a[k], b[k], 5)
        assert np.array_equal(bits(u[k]), bits(uo)), f"u, pair {k}"
        assert np.array_equal(bits(v[k]), bits(vo)), f"v, pair {k}"


def test_signed_zeros_and_denormals(emul):
    rng = np.random.default_rng(9)
    a = (rng.standard_normal((1, 30, 70)) * 1e-19).astype(f32)  # products underflow into the denormal range
    b = (rng.standard_normal((1, 30, 70)) * 1e-19).astype(f32)
    a[0, 3:9, 4:20] = 0.0
    b[0, 3:9, 4:20] = -0.0
    u, v = run_frames(emul, a, b)
    uo, vo = orc.lucas_kanade_single_scale(a[0], b[0], 5)
    assert np.array_equal(bits(u[0]), bits(uo)) and np.array_equal(bits(v[0]), bits(vo))
    same = np.full((1, 20, 70), 7.0, f32)  # every product is +-0: np.sum's +0.0 identity decides the sign
    u, v = run_frames(emul, same, same)
    uo, vo = orc.lucas_kanade_single_scale(same[0], same[0], 5)
    assert np.array_equal(bits(u[0]), bits(uo)) and np.array_equal(bits(v[0]), bits(vo))


@pytest.mark.parametrize("rows", [None, (10, 44, 12, 40)])
def test_refinement_iteration_source_on_cpu_equals_oracle(emul, rows):
    """SRC_WARPED: (prev, warped) -> flow_out = flow_in + d, per-block sums of |du|, |dv|, ping-pong selection,
    converged pairs skipped, row-band limits (rows computed / rows owned)."""
    rng = np.random.default_rng(21)
    B, H, W = 3, 50, 139
    from scipy.ndimage import gaussian_filter

    prev = gaussian_filter((rng.random((B, H, W)) * 255).astype(f32), (0, 1.2, 1.2)).astype(f32)
    curr = (prev + rng.standard_normal((B, H, W)).astype(f32)).astype(f32)
    fu = (rng.standard_normal((B, H, W)) * 1.5).astype(f32)
    fv = (rng.standard_normal((B, H, W)) * 1.5).astype(f32)
    warped = np.stack([orc.warp_image(curr[b], fu[b], fv[b]) for b in range(B)]).astype(f32)
    row_lo, row_hi, own_lo, own_hi = rows if rows else (0, H, 0, H)
    # pair 0: current buffer 0; pair 1: current buffer 1 (sel = 1); pair 2: converged, must not be touched
    sel = np.array([0, 1, 0], np.int32)
    done = np.array([0, 0, 1], np.int32)
    buf_u = [np.full((B, H, W), 123.0, f32), np.full((B, H, W), 123.0, f32)]
    buf_v = [np.full((B, H, W), 123.0, f32), np.full((B, H, W), 123.0, f32)]
    for b in range(B):
        buf_u[sel[b]][b] = fu[b]
        buf_v[sel[b]][b] = fv[b]
    nblk = ((W + 63) // 64) * ((row_hi - row_lo + 15) // 16)
    partial = np.full((B, nblk, 2), np.nan)
    emul.emul_lk_tile5_warped(ptr(prev), ptr(warped), ptr(buf_u[0]), ptr(buf_v[0]), ptr(buf_u[1]), ptr(buf_v[1]),
                              ptr(sel), 0, ptr(done), ptr(partial), B, H, W, row_lo, row_hi, own_lo, own_hi)
    for b in range(2):
        du, dv = orc.lucas_kanade_single_scale(prev[b], warped[b], 5)
        out_u, out_v = buf_u[sel[b] ^ 1][b], buf_v[sel[b] ^ 1][b]
        assert np.array_equal(bits(out_u[row_lo:row_hi]), bits((fu[b] + du)[row_lo:row_hi])), f"u, pair {b}"
        assert np.array_equal(bits(out_v[row_lo:row_hi]), bits((fv[b] + dv)[row_lo:row_hi])), f"v, pair {b}"
        # rows outside the band are not written
        assert (out_u[:row_lo] == 123.0).all() and (out_u[row_hi:] == 123.0).all()
        su = np.abs(du[own_lo:own_hi].astype(np.float64)).sum()
        sv = np.abs(dv[own_lo:own_hi].astype(np.float64)).sum()
        assert partial[b, :, 0].sum() == pytest.approx(su, rel=1e-12)
        assert partial[b, :, 1].sum() == pytest.approx(sv, rel=1e-12)
    assert (buf_u[1][2] == 123.0).all() and np.isnan(partial[2]).all()  # converged pair untouched


# ---------------------------------------------------------------------------------------
# warp_rows_kernel
# ---------------------------------------------------------------------------------------
def run_warp(lib, img, fu, fv, exact, rows=None):
    h, w = img.shape
    out = np.full((h, w), np.nan, f32)
    lo, hi = rows if rows else (0, h)
    lib.emul_warp_rows(ptr(img), ptr(fu), ptr(fv), ptr(out), 1, h, w, lo, hi, int(exact))
    return out


def flow_cases(shape, rng):
    moderate = (rng.standard_normal(shape) * 1.5).astype(f32), (rng.standard_normal(shape) * 1.5).astype(f32)
    tiny = (rng.standard_normal(shape) * 1e-6).astype(f32), (rng.standard_normal(shape) * 1e-5).astype(f32)
    mixed_u, mixed_v = moderate[0].copy(), moderate[1].copy()
    m = rng.random(shape) < 0.02  # a few tiny values: their warps take the rounded-coordinate path
    mixed_u[m] = tiny[0][m]
    mixed_v[m] = 0.0
    mixed_u[0, 0], mixed_v[1, 1], mixed_u[2, 2], mixed_v[3, 3] = np.nan, np.inf, -5e6, 1e30
    return {
        "moderate": moderate,
        "leaves_the_frame": ((rng.standard_normal(shape) * 30).astype(f32), (rng.standard_normal(shape) * 30).astype(f32)),
        "tiny": tiny,
        "mixed_with_nan_inf_huge": (mixed_u, mixed_v),
        "zero": (np.zeros(shape, f32), np.zeros(shape, f32)),
        "integer": (rng.integers(-3, 4, shape).astype(f32), rng.integers(-3, 4, shape).astype(f32)),
    }


@pytest.mark.parametrize("shape", [(7, 5), (40, 300), (21, 1100)])
def test_warp_rows_double_source_on_cpu_is_warp_image(emul_warp, shape):
    """Exact flavour: the reference's bits for every flow, including the values whose float64 coordinate sum is
    rounded (bilinear_f64 path), NaN / inf / huge flow, samples outside the frame, widths that are not a multiple of
    the 1024 columns a CTA covers."""
    rng = np.random.default_rng(shape[1])
    img = (rng.random(shape) * 255).astype(f32)
    for name, (fu, fv) in flow_cases(shape, rng).items():
        with np.errstate(all="ignore"):
            ref = orc.warp_image(img, fu, fv)
        got = run_warp(emul_warp, img, fu, fv, exact=True)
        assert np.array_equal(bits(got), bits(ref)), name


def test_warp_rows_float_flavour_is_within_one_rounding(emul_warp):
    """Fast flavour (float32 sample fractions): identical to warp_image for non-negative flow, a rounding of the
    fraction away from it for negative flow (DESIGN.md K3 fast)."""
    rng = np.random.default_rng(4)
    shape = (40, 300)
    img = (rng.random(shape) * 255).astype(f32)
    fu = np.abs(rng.standard_normal(shape) * 2).astype(f32)
    fv = np.abs(rng.standard_normal(shape) * 2).astype(f32)
    assert np.array_equal(bits(run_warp(emul_warp, img, fu, fv, exact=False)), bits(orc.warp_image(img, fu, fv)))
    fu, fv = -fu, -fv
    got, ref = run_warp(emul_warp, img, fu, fv, exact=False), orc.warp_image(img, fu, fv)
    differ = bits(got) != bits(ref)
    assert 0 < differ.mean() < 0.1
    assert np.abs(got - ref).max() <= 255 * 2.0**-23  # the fraction moves by <= 2^-24, the taps are <= 255 apart


def test_warp_rows_row_range(emul_warp):
    rng = np.random.default_rng(8)
    shape = (30, 70)
    img = (rng.random(shape) * 255).astype(f32)
    fu, fv = (rng.standard_normal(shape)).astype(f32), (rng.standard_normal(shape)).astype(f32)
    got = run_warp(emul_warp, img, fu, fv, exact=True, rows=(7, 19))
    ref = orc.warp_image(img, fu, fv)
    assert np.array_equal(bits(got[7:19]), bits(ref[7:19]))
    assert np.isnan(got[:7]).all() and np.isnan(got[19:]).all()
