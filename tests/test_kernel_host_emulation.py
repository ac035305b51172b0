"""Exact-mode kernels run ON THE CPU from their own source files: lk_tile5_kernel (the default for window 5),
lk_tile_kernel (every window; frames, gradients, in-tile warp), iter_finalize_kernel, warp_rows_kernel<double> /
<float> (warp_image of the split refinement iteration), pyramid.cu / pyramid_march.cu with their launchers
(gradients, both Gaussian-pyramid kernels, warp_kernel, both upsample kernels), and motion.cu, lk_fixed.cu, metrics.cu
(the fixture generators' warps, the RTL's integer datapath, the verifier's metrics), and lk_march.cu -- the fast marching
kernels with their TMA ring, mbarriers, packed pairs and warp shuffles (stand-ins in tests/host_emul/emul_lk_march.cpp).

tests/host_emul/emul_*.cpp #include optical-flow-fpga_b200/csrc/lk_tile5.cu / lk_tile.cu / warp_rows.cuh and compile them with
g++ on top of tests/host_emul/cuda_on_host.h (every CUDA thread of a block is an OS thread, __syncthreads() a barrier,
__shared__ arrays statics, warp votes / shuffles exchanges through a scratch line, the *_rn / *_rd intrinsics single
IEEE operations).  What the GPU tests establish on the device
-- the kernel's bits are the oracle's -- is established here for the kernel's SOURCE: tile / halo indexing, the
scaled-tap Sobel, the product planes, np.sum's order, the border rule, the flow update and the per-block sums,
on ragged shapes.  It cannot see what only the device has (register allocation, the SASS the compiler emits for
a division); tests/test_gpu_parity.py covers that.
"""

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import flow_metrics_oracle as fmo
from oracle import lk_fixed_oracle as fxo
from oracle import lk_float_oracle as orc
from oracle import pattern_oracle as pto

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "optical-flow-fpga_b200" / "csrc"
CUDA_INC = Path("/usr/local/cuda/include")
f32 = np.float32
_vp, _i = C.c_void_p, C.c_int


def _build(tmp_path_factory, name, *more, defines=()):
    gxx = shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers (vector types only; nothing CUDA is linked or run)")
    out = tmp_path_factory.mktemp(name) / f"lib{name}.so"
    srcs = [str(ROOT / "tests" / "host_emul" / f"{n}.cpp") for n in (name,) + more]
    extra = os.environ.get("OF_EMUL_EXTRA_FLAGS", "").split()  # tests/host_emul/asan_check.sh: AddressSanitizer
    cmd = [gxx, "-O1", "-ffp-contract=off", "-frounding-math", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w",
           "-DOF_HOST_EMULATION", *defines, *extra, "-I", str(CSRC), "-I", str(CUDA_INC), *srcs, "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(str(out))


@pytest.fixture(scope="module", params=["lk_exact_march", "lk_tile5"])
def emul(tmp_path_factory, request):
    """The two exact window-5 kernels behind the same two calls: lk_exact_march_kernel (the default; through its
    launcher) and lk_tile5_kernel (OF_B200_EXACT=tile).  Every test that takes this fixture runs on both."""
    name = request.param
    lib = _build(tmp_path_factory, f"emul_{name}")
    frames, warped = getattr(lib, f"emul_{name}_frames"), getattr(lib, f"emul_{name}_warped")
    frames.argtypes = [_vp] * 4 + [_i] * 3
    warped.argtypes = [_vp] * 6 + [_vp, _i, _vp, _vp] + [_i] * 7
    lib.emul_lk_tile5_frames, lib.emul_lk_tile5_warped = frames, warped  # one spelling in the tests below
    return lib


@pytest.fixture(scope="module")
def emul_v1(tmp_path_factory):
    # launch_lk_tile dispatches to launch_lk_exact_march / launch_lk_tile5
    lib = _build(tmp_path_factory, "emul_lk_tile", "emul_lk_tile5", "emul_lk_exact_march", defines=("-DEMUL_EXACT_MARCH_NO_TILE_GEOMETRY",))
    lib.emul_lk_tile.argtypes = [_i, _i] + [_vp] * 5 + [_i] * 3
    lib.emul_lk_tile_refine.argtypes = [_i, _i] + [_vp] * 6 + [_vp, _i, _vp, _vp] + [_i] * 7
    lib.emul_iter_finalize.argtypes = [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, C.c_long, _i, _i]
    return lib


@pytest.fixture(scope="module")
def emul_pyr(tmp_path_factory):
    lib = _build(tmp_path_factory, "emul_pyramid", "emul_pyramid_march")
    lib.emul_pyramid_down.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i]
    lib.emul_gradients.argtypes = [_vp] * 5 + [_i] * 3
    lib.emul_warp.argtypes = [_vp] * 4 + [_i] * 3
    lib.emul_upsample_flow.argtypes = [_vp] * 4 + [_i] * 7
    return lib


@pytest.fixture(scope="module")
def emul_warp(tmp_path_factory):
    lib = _build(tmp_path_factory, "emul_warp_rows")
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


def test_refinement_band_needs_the_warped_plane_on_three_halo_rows_only(emul):
    """Row-band use of the split form: warp_rows_kernel has to cover rows [row_lo - 3, row_hi + 3) (window // 2 + 1
    for the Sobel tap); whatever lies outside -- here NaN -- must not reach the rows that are written."""
    rng = np.random.default_rng(5)
    H, W = 64, 100
    prev = (rng.random((1, H, W)) * 255).astype(f32)
    warped_full = (prev + rng.standard_normal((1, H, W))).astype(f32)
    fu = rng.standard_normal((1, H, W)).astype(f32)
    fv = rng.standard_normal((1, H, W)).astype(f32)
    row_lo, row_hi = 18, 41
    warped = np.full_like(warped_full, np.nan)
    warped[:, row_lo - 3:row_hi + 3] = warped_full[:, row_lo - 3:row_hi + 3]
    out_u, out_v = np.full_like(fu, 7.0), np.full_like(fv, 7.0)
    nblk = ((W + 63) // 64) * ((row_hi - row_lo + 15) // 16)
    partial = np.zeros((1, nblk, 2))
    emul.emul_lk_tile5_warped(ptr(prev), ptr(warped), ptr(fu), ptr(fv), ptr(out_u), ptr(out_v), None, 0, None, ptr(partial),
                              1, H, W, row_lo, row_hi, row_lo, row_hi)
    du, dv = orc.lucas_kanade_single_scale(prev[0], warped_full[0], 5)
    assert np.array_equal(bits(out_u[0, row_lo:row_hi]), bits((fu[0] + du)[row_lo:row_hi]))
    assert np.array_equal(bits(out_v[0, row_lo:row_hi]), bits((fv[0] + dv)[row_lo:row_hi]))
    assert np.isfinite(partial).all()


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


# ---------------------------------------------------------------------------------------
# lk_tile_kernel (first exact kernel: every window, gradients, in-tile warp) and the iteration loop
# ---------------------------------------------------------------------------------------
SRC_FRAMES, SRC_WARP, SRC_GRADS, SRC_WARPED = 0, 1, 2, 3


@pytest.mark.parametrize("w", [1, 3, 5, 7, 9, 11])
def test_first_tile_kernel_source_on_cpu_every_window(emul_v1, w):
    rng = np.random.default_rng(w)
    shape = (2, 27, 75)
    a = (rng.standard_normal(shape) * 50).astype(f32)
    b = (a + rng.standard_normal(shape) * 3).astype(f32)
    u, v = np.full_like(a, np.nan), np.full_like(a, np.nan)
    assert emul_v1.emul_lk_tile(SRC_FRAMES, w, ptr(a), ptr(b), None, ptr(u), ptr(v), 2, shape[1], shape[2]) == 0
    for k in range(2):
        uo, vo = orc.lucas_kanade_single_scale(a[k], b[k], w)
        assert np.array_equal(bits(u[k]), bits(uo)) and np.array_equal(bits(v[k]), bits(vo)), f"pair {k}"
    # the same windows from gradient planes (lucas_kanade_from_gradients)
    ix, iy, it = (rng.standard_normal(shape[1:]).astype(f32) for _ in range(3))
    u1, v1 = np.full_like(ix, np.nan), np.full_like(ix, np.nan)
    assert emul_v1.emul_lk_tile(SRC_GRADS, w, ptr(ix), ptr(iy), ptr(it), ptr(u1), ptr(v1), 1, shape[1], shape[2]) == 0
    uo, vo = orc.lucas_kanade_from_gradients(ix, iy, it, w)
    assert np.array_equal(bits(u1), bits(uo)) and np.array_equal(bits(v1), bits(vo))


def _level_loop(emul_v1, emul_warp, emul5, prev, curr, iters, form):
    """The exact-mode iteration loop of one pyramid level as of_api.cu's driver runs it, kernel by kernel:
    form 'fused' = lk_tile_kernel<SRC_WARP>; 'split' = warp_rows_kernel<double> + lk_tile5_kernel<SRC_WARPED>;
    'split_v1' = warp_rows_kernel<double> + lk_tile_kernel<SRC_WARPED>.  Then iter_finalize_kernel."""
    B, H, W = prev.shape
    start = iters & 1
    bu = [np.zeros((B, H, W), f32), np.zeros((B, H, W), f32)]
    bv = [np.zeros((B, H, W), f32), np.zeros((B, H, W), f32)]
    sel = np.zeros(B, np.int32)
    done = np.zeros(B, np.int32)
    executed = np.zeros(B, np.int32)
    resid = np.zeros((B, iters, 2), f32)
    nblk = ((W + 63) // 64) * ((H + 15) // 16)
    partial = np.zeros((B, nblk, 2))
    warped = np.zeros((B, H, W), f32)
    for it in range(iters):
        if form == "fused":
            emul_v1.emul_lk_tile_refine(SRC_WARP, 5, ptr(prev), ptr(curr), ptr(bu[0]), ptr(bv[0]), ptr(bu[1]), ptr(bv[1]),
                                        ptr(sel), start, ptr(done), ptr(partial), B, H, W, 0, H, 0, H)
        else:
            for b in range(B):
                if not done[b]:
                    cur = sel[b] ^ start
                    emul_warp.emul_warp_rows(ptr(curr[b]), ptr(bu[cur][b]), ptr(bv[cur][b]), ptr(warped[b]), 1, H, W, 0, H, 1)
            if form == "split":
                emul5.emul_lk_tile5_warped(ptr(prev), ptr(warped), ptr(bu[0]), ptr(bv[0]), ptr(bu[1]), ptr(bv[1]), ptr(sel),
                                           start, ptr(done), ptr(partial), B, H, W, 0, H, 0, H)
            else:
                emul_v1.emul_lk_tile_refine(SRC_WARPED, 5, ptr(prev), ptr(warped), ptr(bu[0]), ptr(bv[0]), ptr(bu[1]),
                                            ptr(bv[1]), ptr(sel), start, ptr(done), ptr(partial), B, H, W, 0, H, 0, H)
        emul_v1.emul_iter_finalize(ptr(partial), nblk, H, W, ptr(sel), ptr(done), ptr(executed), 1, ptr(resid), iters * 2,
                                   it, B)
    out_u = np.stack([bu[sel[b] ^ start][b] for b in range(B)])
    out_v = np.stack([bv[sel[b] ^ start][b] for b in range(B)])
    return out_u, out_v, executed, resid


@pytest.mark.parametrize("form", ["fused", "split", "split_v1"])
def test_exact_iteration_loop_source_on_cpu_equals_oracle(emul_v1, emul_warp, emul, form):
    """One level, three iterations, three pairs: a moving pair, a pair with sub-pixel motion of the other sign, and a
    pair without motion that must stop after its first iteration (the reference's early exit) while the others go on."""
    from scipy.ndimage import gaussian_filter, shift

    rng = np.random.default_rng(33)
    H, W = 40, 90
    base = gaussian_filter((rng.random((3, H, W)) * 255).astype(f32), (0, 1.5, 1.5)).astype(f32)
    prev = base.copy()
    curr = np.stack([shift(base[0], (0.4, -0.7), order=1, mode="nearest"), shift(base[1], (-0.3, 0.2), order=1, mode="nearest"),
                     base[2]]).astype(f32)
    u, v, executed, resid = _level_loop(emul_v1, emul_warp, emul, prev, curr, 3, form)
    for b in range(3):
        trace = []
        uo, vo = orc.lucas_kanade_pyramidal(prev[b], curr[b], 1, 5, 3, trace=trace)
        assert np.array_equal(bits(u[b]), bits(uo)), f"u, pair {b}"
        assert np.array_equal(bits(v[b]), bits(vo)), f"v, pair {b}"
        assert executed[b] == len(trace)
        for (_, it, mu, mv) in trace:  # float64 sums vs the reference's float32 pairwise mean
            assert resid[b, it, 0] == pytest.approx(mu, rel=1e-5, abs=1e-9)
            assert resid[b, it, 1] == pytest.approx(mv, rel=1e-5, abs=1e-9)
    assert executed.tolist() == [3, 3, 1]


# ---------------------------------------------------------------------------------------
# pyramid.cu / pyramid_march.cu through their own launchers
# ---------------------------------------------------------------------------------------
def run_pyramid_down(lib, img, rows=None, fast=False):
    wts = orc.gaussian_weights(2.0)
    b, h, w = img.shape
    oh, ow = h // 2, w // 2
    out = np.full((b, oh, ow), np.nan, f32)
    lo, hi = rows if rows else (0, oh)
    assert lib.emul_pyramid_down(ptr(img), ptr(out), b, h, w, oh, ow, ptr(wts), (len(wts) - 1) // 2, lo, hi, int(fast)) == 0
    return out


def test_pyramid_kernels_source_on_cpu_random_shapes(emul_pyr):
    """One pyramid level (gaussian_filter sigma 2 + bilinear decimation) for 40 random shapes: the marching kernel
    (>= 16 x 248; strips of 240 columns, bands planned per launch) and the tile kernel (everything smaller),
    chosen by launch_pyramid_down exactly as on the device."""
    rng = np.random.default_rng(17)
    shapes = [(16, 248), (16, 247), (8, 8), (2, 2), (3, 9), (45, 67), (17, 249), (64, 481), (31, 720), (130, 250)]
    while len(shapes) < 40:
        shapes.append((int(rng.integers(2, 120)), int(rng.integers(2, 700))))
    for h, w in shapes:
        img = (rng.random((1, h, w)) * 255).astype(f32)
        got = run_pyramid_down(emul_pyr, img)
        want = orc.build_gaussian_pyramid(img[0], 2)[0]
        assert np.array_equal(bits(got[0]), bits(want)), (h, w)


def test_pyramid_march_batch_row_range_and_fma_flavour(emul_pyr):
    rng = np.random.default_rng(3)
    img = (rng.random((3, 96, 530)) * 255).astype(f32)
    want = np.stack([orc.build_gaussian_pyramid(img[b], 2)[0] for b in range(3)])
    assert np.array_equal(bits(run_pyramid_down(emul_pyr, img)), bits(want))
    part = run_pyramid_down(emul_pyr, img, rows=(11, 37))  # row-band mode: only these coarse rows
    assert np.array_equal(bits(part[:, 11:37]), bits(want[:, 11:37]))
    assert np.isnan(part[:, :11]).all() and np.isnan(part[:, 37:]).all()
    fma = run_pyramid_down(emul_pyr, img, fast=True)  # fast mode: fused multiply-adds, last-bit differences only
    assert np.abs(fma - want).max() <= np.spacing(f32(255.0))
    assert (bits(fma) != bits(want)).mean() < 1e-3


def test_gradients_warp_upsample_launchers_source_on_cpu(emul_pyr):
    rng = np.random.default_rng(12)
    for h, w in [(5, 7), (33, 100), (64, 257)]:
        p = (rng.standard_normal((2, h, w)) * 50).astype(f32)
        c = (p + rng.standard_normal((2, h, w))).astype(f32)
        ix, iy, it = (np.full_like(p, np.nan) for _ in range(3))
        assert emul_pyr.emul_gradients(ptr(p), ptr(c), ptr(ix), ptr(iy), ptr(it), 2, h, w) == 0
        for b in range(2):
            for got, want in zip((ix[b], iy[b], it[b]), orc.compute_gradients(p[b], c[b])):
                assert np.array_equal(bits(got), bits(want))
        fu = (rng.standard_normal((2, h, w)) * 2).astype(f32)
        fv = (rng.standard_normal((2, h, w)) * 1e-6).astype(f32)  # tiny: the rounded float64 coordinate sum
        out = np.full_like(p, np.nan)
        assert emul_pyr.emul_warp(ptr(p), ptr(fu), ptr(fv), ptr(out), 2, h, w) == 0
        for b in range(2):
            assert np.array_equal(bits(out[b]), bits(orc.warp_image(p[b], fu[b], fv[b])))
    # upsample_flow: the tile kernel (true upsampling, step <= ~0.5) and the generic kernel (any other ratio)
    for (ch, cw), (th, tw) in [((20, 31), (40, 62)), ((20, 31), (41, 63)), ((33, 40), (33, 40)), ((30, 50), (12, 21)),
                               ((7, 300), (15, 601)), ((1, 5), (3, 9))]:
        cu = (rng.standard_normal((2, ch, cw)) * 3).astype(f32)
        cv = (rng.standard_normal((2, ch, cw)) * 3).astype(f32)
        fu, fv = np.full((2, th, tw), np.nan, f32), np.full((2, th, tw), np.nan, f32)
        assert emul_pyr.emul_upsample_flow(ptr(cu), ptr(cv), ptr(fu), ptr(fv), 2, ch, cw, th, tw, 0, th) == 0
        for b in range(2):
            wu, wv = orc.upsample_flow(cu[b], cv[b], (th, tw))
            assert np.array_equal(bits(fu[b]), bits(wu)) and np.array_equal(bits(fv[b]), bits(wv)), ((ch, cw), (th, tw))


# ---------------------------------------------------------------------------------------
# motion.cu, lk_fixed.cu, metrics.cu
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emul_motion(tmp_path_factory):
    lib = _build(tmp_path_factory, "emul_motion")
    lib.emul_apply_motion.argtypes = [_vp] * 4 + [_i] * 3 + [C.c_double]
    lib.emul_warp_affine.argtypes = [_vp] * 3 + [_i] * 4
    return lib


def test_fixture_generator_warps_source_on_cpu(emul_motion):
    """apply_motion (scipy.ndimage.shift, order 1, constant 128) and apply_motion_opencv (cv2.warpAffine's fixed-point
    bilinear warp) against the pattern oracle, which tests/golden/motion.npz pins to the reference's generators."""
    rng = np.random.default_rng(6)
    shifts = [(0.0, 0.0), (0.5, -0.25), (-1.75, 2.0), (3.0, 3.0), (-0.001, 0.999), (40.0, 0.0), (0.3, -60.0), (1e-9, -1e-9)]
    for h, w in [(24, 40), (7, 300), (33, 257)]:
        src = rng.integers(0, 256, (len(shifts), h, w)).astype(np.uint8)
        dx = np.array([s[0] for s in shifts])
        dy = np.array([s[1] for s in shifts])
        dst = np.zeros_like(src)
        assert emul_motion.emul_apply_motion(ptr(src), ptr(dst), ptr(dx), ptr(dy), len(shifts), h, w, 128.0) == 0
        for k, (sx, sy) in enumerate(shifts):
            assert np.array_equal(dst[k], pto.apply_motion(src[k], sx, sy)), (h, w, sx, sy)
        params = [(0, 0, 0, 1), (1.5, -0.5, 0, 1), (0, 0, 2.0, 1), (0, 0, -5.0, 1.02), (3.2, 1.1, 0.5, 0.98), (-10, 7, 30, 1.2)]
        params += [tuple(rng.uniform(-4, 4, 2)) + (float(rng.uniform(-10, 10)), float(rng.uniform(0.9, 1.1))) for _ in range(50)]
        mats = [pto.motion_matrix(w, h, *q) for q in params]  # > 48 frames: the launcher splits the batch
        minv = np.ascontiguousarray(np.stack([pto.invert_affine(m) for m in mats]).reshape(len(mats), 6))
        src = rng.integers(0, 256, (len(mats), h, w)).astype(np.uint8)
        dst = np.zeros_like(src)
        assert emul_motion.emul_warp_affine(ptr(src), ptr(dst), ptr(minv), len(mats), h, w, 128) == 0
        for k, m in enumerate(mats):
            assert np.array_equal(dst[k], pto.warp_affine_u8(src[k], m)), (h, w, params[k])


@pytest.mark.parametrize("quirk", [True, False])
def test_fixed_point_tile_kernel_source_on_cpu(tmp_path_factory, quirk):
    lib = _build(tmp_path_factory, "emul_fixed")
    lib.emul_lk_fixed.argtypes = [_vp] * 4 + [_i] * 4
    rng = np.random.default_rng(int(quirk) + 40)
    for h, w in [(7, 7), (8, 9), (33, 70), (50, 129)]:
        p8 = rng.integers(0, 256, (2, h, w)).astype(np.uint8)
        c8 = np.clip(np.roll(p8, 1, axis=2).astype(np.int32) + rng.integers(-3, 4, p8.shape), 0, 255).astype(np.uint8)
        p8[1], c8[1] = rng.integers(0, 256, (h, w)), rng.integers(0, 256, (h, w))  # unrelated frames: wrap-arounds, clamps
        u = np.full((2, h, w), 77, np.int16)
        v = np.full((2, h, w), 77, np.int16)
        assert lib.emul_lk_fixed(ptr(p8), ptr(c8), ptr(u), ptr(v), 2, h, w, int(quirk)) == 0
        for b in range(2):
            uo, vo = fxo.lk_single_scale_fx(p8[b], c8[b], mirror_avg_quirk=quirk)
            assert np.array_equal(u[b], uo) and np.array_equal(v[b], vo), (h, w, b)


def test_metrics_kernels_source_on_cpu(tmp_path_factory):
    lib = _build(tmp_path_factory, "emul_metrics")
    lib.emul_metrics_blocks_per_pair.argtypes = [_i, _i]
    lib.emul_flow_metrics.argtypes = [_vp] * 4 + [_i] * 7 + [_vp, _vp]
    rng = np.random.default_rng(2)
    B, H, W = 3, 60, 90
    truth = np.array([[1.5, -0.5], [0.0, 0.0], [-2.0, 3.0]], f32)
    u = (truth[:, 0, None, None] + rng.standard_normal((B, H, W)) * 0.3).astype(f32)
    v = (truth[:, 1, None, None] + rng.standard_normal((B, H, W)) * 0.3).astype(f32)
    u[1], v[1] = 0.0, 0.0  # no motion in truth and prediction: angular error 0 by definition (flow_metrics.py:143-147)
    y0, y1, x0, x1 = 10, 50, 5, 85
    ut, vt = np.ascontiguousarray(truth[:, 0]), np.ascontiguousarray(truth[:, 1])
    partial = np.zeros((B, lib.emul_metrics_blocks_per_pair(y1 - y0, x1 - x0), 8))
    out = np.zeros((B, 5))
    assert lib.emul_flow_metrics(ptr(u), ptr(v), ptr(ut), ptr(vt), B, H, W, y0, y1, x0, x1, ptr(partial), ptr(out)) == 0
    mask = np.zeros((H, W), bool)
    mask[y0:y1, x0:x1] = True
    for b in range(B):
        want = fmo.all_metrics(u[b], v[b], float(truth[b, 0]), float(truth[b, 1]), mask)
        for k, name in enumerate(("mae_u", "mae_v", "rmse", "epe", "aae")):
            assert out[b, k] == pytest.approx(want[name], rel=2e-6, abs=1e-7), (b, name)


# ---------------------------------------------------------------------------------------
# lk_march.cu: the fast marching kernels (single scale float / uint8 / fixed point, refinement iteration)
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emul_march(tmp_path_factory):
    lib = _build(tmp_path_factory, "emul_lk_march")
    lib.emul_lk_march.argtypes = [_vp] * 4 + [_i] * 5
    lib.emul_lk_march_u8.argtypes = [_vp] * 4 + [_i] * 4
    lib.emul_lk_march_fx.argtypes = [_vp] * 4 + [_i] * 4
    lib.emul_lk_refine.argtypes = [_i] + [_vp] * 6 + [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp] + [_i] * 10 + [_vp, _i]
    lib.emul_lk_refine_units_per_pair.argtypes = [_i, _i, _i]
    return lib


TMA_OR_ERROR, PLAIN_LOADS = 1, 2


@pytest.mark.parametrize("window", [5, 7])
@pytest.mark.parametrize("shape", [(5, 8), (9, 120), (24, 128), (17, 132), (40, 248), (31, 364)])
def test_marching_kernel_source_on_cpu_uint8_valued_frames(emul_march, shape, window):
    """Fast mode is bit-identical to the reference on uint8-valued frames (every partial sum is exact, so the
    separable association does not matter): both load paths -- the TMA ring (tensor-map boxes with zero fill, row-by-row
    edge chunks, patched border columns) and plain global loads -- on widths that end inside a strip, one warp wide,
    several strips wide, heights that are not a multiple of the 8-row chunk.  Window 5 and 7 (the large_window preset,
    verification_config.yaml:99-103) are the same kernel with a 5- or 7-row state."""
    h, w = shape
    rng = np.random.default_rng(h * w)
    # exactness needs every window sum below 2^16 (multiples of 1/256 in a 24-bit significand): full-range white noise
    # stays below with 25 taps, with 49 taps a few pixels per frame cross it (and then differ in the last bit), so the
    # window-7 frames use half the range -- any natural frame has far less gradient energy than either
    top = 256 if window == 5 else 128
    p = rng.integers(0, top, (2, h, w)).astype(f32)
    c = rng.integers(0, top, (2, h, w)).astype(f32)
    want = [orc.lucas_kanade_single_scale(p[b], c[b], window) for b in range(2)]
    for path in (TMA_OR_ERROR, PLAIN_LOADS):
        u, v = np.full_like(p, np.nan), np.full_like(p, np.nan)
        assert emul_march.emul_lk_march(ptr(p), ptr(c), ptr(u), ptr(v), 2, h, w, path, window) == 0
        for b in range(2):
            assert np.array_equal(bits(u[b]), bits(want[b][0])) and np.array_equal(bits(v[b]), bits(want[b][1])), (path, b)


@pytest.mark.parametrize("window", [5, 7])
@pytest.mark.parametrize("shape", [(24, 128), (33, 256), (20, 144)])
def test_marching_kernel_uint8_ingest_and_fixed_point_flavours(emul_march, shape, window):
    h, w = shape
    rng = np.random.default_rng(h + w)
    top = 256 if window == 5 else 128  # see test_marching_kernel_source_on_cpu_uint8_valued_frames
    p8 = rng.integers(0, top, (2, h, w)).astype(np.uint8)
    c8 = rng.integers(0, top, (2, h, w)).astype(np.uint8)
    u, v = np.full((2, h, w), np.nan, f32), np.full((2, h, w), np.nan, f32)
    assert emul_march.emul_lk_march_u8(ptr(p8), ptr(c8), ptr(u), ptr(v), 2, h, w, window) == 0
    for b in range(2):
        uo, vo = orc.lucas_kanade_single_scale(p8[b].astype(f32), c8[b].astype(f32), window)
        assert np.array_equal(bits(u[b]), bits(uo)) and np.array_equal(bits(v[b]), bits(vo))
    for quirk in ((1, 0) if window == 5 else ()):  # the RTL's window is 5 x 5
        u16, v16 = np.full((2, h, w), 77, np.int16), np.full((2, h, w), 77, np.int16)
        assert emul_march.emul_lk_march_fx(ptr(p8), ptr(c8), ptr(u16), ptr(v16), 2, h, w, quirk) == 0
        for b in range(2):
            uo, vo = fxo.lk_single_scale_fx(p8[b], c8[b], mirror_avg_quirk=bool(quirk))
            assert np.array_equal(u16[b], uo) and np.array_equal(v16[b], vo), (quirk, b)


def test_uint8_flavour_with_the_sobel_stage_on_integer_fields(tmp_path_factory):
    """-DOF_U8_INT_SOBEL=1 (a measured, switched-off option of lk_march.cu): q = prev + curr stays in packed 16-bit
    fields and the Sobel sums are 32-bit additions on both fields at once.  Same bits as the reference, windows 5 and 7,
    strips that end inside the frame, a height that is not a multiple of the chunk."""
    lib = _build(tmp_path_factory, "emul_lk_march", defines=("-DOF_U8_INT_SOBEL=1",))
    lib.emul_lk_march_u8.argtypes = [_vp] * 4 + [_i] * 4
    for (h, w), window in (((33, 256), 5), ((20, 144), 7)):
        rng = np.random.default_rng(h * w + window)
        top = 256 if window == 5 else 128
        p8 = rng.integers(0, top, (2, h, w)).astype(np.uint8)
        c8 = rng.integers(0, top, (2, h, w)).astype(np.uint8)
        u, v = np.full((2, h, w), np.nan, f32), np.full((2, h, w), np.nan, f32)
        assert lib.emul_lk_march_u8(ptr(p8), ptr(c8), ptr(u), ptr(v), 2, h, w, window) == 0
        for b in range(2):
            uo, vo = orc.lucas_kanade_single_scale(p8[b].astype(f32), c8[b].astype(f32), window)
            assert np.array_equal(bits(u[b]), bits(uo)) and np.array_equal(bits(v[b]), bits(vo)), (h, w, window, b)


def _refine_once(lib, form, prev, curr, fu, fv, iteration=0, counter=None, state=None, rows=None, window=5,
                 warped=None, warped_next=None, warped_ready=0):
    b, h, w = prev.shape
    out_u, out_v = np.full_like(fu, 9.0), np.full_like(fv, 9.0)
    sel = np.zeros(b, np.int32) if state is None else state["sel"]
    done = np.zeros(b, np.int32) if state is None else state["done"]
    executed = np.zeros(b, np.int32) if state is None else state["executed"]
    resid = np.zeros((b, 4, 2), f32) if state is None else state["resid"]
    lo, hi = rows if rows else (0, h)
    units = lib.emul_lk_refine_units_per_pair(b, hi - lo, w)  # the kernels index partial[pair][unit][2]
    partial = np.zeros((b, units, 2))
    warped = np.zeros_like(prev) if warped is None else warped
    cnt = np.zeros(b, np.uint32)
    rc = lib.emul_lk_refine(form, ptr(prev), ptr(curr), ptr(fu), ptr(fv), ptr(out_u), ptr(out_v), ptr(sel), 0, ptr(done),
                            ptr(partial), ptr(warped), ptr(cnt), ptr(executed), ptr(resid), 4, iteration, b, h, w, lo, hi, lo, hi, window,
                            None if warped_next is None else ptr(warped_next), warped_ready)
    assert rc == 0
    return out_u, out_v, partial, dict(sel=sel, done=done, executed=executed, resid=resid)


@pytest.mark.parametrize("window", [5, 7])
def test_fast_refinement_iteration_source_on_cpu(emul_march, window):
    """One fast-mode iteration in its three forms.  Split and fused give the same bits; against the reference the
    iteration is tolerance-level (float32 warp fraction, separable window sums), so the comparison is statistical:
    flow_out - flow_in equals the oracle's increment except where the 2 x 2 system is ill-conditioned."""
    from scipy.ndimage import gaussian_filter, shift

    rng = np.random.default_rng(14)
    b, h, w = 2, 40, 248
    prev = gaussian_filter((rng.random((b, h, w)) * 255).astype(f32), (0, 1.5, 1.5)).astype(f32)
    curr = np.stack([shift(prev[0], (0.4, -0.7), order=1, mode="nearest"), shift(prev[1], (-0.6, 0.3), order=1, mode="nearest")]).astype(f32)
    fu = (rng.standard_normal((b, h, w)) * 0.2).astype(f32)
    fv = (rng.standard_normal((b, h, w)) * 0.2).astype(f32)
    split_u, split_v, split_part, _ = _refine_once(emul_march, 0, prev, curr, fu, fv, window=window)
    fused_u, fused_v, fused_part, _ = _refine_once(emul_march, 2, prev, curr, fu, fv, window=window)
    assert np.array_equal(bits(split_u), bits(fused_u)) and np.array_equal(bits(split_v), bits(fused_v))
    if window == 5:
        # the warp-specialised form (producer warps fill the marching warps' ring stages: no warped plane) -- same bits,
        # same per-unit sums as the split form
        ws_u, ws_v, ws_part, _ = _refine_once(emul_march, 3, prev, curr, fu, fv, window=window)
        assert np.array_equal(bits(split_u), bits(ws_u)) and np.array_equal(bits(split_v), bits(ws_v))
        assert np.array_equal(ws_part, split_part)
        wb_u, wb_v, _, _ = _refine_once(emul_march, 3, prev, curr, fu, fv, rows=(8, 30), window=window)
        assert np.array_equal(bits(wb_u[:, 8:30]), bits(split_u[:, 8:30])) and (wb_u[:, :8] == 9.0).all() and (wb_u[:, 30:] == 9.0).all()
    # the sums of |du|, |dv| are formed differently (four float32 magnitudes are added before widening in one form)
    assert split_part.sum(axis=1) == pytest.approx(fused_part.sum(axis=1), rel=1e-8)
    for k in range(b):
        du, dv = orc.lucas_kanade_single_scale(prev[k], orc.warp_image(curr[k], fu[k], fv[k]), window)
        err = np.maximum(np.abs(split_u[k] - (fu[k] + du)), np.abs(split_v[k] - (fv[k] + dv)))
        assert np.median(err) < 1e-5 and (err > 1e-3).mean() < 0.02
        assert split_part[k, :, 0].sum() == pytest.approx(np.abs(du).astype(np.float64).sum(), rel=1e-3)
    # rows [row_lo, row_hi) only (row-band mode): the same bits on those rows, nothing written outside
    band_u, band_v, _, _ = _refine_once(emul_march, 0, prev, curr, fu, fv, rows=(8, 30), window=window)
    assert np.array_equal(bits(band_u[:, 8:30]), bits(split_u[:, 8:30])) and np.array_equal(bits(band_v[:, 8:30]), bits(split_v[:, 8:30]))
    assert (band_u[:, :8] == 9.0).all() and (band_u[:, 30:] == 9.0).all()


@pytest.mark.parametrize("window", [5, 7])
def test_marching_kernel_warps_the_next_iteration_itself(emul_march, emul_warp, window):
    """The refinement flavour's epilogue: warped_next = warp(curr, flow_out) from the flow the kernel has in registers
    (asynchronous copies into a shared-memory landing zone, blended a step later).  It is warp_rows_kernel<float>'s
    plane bit for bit -- on every row of the frame, with flow that leaves the frame, on a row band -- and a second
    iteration that consumes it (no warp_rows launch) gives the classic two-launch iteration's bits."""
    from scipy.ndimage import gaussian_filter

    rng = np.random.default_rng(21)
    b, h, w = 2, 43, 248
    prev = gaussian_filter((rng.random((b, h, w)) * 255).astype(f32), (0, 1.5, 1.5)).astype(f32)
    curr = np.roll(prev, (1, -1), axis=(1, 2)) + rng.standard_normal(prev.shape).astype(f32)
    fu = (rng.standard_normal((b, h, w)) * 0.8).astype(f32)
    fv = (rng.standard_normal((b, h, w)) * 0.8).astype(f32)
    fu[0, :3, :5] = -7.5  # samples outside the frame
    fv[1, -2:, -6:] = 9.25
    # classic: iteration 1, then warp_rows + march for iteration 2
    u1, v1, _, _ = _refine_once(emul_march, 0, prev, curr, fu, fv, window=window)
    u2, v2, part2, _ = _refine_once(emul_march, 0, prev, curr, u1, v1, window=window)
    # chained: iteration 1 also writes the warped plane of iteration 2, which then launches no warp_rows
    wn = np.full_like(prev, np.nan)
    cu1, cv1, _, _ = _refine_once(emul_march, 0, prev, curr, fu, fv, window=window, warped_next=wn)
    assert np.array_equal(bits(cu1), bits(u1)) and np.array_equal(bits(cv1), bits(v1))
    want = np.stack([run_warp(emul_warp, curr[k], u1[k], v1[k], exact=False) for k in range(b)])
    assert np.array_equal(bits(wn), bits(want))
    cu2, cv2, cpart2, _ = _refine_once(emul_march, 0, prev, curr, cu1, cv1, window=window, warped=wn, warped_ready=1)
    assert np.array_equal(bits(cu2), bits(u2)) and np.array_equal(bits(cv2), bits(v2)) and np.array_equal(cpart2, part2)
    # a row band: the plane is written on the band's rows only
    wb = np.full_like(prev, 7.0)
    _refine_once(emul_march, 0, prev, curr, fu, fv, window=window, rows=(8, 30), warped_next=wb)
    assert np.array_equal(bits(wb[:, 8:30]), bits(want[:, 8:30])) and (wb[:, :8] == 7.0).all() and (wb[:, 30:] == 7.0).all()


def test_fast_refinement_fused_tail_source_on_cpu(emul_march):
    """form 1: the pair's last warp reduces the partial sums, applies the reference's early-exit test
    (mean|du| < 0.01 and mean|dv| < 0.01, lucas_kanade_pyramidal.py:213-223) and flips the ping-pong selector."""
    rng = np.random.default_rng(15)
    b, h, w = 2, 24, 128
    prev = (rng.integers(0, 256, (b, h, w))).astype(f32)
    curr = prev.copy()
    curr[0] = rng.integers(0, 256, (h, w))  # pair 0 moves, pair 1 does not: its increment is exactly 0
    zeros = np.zeros_like(prev)
    out_u, out_v, _, st = _refine_once(emul_march, 1, prev, curr, zeros, zeros.copy())
    assert st["sel"].tolist() == [1, 1] and st["done"].tolist() == [0, 1] and st["executed"].tolist() == [1, 1]
    ws_u, ws_v, _, ws_st = _refine_once(emul_march, 4, prev, curr, zeros, zeros.copy())  # warp-specialised form, same tail
    assert np.array_equal(bits(ws_u), bits(out_u)) and np.array_equal(bits(ws_v), bits(out_v))
    assert ws_st["sel"].tolist() == [1, 1] and ws_st["done"].tolist() == [0, 1] and np.array_equal(ws_st["resid"], st["resid"])
    du, dv = orc.lucas_kanade_single_scale(prev[0], curr[0], 5)  # flow_in = 0: the warp is the identity
    assert np.array_equal(bits(out_u[0]), bits(du)) and np.array_equal(bits(out_v[0]), bits(dv))
    assert st["resid"][0, 0, 0] == pytest.approx(np.abs(du).mean(), rel=1e-5)
    assert (out_u[1] == 0).all() and st["resid"][1, 0].tolist() == [0.0, 0.0]
