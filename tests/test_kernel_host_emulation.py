"""The default exact-mode kernel for window 5, lk_tile5_kernel, run ON THE CPU from its own source file.

tests/host_emul/emul_lk_tile5.cpp #includes optical-flow-fpga_b200/csrc/lk_tile5.cu and compiles it with g++ on
top of tests/host_emul/cuda_on_host.h (every CUDA thread of a block is an OS thread, __syncthreads() a barrier,
__shared__ arrays statics, the *_rn intrinsics single IEEE operations).  What the GPU tests establish on the device
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
