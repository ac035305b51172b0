"""The whole product on the CPU: Python wrapper -> C ABI -> drivers -> kernels, with tests/host_emul/'s emulated
library (every product .cu compiled by g++ on cuda_on_host.h, fake_cudart.cpp for the runtime calls) standing in for
libof_b200.so.  What tests/test_kernel_host_emulation.py does kernel by kernel, this does for the code AROUND the kernels:
the host-buffer pipeline, the pyramidal level loop (workspace carving, ping-pong buffers, per-level kernel selection,
early exit, final select), argument checks.  The checks run in a child process because the wrapper loads its library
once; the library is selected through OF_B200_LIB_NAME, the wrapper's hook for experimental builds, and lives under
tests/ -- the product itself still has no CPU path (tests/test_abi.py).

The same switch runs most of the GPU suite here, slowly:
    python tests/host_emul/build_emulated_library.py
    OF_B200_LIB_NAME=../tests/host_emul/_build/libof_b200_emulated.so python -m pytest tests/test_gpu_parity.py -m gpu \\
        -k "not 1080p and not 4k and not 8k and not rowband and not lanes and not times_out and not variants"
"""

import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "host_emul"))

CHILD = r"""
import os
import sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {backend!r})
import of_b200
import lucas_kanade_core, lucas_kanade_pyramidal
from oracle import lk_float_oracle as orc
from oracle import lk_fixed_oracle as fxo
from scipy.ndimage import gaussian_filter, shift

assert of_b200.LIB_PATH.name == "libof_b200_emulated.so" and of_b200.device_count() == 1
bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
same = lambda a, b: a.shape == b.shape and np.array_equal(bits(a), bits(b))
rng = np.random.default_rng(0)

# single scale through the host-buffer entry point: fast (marching kernel) and exact (lk_tile5_kernel), a batch, the
# drop-in module, a window the second tile kernel does not take, a width the marching kernel does not take
p = rng.integers(0, 256, (3, 40, 128)).astype(np.float32)
c = rng.integers(0, 256, (3, 40, 128)).astype(np.float32)
want = [orc.lucas_kanade_single_scale(p[b], c[b], 5) for b in range(3)]
for mode in (of_b200.MODE_FAST, of_b200.MODE_EXACT):
    u, v = of_b200.lk_single_scale_batch(p, c, 5, mode)
    assert all(same(u[b], want[b][0]) and same(v[b], want[b][1]) for b in range(3)), mode
u, v = lucas_kanade_core.lucas_kanade_single_scale(p[0], c[0], window_size=5)
assert same(u, want[0][0]) and same(v, want[0][1])
g = (rng.standard_normal((2, 37, 61)) * 40).astype(np.float32)
for w in (3, 7):
    u, v = of_b200.lk_single_scale(g[0], g[1], w, mode=of_b200.MODE_EXACT)
    uo, vo = orc.lucas_kanade_single_scale(g[0], g[1], w)
    assert same(u, uo) and same(v, vo), w
u, v = of_b200.lk_single_scale(g[0], g[1], 5, mode=of_b200.MODE_FAST)  # width 61: falls back to the reference-order kernel
uo, vo = orc.lucas_kanade_single_scale(g[0], g[1], 5)
assert same(u, uo) and same(v, vo)

# device entry point with planes that are not 16-byte aligned (base pointer offset by one float): the marching kernel
# moves 128-bit words, so the driver must route them to the reference-order tile kernel
h_, w_ = 40, 128
buf_p, buf_c = np.zeros(h_ * w_ + 1, np.float32), np.zeros(h_ * w_ + 1, np.float32)
buf_p[1:], buf_c[1:] = p[0].ravel(), c[0].ravel()
out_u, out_v = np.full(h_ * w_ + 1, np.nan, np.float32), np.full(h_ * w_ + 1, np.nan, np.float32)
of_b200.lk_single_scale_dev(buf_p.ctypes.data + 4, buf_c.ctypes.data + 4, out_u.ctypes.data + 4, out_v.ctypes.data + 4, 1, h_, w_,
                            5, of_b200.MODE_FAST)
assert same(out_u[1:].reshape(h_, w_), want[0][0]) and same(out_v[1:].reshape(h_, w_), want[0][1])

# uint8 ingest and the fixed-point mode
p8, c8 = p.astype(np.uint8), c.astype(np.uint8)
u, v = of_b200.lk_single_scale_u8_batch(p8, c8, 5, of_b200.MODE_FAST)
assert all(same(u[b], want[b][0]) and same(v[b], want[b][1]) for b in range(3))
u16, v16 = of_b200.lk_single_scale_fx(p8[0], c8[0])
uo, vo = fxo.lk_single_scale_fx(p8[0], c8[0])
assert np.array_equal(u16, uo) and np.array_equal(v16, vo)
# the fixed-point kernels' sources against vectors produced by executing the reference's RTL text (oracle/sv_eval.py):
# marching kernel (width % 16 == 0) and tile kernel (one column more)
zz = np.load(os.path.join(os.path.dirname(os.path.abspath(fxo.__file__)), "..", "tests", "golden", "fx_rtl_text_vectors.npz"))
for extra in (0, 1):
    fpv = np.ascontiguousarray(np.pad(zz["frame_prev"], ((0, 0), (0, extra)), constant_values=128))
    fcv = np.ascontiguousarray(np.pad(zz["frame_curr"], ((0, 0), (0, extra)), constant_values=128))
    u16, v16 = of_b200.lk_single_scale_fx(fpv, fcv)
    assert np.array_equal(u16[zz["centres"][:, 0], zz["centres"][:, 1]], zz["u"]), "RTL-text vectors, u"
    assert np.array_equal(v16[zz["centres"][:, 0], zz["centres"][:, 1]], zz["v"]), "RTL-text vectors, v"

# the pyramidal driver: 3 levels x 3 iterations, exact mode bit for bit (levels 64x96, 32x48, 16x24: marching pyramid
# kernel / tile pyramid kernel, split refinement), a batch whose second pair converges early, the drop-in module
pp = gaussian_filter((rng.random((64, 96)) * 255).astype(np.float32), 1.5).astype(np.float32)
cc = shift(pp, (0.6, -0.8), order=1, mode="nearest").astype(np.float32)
uo, vo = orc.lucas_kanade_pyramidal(pp, cc, 3, 5, 3)
u, v = of_b200.lk_pyramidal(pp, cc, 3, 5, 3, mode=of_b200.MODE_EXACT)
assert same(u, uo) and same(v, vo)
ub, vb, (iters, resid) = of_b200.lk_pyramidal_batch(np.stack([pp, pp]), np.stack([cc, pp]), 3, 5, 3, mode=of_b200.MODE_EXACT,
                                                    return_trace=True)
assert same(ub[0], uo) and same(vb[0], vo) and iters[0].tolist() == [3, 3, 3] and iters[1].tolist() == [1, 1, 1]
assert not ub[1].any() and not vb[1].any()
u, v = of_b200.lk_pyramidal(pp, cc, 3, 5, 3, mode=of_b200.MODE_FAST)  # fast mode: tolerance-level
d = np.maximum(np.abs(u - uo), np.abs(v - vo))
assert np.median(d) < 1e-5 and (d > 1e-3).mean() < 0.01
odd_p = gaussian_filter((rng.random((45, 67)) * 255).astype(np.float32), 1.0).astype(np.float32)
odd_c = shift(odd_p, (0.7, -1.3), order=1, mode="nearest").astype(np.float32)
for mode in (of_b200.MODE_EXACT, of_b200.MODE_FAST):  # no level is TMA-able: every level on the tile kernels, both modes
    u, v = of_b200.lk_pyramidal(odd_p, odd_c, 2, 5, 2, mode=mode)
    uo2, vo2 = orc.lucas_kanade_pyramidal(odd_p, odd_c, 2, 5, 2)
    assert same(u, uo2) and same(v, vo2), mode

# helpers of the drop-in module
lv = lucas_kanade_pyramidal.build_gaussian_pyramid(pp, 3)
assert all(same(a, b) for a, b in zip(lv, orc.build_gaussian_pyramid(pp, 3)))
fu = (rng.standard_normal(pp.shape) * 2).astype(np.float32)
fv = (rng.standard_normal(pp.shape) * 2).astype(np.float32)
assert same(lucas_kanade_pyramidal.warp_image(pp, fu, fv), orc.warp_image(pp, fu, fv))
gu, gv = lucas_kanade_pyramidal.upsample_flow(fu[:32, :48], fv[:32, :48], (64, 96))
wu, wv = orc.upsample_flow(fu[:32, :48], fv[:32, :48], (64, 96))
assert same(gu, wu) and same(gv, wv)

# the refinement iteration as a building block (of_lk_refine_f32_dev, what the Python row-band driver calls):
# rows [row_lo, row_hi) only, sums of |du|, |dv| over the owned rows
fu32 = (rng.standard_normal(pp.shape) * 0.5).astype(np.float32)
fv32 = (rng.standard_normal(pp.shape) * 0.5).astype(np.float32)
du, dv = orc.lucas_kanade_single_scale(pp, orc.warp_image(cc, fu32, fv32), 5)
H, W = pp.shape
ws_bytes = of_b200.lk_refine_workspace_bytes(1, H, W)
ws = np.zeros(ws_bytes, np.uint8)
for mode in (of_b200.MODE_EXACT, of_b200.MODE_FAST):
    ou, ov = np.full_like(pp, 5.0), np.full_like(pp, 5.0)
    sums = np.zeros(2)
    lo, hi, own_lo, own_hi = 10, 44, 14, 40
    of_b200.lk_refine_dev(pp.ctypes.data, cc.ctypes.data, fu32.ctypes.data, fv32.ctypes.data, ou.ctypes.data, ov.ctypes.data,
                          1, H, W, 5, mode, lo, hi, own_lo, own_hi, sums.ctypes.data, ws.ctypes.data, ws_bytes)
    assert (ou[:lo] == 5.0).all() and (ou[hi:] == 5.0).all()
    if mode == of_b200.MODE_EXACT:
        assert same(ou[lo:hi], (fu32 + du)[lo:hi]) and same(ov[lo:hi], (fv32 + dv)[lo:hi])
        assert abs(sums[0] - np.abs(du[own_lo:own_hi]).astype(np.float64).sum()) <= 1e-9 * sums[0]
    else:
        err = np.maximum(np.abs(ou[lo:hi] - (fu32 + du)[lo:hi]), np.abs(ov[lo:hi] - (fv32 + dv)[lo:hi]))
        assert np.median(err) < 1e-5

# errors come back as exceptions, not crashes
for bad in (lambda: of_b200.lk_single_scale(p[0], c[0], 4), lambda: of_b200.lk_pyramidal(pp[:4, :4], cc[:4, :4], 5, 5, 3)):
    try:
        bad()
        raise SystemExit("expected ValueError")
    except ValueError:
        pass
print("emulated library ok:", of_b200.kernel_launches(), "kernel launches")
"""


def test_product_runs_end_to_end_on_the_emulated_library():
    import build_emulated_library

    try:
        lib = build_emulated_library.build()
    except RuntimeError as e:
        if "needs g++" in str(e):
            pytest.skip(str(e))
        raise
    backend = ROOT / "optical-flow-fpga_b200"
    env = dict(os.environ, OF_B200_LIB_NAME=os.path.relpath(lib, backend))
    code = CHILD.format(root=str(ROOT), backend=str(backend))
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    assert "emulated library ok" in res.stdout


ROWBAND_CHILD = r"""
import ctypes as C, multiprocessing as mp, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {backend!r})
import of_b200, synthetic

bits = lambda a: np.ascontiguousarray(a).view(np.uint32)


def rank_main(r, world, shape, levels, iters, repl_px, mode, p, c, conn):
    assert of_b200.lib().cuda_on_host_use_slice(r, world) == 0  # this rank's part of the shared "device memory"
    cx = of_b200.RowbandContext(r, world, shape[0], shape[1], levels, 5, iters, mode)
    conn.send(cx.arena_ptr)
    cx.set_peers(conn.recv())
    cx.set_replicate_pixels(repl_px)
    cx.set_timeout_ms(120000)  # OS threads on a few cores: the ranks drift far more than GPUs do
    out = []
    for rep in range(2):  # the second run reuses the arenas and the flag sequence numbers
        u, v = np.full_like(p, np.nan), np.full_like(p, np.nan)
        cx.run(p.ctypes.data, c.ctypes.data, u.ctypes.data, v.ctypes.data, 0)
        it_r, _, err = cx.trace(0)
        out.append((u, v, it_r.tolist(), err))
    conn.send(out)
    conn.recv()
    cx.close()


def run(world, shape, levels, iters, repl_px, mode):
    prev, curr, _ = synthetic.make_pairs_numpy(1, shape[0], shape[1], seed=31)
    p, c = np.ascontiguousarray(prev[0]), np.ascontiguousarray(np.roll(curr[0], 3, axis=0))
    u1, v1, (iters_exec, _) = of_b200.lk_pyramidal(p, c, levels, 5, iters, mode=mode, return_trace=True)
    ctx = mp.get_context("fork")
    pipes = [ctx.Pipe() for _ in range(world)]
    procs = [ctx.Process(target=rank_main, args=(r, world, shape, levels, iters, repl_px, mode, p, c, pipes[r][1])) for r in range(world)]
    for x in procs:
        x.start()
    arenas = [pipes[r][0].recv() for r in range(world)]
    for r in range(world):
        pipes[r][0].send(arenas)
    results = [pipes[r][0].recv() for r in range(world)]
    for r in range(world):
        pipes[r][0].send("done")
    for x in procs:
        x.join()
    for r, res in enumerate(results):
        for rep, (u, v, it_r, err) in enumerate(res):
            what = (world, shape, levels, iters, repl_px, mode, "rank", r, "run", rep)
            assert err == 0, what
            assert it_r == np.asarray(iters_exec).reshape(-1).tolist(), what
            assert np.array_equal(bits(u), bits(u1)) and np.array_equal(bits(v), bits(v1)), what


lib = of_b200.lib()
lib.cuda_on_host_shared_init.argtypes = [C.c_size_t]
assert lib.cuda_on_host_shared_init(1 << 30) == 0  # before the ranks fork: they inherit the mapping at one address
for mode in (of_b200.MODE_EXACT, of_b200.MODE_FAST):
    run(2, (96, 128), 1, 1, 0, mode)            # one level, split in two
    run(3, (96, 128), 2, 2, 0, mode)            # every level in row bands
    run(4, (120, 248), 3, 2, 2000, mode)        # the coarsest level whole on every rank
print("row bands ok")
"""


def test_row_band_driver_with_ranks_as_processes():
    """of_rowband_run on 2, 3 and 4 ranks: each rank is a forked process, "peer-mapped device memory" is one shared
    mapping the fake runtime allocates the arenas from, so the ranks' kernels meet through the flag words and push rows
    into each other's arenas as they do over NVLink -- peer.cu's collectives, the fused tail's peer all-reduce, the
    band / halo bookkeeping of of_rowband.inl.  Every rank's gathered flow must equal the whole-frame driver's, bit for
    bit, with the same early-exit decisions, on two consecutive runs."""
    import build_emulated_library

    try:
        lib = build_emulated_library.build()
    except RuntimeError as e:
        if "needs g++" in str(e):
            pytest.skip(str(e))
        raise
    backend = ROOT / "optical-flow-fpga_b200"
    env = dict(os.environ, OF_B200_LIB_NAME=os.path.relpath(lib, backend))
    code = ROWBAND_CHILD.format(root=str(ROOT), backend=str(backend))
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=1500)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    assert "row bands ok" in res.stdout
