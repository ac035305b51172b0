"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference goldens.

All comparisons of the float path in EXACT mode are bit-exact.  FAST mode is bit-exact on
uint8-valued frames (the verifier's inputs and the benchmark's synthetic frames) and is
held to the north star's tolerance (max |du|, |dv| <= 1e-3 px, MAE/EPE equal to 3 decimals)
where floats are not exactly summable.  The fixed-point mode is bit-exact against the
integer oracle.
"""

import hashlib
import json
import sys

import numpy as np
import pytest

from conftest import BACKEND_DIR, ROOT

sys.path.insert(0, str(BACKEND_DIR))

from oracle import flow_metrics_oracle as fm  # noqa: E402
from oracle import lk_fixed_oracle as fxo  # noqa: E402
from oracle import lk_float_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ofb():
    import build as of_build

    of_build.build()
    import of_b200

    assert of_b200.device_count() > 0, "GPU tests need a CUDA device (no CPU fallback exists)"
    return of_b200


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, what
    if a.dtype == np.float32:
        neq = bits(a) != bits(b)
    else:
        neq = a != b
    assert not neq.any(), f"{what}: {int(neq.sum())} of {a.size} values differ, first at {np.argwhere(neq)[:3].tolist()}"


# ---------------------------------------------------------------------------------------
# drop-in modules against the reference's golden outputs
# ---------------------------------------------------------------------------------------
def test_single_scale_13_patterns_bit_exact(ofb, golden_index, golden_frames):
    import lucas_kanade_core as core

    for name, entry in golden_index["patterns"].items():
        p, c = (f.astype(np.float32) for f in golden_frames[name])
        for mode in (ofb.MODE_EXACT, ofb.MODE_FAST):
            u, v = ofb.lk_single_scale(p, c, 5, mode=mode)
            assert sha(u) == entry["single_scale"]["sha256_u"], (name, mode)
            assert sha(v) == entry["single_scale"]["sha256_v"], (name, mode)
        u, v = core.lucas_kanade_single_scale(p, c)  # the module the verifier imports
        assert sha(u) == entry["single_scale"]["sha256_u"]
        assert u.dtype == np.float32 and u.flags["C_CONTIGUOUS"]


def test_pyramidal_13_patterns_bit_exact_and_metrics(ofb, golden_index, golden_frames, capsys):
    import lucas_kanade_pyramidal as pyr

    for name, entry in golden_index["patterns"].items():
        p, c = (f.astype(np.float32) for f in golden_frames[name])
        u, v, (iters, resid) = ofb.lk_pyramidal(p, c, 3, 5, 3, mode=ofb.MODE_EXACT, return_trace=True)
        assert sha(u) == entry["pyramidal"]["sha256_u"], name
        assert sha(v) == entry["pyramidal"]["sha256_v"], name
        ref_iters = sum(1 for ln in entry["pyramidal"]["reference_log"] if ln.startswith("Iteration"))
        assert int(iters.sum()) == ref_iters, name
        mask = fm.test_region_mask(p.shape, name, golden_index["center_crop"])
        gt = entry["ground_truth"]
        m = fm.all_metrics(u, v, gt["u"], gt["v"], mask)
        for k, val in entry["verification_baseline"]["pyramidal"].items():
            assert m[k] == val, (name, k)
    u2, v2 = pyr.lucas_kanade_pyramidal(p, c, num_levels=3, window_size=5, num_iterations=3)
    capsys.readouterr()
    assert_bit_equal(u2, u, "drop-in module")
    assert_bit_equal(v2, v, "drop-in module")


def test_pyramidal_fast_mode_13_patterns_within_tolerance(ofb, golden_index, golden_frames):
    """FAST pyramidal (register-marching refinement kernel): the warp blends in float64 with float32
    sample fractions (one rounding away from the reference's for negative flow), the window sums
    are separable float32.  Contract: MAE / EPE equal to the reference's to 3 decimals on every
    verifier pattern and the same early-exit decisions; the per-pixel deviations are confined to ill-conditioned
    pixels and are held to what was MEASURED per pattern (tests/golden/fast_mode_deviation.json,
    written from bench.py's verifier_patterns report): no more pixels beyond 1e-3 px than recorded (+ 2 % slack
    for a different summation order in a future kernel), no larger maximum than twice the recorded one.  Exact
    mode is the one that meets the north star's per-pixel bound (it is bit-identical)."""
    from conftest import GOLDEN

    recorded = json.load(open(GOLDEN / "fast_mode_deviation.json"))["patterns"]
    report = {}
    for name, entry in golden_index["patterns"].items():
        p, c = (f.astype(np.float32) for f in golden_frames[name])
        ue, ve = ofb.lk_pyramidal(p, c, 3, 5, 3, mode=ofb.MODE_EXACT)
        uf, vf, (iters, _) = ofb.lk_pyramidal(p, c, 3, 5, 3, mode=ofb.MODE_FAST, return_trace=True)
        ref_iters = sum(1 for ln in entry["pyramidal"]["reference_log"] if ln.startswith("Iteration"))
        assert int(iters.sum()) == ref_iters, name  # same early-exit decisions
        mask = fm.test_region_mask(p.shape, name, golden_index["center_crop"])
        gt = entry["ground_truth"]
        m = fm.all_metrics(uf, vf, gt["u"], gt["v"], mask)
        for k in ("mae_u", "mae_v", "epe"):
            assert abs(m[k] - entry["verification_baseline"]["pyramidal"][k]) < 5e-4, (name, k, m[k])
        d = np.maximum(np.abs(uf - ue), np.abs(vf - ve))
        report[name] = (float(d.max()), int((d > 1e-3).sum()))
        rec = recorded[name]
        assert report[name][1] <= rec["pixels_gt_1e-3"] * 1.02 + 8, (name, report[name], rec)
        assert report[name][0] <= 2.0 * rec["max_abs_diff_px"] + 1e-6, (name, report[name], rec)
    print("fast-vs-exact pyramidal: max |d| px, pixels > 1e-3:", report)
    assert report["no_motion"] == (0.0, 0)


def test_pyramidal_fast_mode_mixed_levels(ofb):
    """4 levels where the coarsest widths are not multiples of 4: the driver mixes the marching
    kernel (fine levels) with the tile kernel (coarse levels) in one run."""
    rng = np.random.default_rng(3)
    from scipy.ndimage import gaussian_filter, shift

    p = gaussian_filter((rng.random((200, 328)) * 255).astype(np.float32), 1.2)  # 328, 164, 82, 41
    c = shift(p, (1.4, -2.2), order=1, mode="nearest").astype(np.float32)
    ue, ve = ofb.lk_pyramidal(p, c, 4, 5, 2, mode=ofb.MODE_EXACT)
    uo, vo = orc.lucas_kanade_pyramidal(p, c, 4, 5, 2)
    assert_bit_equal(ue, uo, "exact u")
    uf, vf = ofb.lk_pyramidal(p, c, 4, 5, 2, mode=ofb.MODE_FAST)
    d = np.maximum(np.abs(uf - ue), np.abs(vf - ve))
    assert np.median(d) < 1e-5 and (d > 1e-3).mean() < 0.05
    uf2, vf2 = ofb.lk_pyramidal(p, c, 4, 5, 2, mode=ofb.MODE_FAST)
    assert_bit_equal(uf, uf2, "determinism")


def test_no_motion_converges_after_one_iteration_per_level(ofb, golden_frames):
    p, c = (f.astype(np.float32) for f in golden_frames["no_motion"])
    u, v, (iters, resid) = ofb.lk_pyramidal(p, c, 3, 5, 3, return_trace=True)
    assert iters.tolist() == [1, 1, 1]
    assert not u.any() and not v.any()


def test_helpers_against_reference_units(ofb, golden_units):
    import lucas_kanade_core as core
    import lucas_kanade_pyramidal as pyr

    g = golden_units
    ix, iy, it = core.compute_gradients(g["grad_prev"], g["grad_curr"])
    assert_bit_equal(ix, g["grad_ix"], "Ix")
    assert_bit_equal(iy, g["grad_iy"], "Iy")
    assert_bit_equal(it, g["grad_it"], "It")

    u, v = core.lucas_kanade_from_gradients(g["fg_ix"], g["fg_iy"], g["fg_it"], 5)
    assert_bit_equal(u, g["fg_u"], "from_gradients u")
    assert_bit_equal(v, g["fg_v"], "from_gradients v")

    levels = pyr.build_gaussian_pyramid(g["grad_prev"], 4)
    for i, lvl in enumerate(levels):
        assert_bit_equal(lvl, g[f"pyr4_level{i}"], f"pyramid level {i}")
    levels = pyr.build_gaussian_pyramid(g["pyr_odd_in"], 3)
    for i, lvl in enumerate(levels):
        assert_bit_equal(lvl, g[f"pyr_odd_level{i}"], f"odd pyramid level {i}")

    out = pyr.warp_image(g["warp_img"], g["warp_u"], g["warp_v"])
    assert_bit_equal(out, g["warp_out"], "warp")
    for shape in ((60, 80), (61, 83)):
        uu, vv = pyr.upsample_flow(g["up_u"], g["up_v"], shape)
        assert_bit_equal(uu, g[f"up_out_u_{shape[0]}x{shape[1]}"], "upsample u")
        assert_bit_equal(vv, g[f"up_out_v_{shape[0]}x{shape[1]}"], "upsample v")


@pytest.mark.parametrize("w", [3, 5, 7])
def test_general_float_frames_exact_mode(ofb, golden_units, w):
    g = golden_units
    u, v = ofb.lk_single_scale(g[f"float_w{w}_prev"], g[f"float_w{w}_curr"], w, mode=ofb.MODE_EXACT)
    assert_bit_equal(u, g[f"float_w{w}_u"], f"w={w} u")
    assert_bit_equal(v, g[f"float_w{w}_v"], f"w={w} v")


def test_small_pyramidal_window7(ofb, golden_units):
    g = golden_units
    u, v = ofb.lk_pyramidal(g["pyr_small_prev"], g["pyr_small_curr"], 2, 7, 2)
    assert_bit_equal(u, g["pyr_small_u"], "u")
    assert_bit_equal(v, g["pyr_small_v"], "v")


# ---------------------------------------------------------------------------------------
# CUDA path against the oracle on seeded inputs (shapes the goldens do not cover)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("w", [5, 7])
@pytest.mark.parametrize("shape", [(7, 8), (33, 8), (52, 124), (100, 248), (64, 120), (61, 364), (240, 320)])
def test_fast_kernel_ragged_shapes_uint8_frames(ofb, shape, w):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    # smooth-ish uint8 texture so that the window sums stay exactly representable
    base = rng.integers(0, 256, size=(shape[0] + 8, shape[1] + 8)).astype(np.float32)
    k = np.ones((5, 5), np.float32) / 25
    from scipy.signal import convolve2d

    sm = np.rint(convolve2d(base, k, mode="same")).astype(np.float32)
    p = sm[4:-4, 4:-4].copy()
    c = sm[3:-5, 5:-3].copy()
    uo, vo = orc.lucas_kanade_single_scale(p, c, w)
    for mode in (ofb.MODE_FAST, ofb.MODE_EXACT):
        u, v = ofb.lk_single_scale(p, c, w, mode=mode)
        assert_bit_equal(u, uo, f"{shape} w {w} mode {mode} u")
        assert_bit_equal(v, vo, f"{shape} w {w} mode {mode} v")


def test_large_window_preset_fast_mode(ofb, golden_index, golden_frames):
    """verification_config.yaml:99-103 (`large_window`: 3 levels, window 7, 3 iterations) on the marching kernels
    (lk_march_kernel<..., WIN = 7>).  Single scale, window 7, on the verifier's uint8 patterns: fast mode gives the
    reference's bits (float and uint8 ingest).  Pyramidal: exact mode equals the oracle bit for bit; fast mode takes
    the same early-exit decisions and stays at tolerance level against it."""
    names = ["translate_medium", "rotate_small", "zoom_in", "translate_extreme", "no_motion"]
    for name in names:
        p8, c8 = golden_frames[name]
        p, c = p8.astype(np.float32), c8.astype(np.float32)
        uo, vo = orc.lucas_kanade_single_scale(p, c, 7)
        u, v = ofb.lk_single_scale(p, c, 7, mode=ofb.MODE_FAST)
        assert_bit_equal(u, uo, f"{name} single-scale w7 fast u")
        assert_bit_equal(v, vo, f"{name} single-scale w7 fast v")
        u8_, v8_ = ofb.lk_single_scale_u8_batch(p8[None], c8[None], 7, ofb.MODE_FAST)
        assert_bit_equal(u8_[0], uo, f"{name} uint8 ingest w7 u")
        assert_bit_equal(v8_[0], vo, f"{name} uint8 ingest w7 v")
        ue, ve, (it_e, _) = ofb.lk_pyramidal(p, c, 3, 7, 3, mode=ofb.MODE_EXACT, return_trace=True)
        if name in ("translate_medium", "rotate_small"):
            po, qo = orc.lucas_kanade_pyramidal(p, c, 3, 7, 3)
            assert_bit_equal(ue, po, f"{name} pyramidal w7 exact u")
            assert_bit_equal(ve, qo, f"{name} pyramidal w7 exact v")
        uf, vf, (it_f, _) = ofb.lk_pyramidal(p, c, 3, 7, 3, mode=ofb.MODE_FAST, return_trace=True)
        assert it_f.tolist() == it_e.tolist(), name
        d = np.maximum(np.abs(uf - ue), np.abs(vf - ve))
        # measured on the 13 patterns (round 2): at most 0.5 % of the pixels beyond 1e-3 px, largest 0.037 px -- the
        # 7 x 7 system is better conditioned than the 5 x 5 one (tests/golden/fast_mode_deviation.json)
        assert np.median(d) < 1e-4 and (d > 1e-3).mean() < 0.01 and d.max() < 0.1, (name, float(d.max()), float((d > 1e-3).mean()))
    assert float(d.max()) == 0.0  # no_motion


@pytest.mark.parametrize("w", [1, 3, 5, 7, 9, 11])
def test_exact_mode_all_windows_random_floats(ofb, w):
    rng = np.random.default_rng(w)
    p = (rng.random((41, 67)) * 255).astype(np.float32)
    c = (p + rng.standard_normal((41, 67)).astype(np.float32) * 3).astype(np.float32)
    uo, vo = orc.lucas_kanade_single_scale(p, c, w)
    u, v = ofb.lk_single_scale(p, c, w, mode=ofb.MODE_EXACT)
    assert_bit_equal(u, uo, f"w={w} u")
    assert_bit_equal(v, vo, f"w={w} v")


def test_fast_mode_on_general_floats_is_within_tolerance(ofb):
    """Not exactly summable inputs: FAST differs from the reference order only by rounding."""
    rng = np.random.default_rng(5)
    from scipy.ndimage import gaussian_filter

    p = gaussian_filter((rng.random((120, 240)) * 255).astype(np.float32), 1.5)
    c = np.roll(p, 1, axis=1) + rng.standard_normal(p.shape).astype(np.float32) * 0.01
    c = c.astype(np.float32)
    uo, vo = orc.lucas_kanade_single_scale(p, c, 5)
    u, v = ofb.lk_single_scale(p, c, 5, mode=ofb.MODE_FAST)
    # well-conditioned pixels only: |det| >> eps, where 1e-3 px is meaningful
    ix, iy, it = orc.compute_gradients(p, c)
    sxx, syy, sxy, _, _ = orc.window_sums(ix, iy, it, 5)
    det = np.zeros_like(p)
    det[2:-2, 2:-2] = sxx * syy - sxy * sxy
    good = det > 1.0
    assert good.mean() > 0.5
    assert np.max(np.abs(u - uo)[good]) <= 1e-3
    assert np.max(np.abs(v - vo)[good]) <= 1e-3


def test_batch_equals_single_pairs_and_is_deterministic(ofb):
    import synthetic

    prev, curr, _ = synthetic.make_pairs_numpy(5, 96, 248, seed=3)
    for mode in (ofb.MODE_FAST, ofb.MODE_EXACT):
        ub, vb = ofb.lk_single_scale_batch(prev, curr, 5, mode=mode)
        ub2, vb2 = ofb.lk_single_scale_batch(prev, curr, 5, mode=mode)
        assert_bit_equal(ub, ub2, "determinism")
        for b in range(5):
            u, v = ofb.lk_single_scale(prev[b], curr[b], 5, mode=mode)
            assert_bit_equal(ub[b], u, f"pair {b}")
            assert_bit_equal(vb[b], v, f"pair {b}")
    uo, vo = orc.lucas_kanade_single_scale(prev[4], curr[4], 5)
    assert_bit_equal(ub[4], uo, "oracle")
    assert_bit_equal(vb[4], vo, "oracle")


def test_pyramidal_batch_with_mixed_convergence(ofb, golden_frames):
    """Pairs of one batch stop iterating independently (no_motion exits early, the others do not)."""
    names = ["translate_small", "no_motion", "rotate_small"]
    prev = np.stack([golden_frames[n][0] for n in names]).astype(np.float32)
    curr = np.stack([golden_frames[n][1] for n in names]).astype(np.float32)
    u, v, (iters, resid) = ofb.lk_pyramidal_batch(prev, curr, 3, 5, 3, return_trace=True)
    assert iters[1].tolist() == [1, 1, 1] and iters[0].tolist() == [3, 3, 3]
    for b, n in enumerate(names):
        uo, vo = orc.lucas_kanade_pyramidal(prev[b], curr[b], 3, 5, 3)
        assert_bit_equal(u[b], uo, n)
        assert_bit_equal(v[b], vo, n)


@pytest.mark.parametrize("levels,iters,shape", [(1, 2, (40, 56)), (2, 1, (45, 67)), (4, 2, (128, 160)), (3, 0, (64, 64))])
def test_pyramidal_other_presets_against_oracle(ofb, levels, iters, shape):
    rng = np.random.default_rng(levels * 10 + iters)
    from scipy.ndimage import gaussian_filter, shift

    p = gaussian_filter((rng.random(shape) * 255).astype(np.float32), 1.0)
    c = shift(p, (0.7, -1.3), order=1, mode="nearest").astype(np.float32)
    uo, vo = orc.lucas_kanade_pyramidal(p, c, levels, 5, iters)
    u, v = ofb.lk_pyramidal(p, c, levels, 5, iters)
    assert_bit_equal(u, uo, "u")
    assert_bit_equal(v, vo, "v")


def test_exact_mode_kernel_variants_give_the_same_bits(ofb):
    """Exact mode has three switches, each read once per process, so every combination runs in a child:
    OF_B200_EXACT_REFINE = split (warp_rows_kernel<double> + tile kernel on (prev, warped)) | fused (the tile
    kernel gathers its own halo), OF_B200_TILE = v1 (lk_tile_kernel) | v2 (lk_tile5_kernel, window 5),
    OF_B200_EXACT = march (lk_exact_march_kernel, the default: packed pairs, shared-memory rings) | tile.
    Single-scale on float frames (ragged shapes, one narrower than a tile) and a 3-level pyramidal run whose
    flow has both signs and leaves the frame (warp's outside -> 0 rule) must hash identically."""
    import os
    import subprocess

    code = (
        "import sys, hashlib, numpy as np\n"
        f"sys.path.insert(0, {str(BACKEND_DIR)!r})\n"
        "import of_b200\n"
        "from scipy.ndimage import gaussian_filter, shift\n"
        "rng = np.random.default_rng(77)\n"
        "h = hashlib.sha256()\n"
        "for shape in ((150, 203), (37, 61), (16, 64), (5, 5), (129, 260)):\n"
        "    a = (rng.standard_normal(shape) * 50).astype(np.float32)\n"
        "    b = (a + rng.standard_normal(shape) * 3).astype(np.float32)\n"
        "    u, v = of_b200.lk_single_scale(a, b, 5, mode=of_b200.MODE_EXACT)\n"
        "    h.update(u.tobytes() + v.tobytes())\n"
        "p = gaussian_filter((rng.random((2, 150, 203)) * 255).astype(np.float32), (0, 1.5, 1.5))\n"
        "c = np.stack([shift(p[0], (-2.6, 3.4), order=1, mode='nearest'), shift(p[1], (1.2, -0.3), order=1, mode='nearest')]).astype(np.float32)\n"
        "u, v = of_b200.lk_pyramidal_batch(p, c, 3, 5, 4, mode=of_b200.MODE_EXACT)\n"
        "h.update(u.tobytes() + v.tobytes())\n"
        "print(h.hexdigest())\n"
    )
    got = {}
    for refine, tile, exact in (("fused", "v1", "tile"), ("split", "v1", "tile"), ("split", "v2", "tile"), ("split", "v2", "march")):
        env = dict(os.environ, OF_B200_EXACT_REFINE=refine, OF_B200_TILE=tile, OF_B200_EXACT=exact)
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        assert res.returncode == 0, res.stderr[-2000:]
        got[(refine, tile, exact)] = res.stdout.strip().splitlines()[-1]
    assert len(set(got.values())) == 1, got


def test_fast_mode_refinement_variants_give_the_same_bits(ofb):
    """Fast mode's refinement iteration exists in four forms -- warp_rows + marching kernel, the marching
    kernel warping the next iteration's input in its epilogue (OF_B200_REFINE_WARP=chain), the kernel that
    gathers inside the marching warps (OF_B200_REFINE=fused), and the warp-specialised marching kernel whose producer
    warps fill the ring stages (OF_B200_REFINE=ws).  All use the same sample and blend routines:
    the pyramidal flow (window 5 and 7, early exits included) must hash identically."""
    import os
    import subprocess

    code = (
        "import sys, hashlib, numpy as np\n"
        f"sys.path.insert(0, {str(BACKEND_DIR)!r})\n"
        "import of_b200\n"
        "from scipy.ndimage import gaussian_filter, shift\n"
        "rng = np.random.default_rng(78)\n"
        "h = hashlib.sha256()\n"
        "p = gaussian_filter((rng.random((3, 152, 248)) * 255).astype(np.float32), (0, 1.5, 1.5))\n"
        "c = np.stack([shift(p[0], (-2.6, 3.4), order=1, mode='nearest'), shift(p[1], (1.2, -0.3), order=1, mode='nearest'), p[2]]).astype(np.float32)\n"
        "for w, it in ((5, 4), (7, 3), (5, 1)):\n"
        "    u, v = of_b200.lk_pyramidal_batch(p, c, 3, w, it, mode=of_b200.MODE_FAST)\n"
        "    h.update(u.tobytes() + v.tobytes())\n"
        "print(h.hexdigest())\n"
    )
    got = {}
    for refine, warp in (("split", "rows"), ("split", "chain"), ("fused", "rows"), ("ws", "rows")):
        env = dict(os.environ, OF_B200_REFINE=refine, OF_B200_REFINE_WARP=warp)
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        got[(refine, warp)] = res.stdout.strip().splitlines()[-1]
    assert len(set(got.values())) == 1, got


# ---------------------------------------------------------------------------------------
# full-size frames: oracle on one pair + size-independent properties
# ---------------------------------------------------------------------------------------
def test_1080p_batch_against_oracle_and_locality(ofb):
    import synthetic

    prev, curr, _ = synthetic.make_pairs_numpy(3, 1080, 1920, seed=11)
    u, v = ofb.lk_single_scale_batch(prev, curr, 5, mode=ofb.MODE_FAST)
    uo, vo = orc.lucas_kanade_single_scale(prev[1], curr[1], 5)
    assert_bit_equal(u[1], uo, "1080p u")
    assert_bit_equal(v[1], vo, "1080p v")
    # locality: flow at a pixel depends on a 7x7 neighbourhood only, so a crop computed on
    # its own agrees with the full frame 3 pixels inside the crop
    ys, xs = slice(401, 701), slice(1000, 1400)
    uc, vc = ofb.lk_single_scale(prev[2][ys, xs], curr[2][ys, xs], 5, mode=ofb.MODE_FAST)
    assert_bit_equal(uc[3:-3, 3:-3], u[2][ys, xs][3:-3, 3:-3], "crop u")
    assert_bit_equal(vc[3:-3, 3:-3], v[2][ys, xs][3:-3, 3:-3], "crop v")
    # identical frames -> exactly zero flow; the window_size // 2 border is zero
    uz, vz = ofb.lk_single_scale(prev[0], prev[0], 5, mode=ofb.MODE_FAST)
    assert not uz.any() and not vz.any()
    for arr in (u, v):
        assert not arr[:, :2].any() and not arr[:, -2:].any() and not arr[:, :, :2].any() and not arr[:, :, -2:].any()


def test_4k_fast_equals_exact_kernel(ofb):
    import synthetic

    prev, curr, _ = synthetic.make_pairs_numpy(1, 2160, 3840, seed=12)
    uf, vf = ofb.lk_single_scale(prev[0], curr[0], 5, mode=ofb.MODE_FAST)
    ue, ve = ofb.lk_single_scale(prev[0], curr[0], 5, mode=ofb.MODE_EXACT)
    assert_bit_equal(uf, ue, "4K u")
    assert_bit_equal(vf, ve, "4K v")


# ---------------------------------------------------------------------------------------
# fixed-point (RTL) mode
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("quirk", [True, False])
def test_fixed_point_mode_bit_exact(ofb, golden_frames, quirk):
    for name in ("translate_medium", "rotate_large", "no_motion"):
        p, c = golden_frames[name]
        uo, vo = fxo.lk_single_scale_fx(p, c, mirror_avg_quirk=quirk)
        u, v = ofb.lk_single_scale_fx(p, c, mirror_avg_quirk=quirk)
        assert u.dtype == np.int16
        assert_bit_equal(u, uo, f"{name} u")
        assert_bit_equal(v, vo, f"{name} v")
    rng = np.random.default_rng(9)
    p = rng.integers(0, 256, size=(3, 37, 75), dtype=np.uint8)  # worst-case contrast: exercises the 32-bit wraps
    c = rng.integers(0, 256, size=(3, 37, 75), dtype=np.uint8)
    u, v = ofb.lk_single_scale_fx(p, c, mirror_avg_quirk=quirk)
    for b in range(3):
        uo, vo = fxo.lk_single_scale_fx(p[b], c[b], mirror_avg_quirk=quirk)
        assert_bit_equal(u[b], uo, f"random {b} u")
        assert_bit_equal(v[b], vo, f"random {b} v")
    assert np.abs(u).max() <= 1024 and np.abs(v).max() <= 1024
    # worst-case contrast on frames the marching kernel takes (width % 16 == 0): 32-bit wraps of the
    # 64-bit products, quotients that wrap in 16 bits, the +-8 px clamp; ragged heights, tiny frames
    for shape in ((4, 64, 128), (2, 37, 48), (1, 7, 16), (1, 200, 256), (2, 9, 32)):
        p = rng.integers(0, 256, size=shape, dtype=np.uint8)
        c = rng.integers(0, 256, size=shape, dtype=np.uint8)
        u, v = ofb.lk_single_scale_fx(p, c, mirror_avg_quirk=quirk)
        for b in range(shape[0]):
            uo, vo = fxo.lk_single_scale_fx(p[b], c[b], mirror_avg_quirk=quirk)
            assert_bit_equal(u[b], uo, f"noise {shape} pair {b} u")
            assert_bit_equal(v[b], vo, f"noise {shape} pair {b} v")
    # low contrast: small determinants around the |det| > 1000 threshold
    p = (128 + rng.integers(-3, 4, size=(2, 96, 160))).astype(np.uint8)
    c = np.roll(p, 1, axis=2)
    u, v = ofb.lk_single_scale_fx(p, c, mirror_avg_quirk=quirk)
    for b in range(2):
        uo, vo = fxo.lk_single_scale_fx(p[b], c[b], mirror_avg_quirk=quirk)
        assert_bit_equal(u[b], uo, f"low contrast {b} u")
        assert_bit_equal(v[b], vo, f"low contrast {b} v")


def test_error_behaviour(ofb):
    z = np.zeros((16, 16), np.float32)
    with pytest.raises(ValueError):
        ofb.lk_single_scale(z, z, window_size=6)
    with pytest.raises(ofb.OFBackendError):
        ofb.lk_single_scale(z, z, window_size=13)  # unsupported, not silently approximated
    with pytest.raises(ValueError):
        ofb.lk_pyramidal(z, z, num_levels=6)  # 16 -> 8 -> 4 -> 2 -> 1 -> 0
    u, v = ofb.lk_single_scale(np.zeros((3, 3), np.float32), np.zeros((3, 3), np.float32), 5)
    assert u.shape == (3, 3) and not u.any()  # smaller than the window: all border


# ---------------------------------------------------------------------------------------
# row-band pyramidal mode (multi-GPU decomposition), emulated with thread ranks on one GPU
# ---------------------------------------------------------------------------------------
def _rowband_thread(rank, comm, out, args):
    import distributed as ofd

    p, c, levels, iters, mode = args[:5]
    window = args[5] if len(args) > 5 else 5
    try:
        out[rank] = ofd.lk_pyramidal_rowbands(p, c, levels, window, iters, mode=mode, comm=comm.view(rank),
                                              backend=ofd.CudaBackend())
    except Exception as e:  # surface the failure in the main thread
        out[rank] = e
        comm._barrier.abort()


@pytest.mark.parametrize("window", [5, 7])
@pytest.mark.parametrize("mode_name", ["exact", "fast"])
@pytest.mark.parametrize("world", [1, 3])
def test_pyramidal_rowbands_equal_single_gpu(ofb, world, mode_name, window):
    """Each rank computes only its rows (plus the shrinking overlap: 4 rows per iteration for window 5, 6 for
    window 7); the gathered result must equal the whole-frame run bit for bit, in both arithmetic modes."""
    import threading

    import distributed as ofd
    import synthetic

    mode = ofb.MODE_EXACT if mode_name == "exact" else ofb.MODE_FAST
    prev, curr, _ = synthetic.make_pairs_numpy(1, 200, 248, seed=31)
    p, c = prev[0], np.roll(curr[0], 3, axis=0)  # large enough motion for several iterations
    u1, v1, (iters_exec, _) = ofb.lk_pyramidal(p, c, 3, window, 3, mode=mode, return_trace=True)
    if world == 1:
        u, v = ofd.lk_pyramidal_rowbands(p, c, 3, window, 3, mode=mode, comm=ofd.SingleProcessComm(), backend=ofd.CudaBackend())
        results = [(u, v)]
    else:
        comm = ofd.ThreadComm(world)
        results = [None] * world
        ts = [threading.Thread(target=_rowband_thread, args=(r, comm, results, (p, c, 3, 3, mode, window))) for r in range(world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    for r, res in enumerate(results):
        assert not isinstance(res, Exception), res
        assert_bit_equal(res[0], u1, f"rank {r} u ({mode_name}, window {window})")
        assert_bit_equal(res[1], v1, f"rank {r} v ({mode_name}, window {window})")


def test_device_building_blocks_match_host_entry_points(ofb, golden_units):
    """of_pyramid_down_f32_dev / of_upsample_flow_f32_dev (row range) against the host calls."""
    import torch

    g = golden_units
    dev = torch.device("cuda", 0)
    src = torch.from_numpy(g["pyr_odd_in"]).to(dev)
    dst = torch.empty((22, 33), dtype=torch.float32, device=dev)
    ofb.pyramid_down_dev(src.data_ptr(), dst.data_ptr(), 1, 45, 67, 22, 33)
    torch.cuda.synchronize()
    assert_bit_equal(dst.cpu().numpy(), g["pyr_odd_level1"], "pyramid_down_dev")
    cu, cv = torch.from_numpy(g["up_u"]).to(dev), torch.from_numpy(g["up_v"]).to(dev)
    u = torch.full((61, 83), float("nan"), dtype=torch.float32, device=dev)
    v = torch.full((61, 83), float("nan"), dtype=torch.float32, device=dev)
    ofb.upsample_flow_dev(cu.data_ptr(), cv.data_ptr(), u.data_ptr(), v.data_ptr(), 1, 30, 40, 61, 83, 10, 37)
    torch.cuda.synchronize()
    un = u.cpu().numpy()
    assert_bit_equal(un[10:37], g["up_out_u_61x83"][10:37], "upsample rows")
    assert np.isnan(un[:10]).all() and np.isnan(un[37:]).all()  # rows outside the range untouched


@pytest.mark.parametrize("shape", [(240, 320), (64, 128), (37, 48), (1080, 1920), (50, 100), (33, 8)])
def test_uint8_ingest_equals_float_path(ofb, shape, golden_index, golden_frames):
    """of_lk_single_scale_u8: uint8 frames in, the float path's flow out, bit for bit -- through the
    TMA uint8 kernel where the frame allows it (width % 16 == 0), through widening + the float
    kernels elsewhere, in both modes; and equal to the oracle."""
    import torch

    H, W = shape
    rng = np.random.default_rng(H + W)
    if shape == (240, 320):
        p8, c8 = golden_frames["translate_medium"]
        p8, c8 = np.stack([p8, golden_frames["rotate_small"][0]]), np.stack([c8, golden_frames["rotate_small"][1]])
    else:
        p8 = rng.integers(0, 256, (3, H, W)).astype(np.uint8)
        c8 = np.clip(np.roll(p8, 1, axis=2).astype(np.int32) + rng.integers(-3, 4, p8.shape), 0, 255).astype(np.uint8)
    pf, cf = p8.astype(np.float32), c8.astype(np.float32)
    for mode in (ofb.MODE_FAST, ofb.MODE_EXACT):
        uf, vf = ofb.lk_single_scale_batch(pf, cf, 5, mode)
        u8_, v8_ = ofb.lk_single_scale_u8_batch(p8, c8, 5, mode)
        assert_bit_equal(u8_, uf, f"u ({shape}, mode {mode})")
        assert_bit_equal(v8_, vf, f"v ({shape}, mode {mode})")
    uo, vo = orc.lucas_kanade_single_scale(pf[0], cf[0], 5)  # u8_, v8_: the exact-mode run (last of the loop)
    assert_bit_equal(u8_[0], uo, "u vs oracle")
    assert_bit_equal(v8_[0], vo, "v vs oracle")
    if shape == (240, 320):
        assert sha(u8_[0]) == golden_index["patterns"]["translate_medium"]["single_scale"]["sha256_u"]
    # device entry point: takes the TMA-able frames, refuses the others loudly
    dev = torch.device("cuda", 0)
    pd, cd = torch.from_numpy(p8).to(dev), torch.from_numpy(c8).to(dev)
    ud = torch.empty(p8.shape, dtype=torch.float32, device=dev)
    vd = torch.empty_like(ud)
    if W % 16 == 0 and (H * W) % 16 == 0:
        ofb.lk_single_scale_u8_dev(pd.data_ptr(), cd.data_ptr(), ud.data_ptr(), vd.data_ptr(), p8.shape[0], H, W)
        torch.cuda.synchronize()
        uff, vff = ofb.lk_single_scale_batch(pf, cf, 5, ofb.MODE_FAST)  # random noise: fast != exact order
        assert_bit_equal(ud.cpu().numpy(), uff, "dev u")
        assert_bit_equal(vd.cpu().numpy(), vff, "dev v")
    else:
        with pytest.raises(ofb.OFBackendError):
            ofb.lk_single_scale_u8_dev(pd.data_ptr(), cd.data_ptr(), ud.data_ptr(), vd.data_ptr(), p8.shape[0], H, W)


def test_apply_motion_kernel_equals_scipy_shift(ofb):
    """of_apply_motion_u8 against the reference's apply_motion outputs (tests/golden/motion.npz) and,
    on a batch with one random sub-pixel shift per frame, against the oracle."""
    from conftest import GOLDEN
    from oracle import pattern_oracle as po

    z = np.load(GOLDEN / "motion.npz")
    cases = z["cases"]
    for name, src in (("texture", z["texture_128x96"]), ("noise", z["noise_53x37"])):
        batch = np.repeat(src[None], len(cases), axis=0)
        got = ofb.apply_motion_u8_batch(batch, cases[:, 0], cases[:, 1])
        for i in range(len(cases)):
            assert_bit_equal(got[i], z[f"{name}_shift_{i}"], f"{name} shift {cases[i].tolist()}")
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, (6, 270, 481)).astype(np.uint8)
    dx, dy = rng.uniform(-3, 3, 6), rng.uniform(-3, 3, 6)
    got = ofb.apply_motion_u8_batch(frames, dx, dy)
    for b in range(6):
        assert_bit_equal(got[b], po.apply_motion(frames[b], dx[b], dy[b]), f"random shift {b}")
    assert_bit_equal(ofb.apply_motion_u8_batch(frames[0], 0.0, 0.0), frames[0], "zero shift is the identity")
    assert_bit_equal(ofb.apply_motion_u8_batch(frames[0], 1.5, 0.0, cval=7.0), po.apply_motion(frames[0], 1.5, 0.0, 7.0), "cval")


def test_warp_affine_kernel_equals_opencv(ofb):
    """of_warp_affine_u8 against the reference's apply_motion_opencv outputs (cv2.warpAffine; the 13 verifier
    parameter sets + 7 random affine maps on two frames, tests/golden/motion.npz) and against the oracle on
    a larger frame; the matrices come from of_b200.motion_matrix (= cv2.getRotationMatrix2D + translation)."""
    from conftest import GOLDEN
    from oracle import pattern_oracle as po

    z = np.load(GOLDEN / "motion.npz")
    cases = z["affine_cases"]
    for name, src in (("texture", z["texture_128x96"]), ("noise", z["noise_53x37"])):
        h, w = src.shape
        mats = np.stack([ofb.motion_matrix(w, h, *c) for c in cases])
        for i, c in enumerate(cases):
            assert np.array_equal(mats[i], po.motion_matrix(w, h, *c))
        got = ofb.warp_affine_u8_batch(np.repeat(src[None], len(cases), axis=0), mats)
        for i in range(len(cases)):
            assert_bit_equal(got[i], z[f"{name}_affine_{i}"], f"{name} affine {cases[i].tolist()}")
    rng = np.random.default_rng(21)
    frames = rng.integers(0, 256, (60, 270, 481)).astype(np.uint8)  # more pairs than one launch carries matrices for
    params = np.stack([rng.uniform(-9, 9, 60), rng.uniform(-9, 9, 60), rng.uniform(-25, 25, 60), rng.uniform(0.7, 1.4, 60)], axis=1)
    mats = np.stack([ofb.motion_matrix(481, 270, *p) for p in params])
    got = ofb.warp_affine_u8_batch(frames, mats)
    for b in (0, 1, 47, 48, 59):
        assert_bit_equal(got[b], po.warp_affine_u8(frames[b], mats[b]), f"random affine {b}")
    ident = ofb.warp_affine_u8_batch(frames[0], ofb.motion_matrix(481, 270))
    assert_bit_equal(ident, frames[0], "identity")


# motion parameters of the reference's 13 verifier patterns (python/generate_test_suite.py:59-136): dx, dy, rotation, scale
VERIFIER_PATTERNS = {
    "translate_small": (0.5, 0.5, 0.0, 1.0), "translate_medium": (2.0, 0.0, 0.0, 1.0), "translate_large": (15.0, 0.0, 0.0, 1.0),
    "translate_vertical": (0.0, 10.0, 0.0, 1.0), "translate_diagonal": (10.0, 10.0, 0.0, 1.0), "rotate_small": (0.0, 0.0, 2.0, 1.0),
    "rotate_medium": (0.0, 0.0, 5.0, 1.0), "rotate_large": (0.0, 0.0, 15.0, 1.0), "zoom_in": (0.0, 0.0, 0.0, 1.1),
    "zoom_out": (0.0, 0.0, 0.0, 0.9), "translate_rotate": (5.0, 5.0, 3.0, 1.0), "no_motion": (0.0, 0.0, 0.0, 1.0),
    "translate_extreme": (30.0, 20.0, 0.0, 1.0),
}


def test_verifier_pipeline_on_the_device_reproduces_the_baseline(ofb, golden_index, golden_frames):
    """The reference's whole acceptance loop with every step on the GPU backend: second frames by
    of_warp_affine_u8 from the base texture (must equal the reference generator's frame_01 bit for bit),
    single-scale LK on the uint8 frames, metrics by of_flow_metrics_f32 -- against
    python/verification_baseline.json."""
    names = list(golden_index["patterns"])
    assert sorted(names) == sorted(VERIFIER_PATTERNS)
    base = golden_frames[names[0]][0]
    for n in names:
        assert np.array_equal(golden_frames[n][0], base)  # every pattern starts from the same texture
    h, w = base.shape
    mats = np.stack([ofb.motion_matrix(w, h, *VERIFIER_PATTERNS[n]) for n in names])
    second = ofb.warp_affine_u8_batch(np.repeat(base[None], len(names), axis=0), mats)
    for i, n in enumerate(names):
        assert_bit_equal(second[i], golden_frames[n][1], f"generated frame_01 of {n}")
    u, v = ofb.lk_single_scale_u8_batch(np.repeat(base[None], len(names), axis=0), second, 5, ofb.MODE_FAST)
    crop = golden_index["center_crop"]
    for i, n in enumerate(names):
        entry = golden_index["patterns"][n]
        assert sha(u[i]) == entry["single_scale"]["sha256_u"] and sha(v[i]) == entry["single_scale"]["sha256_v"], n
        m = ofb.flow_metrics_batch(u[i], v[i], entry["ground_truth"]["u"], entry["ground_truth"]["v"],
                                   ofb.verifier_test_region((h, w), n, crop))[0]
        for k, val in entry["verification_baseline"]["single_scale"].items():
            assert m[k] == pytest.approx(val, rel=2e-6, abs=1e-6), (n, k)


def test_gpu_flow_metrics_against_reference_baseline(ofb, golden_index, golden_frames):
    """of_flow_metrics_f32 (compute_all_metrics over the verifier's test region, on the device):
    all 13 patterns in ONE batched call per method.  Against the reference's
    verification_baseline.json: equal to 3 decimals (the north star's tolerance) and within 2e-6
    relative (float64 instead of float32-pairwise means); same against the CPU oracle."""
    names = list(golden_index["patterns"])
    crop = golden_index["center_crop"]
    for method, runner in (("single_scale", lambda p, c: ofb.lk_single_scale(p, c, 5, mode=ofb.MODE_EXACT)),
                           ("pyramidal", lambda p, c: ofb.lk_pyramidal(p, c, 3, 5, 3, mode=ofb.MODE_EXACT))):
        by_region = {}
        for name in names:
            p, c = (f.astype(np.float32) for f in golden_frames[name])
            u, v = runner(p, c)
            by_region.setdefault(ofb.verifier_test_region(p.shape, name, crop), []).append((name, u, v))
        for region, items in by_region.items():
            gts = [golden_index["patterns"][n]["ground_truth"] for n, _, _ in items]
            got = ofb.flow_metrics_batch(np.stack([u for _, u, _ in items]), np.stack([v for _, _, v in items]),
                                         [g["u"] for g in gts], [g["v"] for g in gts], region)
            for (name, u, v), g, m in zip(items, gts, got):
                mask = fm.test_region_mask(u.shape, name, crop)
                y0, y1, x0, x1 = region
                assert mask.sum() == (y1 - y0) * (x1 - x0) and mask[y0:y1, x0:x1].all(), name
                want = fm.all_metrics(u, v, g["u"], g["v"], mask)
                base = golden_index["patterns"][name]["verification_baseline"][method]
                for k in ofb.METRIC_NAMES:
                    assert m[k] == pytest.approx(want[k], rel=2e-6, abs=1e-6), (method, name, k)
                    assert m[k] == pytest.approx(base[k], rel=2e-6, abs=1e-6), (method, name, k)
                    assert round(m[k], 3) == round(base[k], 3), (method, name, k)
    # whole-frame region, a single [H, W] field, and the argument checks
    u = np.full((40, 50), 1.5, np.float32)
    m = ofb.flow_metrics_batch(u, -u, 1.0, 0.0)[0]
    assert m["mae_u"] == pytest.approx(0.5) and m["mae_v"] == pytest.approx(1.5)
    assert m["epe"] == pytest.approx(np.sqrt(0.25 + 2.25), rel=1e-6) and m["rmse"] == pytest.approx(m["epe"], rel=1e-6)
    with pytest.raises(ValueError):
        ofb.flow_metrics_batch(u, u, 0.0, 0.0, region=(10, 10, 0, 5))


@pytest.mark.parametrize("mode_name", ["exact", "fast"])
@pytest.mark.parametrize("world,shape,levels,iters,repl_px,window",
                         [(1, (200, 248), 3, 3, 0, 5), (3, (200, 248), 3, 3, 0, 5), (4, (270, 480), 4, 4, 0, 5),
                          (4, (270, 480), 4, 4, 10000, 5), (3, (200, 248), 3, 3, 10 ** 9, 5), (8, (96, 128), 2, 2, 0, 5),
                          (3, (200, 248), 3, 3, 0, 7), (4, (270, 480), 4, 4, 10000, 7)])
def test_native_rowband_driver_equals_single_gpu(ofb, world, shape, levels, iters, repl_px, window, mode_name):
    """of_rowband_run: `world` ranks emulated on ONE device, each with its own arena and stream;
    the peer "mapping" is the other contexts' arena pointers.  Every call only enqueues, so one
    host thread can issue all ranks; the ranks' kernels then meet through the flag words exactly as
    they do over NVLink.  Every rank's gathered flow must equal the whole-frame run bit for bit,
    with the same early-exit decisions."""
    import torch

    import synthetic

    mode = ofb.MODE_EXACT if mode_name == "exact" else ofb.MODE_FAST
    H, W = shape
    prev, curr, _ = synthetic.make_pairs_numpy(1, H, W, seed=31)
    p, c = prev[0], np.roll(curr[0], 3, axis=0)
    u1, v1, (iters_exec, _) = ofb.lk_pyramidal(p, c, levels, window, iters, mode=mode, return_trace=True)
    dev = torch.device("cuda", 0)
    pd, cd = torch.from_numpy(p).to(dev), torch.from_numpy(c).to(dev)
    ctxs = [ofb.RowbandContext(r, world, H, W, levels, window, iters, mode) for r in range(world)]
    try:
        arenas = [cx.arena_ptr for cx in ctxs]
        for cx in ctxs:
            cx.set_peers(arenas)
            # 0: every level in row bands; 10000: the two coarsest levels whole on every rank; 1e9: all
            cx.set_replicate_pixels(repl_px)
        streams = [torch.cuda.Stream(device=dev) for _ in range(world)]  # non-blocking streams
        torch.cuda.synchronize()
        for rep in range(2):  # the second run reuses the arenas and the flag sequence numbers
            outs = []
            for r, cx in enumerate(ctxs):
                uo, vo = torch.empty_like(pd), torch.empty_like(pd)
                cx.run(pd.data_ptr(), cd.data_ptr(), uo.data_ptr(), vo.data_ptr(), streams[r].cuda_stream)
                outs.append((uo, vo))
            for r, cx in enumerate(ctxs):
                it_r, _, err = cx.trace(streams[r].cuda_stream)
                assert err == 0, f"rank {r}: a wait timed out"
                assert it_r.tolist() == np.asarray(iters_exec).reshape(-1).tolist(), (r, it_r, iters_exec)
                assert_bit_equal(outs[r][0].cpu().numpy(), u1, f"rank {r} u ({mode_name}, run {rep})")
                assert_bit_equal(outs[r][1].cpu().numpy(), v1, f"rank {r} v ({mode_name}, run {rep})")
    finally:
        for cx in ctxs:
            cx.close()


def test_peer_rowband_lanes_single_process(ofb):
    """distributed.PeerRowbandLanes without a process group (one rank): three pairs over two lanes,
    plain and replayed from a CUDA graph; every pair equals the single-GPU driver bit for bit."""
    import torch

    import distributed as ofd
    import synthetic

    H, W = 200, 248
    prev, curr, _ = synthetic.make_pairs_numpy(3, H, W, seed=17)
    curr = np.stack([np.roll(c, 2, axis=0) for c in curr])
    want = [ofb.lk_pyramidal(prev[b], curr[b], 3, 5, 3, mode=ofb.MODE_FAST) for b in range(3)]
    dev = torch.device("cuda", 0)
    pd, cd = torch.from_numpy(prev).to(dev), torch.from_numpy(curr).to(dev)
    u, v = torch.zeros_like(pd), torch.zeros_like(pd)
    lanes = ofd.PeerRowbandLanes(H, W, 3, 5, 3, ofb.MODE_FAST, lanes=2)
    try:
        lanes.run_batch(pd, cd, u, v)
        lanes.trace()
        for b in range(3):
            assert_bit_equal(u[b].cpu().numpy(), want[b][0], f"pair {b} u")
            assert_bit_equal(v[b].cpu().numpy(), want[b][1], f"pair {b} v")
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            lanes.run_batch(pd, cd, u, v)
        for _ in range(2):
            u.zero_()
            v.zero_()
            graph.replay()
            lanes.trace()
            for b in range(3):
                assert_bit_equal(u[b].cpu().numpy(), want[b][0], f"graph replay, pair {b} u")
    finally:
        lanes.close()


def test_native_rowband_driver_times_out_instead_of_hanging(ofb):
    """A rank whose peer never shows up must not wedge the GPU: every spin has a time-out that sets
    the error word, the run drains, of_rowband_trace / of_rowband_status report OF_ERR_PEER_TIMEOUT, and the
    run after that starts clean."""
    import time

    import torch

    import synthetic

    H, W = 200, 248
    prev, curr, _ = synthetic.make_pairs_numpy(1, H, W, seed=5)
    dev = torch.device("cuda", 0)
    pd, cd = torch.from_numpy(prev[0]).to(dev), torch.from_numpy(curr[0]).to(dev)
    ctxs = [ofb.RowbandContext(r, 2, H, W, 3, 5, 3, ofb.MODE_FAST) for r in range(2)]
    try:
        for cx in ctxs:
            cx.set_peers([c.arena_ptr for c in ctxs])
            cx.set_replicate_pixels(0)
            cx.set_timeout_ms(40)
        st = torch.cuda.Stream(device=dev)
        t0 = time.perf_counter()
        ctxs[0].run(pd.data_ptr(), cd.data_ptr(), None, None, st.cuda_stream)  # rank 1 never runs
        with pytest.raises(ofb.OFBackendError, match="time-out"):
            ctxs[0].trace(st.cuda_stream)
        assert time.perf_counter() - t0 < 5.0  # one time-out, then every later wait falls through
        # The error has been reported, so the next run starts with a clean error word: both ranks run now
        # (two streams of the one device); rank 0's sequence numbers are one run ahead of rank 1's, so rank 1
        # runs twice -- its first run pairs with the run rank 0 gave up on.
        s1 = torch.cuda.Stream(device=dev)
        ctxs[1].run(pd.data_ptr(), cd.data_ptr(), None, None, s1.cuda_stream)
        try:
            ctxs[1].status(s1.cuda_stream)  # rank 0 abandoned that run half-way: rank 1 times out as well (40 ms)
        except ofb.OFBackendError:
            pass
        for cx in ctxs:
            cx.set_timeout_ms(4000)
        u0, v0 = torch.empty_like(pd), torch.empty_like(pd)
        ctxs[0].run(pd.data_ptr(), cd.data_ptr(), u0.data_ptr(), v0.data_ptr(), st.cuda_stream)
        ctxs[1].run(pd.data_ptr(), cd.data_ptr(), None, None, s1.cuda_stream)
        ctxs[0].status(st.cuda_stream)
        ctxs[1].status(s1.cuda_stream)
        u1, v1 = ofb.lk_pyramidal(prev[0], curr[0], 3, 5, 3, mode=ofb.MODE_FAST)
        assert_bit_equal(u0.cpu().numpy(), u1, "run after a reported time-out, u")
        assert_bit_equal(v0.cpu().numpy(), v1, "run after a reported time-out, v")
    finally:
        for cx in ctxs:
            cx.close()


@pytest.mark.parametrize("shape", [(16, 248), (135, 249), (300, 517), (333, 1000), (1080, 1920)])
def test_pyramid_marching_kernel_against_oracle(ofb, shape):
    """One pyramid level by the marching kernel (frames >= 16 x 248) on general floats, odd sizes,
    a batch, and a row range; bit-exact against the SciPy-order oracle."""
    import torch

    H, W = shape
    rng = np.random.default_rng(H * 7 + W)
    imgs = (rng.standard_normal((3, H, W)) * 60.0 + 100.0).astype(np.float32)
    oh, ow = H // 2, W // 2
    want = np.stack([orc.build_gaussian_pyramid(im, 2)[0] for im in imgs])
    assert want.shape == (3, oh, ow)
    assert_bit_equal(ofb.pyramid_down(imgs[1]), want[1], "host entry point")
    dev = torch.device("cuda", 0)
    src = torch.from_numpy(imgs).to(dev)
    dst = torch.full((3, oh, ow), float("nan"), dtype=torch.float32, device=dev)
    ofb.pyramid_down_dev(src.data_ptr(), dst.data_ptr(), 3, H, W, oh, ow)
    torch.cuda.synchronize()
    assert_bit_equal(dst.cpu().numpy(), want, "batch of 3")
    lo, hi = oh // 3, oh - 1
    dst.fill_(float("nan"))
    ofb.pyramid_down_dev(src.data_ptr(), dst.data_ptr(), 3, H, W, oh, ow, row_lo=lo, row_hi=hi)
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    assert_bit_equal(got[:, lo:hi], want[:, lo:hi], "row range")
    assert np.isnan(got[:, :lo]).all() and np.isnan(got[:, hi:]).all()


def test_upsample_flow_large_against_oracle(ofb):
    rng = np.random.default_rng(5)
    cu = (rng.standard_normal((135, 240)) * 3).astype(np.float32)
    cv = (rng.standard_normal((135, 240)) * 3).astype(np.float32)
    # (270, 480) ... take the staged-tile kernel (grid step <= ~0.5); (180, 300) and (100, 200) have
    # steps of 0.75 / 1.35 and take the generic one
    for shape in ((270, 480), (271, 481), (400, 700), (180, 300), (100, 200)):
        wu, wv = orc.upsample_flow(cu, cv, shape)
        gu, gv = ofb.upsample_flow(cu, cv, shape)
        assert_bit_equal(gu, wu, f"upsample u {shape}")
        assert_bit_equal(gv, wv, f"upsample v {shape}")


# ---------------------------------------------------------------------------------------
# edge cases of the boundary
# ---------------------------------------------------------------------------------------
def test_fast_mode_falls_back_to_reference_order_kernel_on_unaligned_widths(ofb):
    """Widths that are not a multiple of 4 cannot use TMA / 128-bit loads: fast mode then runs the
    exact kernel (never a CPU path), so the result is still the reference's."""
    rng = np.random.default_rng(17)
    for shape in ((45, 67), (9, 13), (64, 126)):
        p = (rng.random(shape) * 255).astype(np.float32)
        c = (p + rng.standard_normal(shape).astype(np.float32)).astype(np.float32)
        uo, vo = orc.lucas_kanade_single_scale(p, c, 5)
        u, v = ofb.lk_single_scale(p, c, 5, mode=ofb.MODE_FAST)
        assert_bit_equal(u, uo, f"{shape} u")
        assert_bit_equal(v, vo, f"{shape} v")


def test_empty_batch_and_tiny_frames(ofb):
    z = np.zeros((0, 32, 32), np.float32)
    u, v = ofb.lk_single_scale_batch(z, z, 5, mode=ofb.MODE_FAST)
    assert u.shape == (0, 32, 32)
    u, v = ofb.lk_pyramidal_batch(z, z, 2, 5, 2)
    assert u.shape == (0, 32, 32)
    for shape in ((1, 1), (1, 8), (8, 1), (5, 5), (6, 8)):
        p = np.arange(shape[0] * shape[1], dtype=np.float32).reshape(shape)
        c = p[::-1].copy()
        for mode in (ofb.MODE_FAST, ofb.MODE_EXACT):
            u, v = ofb.lk_single_scale(p, c, 5, mode=mode)
            uo, vo = orc.lucas_kanade_single_scale(p, c, 5)
            assert_bit_equal(u, uo, f"{shape} mode {mode}")
            assert_bit_equal(v, vo, f"{shape} mode {mode}")


def test_8k_single_scale_locality_and_unaligned_base(ofb):
    """Largest frame of BASELINE.json (7680x4320): full-frame fast kernel against crops computed
    on their own (locality), plus planes whose base pointer is not 16-byte aligned (the marching
    kernel moves 128-bit words, so the driver must route those to the tile kernel)."""
    import synthetic
    import torch

    prev, curr, _ = synthetic.make_pairs_numpy(1, 4320, 7680, seed=41)
    u, v = ofb.lk_single_scale(prev[0], curr[0], 5, mode=ofb.MODE_FAST)
    for ys, xs in ((slice(0, 300), slice(0, 400)), (slice(4000, 4320), slice(7200, 7680)), (slice(2001, 2301), slice(3333, 3733))):
        uo, vo = orc.lucas_kanade_single_scale(prev[0][ys, xs], curr[0][ys, xs], 5)
        h, w = uo.shape
        # rows / columns that are interior to the crop but not to the frame differ (crop border)
        y0 = 0 if ys.start == 0 else 3
        y1 = h if ys.stop == 4320 else h - 3
        x0 = 0 if xs.start == 0 else 3
        x1 = w if xs.stop == 7680 else w - 3
        assert_bit_equal(u[ys, xs][y0:y1, x0:x1], uo[y0:y1, x0:x1], "8K crop u")
        assert_bit_equal(v[ys, xs][y0:y1, x0:x1], vo[y0:y1, x0:x1], "8K crop v")
    # device pointers offset by one float
    dev = torch.device("cuda", 0)
    h, w = 200, 248
    buf_p = torch.zeros(h * w + 1, dtype=torch.float32, device=dev)
    buf_c = torch.zeros(h * w + 1, dtype=torch.float32, device=dev)
    buf_p[1:] = torch.from_numpy(prev[0][:h, :w].copy()).flatten().to(dev)
    buf_c[1:] = torch.from_numpy(curr[0][:h, :w].copy()).flatten().to(dev)
    out_u = torch.empty(h * w, dtype=torch.float32, device=dev)
    out_v = torch.empty(h * w, dtype=torch.float32, device=dev)
    ofb.lk_single_scale_dev(buf_p.data_ptr() + 4, buf_c.data_ptr() + 4, out_u.data_ptr(), out_v.data_ptr(), 1, h, w, 5,
                            ofb.MODE_FAST)
    torch.cuda.synchronize()
    uo, vo = orc.lucas_kanade_single_scale(prev[0][:h, :w], curr[0][:h, :w], 5)
    assert_bit_equal(out_u.cpu().numpy().reshape(h, w), uo, "unaligned base u")
    assert_bit_equal(out_v.cpu().numpy().reshape(h, w), vo, "unaligned base v")


def _visible_gpus() -> int:
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("mode_name", ["fast", "exact"])
def test_rowband_over_real_nvlink_multi_process(ofb, mode_name):
    """The row-band driver between real GPUs: one process per GPU under torch.distributed.run, arenas mapped
    through CUDA IPC, collectives by peer stores inside the kernels (tests/run_rowband_nccl.py --driver peer).
    Rank 0 asserts the gathered flow equals the single-GPU driver bit for bit.  Needs > 1 visible GPU: on the
    one-GPU box of the round-end suite it skips (the emulated-ranks test above covers the logic there)."""
    import socket
    import subprocess

    n = _visible_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "run_rowband_nccl.py"), "--height", "1080", "--width", "1920",
           "--levels", "4", "--iters", "4", "--mode", mode_name, "--driver", "peer", "--repeat", "2"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == world and out["bit_equal_to_single_gpu"] is True


@pytest.mark.parametrize("kernel", ["march", "tile"])
def test_fixed_point_mode_against_rtl_text_vectors(ofb, kernel):
    """a8 against vectors that were produced by executing the reference's RTL text (oracle/sv_eval.py on
    gradient_compute.sv / window_accumulator.sv / flow_solver.sv, tests/golden/make_golden_fx_rtl_text.py): 320
    neighbourhoods tiled into one frame pair, S8.7 flow compared at the patch centres; both kernels (the frame's width
    is a multiple of 16 -> marching kernel; one column more -> tile kernel)."""
    from conftest import GOLDEN

    z = np.load(GOLDEN / "fx_rtl_text_vectors.npz")
    fp, fc = z["frame_prev"], z["frame_curr"]
    if kernel == "tile":
        fp = np.ascontiguousarray(np.pad(fp, ((0, 0), (0, 1)), constant_values=128))
        fc = np.ascontiguousarray(np.pad(fc, ((0, 0), (0, 1)), constant_values=128))
    u, v = ofb.lk_single_scale_fx(fp, fc)
    cy, cx = z["centres"][:, 0], z["centres"][:, 1]
    assert_bit_equal(u[cy, cx], z["u"], f"u at the patch centres ({kernel})")
    assert_bit_equal(v[cy, cx], z["v"], f"v at the patch centres ({kernel})")
    assert int((np.abs(z["u"].astype(np.int32)) + np.abs(z["v"].astype(np.int32)) > 0).sum()) > 100
