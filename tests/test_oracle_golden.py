"""The CPU oracle against the reference's own outputs (committed golden fixtures).

Everything here is bit-exact: the fixtures were produced by running the reference's
modules (tests/golden/make_golden.py), and the oracle restates their arithmetic in the
same operation order.
"""

import hashlib

import numpy as np
import pytest

from oracle import flow_metrics_oracle as fm
from oracle import lk_float_oracle as orc


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_gradients_match_reference(golden_units):
    g = golden_units
    ix, iy, it = orc.compute_gradients(g["grad_prev"], g["grad_curr"])
    assert np.array_equal(ix, g["grad_ix"])
    assert np.array_equal(iy, g["grad_iy"])
    assert np.array_equal(it, g["grad_it"])
    assert ix.dtype == np.float32


@pytest.mark.parametrize("w", [3, 5, 7])
def test_single_scale_general_float_inputs(golden_units, w):
    g = golden_units
    u, v = orc.lucas_kanade_single_scale(g[f"float_w{w}_prev"], g[f"float_w{w}_curr"], w)
    assert np.array_equal(u, g[f"float_w{w}_u"])
    assert np.array_equal(v, g[f"float_w{w}_v"])


def test_from_gradients(golden_units):
    g = golden_units
    u, v = orc.lucas_kanade_from_gradients(g["fg_ix"], g["fg_iy"], g["fg_it"], 5)
    assert np.array_equal(u, g["fg_u"]) and np.array_equal(v, g["fg_v"])
    # the scalar-loop restatement agrees with the vectorised one
    u2, v2 = orc.lucas_kanade_from_gradients_loop(g["fg_ix"], g["fg_iy"], g["fg_it"], 5)
    assert np.array_equal(u, u2) and np.array_equal(v, v2)


def test_pyramid_levels(golden_units):
    g = golden_units
    pyr = orc.build_gaussian_pyramid(g["grad_prev"], 4)
    assert [p.shape for p in pyr] == [(30, 40), (60, 80), (120, 160), (240, 320)]
    for i, lvl in enumerate(pyr):
        assert np.array_equal(lvl, g[f"pyr4_level{i}"]), f"level {i}"
    pyr = orc.build_gaussian_pyramid(g["pyr_odd_in"], 3)
    assert [p.shape for p in pyr] == [(11, 16), (22, 33), (45, 67)]
    for i, lvl in enumerate(pyr):
        assert np.array_equal(lvl, g[f"pyr_odd_level{i}"]), f"odd level {i}"


def test_warp_and_upsample(golden_units):
    g = golden_units
    out = orc.warp_image(g["warp_img"], g["warp_u"], g["warp_v"])
    assert np.array_equal(out, g["warp_out"])
    assert out[7, 7] == 0.0  # 1e-6 past the last column is outside
    assert out[5, 5] == g["warp_img"][47, 63]
    for shape in ((60, 80), (61, 83)):
        u, v = orc.upsample_flow(g["up_u"], g["up_v"], shape)
        assert np.array_equal(u, g[f"up_out_u_{shape[0]}x{shape[1]}"])
        assert np.array_equal(v, g[f"up_out_v_{shape[0]}x{shape[1]}"])


def test_small_pyramidal_window7(golden_units):
    g = golden_units
    u, v = orc.lucas_kanade_pyramidal(g["pyr_small_prev"], g["pyr_small_curr"], 2, 7, 2)
    assert np.array_equal(u, g["pyr_small_u"]) and np.array_equal(v, g["pyr_small_v"])


def test_blur_and_sampler_match_scipy():
    """The restated third-party pieces against the installed SciPy itself."""
    from scipy import signal
    from scipy.ndimage import gaussian_filter, map_coordinates

    rng = np.random.default_rng(7)
    img = (rng.random((53, 47)) * 255).astype(np.float32)
    assert np.array_equal(orc.gaussian_blur_sigma(img, 2.0), gaussian_filter(img, sigma=2.0))
    tiny = (rng.random((5, 9)) * 255).astype(np.float32)  # reflect wraps more than once
    assert np.array_equal(orc.gaussian_blur_sigma(tiny, 2.0), gaussian_filter(tiny, sigma=2.0))
    yy = rng.random((40, 40)) * 60 - 4
    xx = rng.random((40, 40)) * 55 - 4
    ref = map_coordinates(img, [yy, xx], order=1, mode="constant", cval=0.0)
    assert np.array_equal(orc.bilinear_sample(img, yy, xx), ref)
    k = orc.SOBEL_X
    assert np.array_equal(
        orc._conv3x3_symm(img, k), signal.convolve2d(img, k, mode="same", boundary="symm")
    )


def test_all_13_patterns_bit_exact_and_metrics(golden_index, golden_frames):
    """Single-scale and 3x3 pyramidal flow for every verifier pattern: sha256 of the
    oracle's output equals the reference's, and the metrics equal
    python/verification_baseline.json to the last digit."""
    for name, entry in golden_index["patterns"].items():
        f0, f1 = golden_frames[name]
        p, c = f0.astype(np.float32), f1.astype(np.float32)
        us, vs = orc.lucas_kanade_single_scale(p, c, 5)
        assert sha(us) == entry["single_scale"]["sha256_u"], name
        assert sha(vs) == entry["single_scale"]["sha256_v"], name
        trace = []
        up, vp = orc.lucas_kanade_pyramidal(p, c, 3, 5, 3, trace=trace)
        assert sha(up) == entry["pyramidal"]["sha256_u"], name
        assert sha(vp) == entry["pyramidal"]["sha256_v"], name
        n_ref_iters = sum(1 for ln in entry["pyramidal"]["reference_log"] if ln.startswith("Iteration"))
        assert len(trace) == n_ref_iters, name
        mask = fm.test_region_mask(p.shape, name, golden_index["center_crop"])
        assert int(mask.sum()) == entry["num_test_pixels"]
        gt = entry["ground_truth"]
        for flows, key in (((us, vs), "single_scale"), ((up, vp), "pyramidal")):
            m = fm.all_metrics(flows[0], flows[1], gt["u"], gt["v"], mask)
            for k, val in entry["verification_baseline"][key].items():
                assert m[k] == val, (name, key, k)


def test_subset_full_arrays(golden_flows, golden_frames):
    for name in ("translate_small", "rotate_small", "translate_extreme"):
        f0, f1 = golden_frames[name]
        us, vs = orc.lucas_kanade_single_scale(f0.astype(np.float32), f1.astype(np.float32))
        assert np.array_equal(us, golden_flows[f"{name}__single_u"])
        assert np.array_equal(vs, golden_flows[f"{name}__single_v"])


def test_pattern_oracle_equals_reference_generators():
    """oracle/pattern_oracle.py against tests/golden/motion.npz, which holds outputs of the
    reference's own generate_smooth_synthetic and apply_motion (scipy.ndimage.shift)."""
    from conftest import GOLDEN
    from oracle import pattern_oracle as po

    z = np.load(GOLDEN / "motion.npz")
    assert np.array_equal(po.generate_smooth_synthetic(128, 96), z["texture_128x96"])
    for i, (dx, dy) in enumerate(z["cases"]):
        for name, src in (("texture", z["texture_128x96"]), ("noise", z["noise_53x37"])):
            assert np.array_equal(po.apply_motion(src, dx, dy), z[f"{name}_shift_{i}"]), (name, dx, dy)
    for i, (dx, dy, rot, sc) in enumerate(z["affine_cases"]):  # apply_motion_opencv = cv2.warpAffine in fixed point
        for name, src in (("texture", z["texture_128x96"]), ("noise", z["noise_53x37"])):
            assert np.array_equal(po.apply_motion_opencv(src, dx, dy, rot, sc), z[f"{name}_affine_{i}"]), (name, i)

