#!/usr/bin/env python3
"""Generate the committed golden fixtures by running the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # ~1 min on 8 cores

It imports the reference's own modules from /root/reference/python (with a stub
``matplotlib``, which lucas_kanade_pyramidal.py:226 -> :321 needs), regenerates the
13-pattern verifier suite with the reference's generate_test_suite.py, runs the
reference's single-scale and pyramidal LK on every pattern, and writes

    tests/golden/frames.npz          13 x (frame_00, frame_01) uint8 320x240
    tests/golden/flows_subset.npz    full reference u,v (both methods) for 3 patterns
    tests/golden/units.npz           inputs + reference outputs of the helper
                                     functions (gradients, pyramid, warp, upsample,
                                     small float cases with windows 3/5/7)
    tests/golden/golden_index.json   sha256 of every reference output, the metrics
                                     the reference's flow_metrics gives, and the
                                     numbers of python/verification_baseline.json

Nothing on the GPU box reads /root/reference; the tests there read these files.
The frames derive from python/test_data/mountain_texture.jpg (CC BY-SA 3.0, see
the reference README.md:7-18 for the attribution).
"""

from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import sys
import tempfile
import types
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path
from unittest import mock

import numpy as np

REF = Path("/root/reference/python")
OUT = Path(__file__).resolve().parent


def _install_matplotlib_stub() -> None:
    mpl = types.ModuleType("matplotlib")
    pyplot = mock.MagicMock(name="pyplot")
    pyplot.subplots.side_effect = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
    colors = mock.MagicMock(name="colors")
    mpl.pyplot = pyplot
    mpl.colors = colors
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", pyplot)
    sys.modules.setdefault("matplotlib.colors", colors)


def _ref_modules():
    _install_matplotlib_stub()
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import flow_metrics  # noqa: E402
    import lucas_kanade_core  # noqa: E402
    import lucas_kanade_pyramidal  # noqa: E402
    import optical_flow_verifier  # noqa: E402

    return lucas_kanade_core, lucas_kanade_pyramidal, flow_metrics, optical_flow_verifier


def sha(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def _run_pattern(args):
    name, f0, f1 = args
    core, pyr, _, _ = _ref_modules()
    p = f0.astype(np.float32)
    c = f1.astype(np.float32)
    scratch = tempfile.mkdtemp(prefix="ofgold_")
    cwd = os.getcwd()
    os.chdir(scratch)  # the reference writes python/output/*.png relative to cwd
    try:
        with contextlib.redirect_stdout(io.StringIO()) as log:
            us, vs = core.lucas_kanade_single_scale(p, c, window_size=5)
            up, vp = pyr.lucas_kanade_pyramidal(p, c, num_levels=3, window_size=5, num_iterations=3)
    finally:
        os.chdir(cwd)
    executed = [ln.strip() for ln in log.getvalue().splitlines() if "Iteration" in ln or "Converged" in ln]
    return name, us, vs, up, vp, executed


def main() -> None:
    core, pyr, fm, ver = _ref_modules()
    import generate_test_suite as gts  # reference fixture generator

    tmp = Path(tempfile.mkdtemp(prefix="ofsuite_"))
    with contextlib.redirect_stdout(io.StringIO()):
        gts.generate_full_suite(320, 240, tmp)
    suite = json.load(open(tmp / "suite_index.json"))
    names = [p["name"] if isinstance(p, dict) else p for p in suite["patterns"]]
    frames = {}
    meta = {}
    for n in names:
        f0 = np.fromfile(tmp / n / "frame_00.bin", dtype=np.uint8).reshape(240, 320)
        f1 = np.fromfile(tmp / n / "frame_01.bin", dtype=np.uint8).reshape(240, 320)
        frames[n] = (f0, f1)
        meta[n] = json.load(open(tmp / n / "metadata.json"))["motion_parameters"]

    baseline = json.load(open(REF / "verification_baseline.json"))["patterns"]

    with ProcessPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(_run_pattern, [(n, *frames[n]) for n in names]))

    index = {
        "generator": "tests/golden/make_golden.py",
        "reference": "rothej/optical-flow-fpga python/ (run in the build container)",
        "versions": {
            "numpy": np.__version__,
            "scipy": __import__("scipy").__version__,
            "cv2": __import__("cv2").__version__,
            "PIL": __import__("PIL").__version__,
        },
        "shape": [240, 320],
        "pyramid_config": {"levels": 3, "window_size": 5, "iterations": 3},
        "center_crop": 80,
        "patterns": {},
    }
    subset = {}
    keep_full = ("translate_small", "rotate_small", "translate_extreme")
    worst = 0.0
    for name, us, vs, up, vp, executed in results:
        mask = ver.get_test_region_mask((240, 320), name, 80)
        gt_u, gt_v = meta[name]["dx"], meta[name]["dy"]
        m_single = fm.compute_all_metrics(us, vs, gt_u, gt_v, mask)
        m_pyr = fm.compute_all_metrics(up, vp, gt_u, gt_v, mask)
        for k in m_single:
            worst = max(worst, abs(m_single[k] - baseline[name]["single_scale"]["metrics"][k]))
            worst = max(worst, abs(m_pyr[k] - baseline[name]["pyramidal"]["metrics"][k]))
        index["patterns"][name] = {
            "ground_truth": {"u": gt_u, "v": gt_v},
            "num_test_pixels": int(mask.sum()),
            "single_scale": {"sha256_u": sha(us), "sha256_v": sha(vs), "metrics": m_single},
            "pyramidal": {
                "sha256_u": sha(up),
                "sha256_v": sha(vp),
                "metrics": m_pyr,
                "reference_log": executed,
            },
            "verification_baseline": {
                "single_scale": baseline[name]["single_scale"]["metrics"],
                "pyramidal": baseline[name]["pyramidal"]["metrics"],
            },
        }
        if name in keep_full:
            subset[f"{name}__single_u"] = us
            subset[f"{name}__single_v"] = vs
            subset[f"{name}__pyr_u"] = up
            subset[f"{name}__pyr_v"] = vp
    index["max_abs_metric_diff_vs_verification_baseline"] = worst
    print("max |metric - verification_baseline.json| =", worst)

    # ---- helper-function goldens (unit level) --------------------------------
    units = {}
    p = frames["translate_rotate"][0].astype(np.float32)
    c = frames["translate_rotate"][1].astype(np.float32)
    ix, iy, it = core.compute_gradients(p, c)
    units["grad_prev"], units["grad_curr"] = p, c
    units["grad_ix"], units["grad_iy"], units["grad_it"] = ix, iy, it

    pyr_levels = pyr.build_gaussian_pyramid(p, 4)
    for i, lvl in enumerate(pyr_levels):
        units[f"pyr4_level{i}"] = lvl

    rng = np.random.default_rng(20261018)
    # odd-sized image -> int(h*0.5) truncation and non-2x linspace grid
    odd = (rng.random((45, 67)) * 255).astype(np.float32)
    units["pyr_odd_in"] = odd
    for i, lvl in enumerate(pyr.build_gaussian_pyramid(odd, 3)):
        units[f"pyr_odd_level{i}"] = lvl

    # warp with a flow that leaves the frame on every side and lands on exact edges
    img = (rng.random((48, 64)) * 255).astype(np.float32)
    fu = (rng.standard_normal((48, 64)) * 3).astype(np.float32)
    fv = (rng.standard_normal((48, 64)) * 3).astype(np.float32)
    fu[0, :8] = 0.0
    fv[0, :8] = 0.0
    fu[5, 5] = 58.0  # x = 63 = W-1 exactly
    fv[5, 5] = 42.0  # y = 47 = H-1 exactly
    fu[6, 6] = -6.0  # x = 0 exactly
    fv[6, 6] = -6.0
    fu[7, 7] = np.float32(56.000004)  # just past the last column
    units["warp_img"], units["warp_u"], units["warp_v"] = img, fu, fv
    units["warp_out"] = pyr.warp_image(img, fu, fv)

    cu = (rng.standard_normal((30, 40)) * 2).astype(np.float32)
    cv = (rng.standard_normal((30, 40)) * 2).astype(np.float32)
    units["up_u"], units["up_v"] = cu, cv
    uu, vv = pyr.upsample_flow(cu, cv, (60, 80))
    units["up_out_u_60x80"], units["up_out_v_60x80"] = uu, vv
    uu, vv = pyr.upsample_flow(cu, cv, (61, 83))
    units["up_out_u_61x83"], units["up_out_v_61x83"] = uu, vv

    # general float (not uint8-valued) frames: summation order matters here
    for w in (3, 5, 7):
        a = (rng.random((37, 45)) * 255).astype(np.float32)
        b = (a + rng.standard_normal((37, 45)).astype(np.float32) * 4).astype(np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            u, v = core.lucas_kanade_single_scale(a, b, window_size=w)
        units[f"float_w{w}_prev"], units[f"float_w{w}_curr"] = a, b
        units[f"float_w{w}_u"], units[f"float_w{w}_v"] = u, v
    gx = rng.standard_normal((33, 41)).astype(np.float32)
    gy = rng.standard_normal((33, 41)).astype(np.float32)
    gt = rng.standard_normal((33, 41)).astype(np.float32)
    u, v = core.lucas_kanade_from_gradients(gx, gy, gt, 5)
    units["fg_ix"], units["fg_iy"], units["fg_it"] = gx, gy, gt
    units["fg_u"], units["fg_v"] = u, v

    # pyramidal on a small float pair with 2 levels / 2 iterations / window 7
    sp = (rng.random((64, 80)) * 255).astype(np.float32)
    sc = np.roll(sp, 1, axis=1)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="ofgold_"))
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            u, v = pyr.lucas_kanade_pyramidal(sp, sc, num_levels=2, window_size=7, num_iterations=2)
    finally:
        os.chdir(cwd)
    units["pyr_small_prev"], units["pyr_small_curr"] = sp, sc
    units["pyr_small_u"], units["pyr_small_v"] = u, v

    index["units_sha256"] = {k: sha(v) for k, v in units.items()}

    np.savez_compressed(
        OUT / "frames.npz",
        **{f"{n}__0": frames[n][0] for n in names},
        **{f"{n}__1": frames[n][1] for n in names},
    )
    np.savez_compressed(OUT / "flows_subset.npz", **subset)
    np.savez_compressed(OUT / "units.npz", **units)
    index["pattern_order"] = names
    with open(OUT / "golden_index.json", "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)
    for fn in ("frames.npz", "flows_subset.npz", "units.npz", "golden_index.json"):
        print(fn, (OUT / fn).stat().st_size, "bytes")


if __name__ == "__main__":
    main()
