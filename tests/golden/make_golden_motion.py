#!/usr/bin/env python3
"""Golden vectors for the fixture-generator functions, produced by the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_motion.py

Imports python/generate_test_frames_natural.py of the reference, calls its
generate_smooth_synthetic and apply_motion (scipy.ndimage.shift) and writes
tests/golden/motion.npz: the texture, and for every (dx, dy) case the shifted frame.
"""
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/python")
OUT = Path(__file__).resolve().parent
CASES = [(2.0, 0.0), (0.5, 0.5), (-0.75, 1.25), (0.0, 0.0), (3.0, -2.0), (0.1, -0.9), (-20.0, 3.0), (1e-9, 0.0),
         (0.3333333, 0.6666667), (40.25, -30.75), (200.0, 0.0)]


def main():
    sys.path.insert(0, str(REF))
    import generate_test_frames_natural as g  # the reference module

    tex = g.generate_smooth_synthetic(128, 96)
    rng = np.random.default_rng(11)
    noise = rng.integers(0, 256, (37, 53)).astype(np.uint8)
    out = {"texture_128x96": tex, "noise_53x37": noise, "cases": np.asarray(CASES, dtype=np.float64)}
    for i, (dx, dy) in enumerate(CASES):
        out[f"texture_shift_{i}"] = g.apply_motion(tex, dx, dy)
        out[f"noise_shift_{i}"] = g.apply_motion(noise, dx, dy)
    # apply_motion_opencv (generate_test_suite.py:165-204) with the 13 verifier parameter sets and a few
    # extra affine maps, on both frames
    import generate_test_suite as gts  # the reference module (cv2.warpAffine inside)

    affine = [(p.dx, p.dy, p.rotation, p.scale) for p in gts.STANDARD_TEST_PATTERNS.values()] if hasattr(gts, "STANDARD_TEST_PATTERNS") else []
    if not affine:
        table = next(v for v in vars(gts).values() if isinstance(v, dict) and v and all(isinstance(x, gts.MotionParameters) for x in v.values()))
        affine = [(p.dx, p.dy, p.rotation, p.scale) for p in table.values()]
    rng2 = np.random.default_rng(5)
    affine += [(float(rng2.uniform(-8, 8)), float(rng2.uniform(-8, 8)), float(rng2.uniform(-20, 20)), float(rng2.uniform(0.8, 1.25)))
               for _ in range(7)]
    out["affine_cases"] = np.asarray(affine, dtype=np.float64)
    for i, (dx, dy, rot, sc) in enumerate(affine):
        params = gts.MotionParameters(name=f"case{i}", dx=dx, dy=dy, rotation=rot, scale=sc, description="")
        out[f"texture_affine_{i}"] = gts.apply_motion_opencv(tex, params)
        out[f"noise_affine_{i}"] = gts.apply_motion_opencv(noise, params)
    np.savez_compressed(OUT / "motion.npz", **out)
    print("wrote", OUT / "motion.npz", {k: v.shape for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main()
