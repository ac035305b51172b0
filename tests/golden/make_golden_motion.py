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
    np.savez_compressed(OUT / "motion.npz", **out)
    print("wrote", OUT / "motion.npz", {k: v.shape for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main()
