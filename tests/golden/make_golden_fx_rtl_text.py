#!/usr/bin/env python3
"""Writes tests/golden/fx_rtl_text_vectors.npz: known-answer vectors of the fixed-point datapath obtained by
EXECUTING THE REFERENCE'S RTL TEXT (rtl/unopt/gradient_compute.sv, window_accumulator.sv, flow_solver.sv under
/root/reference) with oracle/sv_eval.py.  Run in the build container:

    python tests/golden/make_golden_fx_rtl_text.py

prev / curr [n, 7, 7] uint8 neighbourhoods and the S8.7 flow (u, v) of their centre pixel; sums [m, 5] and the
solver's output for them; and one frame pair with the neighbourhoods tiled (16 per row, 7-pixel pitch, width padded
to a multiple of 16 for the TMA kernel) so that a whole-frame run can be checked at the patch centres."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from rtl_text_harness import RtlDatapath, random_patches, random_sums  # noqa: E402

RTL = Path("/root/reference/rtl/unopt")


def main():
    dp = RtlDatapath(RTL)
    rng = np.random.default_rng(8)
    patches = random_patches(rng, 320)
    prev = np.stack([p for p, _ in patches])
    curr = np.stack([c for _, c in patches])
    u = np.zeros(len(patches), np.int16)
    v = np.zeros(len(patches), np.int16)
    for k, (p, c) in enumerate(patches):
        u[k], v[k], _, _ = dp.pixel(p, c)
    sums = random_sums(rng, 1200)
    su = np.zeros(len(sums), np.int16)
    sv = np.zeros(len(sums), np.int16)
    for k in range(len(sums)):
        su[k], sv[k] = dp.solve(*sums[k])
    per_row = 16
    rows = (len(patches) + per_row - 1) // per_row
    H, W = rows * 7, per_row * 7
    W = (W + 15) // 16 * 16
    fp = np.full((H, W), 128, np.uint8)
    fc = np.full((H, W), 128, np.uint8)
    centres = np.zeros((len(patches), 2), np.int32)
    for k in range(len(patches)):
        r, c = divmod(k, per_row)
        fp[7 * r:7 * r + 7, 7 * c:7 * c + 7] = prev[k]
        fc[7 * r:7 * r + 7, 7 * c:7 * c + 7] = curr[k]
        centres[k] = (7 * r + 3, 7 * c + 3)
    out = ROOT / "tests" / "golden" / "fx_rtl_text_vectors.npz"
    np.savez_compressed(out, prev=prev, curr=curr, u=u, v=v, sums=sums, solve_u=su, solve_v=sv, frame_prev=fp, frame_curr=fc,
                        centres=centres)
    print(out, out.stat().st_size, "bytes;", int((np.abs(u) + np.abs(v) > 0).sum()), "of", len(u), "pixels solvable")


if __name__ == "__main__":
    main()
