#!/usr/bin/env python3
"""The reference's acceptance loop (optical_flow_verifier.py) with every step on the B200 backend.

    python tools/device_verifier.py [--pyramidal]

Uses the committed golden fixtures (tests/golden/: the 13 patterns' base texture, the reference's
ground truth and verification_baseline.json numbers), so it runs on a box without the reference:
second frames by of_warp_affine_u8 (cv2.warpAffine in fixed point), single-scale LK on the uint8
frames (or the 3 x 3 pyramidal path), metrics by of_flow_metrics_f32, and a table against the baseline.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "optical-flow-fpga_b200"))

# python/generate_test_suite.py:59-136: dx, dy, rotation (degrees), scale
PATTERNS = {
    "translate_small": (0.5, 0.5, 0.0, 1.0), "translate_medium": (2.0, 0.0, 0.0, 1.0), "translate_large": (15.0, 0.0, 0.0, 1.0),
    "translate_vertical": (0.0, 10.0, 0.0, 1.0), "translate_diagonal": (10.0, 10.0, 0.0, 1.0), "rotate_small": (0.0, 0.0, 2.0, 1.0),
    "rotate_medium": (0.0, 0.0, 5.0, 1.0), "rotate_large": (0.0, 0.0, 15.0, 1.0), "zoom_in": (0.0, 0.0, 0.0, 1.1),
    "zoom_out": (0.0, 0.0, 0.0, 0.9), "translate_rotate": (5.0, 5.0, 3.0, 1.0), "no_motion": (0.0, 0.0, 0.0, 1.0),
    "translate_extreme": (30.0, 20.0, 0.0, 1.0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pyramidal", action="store_true", help="3-level, 3-iteration pyramidal LK instead of single scale")
    args = ap.parse_args()
    import of_b200 as ofb

    golden = ROOT / "tests" / "golden"
    index = json.load(open(golden / "golden_index.json"))
    frames = np.load(golden / "frames.npz")
    names = list(index["patterns"])
    base = frames[f"{names[0]}__0"]
    h, w = base.shape
    mats = np.stack([ofb.motion_matrix(w, h, *PATTERNS[n]) for n in names])
    first = np.repeat(base[None], len(names), axis=0)
    second = ofb.warp_affine_u8_batch(first, mats)
    same = all(np.array_equal(second[i], frames[f"{n}__1"]) for i, n in enumerate(names))
    print(f"generated frames equal the reference generator's: {same}")
    if args.pyramidal:
        u, v = ofb.lk_pyramidal_batch(first.astype(np.float32), second.astype(np.float32), 3, 5, 3, ofb.MODE_EXACT)
        method = "pyramidal"
    else:
        u, v = ofb.lk_single_scale_u8_batch(first, second, 5, ofb.MODE_FAST)
        method = "single_scale"
    print(f"{'pattern':20s} {'mae_u':>9s} {'mae_v':>9s} {'rmse':>9s} {'epe':>9s} {'aae':>9s}   max |metric - baseline|")
    worst = 0.0
    for i, n in enumerate(names):
        e = index["patterns"][n]
        m = ofb.flow_metrics_batch(u[i], v[i], e["ground_truth"]["u"], e["ground_truth"]["v"],
                                   ofb.verifier_test_region((h, w), n, index["center_crop"]))[0]
        d = max(abs(m[k] - e["verification_baseline"][method][k]) for k in ofb.METRIC_NAMES)
        worst = max(worst, d)
        print(f"{n:20s} " + " ".join(f"{m[k]:9.4f}" for k in ofb.METRIC_NAMES) + f"   {d:.2e}")
    print(f"largest deviation from python/verification_baseline.json ({method}): {worst:.2e}")
    return 0 if same and worst < 1e-4 else 1


if __name__ == "__main__":
    sys.exit(main())
