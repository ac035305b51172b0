"""CPU oracle for the verifier's error metrics -- TEST INFRASTRUCTURE ONLY.

Restates ``compute_all_metrics`` (reference ``python/flow_metrics.py:166-201``)
and ``get_test_region_mask`` (``python/optical_flow_verifier.py:96-138``) so the
GPU box (which has no ``/root/reference``) can recompute MAE / RMSE / EPE / AAE
from a flow field and compare them with ``python/verification_baseline.json``
(copied into ``tests/golden/golden_index.json`` by ``make_golden.py``, which also
checks this file against the reference's own functions).
"""

from __future__ import annotations

import numpy as np


def test_region_mask(shape, pattern_name: str, center_crop_size: int = 80) -> np.ndarray:
    """Translation patterns: frame minus a 10 px border.  Rotation / zoom /
    combined: central crop (optical_flow_verifier.py:115-136)."""
    h, w = shape
    mask = np.zeros((h, w), dtype=bool)
    if "rotate" in pattern_name or "zoom" in pattern_name:
        cy, cx = h // 2, w // 2
        half = center_crop_size // 2
        mask[cy - half : cy + half, cx - half : cx + half] = True
    else:
        mask[10:-10, 10:-10] = True
    return mask


def all_metrics(u, v, u_true: float, v_true: float, mask=None) -> dict:
    """mae_u, mae_v, rmse, epe, aae with the reference's float32 semantics."""
    if mask is None:
        mask = np.ones_like(u, dtype=bool)
    up = u[mask]
    vp = v[mask]
    eu = up - u_true  # float32 array minus Python float stays float32
    ev = vp - v_true
    mae_u = float(np.mean(np.abs(eu)))
    mae_v = float(np.mean(np.abs(ev)))
    sq = eu**2 + ev**2
    rmse = float(np.sqrt(np.mean(sq)))
    epe = float(np.mean(np.sqrt(sq)))
    # angular error between (u, v, 1) vectors, degrees (flow_metrics.py:108-163)
    mag_true = np.sqrt(u_true**2 + v_true**2)
    mag_pred = np.sqrt(up**2 + vp**2)
    if mag_true < 1e-6 and np.all(mag_pred < 1e-6):
        aae = 0.0
    else:
        ones = np.ones_like(up)
        ut = np.full_like(up, u_true)
        vt = np.full_like(vp, v_true)
        norm_pred = np.sqrt(up**2 + vp**2 + ones**2)
        norm_true = np.sqrt(ut**2 + vt**2 + ones**2)
        dot = (up * ut + vp * vt + ones * ones) / (norm_pred * norm_true)
        dot = np.clip(dot, -1.0, 1.0)
        aae = float(np.mean(np.rad2deg(np.arccos(dot))))
    return {"mae_u": mae_u, "mae_v": mae_v, "rmse": rmse, "epe": epe, "aae": aae}
