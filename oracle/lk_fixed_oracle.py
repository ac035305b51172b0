"""CPU oracle for the fixed-point (RTL-mirroring) LK mode -- TEST INFRASTRUCTURE ONLY.

Integer restatement of the reference's single-scale RTL datapath
(``rtl/unopt/gradient_compute.sv``, ``window_accumulator.sv``, ``flow_solver.sv``)
on geometrically correct 3x3 / 5x5 windows.  Only ``tests/``, ``smoke()`` and
``bench.py``'s CPU-baseline leg import it.

PARITY: PINNED BY THE RTL TEXT, NOT BY A SIMULATOR RUN ("parity unpinned" in the task's strict sense).
The reference holds no runnable golden for this datapath: there is no SystemVerilog simulator in the
build image, and the only known-answer data (the XSim log pasted in the reference's
``README.md:455-532``) predates the committed RTL and cannot be reproduced from it.
What pins the arithmetic below is mechanical all the same: ``oracle/sv_eval.py`` parses
``gradient_compute.sv`` / ``window_accumulator.sv`` / ``flow_solver.sv`` and EXECUTES their
``always_comb`` / ``always_ff`` bodies from the source text under IEEE 1800-2017's expression sizing and
signedness rules (checked on the standard's own examples), and ``tests/test_fixed_oracle_vs_rtl_text.py``
diffs this file against it on randomised neighbourhoods and solver inputs -- every integer equal; the
vectors are committed as ``tests/golden/fx_rtl_text_vectors.npz`` and the GPU kernels are checked against
them.  Until an XSim / Verilator dump of the committed RTL is diffed (``of_export_flow_fx_txt`` writes the
testbench's format for that), the line-by-line reading that the evaluator reproduces is:

* pixels sit in ``logic signed [7:0]`` windows; ``avg = (curr + prev) >> 1`` is
  evaluated in 9 bits with both operands SIGN-extended and a LOGICAL shift
  (``gradient_compute.sv:47-48,109,116``), so avg = ((s8(c)+s8(p)) mod 512) >> 1.
  That equals floor((c+p)/2) when both pixels are on the same side of 128 and is
  off by 128 otherwise.  ``mirror_avg_quirk=False`` gives the intended
  floor((c+p)/2) instead.
* Sobel in correlation form on the zero-extended averages, arithmetic ``>>> 3``
  (``:121-136``); It = prev - curr on the zero-extended centre pixels (``:139``).
* 12x12 -> 24-bit products, 25-term 32-bit sums (``window_accumulator.sv:128-167``).
* 64-bit products, LOW 32 BITS kept, subtracted with 32-bit wrap
  (``flow_solver.sv:83-92,116-120``); solvable iff det > 1000 or det < -1000
  (``:45,123``); quotient of (num <<< 7) / det truncated toward zero in 39 bits,
  low 16 bits kept, then clamped to +-1024 = +-8 px in S8.7 (``:113,127-144``).

Geometry (our definition; the RTL's streaming quirks of SURVEY.md App. B.2 are
not mirrored): gradients exist for 1 <= y <= H-2, 1 <= x <= W-2; flow for
3 <= y <= H-4, 3 <= x <= W-4; everything else is 0.  Output int16, real = int/128.
"""

from __future__ import annotations

import numpy as np

DET_THRESHOLD = 1000
FRAC_BITS = 7
CLAMP = 1024


def _s8(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.int64)
    return np.where(x >= 128, x - 256, x)


def _wrap32(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.int64) & 0xFFFFFFFF
    return np.where(x >= 2**31, x - 2**32, x)


def _wrap16(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.int64) & 0xFFFF
    return np.where(x >= 2**15, x - 2**16, x)


def average_frame(prev_u8: np.ndarray, curr_u8: np.ndarray, mirror_avg_quirk: bool = True):
    if mirror_avg_quirk:
        return ((_s8(curr_u8) + _s8(prev_u8)) & 0x1FF) >> 1
    return (curr_u8.astype(np.int64) + prev_u8.astype(np.int64)) >> 1


def gradients_fx(prev_u8: np.ndarray, curr_u8: np.ndarray, mirror_avg_quirk: bool = True):
    """Integer Ix, Iy, It (int64 arrays, zero on the 1-pixel frame border)."""
    prev_u8 = np.asarray(prev_u8, dtype=np.uint8)
    curr_u8 = np.asarray(curr_u8, dtype=np.uint8)
    h, w = prev_u8.shape
    a = average_frame(prev_u8, curr_u8, mirror_avg_quirk)
    ix = np.zeros((h, w), dtype=np.int64)
    iy = np.zeros((h, w), dtype=np.int64)
    it = np.zeros((h, w), dtype=np.int64)
    if h < 3 or w < 3:
        return ix, iy, it
    tl, tc, tr = a[:-2, :-2], a[:-2, 1:-1], a[:-2, 2:]
    ml, mr = a[1:-1, :-2], a[1:-1, 2:]
    bl, bc, br = a[2:, :-2], a[2:, 1:-1], a[2:, 2:]
    left = -tl - 2 * ml - bl
    right = tr + 2 * mr + br
    top = -tl - 2 * tc - tr
    bottom = bl + 2 * bc + br
    ix[1:-1, 1:-1] = (left + right) >> 3  # arithmetic shift = floor division
    iy[1:-1, 1:-1] = (top + bottom) >> 3
    it[1:-1, 1:-1] = prev_u8[1:-1, 1:-1].astype(np.int64) - curr_u8[1:-1, 1:-1].astype(np.int64)
    return ix, iy, it


def _trunc_div(num: np.ndarray, den: np.ndarray) -> np.ndarray:
    q = np.abs(num) // np.abs(den)
    return np.where((num < 0) != (den < 0), -q, q)


def solve_fx(sxx, syy, sxy, sxt, syt):
    """flow_solver.sv arithmetic on int64 arrays of 32-bit sums -> int16 S8.7."""
    det = _wrap32(_wrap32(sxx * syy) - _wrap32(sxy * sxy))
    nu = _wrap32(_wrap32(syy * sxt) - _wrap32(sxy * syt))
    nv = _wrap32(_wrap32(sxx * syt) - _wrap32(sxy * sxt))
    ok = (det > DET_THRESHOLD) | (det < -DET_THRESHOLD)
    safe = np.where(ok, det, 1)
    qu = _wrap16(_trunc_div(nu << FRAC_BITS, safe))
    qv = _wrap16(_trunc_div(nv << FRAC_BITS, safe))
    qu = np.clip(qu, -CLAMP, CLAMP)
    qv = np.clip(qv, -CLAMP, CLAMP)
    return np.where(ok, qu, 0).astype(np.int16), np.where(ok, qv, 0).astype(np.int16)


def lk_single_scale_fx(prev_u8, curr_u8, mirror_avg_quirk: bool = True):
    """uint8 frame pair -> (u, v) int16 S8.7, window 5x5."""
    prev_u8 = np.asarray(prev_u8, dtype=np.uint8)
    curr_u8 = np.asarray(curr_u8, dtype=np.uint8)
    h, w = prev_u8.shape
    u = np.zeros((h, w), dtype=np.int16)
    v = np.zeros((h, w), dtype=np.int16)
    if h < 7 or w < 7:
        return u, v
    ix, iy, it = gradients_fx(prev_u8, curr_u8, mirror_avg_quirk)
    oh, ow = h - 6, w - 6

    def box(p):
        s = np.zeros((oh, ow), dtype=np.int64)
        for i in range(5):
            for j in range(5):
                s += p[1 + i : 1 + i + oh, 1 + j : 1 + j + ow]
        return s

    sums = [box(p) for p in (ix * ix, iy * iy, ix * iy, ix * it, iy * it)]
    ui, vi = solve_fx(*sums)
    u[3 : h - 3, 3 : w - 3] = ui
    v[3 : h - 3, 3 : w - 3] = vi
    return u, v


def lk_single_scale_fx_scalar(prev_u8, curr_u8, mirror_avg_quirk: bool = True):
    """Per-pixel Python-int restatement (arbitrary precision, explicit bit widths).
    Small inputs only; pins the vectorised version above."""
    prev_u8 = np.asarray(prev_u8, dtype=np.uint8)
    curr_u8 = np.asarray(curr_u8, dtype=np.uint8)
    h, w = prev_u8.shape
    P = [[int(x) for x in row] for row in prev_u8]
    C = [[int(x) for x in row] for row in curr_u8]

    def s8(x):
        return x - 256 if x >= 128 else x

    def wrap(x, bits):
        x &= (1 << bits) - 1
        return x - (1 << bits) if x >> (bits - 1) else x

    def avg(y, x):
        if mirror_avg_quirk:
            return ((s8(C[y][x]) + s8(P[y][x])) & 0x1FF) >> 1
        return (C[y][x] + P[y][x]) >> 1

    def grad(y, x):
        a = [[avg(y + i - 1, x + j - 1) for j in range(3)] for i in range(3)]
        left = -a[0][0] - (a[1][0] << 1) - a[2][0]
        right = a[0][2] + (a[1][2] << 1) + a[2][2]
        top = -a[0][0] - (a[0][1] << 1) - a[0][2]
        bottom = a[2][0] + (a[2][1] << 1) + a[2][2]
        return (left + right) >> 3, (top + bottom) >> 3, P[y][x] - C[y][x]

    u = np.zeros((h, w), dtype=np.int16)
    v = np.zeros((h, w), dtype=np.int16)
    for y in range(3, h - 3):
        for x in range(3, w - 3):
            sxx = syy = sxy = sxt = syt = 0
            for i in range(-2, 3):
                for j in range(-2, 3):
                    gx, gy, gt = grad(y + i, x + j)
                    sxx += gx * gx
                    syy += gy * gy
                    sxy += gx * gy
                    sxt += gx * gt
                    syt += gy * gt
            det = wrap(wrap(sxx * syy, 32) - wrap(sxy * sxy, 32), 32)
            nu = wrap(wrap(syy * sxt, 32) - wrap(sxy * syt, 32), 32)
            nv = wrap(wrap(sxx * syt, 32) - wrap(sxy * sxt, 32), 32)
            if det > DET_THRESHOLD or det < -DET_THRESHOLD:
                def q(num):
                    n = num << FRAC_BITS
                    mag = abs(n) // abs(det)
                    r = -mag if (n < 0) != (det < 0) else mag
                    r = wrap(r, 16)
                    return max(-CLAMP, min(CLAMP, r))

                u[y, x] = q(nu)
                v[y, x] = q(nv)
    return u, v
