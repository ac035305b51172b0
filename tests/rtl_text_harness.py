"""Drives oracle/sv_eval.py over the reference's three datapath modules (test infrastructure).

`RtlDatapath(rtl_dir)` parses gradient_compute.sv, window_accumulator.sv and flow_solver.sv from their text and
chains their procedural blocks for ONE output pixel: 3x3 pixel windows -> gradients (gradient_compute.sv:108-141),
5x5 gradient windows -> five sums (window_accumulator.sv:112-167), sums -> S8.7 flow (flow_solver.sv:83-149).
The line buffers (module instantiations) are not evaluated: the windows they would deliver are set by the caller,
in the orientation rtl/common/line_buffer_5x5.sv:106-140 gives them (window[i][j]: i = row, oldest = top first,
j = column, oldest = left first)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from oracle.sv_eval import Module


class RtlDatapath:
    def __init__(self, rtl_dir):
        rtl_dir = Path(rtl_dir)
        self.grad = Module((rtl_dir / "gradient_compute.sv").read_text())
        self.acc = Module((rtl_dir / "window_accumulator.sv").read_text())
        self.sol = Module((rtl_dir / "flow_solver.sv").read_text())
        # the always_comb block that holds the Sobel / temporal expressions is the second one of the module
        kinds = [k for k, _ in self.grad.blocks]
        assert kinds == ["always_comb", "always_comb", "always_ff"], kinds
        assert [k for k, _ in self.acc.blocks] == ["always_ff", "always_comb", "always_ff"]
        assert [k for k, _ in self.sol.blocks] == ["always_ff", "always_comb", "always_ff"]

    def gradients(self, prev3, curr3):
        """3x3 uint8 windows (row-major, [row][col]) -> (Ix, Iy, It) as signed integers"""
        g = self.grad
        for i in range(3):
            for j in range(3):
                g.set("window_curr", int(curr3[i][j]), i, j)
                g.set("window_prev", int(prev3[i][j]), i, j)
        g.run("always_comb", 1)
        return g.get("sobel_x_comb"), g.get("sobel_y_comb"), g.get("temporal_comb")

    def sums(self, ix5, iy5, it5):
        """5x5 gradient windows -> (sum_IxIx, sum_IyIy, sum_IxIy, sum_IxIt, sum_IyIt): stage 1 (always_ff:
        the 125 products) then stage 2 (always_comb: the adder chains)"""
        a = self.acc
        a.set("rst_n", 1)
        for i in range(5):
            for j in range(5):
                a.set("window_Ix", int(ix5[i][j]), i, j)
                a.set("window_Iy", int(iy5[i][j]), i, j)
                a.set("window_It", int(it5[i][j]), i, j)
        a.run("always_ff", 0)
        a.run("always_comb", 0)
        return tuple(a.get(n) for n in ("accum_IxIx", "accum_IyIy", "accum_IxIy", "accum_IxIt", "accum_IyIt"))

    def solve(self, sxx, syy, sxy, sxt, syt):
        """five 32-bit sums -> (flow_u_comb, flow_v_comb): stage 1 (always_ff: the six 64-bit products), stage 2
        (always_comb: determinant, numerators, threshold, division, clamp)"""
        s = self.sol
        s.set("rst_n", 1)
        for n, val in (("sum_IxIx", sxx), ("sum_IyIy", syy), ("sum_IxIy", sxy), ("sum_IxIt", sxt), ("sum_IyIt", syt)):
            s.set(n, int(val))
        s.run("always_ff", 0)
        s.run("always_comb", 0)
        return s.get("flow_u_comb"), s.get("flow_v_comb")

    def pixel(self, prev7, curr7):
        """7x7 uint8 neighbourhoods of one output pixel -> (u, v, sums, gradient windows)"""
        ix = np.zeros((5, 5), np.int64)
        iy = np.zeros((5, 5), np.int64)
        it = np.zeros((5, 5), np.int64)
        for i in range(5):
            for j in range(5):
                ix[i, j], iy[i, j], it[i, j] = self.gradients(prev7[i:i + 3, j:j + 3], curr7[i:i + 3, j:j + 3])
        sums = self.sums(ix, iy, it)
        u, v = self.solve(*sums)
        return u, v, sums, (ix, iy, it)


def random_patches(rng, n):
    """n pairs of 7x7 uint8 neighbourhoods: smooth textures with sub-pixel-like differences, pure noise, values
    straddling 128 (the 9-bit signed average), extremes."""
    out = []
    yy, xx = np.mgrid[0:7, 0:7].astype(np.float64)
    for k in range(n):
        kind = k % 5
        if kind == 0:  # uniform noise
            p = rng.integers(0, 256, (7, 7))
            c = rng.integers(0, 256, (7, 7))
        elif kind == 1:  # smooth ramp + small noise, curr = shifted
            a, b, o = rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(40, 215)
            dx, dy = rng.uniform(-1, 1, 2)
            p = o + a * xx + b * yy + rng.normal(0, 2, (7, 7))
            c = o + a * (xx - dx) + b * (yy - dy) + rng.normal(0, 2, (7, 7))
        elif kind == 2:  # around 128: the sign-extension quirk fires on mixed pairs
            p = 128 + rng.integers(-6, 7, (7, 7))
            c = 128 + rng.integers(-6, 7, (7, 7))
        elif kind == 3:  # extremes
            p = rng.choice([0, 1, 127, 128, 254, 255], (7, 7))
            c = rng.choice([0, 1, 127, 128, 254, 255], (7, 7))
        else:  # textured: large gradients, large sums (32-bit truncation of the 64-bit products)
            f = rng.uniform(0.5, 2.5)
            ph = rng.uniform(0, 6.28)
            dx = rng.uniform(-1.5, 1.5)
            p = 128 + 120 * np.sin(f * xx + ph) * np.cos(0.7 * f * yy)
            c = 128 + 120 * np.sin(f * (xx - dx) + ph) * np.cos(0.7 * f * yy)
        out.append((np.clip(np.rint(p), 0, 255).astype(np.uint8), np.clip(np.rint(c), 0, 255).astype(np.uint8)))
    return out


def random_sums(rng, n):
    """n x 5 signed 32-bit sums for the solver alone: realistic magnitudes, values whose 64-bit products do not fit
    32 bits, determinants around the +-1000 threshold, quotients beyond +-8 px and beyond 16 bits."""
    out = np.zeros((n, 5), np.int64)
    for k in range(n):
        kind = k % 6
        if kind == 0:
            out[k] = rng.integers(-(1 << 31), 1 << 31, 5)
        elif kind == 1:
            sxx, syy = rng.integers(0, 400000, 2)
            sxy = rng.integers(-300000, 300000)
            out[k] = [sxx, syy, sxy, rng.integers(-800000, 800000), rng.integers(-800000, 800000)]
        elif kind == 2:  # det within a few counts of the threshold
            sxx = rng.integers(1, 60)
            syy = rng.integers(1, 60)
            target = rng.choice([-1001, -1000, -999, 999, 1000, 1001])
            sxy = int(np.sqrt(max(0, sxx * syy - target)))
            out[k] = [sxx, syy, sxy, rng.integers(-5000, 5000), rng.integers(-5000, 5000)]
        elif kind == 3:  # small det, big numerators: quotient beyond the clamp and beyond 16 bits
            out[k] = [rng.integers(30, 80), rng.integers(30, 80), rng.integers(-10, 10), rng.integers(-(1 << 24), 1 << 24),
                      rng.integers(-(1 << 24), 1 << 24)]
        elif kind == 4:  # products around 2^31 .. 2^33
            out[k] = rng.integers(-(1 << 17), 1 << 17, 5)
        else:
            out[k] = rng.integers(-3000, 3000, 5)
    return out
