"""Pins the fixed-point oracle (a8) mechanically: oracle/sv_eval.py executes the procedural blocks of the reference's
gradient_compute.sv / window_accumulator.sv / flow_solver.sv FROM THEIR TEXT with IEEE 1800-2017's sizing and
signedness rules, and oracle/lk_fixed_oracle.py must give the same integers on randomised vectors.

Two layers:
  * known-answer tests of the evaluator itself, taken from the examples of IEEE 1800-2017 (clauses 11.3.3, 11.6.2,
    11.7, 11.8.1) -- they run everywhere;
  * the oracle against the RTL text -- needs /root/reference (build container); on the GPU box the same vectors are
    checked from tests/golden/fx_rtl_text_vectors.npz, which tests/golden/make_golden_fx_rtl_text.py wrote from the
    RTL text."""
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import lk_fixed_oracle as fxo
from oracle.sv_eval import Module

RTL = Path("/root/reference/rtl/unopt")
needs_rtl = pytest.mark.skipif(not (RTL / "flow_solver.sv").exists(), reason="the reference's RTL is only present in the build container")


def _run(body: str, decls: str):
    m = Module(f"module t (input logic clk);\n{decls}\nalways_comb begin\n{body}\nend\nendmodule")
    m.run("always_comb", 0)
    return m


def test_evaluator_lrm_expression_bit_lengths():
    # IEEE 1800-2017 11.6.2: the intermediate carry is lost at 16 bits, kept when an integer enters the expression
    m = _run("a = 16'hFFFF; b = 16'hFFFF; answer1 = (a + b) >> 1; answer2 = (a + b + 0) >> 1;",
             "logic [15:0] a, b, answer1, answer2;")
    assert m.get("answer1") == 0x7FFF and m.get("answer2") == 0xFFFF


def test_evaluator_lrm_integer_division_examples():
    # 11.3.3 "Using integer literals in expressions": -12 / 3, -'d12 / 3, -'sd12 / 3, -4'sd12 / 3
    m = _run("i1 = -12 / 3; i2 = -'d12 / 3; i3 = -'sd12 / 3; i4 = -4'sd12 / 3;", "logic signed [31:0] i1, i2, i3, i4;")
    assert m.get("i1") == -4
    assert m.get("i2") == 1431655761
    assert m.get("i3") == -4
    assert m.get("i4") == 1  # -4'sd12 is the 4-bit value 4


def test_evaluator_lrm_signed_casts_and_shifts():
    # 11.7: regA = $unsigned(-4) -> 8'b11111100; regS = $signed(4'b1100) -> -4
    m = _run("regA = $unsigned(-4); regS = $signed(4'b1100);", "logic [7:0] regA; logic signed [7:0] regS;")
    assert m.get("regA") == 0b11111100 and m.get("regS") == -4
    # 11.4.10: >>> is arithmetic only for a signed operand; >> never
    m = _run("s = 8'sb10000000; r1 = s >>> 2; r2 = s >> 2; u = 8'b10000000; r3 = u >>> 2;",
             "logic signed [7:0] s, r1, r2; logic [7:0] u, r3;")
    assert m.get("r1") == -32 and m.get("r2", signed=False) == 0b00100000 and m.get("r3") == 0b00100000
    # 11.8.1: one unsigned operand makes the whole expression unsigned -- the signed operand is ZERO-extended
    m = _run("s = -1; u = 1; wide = s + u; cmp = (s < u);", "logic signed [3:0] s; logic [3:0] u; logic [7:0] wide; logic cmp;")
    assert m.get("wide") == 0x10 and m.get("cmp") == 0
    # part selects are unsigned, concatenations are unsigned, operands of a concatenation are self-determined
    m = _run("s = -2; r = s[7:0] + 9'sd0; c = {1'b0, s};", "logic signed [7:0] s; logic signed [11:0] r; logic signed [11:0] c;")
    assert m.get("r") == 254 and m.get("c") == 254


def test_evaluator_truncating_division_and_assignment_context():
    # the right-hand side is evaluated at the width of the wider of the two sides, then truncated (10.7, 11.8.2)
    m = _run("a = 39'sd100000000000; b = -32'sd7; q = a / b; n = -39'sd100000000000; q2 = n / b; q3 = n / 32'sd7;",
             "logic signed [38:0] a, n; logic signed [31:0] b; logic signed [15:0] q, q2, q3;")
    full = -(100000000000 // 7)
    wrap = lambda x: ((x + (1 << 15)) & 0xFFFF) - (1 << 15)
    assert m.get("q") == wrap(full) and m.get("q2") == wrap(-full) and m.get("q3") == wrap(full)


@needs_rtl
def test_rtl_modules_parse_with_the_expected_shapes():
    from rtl_text_harness import RtlDatapath

    dp = RtlDatapath(RTL)
    g, a, s = dp.grad, dp.acc, dp.sol
    assert (g.vars["window_curr"].width, g.vars["window_curr"].signed, g.vars["window_curr"].dims) == (8, True, (3, 3))
    assert (g.vars["sobel_x_comb"].width, g.vars["sobel_x_comb"].signed) == (12, True)
    assert a.vars["prod_IxIx_pipe"].width == 24 and a.vars["accum_IxIx"].width == 32
    assert s.vars["prod_det1"].width == 64 and s.params["DET_THRESHOLD"] == 1000 and s.vars["flow_u_comb"].width == 16


@needs_rtl
def test_frame_average_quirk_falls_out_of_the_text():
    """gradient_compute.sv:116 in 9 bits with sign-extended operands and a logical shift: pixels on opposite sides of
    128 average to a value off by 128 -- found here by executing the text, not by reading it."""
    from rtl_text_harness import RtlDatapath

    dp = RtlDatapath(RTL)
    prev = np.full((3, 3), 130, np.uint8)
    curr = np.full((3, 3), 126, np.uint8)
    dp.gradients(prev, curr)
    avg = dp.grad.last_locals["avg_window"]
    got = avg.get((1, 1))
    assert got == ((130 - 256 + 126) & 0x1FF) >> 1 == 0  # intended: 128
    assert got == int(fxo.average_frame(prev, curr, True)[1, 1])
    assert int(fxo.average_frame(prev, curr, False)[1, 1]) == 128


@needs_rtl
def test_oracle_equals_rtl_text_on_random_neighbourhoods():
    from rtl_text_harness import RtlDatapath, random_patches

    dp = RtlDatapath(RTL)
    rng = np.random.default_rng(20261019)
    n_bad = 0
    for k, (p, c) in enumerate(random_patches(rng, 400)):
        u, v, sums, (ix, iy, it) = dp.pixel(p, c)
        gx, gy, gt = fxo.gradients_fx(p, c, True)
        assert np.array_equal(gx[1:6, 1:6], ix) and np.array_equal(gy[1:6, 1:6], iy) and np.array_equal(gt[1:6, 1:6], it), k
        want = [int((a_ * b_).sum()) for a_, b_ in ((ix, ix), (iy, iy), (ix, iy), (ix, it), (iy, it))]
        assert list(sums) == want, (k, sums, want)
        uo, vo = fxo.lk_single_scale_fx(p, c, True)
        us, vs = fxo.lk_single_scale_fx_scalar(p, c, True)
        assert (int(uo[3, 3]), int(vo[3, 3])) == (u, v) == (int(us[3, 3]), int(vs[3, 3])), (k, u, v, uo[3, 3], vo[3, 3])
        n_bad += (u, v) == (0, 0)
    assert n_bad < 300  # most vectors are solvable: the comparison is not vacuous


@needs_rtl
def test_oracle_solver_equals_rtl_text_on_random_sums():
    from rtl_text_harness import RtlDatapath, random_sums

    dp = RtlDatapath(RTL)
    rng = np.random.default_rng(7)
    sums = random_sums(rng, 3000)
    uo, vo = fxo.solve_fx(*[sums[:, i] for i in range(5)])
    clamped = wrapped = 0
    for k in range(len(sums)):
        u, v = dp.solve(*sums[k])
        assert (u, v) == (int(uo[k]), int(vo[k])), (k, sums[k].tolist(), u, v, int(uo[k]), int(vo[k]))
        clamped += abs(u) == 1024 or abs(v) == 1024
    assert clamped > 60


def test_oracle_equals_committed_rtl_text_vectors():
    """The same comparison from the committed vectors (written from the RTL text by make_golden_fx_rtl_text.py):
    runs on the GPU box too, where the reference is absent."""
    z = np.load(GOLDEN / "fx_rtl_text_vectors.npz")
    prev, curr, u, v = z["prev"], z["curr"], z["u"], z["v"]
    for k in range(prev.shape[0]):
        uo, vo = fxo.lk_single_scale_fx(prev[k], curr[k], True)
        assert (int(uo[3, 3]), int(vo[3, 3])) == (int(u[k]), int(v[k])), k
    su, sv = fxo.solve_fx(*[z["sums"][:, i] for i in range(5)])
    assert np.array_equal(su, z["solve_u"]) and np.array_equal(sv, z["solve_v"])
    # the frame the GPU test feeds the kernels: the patches tiled, flow at the patch centres
    uo, vo = fxo.lk_single_scale_fx(z["frame_prev"], z["frame_curr"], True)
    cy, cx = z["centres"][:, 0], z["centres"][:, 1]
    assert np.array_equal(uo[cy, cx], u) and np.array_equal(vo[cy, cx], v)
