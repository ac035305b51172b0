"""The oracle against the REFERENCE ITSELF on randomised inputs -- build container only.

tests/golden/ pins the oracle on committed outputs of the reference; this file widens that pin where the reference
can be imported (/root/reference/python, matplotlib stubbed as in tests/golden/make_golden.py): random shapes (odd,
tiny, wider than tall), general float frames, flow of both signs, tiny flow (where the reference's float64
coordinate sum is itself rounded), flow that leaves the frame.  Everything is compared bit for bit.  On the GPU box
(no /root/reference) the whole file is skipped; nothing else reads the reference at run time.
"""

import contextlib
import io
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

from oracle import lk_float_oracle as orc

REF = Path("/root/reference/python")
pytestmark = pytest.mark.skipif(not (REF / "lucas_kanade_core.py").exists(), reason="the reference is not on this machine")
f32 = np.float32


@pytest.fixture(scope="module")
def ref():
    """The reference's lucas_kanade_core / lucas_kanade_pyramidal.  The backend's drop-in modules carry the same
    flat names (that is the point of a drop-in), so whatever other tests imported is set aside while the reference's
    files are loaded, and put back afterwards."""
    sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
    import make_golden

    names = ("lucas_kanade_core", "lucas_kanade_pyramidal", "lucas_kanade_reference")
    saved = {n: sys.modules.pop(n) for n in names if n in sys.modules}
    saved_path = list(sys.path)
    try:
        make_golden._install_matplotlib_stub()
        sys.path.insert(0, str(REF))
        import lucas_kanade_core as core
        import lucas_kanade_pyramidal as pyr
    finally:
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
        sys.path[:] = saved_path
    assert Path(core.__file__).resolve().parent == REF and Path(pyr.__file__).resolve().parent == REF
    assert pyr.lucas_kanade_single_scale is core.lucas_kanade_single_scale
    return core, pyr


def bits(a):
    return np.ascontiguousarray(a, dtype=f32).view(np.uint32)


@contextlib.contextmanager
def quiet_in_scratch_dir():
    """The reference prints per iteration and writes python/output/*.png relative to the cwd."""
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory(prefix="ofref_") as scratch:
        os.chdir(scratch)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                yield
        finally:
            os.chdir(cwd)


SHAPES = [(7, 9), (16, 16), (23, 40), (41, 17), (30, 64), (65, 33)]


def frames(shape, rng, kind):
    if kind == "uint8":
        return rng.integers(0, 256, shape).astype(f32), rng.integers(0, 256, shape).astype(f32)
    p = (rng.standard_normal(shape) * 40 + 100).astype(f32)
    return p, (p + rng.standard_normal(shape) * 2).astype(f32)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", ["uint8", "float"])
def test_gradients_and_single_scale(ref, shape, kind):
    core, _ = ref
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    p, c = frames(shape, rng, kind)
    for got, want in zip(orc.compute_gradients(p, c), core.compute_gradients(p, c)):
        assert got.dtype == want.dtype and np.array_equal(bits(got), bits(want))
    for w in (3, 5):
        with quiet_in_scratch_dir():
            want = core.lucas_kanade_single_scale(p, c, window_size=w)
        got = orc.lucas_kanade_single_scale(p, c, w)
        assert np.array_equal(bits(got[0]), bits(want[0])) and np.array_equal(bits(got[1]), bits(want[1])), w


@pytest.mark.parametrize("shape", SHAPES)
def test_warp_image_every_kind_of_flow(ref, shape):
    _, pyr = ref
    rng = np.random.default_rng(shape[1])
    img = (rng.random(shape) * 255).astype(f32)
    flows = {
        "moderate": (rng.standard_normal(shape) * 1.5, rng.standard_normal(shape) * 1.5),
        "leaves_the_frame": (rng.standard_normal(shape) * 30, rng.standard_normal(shape) * 30),
        "tiny": (rng.standard_normal(shape) * 1e-6, rng.standard_normal(shape) * 1e-9),
        "integer": (rng.integers(-3, 4, shape), rng.integers(-3, 4, shape)),
        "zero": (np.zeros(shape), np.zeros(shape)),
    }
    for name, (u, v) in flows.items():
        u, v = u.astype(f32), v.astype(f32)
        want = pyr.warp_image(img, u, v)
        got = orc.warp_image(img, u, v)
        assert got.dtype == want.dtype and np.array_equal(bits(got), bits(want)), name


@pytest.mark.parametrize("shape", [(16, 16), (23, 40), (41, 17), (65, 33), (64, 48)])
def test_pyramid_and_upsample(ref, shape):
    _, pyr = ref
    rng = np.random.default_rng(shape[0])
    img = (rng.random(shape) * 255).astype(f32)
    levels = 3 if min(shape) >= 16 else 2
    want = pyr.build_gaussian_pyramid(img, levels)
    got = orc.build_gaussian_pyramid(img, levels)
    assert [g.shape for g in got] == [w.shape for w in want]
    for k, (g, w) in enumerate(zip(got, want)):
        assert g.dtype == w.dtype and np.array_equal(bits(g), bits(w)), f"level {k}"
    cu = (rng.standard_normal(want[0].shape) * 2).astype(f32)
    cv = (rng.standard_normal(want[0].shape) * 2).astype(f32)
    for target in (want[1].shape, shape, (shape[0] + 3, shape[1] + 5)):
        wu, wv = pyr.upsample_flow(cu, cv, target)
        gu, gv = orc.upsample_flow(cu, cv, target)
        assert np.array_equal(bits(gu), bits(wu)) and np.array_equal(bits(gv), bits(wv)), target


@pytest.mark.parametrize("shape,levels,iters", [((24, 32), 2, 2), ((33, 47), 2, 3), ((40, 40), 3, 2)])
def test_small_pyramidal_runs(ref, shape, levels, iters):
    from scipy.ndimage import gaussian_filter, shift

    _, pyr = ref
    rng = np.random.default_rng(shape[0] + levels)
    p = gaussian_filter((rng.random(shape) * 255).astype(f32), 1.2).astype(f32)
    c = shift(p, (0.6, -0.8), order=1, mode="nearest").astype(f32)
    with quiet_in_scratch_dir():
        want = pyr.lucas_kanade_pyramidal(p, c, num_levels=levels, window_size=5, num_iterations=iters)
    got = orc.lucas_kanade_pyramidal(p, c, levels, 5, iters)
    assert np.array_equal(bits(got[0]), bits(want[0])) and np.array_equal(bits(got[1]), bits(want[1]))
