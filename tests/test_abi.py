"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/of_b200.h declares, the ctypes table covers them, the drop-in modules keep the
reference's signatures, and nothing falls back to CPU arithmetic when no GPU is present."""

import inspect
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import BACKEND_DIR, ROOT


@pytest.fixture(scope="module")
def of_b200():
    sys.path.insert(0, str(BACKEND_DIR))
    import build as of_build  # optical-flow-fpga_b200/build.py

    of_build.build()
    import of_b200 as mod

    return mod


def declared_symbols():
    text = (ROOT / "include" / "of_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(of_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(of_b200):
    names = declared_symbols()
    assert len(names) >= 19
    lib = of_b200.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/of_b200.h but not exported"
    assert sorted(of_b200.SIGNATURES) == names
    out = subprocess.run(["nm", "-D", "--defined-only", str(of_b200.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(of_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_version_and_error_string(of_b200):
    assert of_b200.lib().of_version() == 100
    assert isinstance(of_b200.lib().of_last_error(), bytes)


def test_library_contains_sm100a_tma_code(of_b200):
    sass = subprocess.run(["cuobjdump", "-sass", str(of_b200.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTMALDG" in sass, "the fast kernel must stage frames with TMA"


def test_argument_errors_come_before_any_device_work(of_b200):
    z = np.zeros((16, 16), np.float32)
    with pytest.raises(ValueError):
        of_b200.lk_single_scale(z, z, window_size=4)
    with pytest.raises(ValueError):
        of_b200.lk_single_scale(z, np.zeros((16, 20), np.float32))
    with pytest.raises(ValueError):
        of_b200.lk_single_scale(np.zeros((4, 4, 4), np.float32), np.zeros((4, 4, 4), np.float32))


def test_caller_supplied_result_buffers_are_validated(of_b200):
    """out=(u, v) of the batched calls reaches the C library as raw pointers: anything but two distinct writable
    C-contiguous arrays of the batch's shape and float32 must be refused before the call."""
    p = np.zeros((2, 16, 20), np.float32)
    good = (np.empty_like(p), np.empty_like(p))
    cases = {
        "wrong dtype": (np.empty(p.shape, np.float64), good[1]),
        "wrong shape": (np.empty((2, 16, 19), np.float32), good[1]),
        "not contiguous": (np.empty((2, 16, 40), np.float32)[:, :, ::2], good[1]),
        "same array twice": (good[0], good[0]),
        "overlapping views": (np.empty((3, 16, 20), np.float32)[:2], None),
        "read-only": (np.broadcast_to(np.float32(0), p.shape), good[1]),
        "not a pair": (good[0],),
    }
    base = np.empty((3, 16, 20), np.float32)
    cases["overlapping views"] = (base[:2], base[1:])
    for name, out in cases.items():
        for call in (lambda o: of_b200.lk_single_scale_batch(p, p, 5, of_b200.MODE_FAST, out=o),
                     lambda o: of_b200.lk_pyramidal_batch(p, p, 2, 5, 2, of_b200.MODE_FAST, out=o),
                     lambda o: of_b200.lk_single_scale_u8_batch(p.astype(np.uint8), p.astype(np.uint8), 5, of_b200.MODE_FAST, out=o)):
            with pytest.raises(ValueError):
                call(out)


def test_no_cpu_fallback_without_a_device(of_b200):
    if of_b200.device_count() > 0:
        pytest.skip("a CUDA device is present")
    z = np.zeros((16, 16), np.float32)
    with pytest.raises(of_b200.OFBackendError):
        of_b200.lk_single_scale(z, z)
    with pytest.raises(of_b200.OFBackendError):
        of_b200.lk_pyramidal(z, z)
    with pytest.raises(of_b200.OFBackendError):
        of_b200.lk_single_scale_fx(np.zeros((16, 16), np.uint8), np.zeros((16, 16), np.uint8))


def test_product_never_imports_the_oracle():
    for py in BACKEND_DIR.glob("*.py"):
        src = py.read_text()
        assert "oracle" not in src.replace("oracle/", ""), f"{py.name} mentions the oracle"
    for cu in (BACKEND_DIR / "csrc").iterdir():
        assert "oracle" not in cu.read_text()


REFERENCE_SIGNATURES = {
    # python/lucas_kanade_core.py:15,48,73
    ("lucas_kanade_core", "compute_gradients"): ["frame_prev", "frame_curr"],
    ("lucas_kanade_core", "lucas_kanade_single_scale"): ["frame_prev", "frame_curr", ("window_size", 5)],
    ("lucas_kanade_core", "lucas_kanade_from_gradients"): ["Ix", "Iy", "It", ("window_size", 5)],
    # python/lucas_kanade_pyramidal.py:23,66,100,141,231,313,354
    ("lucas_kanade_pyramidal", "build_gaussian_pyramid"): ["image", "num_levels", ("scale_factor", 0.5)],
    ("lucas_kanade_pyramidal", "warp_image"): ["image", "flow_u", "flow_v"],
    ("lucas_kanade_pyramidal", "upsample_flow"): ["flow_u", "flow_v", "target_shape"],
    ("lucas_kanade_pyramidal", "lucas_kanade_pyramidal"): [
        "frame_prev", "frame_curr", ("num_levels", 3), ("window_size", 5), ("num_iterations", 3)],
    ("lucas_kanade_pyramidal", "visualize_flow_comparison"): [
        "flow_u_single", "flow_v_single", "flow_u_pyr", "flow_v_pyr", "output_path", ("scale", 1.0)],
    ("lucas_kanade_pyramidal", "visualize_pyramid_level"): [
        "flow_u", "flow_v", "level", ("num_levels", 3), ("output_dir", "python/output")],
    ("lucas_kanade_pyramidal", "main"): [],
    # python/lucas_kanade_reference.py:22,78,106
    ("lucas_kanade_reference", "visualize_flow"): ["u", "v", "output_path", ("scale", 10.0)],
    ("lucas_kanade_reference", "export_flow_field_txt"): [
        "u", "v", "output_path", "width", "height", ("test_region", None)],
    ("lucas_kanade_reference", "main"): [],
}


def test_drop_in_modules_keep_reference_signatures(of_b200):
    import importlib

    for (mod_name, fn_name), expected in REFERENCE_SIGNATURES.items():
        mod = importlib.import_module(mod_name)
        assert str(BACKEND_DIR) in mod.__file__
        params = list(inspect.signature(getattr(mod, fn_name)).parameters.values())
        got = [p.name if p.default is inspect.Parameter.empty else (p.name, p.default) for p in params]
        assert got == expected, (mod_name, fn_name)
    ref = importlib.import_module("lucas_kanade_reference")
    for const in ("SCRIPT_DIR", "PROJECT_ROOT", "DEFAULT_FRAME_DIR", "DEFAULT_OUTPUT_DIR"):
        assert hasattr(ref, const)


def test_flow_field_text_export_format(of_b200, tmp_path):
    import lucas_kanade_reference as ref

    u = np.arange(6, dtype=np.float32).reshape(2, 3) / 4
    v = -u
    out = tmp_path / "flow.txt"
    ref.export_flow_field_txt(u, v, out, width=3, height=2, test_region={"x_min": 0, "x_max": 1, "y_min": 0, "y_max": 1})
    lines = out.read_text().splitlines()
    assert lines[:4] == [
        "# Optical flow field data (Python reference)",
        "# Format: x y u v",
        "# Image size: 3x2",
        "# Test region: x[0:1], y[0:1]",
    ]
    assert lines[4] == "0 0 0.000000 -0.000000" or lines[4] == "0 0 0.000000 0.000000"
    assert lines[-1] == "2 1 1.250000 -1.250000"
    assert len(lines) == 4 + 6


def test_frame_loader_reads_the_reference_on_disk_formats(of_b200, tmp_path):
    """of_load_frame_u8: frame_XX.bin is `ndarray.tofile` of uint8, frame_XX.mem is one `{val:02x}`
    line per pixel (python/generate_test_suite.py:259-271).  Host-side I/O: runs without a GPU."""
    fr = np.random.default_rng(3).integers(0, 256, (24, 40)).astype(np.uint8)
    fr.tofile(tmp_path / "frame_00.bin")
    with open(tmp_path / "frame_00.mem", "w") as f:
        for val in fr.flatten():
            f.write(f"{val:02x}\n")
    assert np.array_equal(of_b200.load_frame_u8(tmp_path / "frame_00.bin", 24, 40), fr)
    assert np.array_equal(of_b200.load_frame_u8(tmp_path / "frame_00.mem", 24, 40), fr)
    (tmp_path / "c.mem").write_text("// comment\n\nff\n 0A\n")
    assert of_b200.load_frame_u8(tmp_path / "c.mem", 1, 2).tolist() == [[255, 10]]
    for shape in ((23, 40), (25, 40)):  # truncated / oversized files are errors, not partial frames
        for name in ("frame_00.bin", "frame_00.mem"):
            with pytest.raises(ValueError):
                of_b200.load_frame_u8(tmp_path / name, *shape)
    (tmp_path / "bad.mem").write_text("zz\n")
    with pytest.raises(ValueError):
        of_b200.load_frame_u8(tmp_path / "bad.mem", 1, 1)
    with pytest.raises(ValueError):
        of_b200.load_frame_u8(tmp_path / "missing.bin", 1, 1)


def test_native_flow_export_matches_the_reference_format(of_b200, tmp_path):
    """of_export_flow_txt writes what export_flow_field_txt (python/lucas_kanade_reference.py:78-103)
    writes: same header, same "x y u v" lines with Python's "%.6f" rounding.  Host-side I/O."""
    rng = np.random.default_rng(1)
    u = (rng.standard_normal((7, 9)) * 3).astype(np.float32)
    v = (rng.standard_normal((7, 9)) * 3).astype(np.float32)
    u[0, 0], v[0, 0], u[1, 1] = 0.0, -0.0, np.float32(0.0000005)
    region = {"x_min": 2, "x_max": 6, "y_min": 1, "y_max": 5}
    for reg in (region, None):
        want = ["# Optical flow field data (Python reference)", "# Format: x y u v", "# Image size: 9x7"]
        if reg:
            want.append("# Test region: x[2:6], y[1:5]")
        for y in range(7):
            for x in range(9):
                want.append(f"{x} {y} {u[y, x]:.6f} {v[y, x]:.6f}")
        of_b200.export_flow_txt(tmp_path / "flow.txt", u, v, reg)
        assert (tmp_path / "flow.txt").read_text().splitlines() == want
    import lucas_kanade_reference as ref_mod  # the drop-in module writes the same file

    ref_mod.export_flow_field_txt(u, v, tmp_path / "flow_py.txt", 9, 7, region)
    of_b200.export_flow_txt(tmp_path / "flow.txt", u, v, region)
    assert (tmp_path / "flow_py.txt").read_text() == (tmp_path / "flow.txt").read_text()
    fx = (rng.integers(-1024, 1025, (4, 5))).astype(np.int16)
    of_b200.export_flow_txt(tmp_path / "fx.txt", fx, -fx)
    lines = (tmp_path / "fx.txt").read_text().splitlines()
    assert lines[0] == "# Optical flow field data" and lines[2] == "# Image size: 5x4"
    assert lines[3] == f"0 0 {fx[0, 0] / 128:.6f} {-fx[0, 0] / 128:.6f}"
    with pytest.raises(ValueError):
        of_b200.export_flow_txt(tmp_path / "no_such_dir" / "f.txt", u, v)


def test_verifier_test_region_matches_the_mask_oracle(of_b200):
    """of_b200.verifier_test_region gives the rectangle of get_test_region_mask
    (python/optical_flow_verifier.py:96-138) for every pattern family."""
    from oracle import flow_metrics_oracle as fm

    for shape in ((240, 320), (1080, 1920), (101, 203)):
        for name in ("translate_small", "translate_vertical", "rotate_small", "zoom_in", "translate_rotate", "no_motion"):
            y0, y1, x0, x1 = of_b200.verifier_test_region(shape, name, 80)
            mask = fm.test_region_mask(shape, name, 80)
            want = np.zeros(shape, bool)
            want[y0:y1, x0:x1] = True
            assert np.array_equal(mask, want), (shape, name)


def test_header_is_plain_c_and_a_c_program_links(of_b200, tmp_path):
    """include/of_b200.h is C99 (no C++ or torch types in the signatures) and a plain C caller links
    against libof_b200.so; without a GPU the compute call fails loudly with OF_ERR_NO_DEVICE."""
    import shutil

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    hdr = ROOT / "include" / "of_b200.h"
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", str(hdr)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    src = tmp_path / "caller.c"
    src.write_text(
        '#include <stdio.h>\n#include "of_b200.h"\n'
        "int main(void) {\n"
        "    float z[49] = {0}, u[49], v[49];\n"
        "    int rc = of_lk_single_scale_f32(z, z, u, v, 1, 7, 7, 5, OF_MODE_EXACT);\n"
        '    printf("%d %d %d\\n", of_version(), of_device_count(), rc);\n'
        "    return 0;\n}\n"
    )
    exe = tmp_path / "caller"
    lib_dir = str(of_b200.LIB_PATH.parent)
    res = subprocess.run([gcc, "-std=c99", f"-I{ROOT / 'include'}", str(src), "-o", str(exe), f"-L{lib_dir}", "-lof_b200",
                          f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()
    version, devices, rc = (int(x) for x in out)
    assert version == 100
    assert rc == (0 if devices > 0 else 4)  # OF_OK with a GPU, OF_ERR_NO_DEVICE without: never a CPU result


def test_motion_matrix_matches_the_pattern_oracle(of_b200):
    """of_b200.motion_matrix (cv2.getRotationMatrix2D + translation, generate_test_suite.py:183-190) equals the
    oracle's restatement, which tests/golden/make_golden_motion.py checked against cv2 itself."""
    from oracle import pattern_oracle as po

    for args in ((320, 240, 0.5, 0.5, 0.0, 1.0), (320, 240, 0.0, 0.0, 15.0, 1.0), (131, 97, -3.25, 7.5, -11.0, 1.17)):
        assert np.array_equal(of_b200.motion_matrix(*args), po.motion_matrix(*args))
