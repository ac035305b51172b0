"""The north star's drop-in claim, executed: the REFERENCE'S OWN optical_flow_verifier.py and flow_metrics.py, unchanged,
on top of this repository's lucas_kanade_core / lucas_kanade_pyramidal modules.

The reference exists only in the build container, which has no GPU, so the library under the drop-in modules is the
host-emulated one (tests/host_emul/): the product's own .cu files compiled for the CPU.  The run
regenerates the 13-pattern suite with the reference's generator, runs `python -m optical_flow_verifier
--compare-baseline` (BASELINE.json configs[0] and configs[1]: single-scale and 3-level pyramidal LK on every pattern)
and requires every metric of verification_results.json to EQUAL python/verification_baseline.json.  Skipped where the
reference is not present (the GPU box).
"""

import json
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/python")
pytestmark = pytest.mark.skipif(not (REF / "optical_flow_verifier.py").exists(), reason="the reference is not on this machine")
sys.path.insert(0, str(ROOT / "tests" / "host_emul"))

MATPLOTLIB_STUB = """import sys
from unittest import mock
pyplot = mock.MagicMock(name="pyplot")
pyplot.subplots.side_effect = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
colors = mock.MagicMock(name="colors")
sys.modules["matplotlib.pyplot"] = pyplot
sys.modules["matplotlib.colors"] = colors
"""


def numeric_leaves(node, path=""):
    if isinstance(node, dict):
        for k, v in node.items():
            yield from numeric_leaves(v, f"{path}/{k}")
    elif isinstance(node, (int, float)) and not isinstance(node, bool):
        yield path, node


def test_reference_verifier_runs_unchanged_on_the_backend_and_reproduces_its_baseline(tmp_path):
    import build_emulated_library

    try:
        lib = build_emulated_library.build()
    except RuntimeError as e:
        if "needs g++" in str(e):
            pytest.skip(str(e))
        raise
    backend = ROOT / "optical-flow-fpga_b200"
    suite, work = tmp_path / "suite", tmp_path / "work"
    (work / "python").mkdir(parents=True)
    (work / "stubs" / "matplotlib").mkdir(parents=True)
    (work / "stubs" / "matplotlib" / "__init__.py").write_text(MATPLOTLIB_STUB)  # not installed here; plots are off anyway
    res = subprocess.run([sys.executable, str(REF / "generate_test_suite.py"), "--output-dir", str(suite)], cwd=work,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    cfg = (REF / "verification_config.yaml").read_text()
    cfg = "\n".join(line[: len(line) - len(line.lstrip())] + f'test_suite_dir: "{suite}"' if line.strip().startswith("test_suite_dir:")
                    else line for line in cfg.splitlines())
    (work / "python" / "verification_config.yaml").write_text(cfg + "\n")
    shutil.copy(REF / "verification_baseline.json", work / "python" / "verification_baseline.json")

    env = dict(os.environ, OF_B200_LIB_NAME=os.path.relpath(lib, backend),
               PYTHONPATH=os.pathsep.join([str(backend), str(REF), str(work / "stubs")]))
    env.pop("OF_B200_MODE", None)  # default: exact mode
    # the modules the verifier will import: its own file and flow_metrics from the reference, the LK modules from here
    probe = subprocess.run([sys.executable, "-c", "import optical_flow_verifier as v, lucas_kanade_core as c, "
                            "lucas_kanade_pyramidal as p, flow_metrics as f; print(v.__file__, c.__file__, p.__file__, f.__file__)"],
                           cwd=work, env=env, capture_output=True, text=True, timeout=120)
    assert probe.returncode == 0, probe.stderr[-2000:]
    v_file, c_file, p_file, f_file = probe.stdout.split()
    assert Path(v_file).parent == REF and Path(f_file).parent == REF
    assert Path(c_file).parent == backend and Path(p_file).parent == backend

    run = subprocess.run([sys.executable, "-m", "optical_flow_verifier", "--no-visualizations", "--compare-baseline"], cwd=work,
                         env=env, capture_output=True, text=True, timeout=2400)
    assert run.returncode == 0, (run.stdout[-3000:], run.stderr[-3000:])
    assert "All patterns pass regression check" in run.stdout
    got = json.loads((work / "python" / "verification_results.json").read_text())["patterns"]
    want = json.loads((REF / "verification_baseline.json").read_text())["patterns"]
    assert set(got) == set(want) and len(got) == 13
    want_leaves = dict(numeric_leaves(want))
    compared = 0
    for path, value in numeric_leaves(got):
        if path in want_leaves:
            assert value == want_leaves[path], (path, value, want_leaves[path])
            compared += 1
    assert compared >= 13 * 2 * 5  # five metrics, two methods, thirteen patterns


def test_reference_cli_and_drop_in_cli_write_the_same_files(tmp_path):
    """scripts/regenerate_flow_plots.sh's first step: `python lucas_kanade_reference.py` -- the reference's file and this
    repository's file of the same name on the same frames must print the same report and write byte-identical
    flow_u.bin, flow_v.bin and flow_field_python.txt (the file scripts/visualize_flow.py consumes)."""
    import build_emulated_library

    try:
        lib = build_emulated_library.build()
    except RuntimeError as e:
        if "needs g++" in str(e):
            pytest.skip(str(e))
        raise
    backend = ROOT / "optical-flow-fpga_b200"
    suite, stubs = tmp_path / "suite", tmp_path / "stubs"
    (stubs / "matplotlib").mkdir(parents=True)
    (stubs / "matplotlib" / "__init__.py").write_text(MATPLOTLIB_STUB)
    res = subprocess.run([sys.executable, str(REF / "generate_test_suite.py"), "--output-dir", str(suite)], cwd=tmp_path,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    logs = {}
    for who, script, path in (("ref", REF / "lucas_kanade_reference.py", [str(REF), str(stubs)]),
                              ("ours", backend / "lucas_kanade_reference.py", [str(backend), str(stubs)])):
        out = tmp_path / who
        env = dict(os.environ, OF_B200_LIB_NAME=os.path.relpath(lib, backend), PYTHONPATH=os.pathsep.join(path))
        run = subprocess.run([sys.executable, str(script), "--frame-dir", str(suite / "rotate_small"), "--output-dir", str(out)],
                             cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
        assert run.returncode == 0, (who, run.stderr[-2000:])
        logs[who] = run.stdout.replace(str(out), "OUT")
    assert logs["ours"] == logs["ref"]
    for name in ("flow_u.bin", "flow_v.bin", "flow_field_python.txt"):
        assert (tmp_path / "ours" / name).read_bytes() == (tmp_path / "ref" / name).read_bytes(), name

    # the second step of the script: `python lucas_kanade_pyramidal.py` -- the per-level / per-iteration report and the
    # saved flow fields, on a moving pattern and on one that converges in its first iteration
    for pattern in ("translate_medium", "no_motion"):
        logs = {}
        for who, script, path in (("ref", REF / "lucas_kanade_pyramidal.py", [str(REF), str(stubs)]),
                                  ("ours", backend / "lucas_kanade_pyramidal.py", [str(backend), str(stubs)])):
            out = tmp_path / f"{who}_{pattern}"
            out.mkdir()
            env = dict(os.environ, OF_B200_LIB_NAME=os.path.relpath(lib, backend), PYTHONPATH=os.pathsep.join(path))
            run = subprocess.run([sys.executable, str(script), "--frame-dir", str(suite / pattern), "--output-dir", str(out / "o")],
                                 cwd=out, env=env, capture_output=True, text=True, timeout=900)
            assert run.returncode == 0, (who, pattern, run.stderr[-2000:])
            logs[who] = [ln for ln in run.stdout.replace(str(out), "OUT").splitlines() if "visualization saved" not in ln]
        assert logs["ours"] == logs["ref"], pattern
        for name in ("flow_u_pyramidal.bin", "flow_v_pyramidal.bin"):
            a = (tmp_path / f"ours_{pattern}" / "o" / name).read_bytes()
            assert a == (tmp_path / f"ref_{pattern}" / "o" / name).read_bytes(), (pattern, name)
