"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU."""

import json
import os
import sys
from pathlib import Path

# The native row-band test runs up to 8 emulated ranks as 8 streams of one device that wait for
# each other inside kernels: every stream needs its own hardware queue (default: 8 for all streams).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
BACKEND_DIR = ROOT / "optical-flow-fpga_b200"

for p in (str(ROOT), str(BACKEND_DIR)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_index():
    return json.load(open(GOLDEN / "golden_index.json"))


@pytest.fixture(scope="session")
def golden_frames():
    z = np.load(GOLDEN / "frames.npz")
    names = sorted({k.rsplit("__", 1)[0] for k in z.files})
    return {n: (z[f"{n}__0"], z[f"{n}__1"]) for n in names}


@pytest.fixture(scope="session")
def golden_units():
    z = np.load(GOLDEN / "units.npz")
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_flows():
    z = np.load(GOLDEN / "flows_subset.npz")
    return {k: z[k] for k in z.files}
