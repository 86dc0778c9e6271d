import importlib.util
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "low-cost-hardware-accelerated-vision-based-depth-perception-for-real-time-applications_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_binding():
    """The package directory name is not a Python identifier, so it is loaded by path as `elas_b200`."""
    if "elas_b200" in sys.modules:
        return sys.modules["elas_b200"]
    spec = importlib.util.spec_from_file_location("elas_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["elas_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def svb():
    return load_binding().binding


@pytest.fixture(scope="session")
def ref():
    from oracle.ref import RefElas

    return RefElas()


@pytest.fixture(scope="session")
def kitti_gray():
    z = np.load(os.path.join(GOLDEN, "kitti_gray.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(GOLDEN, "kitti_golden.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
        return json.load(f)
