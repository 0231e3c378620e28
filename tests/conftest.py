import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def gold(name):
    return np.load(os.path.join(GOLD, name))


@pytest.fixture(scope="session")
def bun000():
    return np.load(os.path.join(GOLD, "bun000_xyz.npy"))


@pytest.fixture(scope="session")
def bun045():
    return np.load(os.path.join(GOLD, "bun045_xyz.npy"))


@pytest.fixture(scope="session")
def engine():
    import hgmm_b200
    eng = hgmm_b200.Engine(0)
    yield eng
    eng.close()


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
