"""Shared fixtures.  The CPU oracle (oracle/) is test infrastructure: it is loaded here, never by the product."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
    return ctypes.CDLL(path)


@pytest.fixture(scope="session")
def oracle_factory(oracle_lib):
    from difffr_b200.cabi import Context

    def make(**kw):
        return Context(lib=oracle_lib, prefix="orc_", **kw)

    return make


@pytest.fixture(scope="session")
def gpu_factory():
    from difffr_b200.cabi import Context

    def make(**kw):
        return Context(device=0, **kw)  # raises DfrError without a CUDA device: there is no CPU fallback

    return make


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))
