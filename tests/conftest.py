import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "accelerated-3d-acoustic-fdtd-kernel_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package; builds libfdtd_b200.so in-tree if it is missing (nvcc needs no GPU)."""
    mod = importlib.import_module(PKG_NAME)
    if not os.path.exists(mod.lib_path()):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand by oracle/Makefile."""
    from oracle import oracle as O

    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        O.build()
    return O


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(d, "golden.json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(d, "golden.npz"))


def bench_inputs(O, n, T, S):
    """The driver's benchmark inputs (main.cpp:285-356): zero field, m = 1.5, Ricker, lattice sources."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    u = np.zeros((3, nx + 8, ny + 8, nz + 8), np.float32)
    m = np.full((nx + 8, ny + 8, nz + 8), 1.5, np.float32)
    return u, m, O.fill_ricker(T, S), O.fill_source_coords(S, nx, ny, nz)


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))
