"""GPU tests of the staged run (fdtd_b200_plan_run_staged): upload, time loop and download as one pipeline,
the time loop skewed along x so blocks of planes run while later chunks are still on the wire.

Every point sees the inputs of the unskewed loop, so the bar is 0 ulp against the oracle (exact arithmetic)
and against the three-phase path upload -> run -> download.
"""
import numpy as np
import pytest

from conftest import bench_inputs, bits_equal
from test_parity_gpu import random_case

pytestmark = pytest.mark.gpu


def interior_sources(crd, shape):
    """Move every source at least two cells inside the grid (no trilinear corner in a halo cell)."""
    hi = (np.array(shape, np.float32) - 3) * np.float32(0.1)
    return np.clip(crd, np.float32(0.2), hi).astype(np.float32)


@pytest.fixture(params=[0, 1], ids=["direct", "bounce"])
def bounce(request):
    """numpy arrays are pageable: FDTD_B200_BOUNCE=1 (the default for pageable arrays) packs chunks into pinned
    buffers with host threads, 0 hands the caller's arrays to cudaMemcpyAsync directly."""
    import os

    os.environ["FDTD_B200_BOUNCE"] = str(request.param)
    yield request.param
    del os.environ["FDTD_B200_BOUNCE"]


@pytest.mark.parametrize("shape,T,planes,time_m", [
    ((70, 24, 64), 7, 8, 0),      # several blocks, skew 2(T-1) = 12 > block length
    ((64, 16, 32), 1, 32, 0),     # a single step: exactly two blocks
    ((96, 20, 72), 12, 16, 4),    # ring phase 1, timed steps
    ((50, 12, 36), 9, 8, 2),      # generic kernel (small grid)
    ((45, 9, 10), 6, 8, 0),       # nz % 4 != 0: generic kernel only
])
def test_staged_matches_oracle_and_three_phase_path(pkg, oracle, bounce, shape, T, planes, time_m):
    rng = np.random.default_rng(sum(shape) + T)
    u, m, src, crd = random_case(rng, shape, time_m + T, 5)
    crd = interior_sources(crd, shape)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port", time_m=time_m, time_M=time_m + T - 1)
    with pkg.Plan(*shape, deviceid=0) as p:
        p.set_option("stage_planes", planes)
        p.set_sources(src, crd)
        out = u.copy()
        t = p.run_staged(out, m, time_m, time_m + T - 1)
        assert t is not None, "staged path refused"
        assert bits_equal(out, ref)
        assert (t.section0 > 0) == (T > 5) and t.section1 == 0.0
        # the device copy is complete as well (a later resident run may continue from it)
        assert bits_equal(p.download(), ref)
        # and the three-phase path on the same plan gives the same bits
        p.upload(u, m)
        p.run(time_m, time_m + T - 1)
        assert bits_equal(p.download(), ref)


def test_staged_contracted_equals_resident_contracted(pkg, oracle):
    shape, T = (80, 32, 64), 10
    rng = np.random.default_rng(77)
    u, m, src, crd = random_case(rng, shape, T, 4)
    crd = interior_sources(crd, shape)
    with pkg.Plan(*shape, deviceid=0) as p:
        p.set_option("exact", 0)
        p.set_option("kernel", 2)
        p.set_option("stage_planes", 16)
        p.set_sources(src, crd)
        a = u.copy()
        assert p.run_staged(a, m, 0, T - 1) is not None
        p.upload(u, m)
        p.run(0, T - 1)
        b = p.download()
    assert bits_equal(a, b)


def test_staged_refuses_what_it_cannot_do(pkg, oracle):
    """Sources in halo cells go through the stand-alone scatter after each step: three-phase path.  The ABI entry
    falls back by itself and stays exact."""
    shape, T = (72, 16, 32), 8
    rng = np.random.default_rng(5)
    u, m, src, crd = random_case(rng, shape, T, 4)
    crd = interior_sources(crd, shape)
    crd[1, 1] = np.float32(-0.03)  # a corner in the y halo
    with pkg.Plan(*shape, deviceid=0) as p:
        p.set_option("stage_planes", 8)
        p.set_sources(src, crd)
        keep = u.copy()
        assert p.run_staged(keep, m, 0, T - 1) is None
        assert bits_equal(keep, u)  # untouched
        p.set_option("stage_planes", 0)
        p.set_sources(src, interior_sources(crd, shape))
        assert p.run_staged(keep, m, 0, T - 1) is None  # switched off
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    out = u.copy()
    rc = pkg.Kernel_B200(m, src, crd, out, shape[0] - 1, 0, shape[1] - 1, 0, shape[2] - 1, 0, 1e-3, 0.1, 0.1, 0.1, 0.0, 0.0, 0.0,
                         3, 0, T - 1, 0, 0, 1)
    assert rc == 0 and bits_equal(out, ref)


def test_abi_takes_the_staged_path_and_matches_golden(pkg, oracle, golden, bounce):
    """Kernel_CUDA_Optimized on the driver's 256^3 benchmark config: pipelined staging, same bits as the reference
    build, and section0 still covers only the 45 timed steps."""
    import hashlib

    meta, _ = golden
    g = meta["bench256_s1"]
    n = g["n"]
    u, m, src, crd = bench_inputs(oracle, n, g["T"], g["S"])
    t = pkg.Profiler(0.0, 0.0)
    rc = pkg.Kernel_CUDA_Optimized(m, src, crd, u, n - 1, 0, n - 1, 0, n - 1, 0, 1e-3, 0.1, 0.1, 0.1, 0.0, 0.0, 0.0,
                                   g["S"] - 1, 0, g["T"] - 1, 0, 0, 1, t)
    assert rc == 0
    assert hashlib.sha256(np.ascontiguousarray(u).tobytes()).hexdigest() == g["sha256"]
    assert 0 < t.section0 < 0.05 and t.section1 == 0.0
