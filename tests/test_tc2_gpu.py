"""GPU tests of the two-step pass on 2-CTA clusters (csrc/stencil_tc2.cu, option "cluster"): CTA 0 computes u^{n+1}
and streams it through distributed shared memory into CTA 1, which computes u^{n+2}.  Same bars as stencil_tb2:
exact arithmetic 0 ulp vs the oracle, contracted arithmetic 0 ulp vs the one-step contracted kernel."""
import numpy as np
import pytest

from conftest import bits_equal
from test_tb2_gpu import REL_L2_TOL, fused_case, run_plan

pytestmark = pytest.mark.gpu

TILES = [(28, 64, 1), (24, 64, 1), (32, 64, 1), (16, 128, 1), (16, 64, 1), (12, 128, 1), (16, 128, 2), (32, 64, 2), (28, 64, 2),
         (24, 64, 2), (40, 64, 2)]  # output tile, rows per thread


@pytest.mark.parametrize("ty,tz,rows", TILES)
def test_cluster_pass_bit_exact(pkg, oracle, ty, tz, rows):
    """Every instantiation on a grid that is not a multiple of the tile, several x chunks (some shorter than the ring),
    random m, sources incl. coincident ones."""
    shape, T, S = (23, 44, 72), 11, 6
    u, m, src, crd = fused_case(300 + ty + tz, shape, T, S)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    opts = {"kernel": 2, "t_fuse": 2, "cluster": 1, "tile_y": ty, "tile_z": tz, "rows": rows, "xchunk": 9}
    out, t, info = run_plan(pkg, u, m, src, crd, options=dict(opts, exact=1))
    assert info["t_fuse_used"] == 2 and (info["tile_y_used"], info["tile_z_used"]) == (ty, tz)
    assert info["launches"] < T + 1
    assert bits_equal(out, ref)
    one, _, _ = run_plan(pkg, u, m, src, crd, options={"kernel": 2, "t_fuse": 1, "exact": 0})
    two, _, _ = run_plan(pkg, u, m, src, crd, options=dict(opts, exact=0))
    assert bits_equal(two, one)
    assert oracle.rel_l2(two, ref) < REL_L2_TOL


@pytest.mark.parametrize("xchunk", [1, 3, 5, 16, 40])
def test_cluster_pass_chunk_lengths(pkg, oracle, xchunk):
    """Chunks shorter and longer than the rings (the A -> B ring has 8 slots, B trails A by 4 planes)."""
    shape, T, S = (40, 56, 128), 9, 4
    u, m, src, crd = fused_case(11 + xchunk, shape, T, S)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    out, _, info = run_plan(pkg, u, m, src, crd, options={"kernel": 2, "t_fuse": 2, "cluster": 1, "xchunk": xchunk})
    assert info["t_fuse_used"] == 2 and info["xchunk_used"] == xchunk
    assert bits_equal(out, ref)


@pytest.mark.parametrize("exact", [1, 0])
def test_cluster_pass_dense_256(pkg, oracle, exact):
    shape, T, S = (256, 256, 256), 12, 6
    u, m, src, crd = fused_case(78, shape, T, S)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port", threads=8)
    out, t, info = run_plan(pkg, u, m, src, crd, options={"exact": exact, "t_fuse": 2, "cluster": 1})
    assert info["kernel_used"] == 2 and info["t_fuse_used"] == 2
    if exact:
        assert bits_equal(out, ref)
    else:
        assert oracle.rel_l2(out, ref) < REL_L2_TOL
        assert float(np.abs(out - ref).max()) <= 1e-5 * float(np.abs(ref).max())


def test_cluster_pass_restart_and_ring_phase(pkg, oracle):
    shape, T, S = (16, 32, 64), 20, 3
    u, m, src, crd = fused_case(5, shape, T, S)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port", time_m=1, time_M=T - 1)
    out, _, info = run_plan(pkg, u, m, src, crd, options={"kernel": 2, "t_fuse": 2, "cluster": 1}, time_m=1, time_M=T - 1, split=8)
    assert info["t_fuse_used"] == 2
    assert bits_equal(out, ref)
