"""GPU tests of the SURVEY 8(f) "next" rows 3-4: space orders 6..12 and receiver sampling.

The reference ships order-4 kernels only and no receivers, so the checker is this repo's generalisation in
oracle/fdtd_oracle.c (oracle_run_order), which is bit-identical to the pinned order-4 oracle at order 4
(tests/test_oracle.py).  Bars as everywhere: exact arithmetic 0 ulp, contracted relative L2 < 1e-4."""
import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def order_case(seed, shape, T, S, so, nrec=7):
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    pad = 2 * so
    u = rng.uniform(-1, 1, (3, nx + pad, ny + pad, nz + pad)).astype(np.float32)
    m = rng.uniform(0.5, 3.0, (nx + pad, ny + pad, nz + pad)).astype(np.float32)
    src = rng.uniform(-20, 20, (T, S)).astype(np.float32)
    ext = (np.array(shape, np.float32) - 1) * np.float32(0.1)
    crd = (rng.uniform(-0.04, 1.04, (S, 3)) * ext).astype(np.float32)
    if S >= 3:
        crd[1] = crd[0]
    rec = (rng.uniform(-0.06, 1.06, (nrec, 3)) * ext).astype(np.float32)
    rec[0] = crd[0]                      # a receiver on top of a source
    rec[1] = (0.0, 0.0, 0.0)             # on the corner cell
    rec[2] = ext                         # base corner on the last cell: +1 corners in the first halo cell
    rec[3] = ext + np.float32(0.25)      # base corner in the halo: corners out of range are skipped
    return u, m, src, crd, rec


@pytest.mark.parametrize("so", [4, 6, 8, 10, 12])
@pytest.mark.parametrize("exact", [1, 0])
def test_space_orders_against_the_oracle(pkg, oracle, so, exact):
    shape, T, S = (21, 18, 37), 9, 5
    u, m, src, crd, rec = order_case(40 + so, shape, T, S, so)
    ref = u.copy()
    _, _, ref_rec = oracle.run_order(ref, m, src, crd, space_order=so, rec_coords=rec)
    with pkg.Plan(*shape, deviceid=0, space_order=so) as p:
        assert p.shape == u.shape and p.get_option("space_order") == so
        p.set_option("exact", exact)
        p.upload(u, m)
        p.set_sources(src, crd)
        p.set_receivers(rec)
        t = p.run(0, T - 1)
        out = p.download()
        got_rec, owned = p.receivers()
        assert p.get_option("kernel_used") == (3 if so != 4 else 1)
    assert owned.all() and got_rec.shape == ref_rec.shape
    if exact:
        assert bits_equal(out, ref)
        assert bits_equal(got_rec, ref_rec)
    else:
        assert oracle.rel_l2(out, ref) < 1e-4
        assert oracle.rel_l2(got_rec, ref_rec) < 1e-4
    assert t.section0 > 0 and t.section1 > 0  # sampling is reported under section1


def test_order_is_read_off_the_padding_by_the_reference_abi(pkg, oracle):
    """A driver built with -DSTENCIL_ORDER=8 pads with 8 cells (main.cpp:27-32) and calls the same 24-argument entry."""
    so, shape, T, S = 8, (24, 24, 24), 8, 3
    u, m, src, crd, _ = order_case(5, shape, T, S, so)
    ref = u.copy()
    oracle.run_order(ref, m, src, crd, space_order=so)
    n = shape[0]
    rc = pkg.Kernel_CUDA_Optimized(m, src, crd, u, n - 1, 0, n - 1, 0, n - 1, 0, 1e-3, 0.1, 0.1, 0.1, 0.0, 0.0, 0.0,
                                   S - 1, 0, T - 1, 0, 0, 1)
    assert rc == 0 and bits_equal(u, ref)


def test_higher_order_is_more_accurate_on_a_smooth_field(pkg):
    """One step on a plane wave: the discrete Laplacian error falls with the order (the weights are right)."""
    n, so_list, errs = 48, (4, 8, 12), []
    k = 2 * np.pi * 3 / n  # three periods across the grid: h*k ~ 0.39
    for so in so_list:
        i = np.arange(-so, n + so, dtype=np.float64)
        w = np.sin(k * i)[:, None, None] * np.cos(k * i)[None, :, None] * np.sin(k * i + 0.3)[None, None, :]
        u = np.stack([w, np.zeros_like(w), w]).astype(np.float32)  # step 0: current = u[0], previous = u[2]
        m = np.ones(u.shape[1:], np.float32)
        with pkg.Plan(n, n, n, deviceid=0, space_order=so, dt=1.0, h=(1.0, 1.0, 1.0)) as p:
            p.set_option("exact", 0)
            p.upload(u, m)
            p.run(0, 0)
            out = p.download()[1, so:-so, so:-so, so:-so].astype(np.float64)
        # u_new = 2 u0 - u_prev + dt^2 lap(u0)/m = w + lap(w); the exact Laplacian of w is -3 k^2 w
        w32 = u[0, so:-so, so:-so, so:-so].astype(np.float64)
        errs.append(np.abs((out - w32) + 3 * k * k * w32).max())
    assert errs[0] > 20 * errs[1] and errs[1] >= 0.5 * errs[2] and errs[0] < 1e-3


@pytest.mark.parametrize("t_fuse", [1, 2])
def test_receivers_with_the_streaming_kernels(pkg, oracle, t_fuse):
    """Order 4 through the TMA kernels (one and two steps per launch) with receivers: traces and field 0 ulp."""
    from test_tb2_gpu import fused_case

    shape, T, S = (40, 48, 128), 13, 4
    u, m, src, crd = fused_case(21, shape, T, S)
    rng = np.random.default_rng(1)
    rec = (rng.uniform(0.0, 1.0, (9, 3)) * (np.array(shape, np.float32) - 1) * np.float32(0.1)).astype(np.float32)
    ref = u.copy()
    _, _, ref_rec = oracle.run_order(ref, m, src, crd, space_order=4, rec_coords=rec)
    with pkg.Plan(*shape, deviceid=0) as p:
        p.set_option("kernel", 2)
        p.set_option("t_fuse", t_fuse)
        p.upload(u, m)
        p.set_sources(src, crd)
        p.set_receivers(rec)
        p.run(0, T - 1)
        assert p.get_option("t_fuse_used") == t_fuse
        got, owned = p.receivers()
        out = p.download()
    assert owned.all() and bits_equal(out, ref) and bits_equal(got, ref_rec)


@pytest.mark.parametrize("nparts", [2, 3])
def test_receivers_on_slabs(pkg, oracle, nparts):
    """x-slabs: every receiver is sampled by exactly one slab (its +1 corner may sit in a ghost plane); traces 0 ulp."""
    from oracle import windows as W

    shape, T, S = (32 * nparts, 24, 64), 10, 6
    u, m, src, crd = W.dense_seam_case(3, shape, T, S, nparts)
    rec = crd.copy()  # receivers on the sources: on and next to every seam
    rec = np.concatenate([rec, [[0.0, 0.1, 0.2], [(shape[0] - 1) * 0.1, 1.0, 2.0]]]).astype(np.float32)
    ref = u.copy()
    _, _, ref_rec = oracle.run_order(ref, m, src, crd, space_order=4, rec_coords=rec)
    ls = pkg.LocalSlabs(*shape, [0] * nparts, options={"t_fuse": 2})
    ls.upload(u, m)
    ls.set_sources(src, crd)
    ls.set_receivers(rec)
    ls.run(0, T - 1)
    assert all(p.get_option("t_fuse_used") == 1 for p in ls.plans)  # linked slabs with receivers run one-step passes
    got = ls.receivers()
    out = np.zeros_like(u)
    ls.download(out)
    ls.close()
    assert bits_equal(out, ref) and bits_equal(got, ref_rec)
