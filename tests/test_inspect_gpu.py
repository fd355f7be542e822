"""GPU tests of the resident-field accessors (SURVEY 8b "_window / _checksum", csrc/fdtd_inspect.cu) and of the
cropped-window parity check built on them (oracle/windows.py): windows of a resident level equal the same slices
of a full download, device-side checksums equal numpy's, the checksums of x-slabs add up to the single-slab ones,
and the 512^3 benchmark field equals the oracle run on a cropped grid bit for bit with nothing non-zero outside."""
import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def np_checksum(a, x_offset_planes=0, full_shape=None, origin=(0, 0, 0)):
    """numpy restatement of fdtd_b200_checksum over the dense array a = level[x0:x1, y0:y1, z0:z1]."""
    bits = a.view(np.uint32).astype(np.uint64)
    nxp, nyp, nzp = full_shape
    X, Y, Z = np.meshgrid(*(np.arange(o, o + s, dtype=np.uint64) for o, s in zip(origin, a.shape)), indexing="ij")
    lin = ((X + np.uint64(x_offset_planes)) * np.uint64(nyp) + Y) * np.uint64(nzp) + Z + np.uint64(1)
    with np.errstate(over="ignore"):
        pos = int((bits * lin).sum(dtype=np.uint64))
    fin = np.isfinite(a)
    return {"bit_sum": int(bits.sum(dtype=np.uint64)), "pos_sum": pos, "nonzero": int(np.count_nonzero(a)),
            "nonfinite": int((~fin).sum()), "max_abs": float(np.abs(a[fin]).max()) if fin.any() else 0.0,
            "sum_sq": float((a[fin].astype(np.float64) ** 2).sum())}


def test_window_and_checksum_match_numpy(pkg):
    rng = np.random.default_rng(7)
    shape = (20, 24, 40)
    u = rng.uniform(-1, 1, (3, 28, 32, 48)).astype(np.float32)
    u[1, 5, 6, 7] = np.inf
    u[1, 9, 9, 9] = np.nan
    u[2, 10:14] = 0.0
    u[2, 11, 3, 3] = -0.0
    m = np.full((28, 32, 48), 1.5, np.float32)
    with pkg.Plan(*shape, deviceid=0) as p:
        p.upload(u, m)
        for lvl in range(3):
            for w in (None, p.interior(), (3, 17, 0, 32, 5, 6), (0, 1, 31, 32, 47, 48), (4, 24, 4, 28, 4, 44)):
                ww = p._window(w)
                ref = np.ascontiguousarray(u[lvl, ww[0]:ww[1], ww[2]:ww[3], ww[4]:ww[5]])
                assert bits_equal(p.download_window(lvl, ww), ref)
                c, r = p.checksum(lvl, w), np_checksum(ref, 0, u.shape[1:], (ww[0], ww[2], ww[4]))
                for k in ("bit_sum", "pos_sum", "nonzero", "nonfinite"):
                    assert c[k] == r[k], (lvl, w, k)
                assert c["max_abs"] == np.float32(r["max_abs"])
                assert abs(c["sum_sq"] - r["sum_sq"]) <= 1e-12 * max(1.0, r["sum_sq"])
        for bad in ((0, 0, 0, 1, 0, 1), (0, 29, 0, 1, 0, 1), (-1, 2, 0, 1, 0, 1)):
            with pytest.raises(pkg.FdtdError):
                p.checksum(0, bad)
            with pytest.raises(pkg.FdtdError):
                p.download_window(0, bad)


def test_window_follows_level_placement_after_two_step_passes(pkg):
    """Two-step passes rotate a spare device level through the ring: windows/checksums must follow it."""
    from test_tb2_gpu import fused_case

    shape, T = (40, 48, 128), 9
    u, m, src, crd = fused_case(3, shape, T, 4)
    with pkg.Plan(*shape, deviceid=0) as p:
        p.set_option("kernel", 2)
        p.set_option("t_fuse", 2)
        p.upload(u, m)
        p.set_sources(src, crd)
        p.run(0, T - 1)
        assert p.get_option("t_fuse_used") == 2
        full = p.download()
        for lvl in range(3):
            assert bits_equal(p.download_window(lvl, None), full[lvl])
            assert p.checksum(lvl)["bit_sum"] == int(full[lvl].view(np.uint32).astype(np.uint64).sum())


@pytest.mark.parametrize("nparts", [2, 3])
def test_slab_checksums_add_up(pkg, nparts):
    """Integer checksums of the slabs' interior windows sum (mod 2^64) to the single-slab interior checksum."""
    from oracle import windows as W

    shape, T, S = (32 * nparts, 24, 64), 8, 6
    u, m, src, crd = W.dense_seam_case(11, shape, T, S, nparts)
    with pkg.Plan(*shape, deviceid=0) as p:
        p.upload(u, m)
        p.set_sources(src, crd)
        p.run(0, T - 1)
        one = [p.checksum(lvl, p.interior()) for lvl in range(3)]
    ls = pkg.LocalSlabs(*shape, [0] * nparts)
    ls.upload(u, m)
    ls.set_sources(src, crd)
    ls.run(0, T - 1)
    for lvl in range(3):
        parts = [q.checksum(lvl, q.interior()) for q in ls.plans]
        for k in ("bit_sum", "pos_sum", "nonzero", "nonfinite"):
            assert sum(c[k] for c in parts) % 2 ** 64 == one[lvl][k], (lvl, k)
        assert max(c["max_abs"] for c in parts) == one[lvl]["max_abs"]
    ls.close()


@pytest.mark.parametrize("t_fuse", [1, 2])
def test_512_windows_bit_identical_to_cropped_oracle(pkg, oracle, t_fuse):
    """BASELINE configs[2] through the checker bench.py uses: exact arithmetic, windows around the source equal the
    oracle on the cropped grid bit for bit, and the checksum of the whole interior says nothing else is non-zero."""
    from oracle import windows as W

    n, T, S = 512, 50, 1
    src, crd = pkg.fill_ricker(T, S), pkg.fill_source_coords(S, n, n, n)
    wins = W.source_windows(crd, (n, n, n))
    refs = [W.run_window(w, src, threads=8) for w in wins]
    with pkg.Plan(n, n, n, deviceid=0) as p:
        p.set_option("t_fuse", t_fuse)
        p.fill(0.0, 1.5)
        p.set_sources(src, crd)
        p.run(0, T - 1)
        assert p.get_option("t_fuse_used") == t_fuse
        acc = W.compare_windows(p, 0, n, wins, refs)
        nz = sum(p.checksum(lvl, p.interior())["nonzero"] for lvl in range(3))
    assert acc.bit_identical and acc.cells == 3 * int(np.prod(wins[0]["size"]))
    assert nz == acc.nonzero_in_windows and nz > 10000
    assert abs(acc.peak - 0.116841748) < 1e-8  # SURVEY 8c known answer at n = 512


def test_27_source_lattice_windows_at_384(pkg, oracle):
    """Several clusters: the 27-source lattice at 384^3 (spacing 96 cells) gives 27 windows."""
    from oracle import windows as W

    n, T, S = 384, 30, 27
    src, crd = pkg.fill_ricker(T, S), pkg.fill_source_coords(S, n, n, n)
    wins = W.source_windows(crd, (n, n, n), half=40)
    assert len(wins) == 27
    refs = [W.run_window(w, src, threads=8) for w in wins]
    with pkg.Plan(n, n, n, deviceid=0) as p:
        p.fill(0.0, 1.5)
        p.set_sources(src, crd)
        p.run(0, T - 1)
        acc = W.compare_windows(p, 0, n, wins, refs)
        nz = sum(p.checksum(lvl, p.interior())["nonzero"] for lvl in range(3))
    assert acc.bit_identical and nz == acc.nonzero_in_windows
