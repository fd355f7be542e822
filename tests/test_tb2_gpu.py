"""GPU tests of the two-step passes (temporal blocking, t_fuse = 2, csrc/stencil_tb2.cu).

A two-step pass computes u^{n+1} and u^{n+2} from one read of u^{n-1}, u^n and m.  Every point goes through
the same point<EXACT>() arithmetic as in the one-step kernels, so the bars are the same:
  * exact = 1: BIT-IDENTICAL to the oracle (and therefore to the reference built for the host);
  * exact = 0: bit-identical to the one-step contracted kernel, relative L2 < 1e-4 vs the oracle.
Also covered: the conditions under which passes fall back to one step (different halo shells, a source in
a halo cell, steps that do not pair up), restarts, and x-slabs with 4-plane ghost zones.
"""
import numpy as np
import pytest

from conftest import bench_inputs, bits_equal

pytestmark = pytest.mark.gpu

REL_L2_TOL = 1e-4  # README.md:33


def fused_case(seed, shape, T, S, *, same_shell=True, interior_sources=True, seam_parts=0):
    """Random field and model.  same_shell: the three levels share one halo shell (what a physical Dirichlet
    boundary means); interior_sources: no trilinear corner touches a halo cell."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    u = rng.uniform(-1, 1, (3, nx + 8, ny + 8, nz + 8)).astype(np.float32)
    if same_shell:
        inner = (slice(4, nx + 4), slice(4, ny + 4), slice(4, nz + 4))
        for lvl in (1, 2):
            keep = u[lvl][inner].copy()
            u[lvl] = u[0]
            u[lvl][inner] = keep
    m = rng.uniform(0.5, 3.0, (nx + 8, ny + 8, nz + 8)).astype(np.float32)
    src = rng.uniform(-20, 20, (T, max(S, 1))).astype(np.float32)
    lo, hi = (0.02, 0.97) if interior_sources else (-0.04, 1.04)
    crd = (rng.uniform(lo, hi, (max(S, 1), 3)) * (np.array(shape, np.float32) - 1) * np.float32(0.1)).astype(np.float32)
    if S >= 3:
        crd[1] = crd[0]  # coincident sources: per-cell ordering
    if seam_parts:
        base, i = nx // seam_parts, 0
        for k in range(1, seam_parts):  # sources straddling every seam, and one/two planes off it
            crd[i % S, 0] = np.float32((k * base - 1) * 0.1) + np.float32(0.04)
            crd[(i + 1) % S, 0] = np.float32(k * base * 0.1)
            crd[(i + 2) % S, 0] = np.float32((k * base - 2) * 0.1) + np.float32(0.03)
            crd[(i + 3) % S, 0] = np.float32((k * base + 1) * 0.1) + np.float32(0.02)
            i += 4
    return u, m, src, crd


def run_plan(pkg, u, m, src, crd, *, options, time_m=0, time_M=None, h=0.1, split=None):
    nxp, nyp, nzp = u.shape[1:]
    out = u.copy()
    with pkg.Plan(nxp - 8, nyp - 8, nzp - 8, h=h, deviceid=0) as p:
        for k, v in options.items():
            p.set_option(k, v)
        p.upload(out, m)
        if src is not None:
            p.set_sources(src, crd)
            if time_M is None:
                time_M = src.shape[0] - 1
        if split is None:
            t = p.run(time_m, time_M)
        else:
            p.run(time_m, split)
            t = p.run(split + 1, time_M)
        p.download(out)
        info = {k: p.get_option(k) for k in ("kernel_used", "t_fuse_used", "tile_y_used", "tile_z_used", "xchunk_used")}
        info["launches"] = p.last_launches
    return out, t, info


# (TY, TZ, rows per thread, lean): lean = 1 is stencil_tb2l.cu (the default), lean = 0 stencil_tb2.cu (the only one with two rows per thread)
TILES = [(32, 64, 1, 1), (28, 64, 1, 1), (16, 128, 1, 1), (16, 64, 1, 1), (32, 64, 1, 0), (28, 64, 1, 0), (16, 128, 1, 0), (16, 64, 1, 0),
         (32, 64, 2, 0), (16, 128, 2, 0)]


@pytest.mark.parametrize("ty,tz,rows,lean", TILES)
def test_two_step_pass_bit_exact(pkg, oracle, ty, tz, rows, lean):
    """Every instantiation of both two-step kernels on a grid that is not a multiple of the tile, several x chunks, random m,
    sources (coincident ones too) -- exact arithmetic: 0 ulp vs the oracle; contracted: 0 ulp vs the one-step kernel."""
    shape, T, S = (23, 44, 72), 11, 6
    u, m, src, crd = fused_case(200 + ty + tz, shape, T, S)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    opts = {"kernel": 2, "t_fuse": 2, "tile_y": ty, "tile_z": tz, "rows": rows, "xchunk": 9, "tb2_lean": lean}
    out, t, info = run_plan(pkg, u, m, src, crd, options=dict(opts, exact=1))
    assert info["t_fuse_used"] == 2 and (info["tile_y_used"], info["tile_z_used"]) == (ty, tz)
    assert info["launches"] < T + 1  # steps were actually paired (T one-step launches + the mbase gather otherwise)
    assert bits_equal(out, ref)
    assert t.section0 > 0 and t.section1 == 0.0
    one, _, i1 = run_plan(pkg, u, m, src, crd, options={"kernel": 2, "t_fuse": 1, "exact": 0})
    two, _, i2 = run_plan(pkg, u, m, src, crd, options=dict(opts, exact=0))
    assert i1["t_fuse_used"] == 1 and i2["t_fuse_used"] == 2
    assert bits_equal(two, one)
    assert oracle.rel_l2(two, ref) < REL_L2_TOL


@pytest.mark.parametrize("T", [1, 2, 3, 5, 6, 7, 8, 12, 13])
def test_pairing_of_steps(pkg, oracle, T):
    """Any number of steps: passes never straddle the untimed/timed boundary (step 5) or the end of the run."""
    shape, S = (16, 32, 64), 3
    u, m, src, crd = fused_case(7, shape, T, S)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    out, t, info = run_plan(pkg, u, m, src, crd, options={"kernel": 2, "t_fuse": 2})
    assert info["t_fuse_used"] == 2
    assert bits_equal(out, ref)
    assert (t.section0 > 0) == (T > 5)


@pytest.mark.parametrize("time_m,split", [(0, 6), (1, 4), (2, 9), (7, 12)])
def test_ring_phase_and_restart(pkg, oracle, time_m, split):
    """Arbitrary ring phase and a run split in two calls on one plan: the placement of the ring in the four
    device levels carries over."""
    shape, T, S = (16, 32, 64), 20, 3
    u, m, src, crd = fused_case(5, shape, T, S)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port", time_m=time_m, time_M=T - 1)
    out, _, info = run_plan(pkg, u, m, src, crd, options={"kernel": 2, "t_fuse": 2}, time_m=time_m, time_M=T - 1, split=split)
    assert info["t_fuse_used"] == 2
    assert bits_equal(out, ref)


def test_falls_back_when_shells_differ(pkg, oracle):
    """The dense correctness field of main.cpp:525-570 has u[2] = 0 but non-zero halos in u[0], u[1]: ring levels
    cannot move between device levels, so every pass is one step -- and the result is still exact."""
    u, m = oracle.fill_dense(32, 32, 64)
    ref = u.copy()
    oracle.run(ref, m, time_M=19, h=1.0, impl="port")
    out, _, info = run_plan(pkg, u, m, None, None, options={"kernel": 2, "t_fuse": 2}, time_M=19, h=1.0)
    assert info["t_fuse_used"] == 1
    assert bits_equal(out, ref)


def test_falls_back_when_a_source_touches_a_halo(pkg, oracle):
    shape, T, S = (16, 32, 64), 9, 5
    u, m, src, crd = fused_case(3, shape, T, S)
    crd[2] = (np.float32(-0.03), crd[2, 1], crd[2, 2])  # base corner at x = -1: its upper corners hit plane X0-1 .. X0
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    out, t, info = run_plan(pkg, u, m, src, crd, options={"kernel": 2, "t_fuse": 2})
    assert info["t_fuse_used"] == 1
    assert bits_equal(out, ref)
    assert t.section1 > 0  # the halo cell went through the stand-alone scatter


def test_benchmark_config_matches_golden_with_two_step_passes(pkg, oracle, golden):
    """The driver's benchmark inputs through the reference ABI with FDTD_SetRuntimeConfig(t_fuse = 2): two-step passes of the lean
    kernel in the library's default bit-exact arithmetic (with FDTD_B200_TB2_LEAN=0 the hook's t_fuse would be advisory: an exact
    pass of the first two-step kernel is slower than two one-step launches); the result is the golden one either way."""
    import hashlib

    meta, _ = golden
    g = meta["bench256_s1"]
    u, m, src, crd = bench_inputs(oracle, g["n"], g["T"], g["S"])
    n = g["n"]
    import os

    pkg.FDTD_SetRuntimeConfig(1, 2, 1)
    os.environ["FDTD_B200_STAGE_PLANES"] = "0"  # three-phase path: the staged run uses one-step launches
    try:
        t = pkg.Profiler(0.0, 0.0)
        rc = pkg.Kernel_CUDA_Optimized(m, src, crd, u, n - 1, 0, n - 1, 0, n - 1, 0, 1e-3, 0.1, 0.1, 0.1, 0.0, 0.0, 0.0,
                                       g["S"] - 1, 0, g["T"] - 1, 0, 0, 1, t)
    finally:
        pkg.FDTD_SetRuntimeConfig(1, 1, 1)
        del os.environ["FDTD_B200_STAGE_PLANES"]
    assert rc == 0
    assert hashlib.sha256(np.ascontiguousarray(u).tobytes()).hexdigest() == g["sha256"]
    assert t.section0 > 0 and t.section1 == 0.0


def test_many_sources_64(pkg, oracle):
    """The 64-source lattice of BASELINE configs[4] (38 sources on one cell) with two-step passes."""
    n, T, S = 64, 21, 64
    u, m, src, crd = bench_inputs(oracle, n, T, S)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    out, _, info = run_plan(pkg, u, m, src, crd, options={"kernel": 2, "t_fuse": 2})
    assert info["t_fuse_used"] == 2
    assert bits_equal(out, ref)


# ------------------------------------------------------------------ x-slabs: 4 ghost planes per pass
@pytest.mark.parametrize("nparts,shape,opts", [
    (2, (64, 16, 64), {}),
    (3, (96, 36, 72), {"exact": 0}),
    (2, (80, 24, 128), {"tile_y": 16, "tile_z": 128, "xchunk": 12}),
    (4, (131, 16, 64), {"xchunk": 20}),
    # the halo protocol is pull by default for lean two-step runs; the push protocol (peer stores from inside the kernel) and the first kernel
    (3, (96, 36, 72), {"halo_pull": 0}),
    (4, (131, 16, 64), {"xchunk": 20, "halo_pull": 0, "exact": 0}),
    (3, (96, 36, 72), {"tb2_lean": 0}),
])
def test_slabs_with_two_step_passes(pkg, oracle, nparts, shape, opts):
    """Several slabs on ONE device (peer pointer = local pointer): the two outermost planes of u^{n+1} and the
    four outermost planes of u^{n+2} go to the neighbour's ghost planes from inside the kernel; sources on and
    next to the seams are injected by both sides.  Bit-identical to the single-slab run and to the oracle."""
    T, S = 12, 9
    u, m, src, crd = fused_case(31 + nparts, shape, T, S, seam_parts=nparts)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    opts = dict(opts, t_fuse=2)
    one, _, info = run_plan(pkg, u, m, src, crd, options=dict(opts, kernel=2))
    assert info["t_fuse_used"] == 2
    ls = pkg.LocalSlabs(shape[0], shape[1], shape[2], [0] * nparts, options=opts)
    ls.upload(u, m)
    ls.set_sources(src, crd)
    t = ls.run(0, T - 1)
    used = [p.get_option("t_fuse_used") for p in ls.plans]
    out = np.zeros_like(u)
    ls.download(out)
    ls.close()
    assert used == [2] * nparts
    assert bits_equal(out, one), "slab run differs from the single-slab run"
    if opts.get("exact", 1):
        assert bits_equal(out, ref)
    else:
        assert oracle.rel_l2(out, ref) < REL_L2_TOL
    assert t.section0 > 0


@pytest.mark.parametrize("pull", [-1, 0])
def test_slabs_two_step_restart(pkg, oracle, pull):
    """A run split in two calls; the second call runs ONE-step launches in the push protocol, which reads the slabs' own ghost
    planes: a pull run must have refreshed them at its end."""
    shape, T, S = (64, 16, 64), 15, 4
    u, m, src, crd = fused_case(9, shape, T, S, seam_parts=2)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port")
    ls = pkg.LocalSlabs(*shape, [0, 0], options={"t_fuse": 2, "halo_pull": pull})
    ls.upload(u, m)
    ls.set_sources(src, crd)
    ls.run(0, 6)
    for p in ls.plans:
        p.set_option("t_fuse", 1)
    ls.run(7, 10)
    for p in ls.plans:
        p.set_option("t_fuse", 2)
    ls.run(11, T - 1)
    out = np.zeros_like(u)
    ls.download(out)
    ls.close()
    assert bits_equal(out, ref)


def test_512_two_step_equals_one_step(pkg):
    """BASELINE configs[2] at full size: two-step passes reproduce the one-step result bit for bit (exact mode),
    i.e. temporal blocking changes the traffic, not the numbers."""
    n, T = 512, 50
    src, crd = pkg.fill_ricker(T, 1), pkg.fill_source_coords(1, n, n, n)
    res = {}
    with pkg.Plan(n, n, n, deviceid=0) as p:
        p.set_sources(src, crd)
        for tf in (1, 2):
            p.set_option("t_fuse", tf)
            p.fill(0.0, 1.5)
            p.run(0, T - 1)
            assert p.get_option("t_fuse_used") == tf
            res[tf] = p.download()
    assert float(np.abs(res[1]).max()) == pytest.approx(0.116841748, rel=1e-6)
    assert bits_equal(res[1], res[2])


def test_512_bench_headline_configuration_within_tolerance(pkg, oracle):
    """bench.py's headline configuration (contracted arithmetic, two time steps per launch) against the bit-exact
    one-step run at the full BASELINE size: relative L2 < 1e-4 (README.md:33) and max-abs <= 1e-5 x peak |u|
    (BASELINE.json north_star), denormals kept."""
    n, T = 512, 50
    src, crd = pkg.fill_ricker(T, 1), pkg.fill_source_coords(1, n, n, n)
    with pkg.Plan(n, n, n, deviceid=0) as p:
        p.set_sources(src, crd)
        p.fill(0.0, 1.5)
        p.run(0, T - 1)
        ref = p.download()
        p.set_option("exact", 0)
        p.set_option("t_fuse", 2)
        p.fill(0.0, 1.5)
        p.run(0, T - 1)
        assert p.get_option("t_fuse_used") == 2 and p.get_option("exact") == 0
        out = p.download()
    # the wavefield lives in the low 168^3 corner (test_512_properties); everything else is exactly zero in both
    w = slice(0, 176)
    assert not out[:, 176:].any() and not ref[:, 176:].any()
    a, b = out[:, w, w, w].astype(np.float64), ref[:, w, w, w].astype(np.float64)
    rel = np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum())
    assert rel < REL_L2_TOL, rel
    assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()
    den = (np.abs(ref) < np.finfo(np.float32).tiny) & (ref != 0)
    assert den.any() and np.count_nonzero(out[den]) > 0.9 * den.sum()


@pytest.mark.parametrize("exact", [1, 0])
@pytest.mark.parametrize("t_fuse", [1, 2])
def test_dense_256_random_field(pkg, oracle, exact, t_fuse):
    """BASELINE configs[1] size with a DENSE random field and model (every cell significant, unlike the benchmark's
    spike), sources included: the TMA kernels in both arithmetic modes, one and two steps per launch.
    exact: 0 ulp vs the oracle; contracted: relative L2 < 1e-4 AND max-abs <= 1e-5 x peak |u| (north_star)."""
    shape, T, S = (256, 256, 256), 12, 6
    u, m, src, crd = fused_case(77, shape, T, S)
    ref = u.copy()
    oracle.run(ref, m, src, crd, impl="port", threads=8)
    out, t, info = run_plan(pkg, u, m, src, crd, options={"exact": exact, "t_fuse": t_fuse})
    assert info["kernel_used"] == 2 and info["t_fuse_used"] == t_fuse
    if exact:
        assert bits_equal(out, ref)
    else:
        assert oracle.rel_l2(out, ref) < REL_L2_TOL
        assert float(np.abs(out - ref).max()) <= 1e-5 * float(np.abs(ref).max())
