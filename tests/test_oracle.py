"""CPU tests: the oracle restatement (oracle/fdtd_oracle.c) against the golden fixtures generated
from the UNMODIFIED reference (tests/golden/make_golden.py) and, when it is built, against the
compiled reference itself (oracle/_ref/libref_openacc.so)."""
import hashlib

import numpy as np
import pytest

from conftest import bench_inputs, bits_equal


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["bench64_s1", "bench64_s64", "bench32_s27"])
def test_port_matches_golden_benchmark(oracle, golden, name):
    meta, arrs = golden
    g = meta[name]
    u, m, src, crd = bench_inputs(oracle, g["n"], g["T"], g["S"])
    oracle.run(u, m, src, crd, impl="port")
    assert sha(u) == g["sha256"]                      # bit-exact with the reference build
    assert float(np.abs(u).max()) == g["max_abs"]
    assert [float(np.abs(u[i]).max()) for i in range(3)] == g["level_max_abs"]
    if name == "bench64_s1":
        w = g["window"]
        assert bits_equal(u[:, w[0]:w[1], w[2]:w[3], w[4]:w[5]], arrs["bench64_s1_window"])
        # everything outside the stored window is still exactly zero
        z = u.copy()
        z[:, w[0]:w[1], w[2]:w[3], w[4]:w[5]] = 0
        assert not z.any()


def test_survey_known_answers(golden):
    """Known answers recorded by the survey from its own host build of openacc.cpp."""
    meta, _ = golden
    assert np.float32(meta["bench64_s1"]["max_abs"]) == np.float32(0.116838083)
    assert abs(meta["bench64_s1"]["l2"] - 0.216876815) < 1e-9
    assert np.float32(meta["bench256_s1"]["max_abs"]) == np.float32(0.116838083)
    assert np.float32(meta["bench64_s64"]["max_abs"]) == np.float32(1.33085442)
    assert abs(meta["bench64_s64"]["l2"] - 6.02156338) < 1e-8
    assert np.float32(meta["bench128_T200_s64"]["max_abs"]) == np.float32(11.0107412)
    assert abs(meta["dense32"]["max_abs"] - 5609.83) < 0.01


def test_port_matches_golden_dense(oracle, golden):
    meta, arrs = golden
    u, m = oracle.fill_dense(16, 16, 16)
    oracle.run(u, m, time_M=49, h=1.0, impl="port")
    assert bits_equal(u, arrs["dense16_u"])
    assert sha(u) == meta["dense16"]["sha256"]
    u, m = oracle.fill_dense(32, 32, 32)
    oracle.run(u, m, time_M=49, h=1.0, impl="port")
    assert sha(u) == meta["dense32"]["sha256"]


def test_port_matches_golden_random(oracle, golden):
    meta, arrs = golden
    u = arrs["rand16_u_in"].copy()
    oracle.run(u, arrs["rand16_m"], arrs["rand16_src"], arrs["rand16_crd"], impl="port")
    assert bits_equal(u, arrs["rand16_u_out"])
    # non-cubic / not a multiple of 4 / ring phase time_m = 4 / source sub-range
    g = meta["odd"]
    u = arrs["odd_u_in"].copy()
    oracle.run(u, arrs["odd_m"], arrs["odd_src"], arrs["odd_crd"], impl="port", time_m=g["time_m"],
               time_M=g["time_M"], p_src_m=g["p_src_m"], p_src_M=g["p_src_M"])
    assert bits_equal(u, arrs["odd_u_out"])


def test_ricker_and_positions(oracle, golden):
    _, arrs = golden
    assert bits_equal(oracle.fill_ricker(200, 1)[:, 0], arrs["ricker_T200"])
    r = arrs["ricker_T200"].view(np.uint32)
    assert [hex(r[i]) for i in (0, 1, 5, 49, 99, 100, 199)] == [
        "0xba7e1556", "0xba975f0e", "0xbb153327", "0xbea279b4", "0x3f7f3e1e", "0x3f800000", "0xba975ef4"]
    for n, s, cbits, pos, fbits in arrs["lattice_pos"]:
        c = oracle.fill_source_coords(27, int(n), int(n), int(n))
        assert int(c[s, 0].view(np.uint32)) == cbits
        p, f = oracle.source_pos(float(c[s, 0]), 0.0, 0.1)
        assert (p, int(np.float32(f).view(np.uint32))) == (pos, fbits)
    for c, o, h, pos, frac in zip(arrs["pos_coord"], arrs["pos_o"], arrs["pos_h"], arrs["pos_pos"], arrs["pos_frac"]):
        p, f = oracle.source_pos(float(c), float(o), float(h))
        assert p == pos and np.float32(f).view(np.uint32) == np.float32(frac).view(np.uint32)


def test_omp_build_is_bit_identical(oracle):
    u, m, src, crd = bench_inputs(oracle, 32, 20, 27)
    v = u.copy()
    oracle.run(u, m, src, crd, impl="port")
    oracle.run(v, m, src, crd, impl="port", threads=4)
    assert bits_equal(u, v)


def test_port_matches_compiled_reference(oracle):
    """Direct check against /root/reference/openacc.cpp built for the host (skipped where it is absent)."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    rng = np.random.default_rng(7)
    for shape, T, S in (((24, 20, 28), 9, 5), ((8, 8, 8), 7, 3)):
        nx, ny, nz = shape
        u = rng.uniform(-1, 1, (3, nx + 8, ny + 8, nz + 8)).astype(np.float32)
        m = rng.uniform(0.5, 3, (nx + 8, ny + 8, nz + 8)).astype(np.float32)
        src = rng.uniform(-9, 9, (T, S)).astype(np.float32)
        crd = (rng.uniform(-0.02, 1.02, (S, 3)) * (np.array(shape) - 1) * 0.1).astype(np.float32)
        a, b, c = u.copy(), u.copy(), u.copy()
        oracle.run(a, m, src, crd, impl="port", time_m=2, time_M=T - 1)
        oracle.run(b, m, src, crd, impl="reference", time_m=2, time_M=T - 1)
        oracle.run(c, m, src, crd, impl="reference", time_m=2, time_M=T - 1, threads=3)
        assert bits_equal(a, b) and bits_equal(a, c)


def test_order_generalisation_is_pinned_at_order_4(oracle):
    """oracle_run_order (orders 4..12, receivers) is this repo's definition; at order 4 without
    receivers it must be bit-identical to the pinned oracle_run, and its weights must be the reference's literals."""
    rng = np.random.default_rng(3)
    shape, T, S = (14, 11, 18), 9, 5
    u = rng.uniform(-1, 1, (3,) + tuple(n + 8 for n in shape)).astype(np.float32)
    m = rng.uniform(0.5, 3.0, u.shape[1:]).astype(np.float32)
    src = rng.uniform(-20, 20, (T, S)).astype(np.float32)
    crd = (rng.uniform(-0.04, 1.04, (S, 3)) * (np.array(shape, np.float32) - 1) * np.float32(0.1)).astype(np.float32)
    a, b = u.copy(), u.copy()
    oracle.run(a, m, src, crd)
    _, _, rec = oracle.run_order(b, m, src, crd, space_order=4, rec_coords=crd)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert rec.shape == (T, S) and np.isfinite(rec).all() and np.abs(rec).max() > 0
    assert oracle.fd_coeffs(4).view(np.uint32).tolist() == [0xc0200000, 0x3faaaaab, 0xbdaaaaab]  # SURVEY 8a4 probe


def test_fd_weights_are_consistent(oracle):
    """Second-derivative consistency of every order: sum of weights 0, second moment 2, higher even moments 0."""
    for so in (4, 6, 8, 10, 12):
        c = oracle.fd_coeffs(so).astype(np.float64)
        R = so // 2
        k = np.arange(1, R + 1, dtype=np.float64)
        assert abs(c[0] + 2 * c[1:].sum()) < 1e-6
        assert abs(2 * (c[1:] * k ** 2).sum() - 2.0) < 1e-5
        for p in range(2, R + 1):
            assert abs((c[1:] * k ** (2 * p)).sum()) < 5e-3 * max(1.0, (np.abs(c[1:]) * k ** (2 * p)).sum())
