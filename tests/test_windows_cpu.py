"""CPU tests of the cropped-window checker (oracle/windows.py): the oracle on a grid cut out around the sources,
with coordinates that reproduce the same (pos - offset, frac) bits, equals the oracle on the full grid bit for bit."""
import numpy as np

from conftest import bits_equal


def test_shifted_coordinates_reproduce_fraction_bits(oracle):
    from oracle import windows as W

    for n in (512, 1024, 2048, 4096):
        for S in (1, 27, 64):
            crd = oracle.fill_source_coords(S, n, 512, 512) if n == 4096 else oracle.fill_source_coords(S, n, n, n)
            shape = (n, 512, 512) if n == 4096 else (n, n, n)
            for w in W.source_windows(crd, shape):
                for k, p in enumerate(w["sources"]):
                    for a in range(3):
                        pos, frac = oracle.source_pos(float(crd[p, a]), 0.0, 0.1)
                        cpos, cfrac = oracle.source_pos(float(w["coords"][k, a]), 0.0, 0.1)
                        assert cpos == pos - w["off"][a]
                        assert np.float32(cfrac).view(np.uint32) == np.float32(frac).view(np.uint32)
                assert all(o >= 0 and o + s <= n_ for o, s, n_ in zip(w["off"], w["size"], shape))


def test_cropped_oracle_equals_full_oracle(oracle):
    from oracle import windows as W

    n, T, S = 200, 40, 27  # lattice spacing 50 cells: windows with half = 20 stay apart -> 27 crops
    u = np.zeros((3, n + 8, n + 8, n + 8), np.float32)
    m = np.full((n + 8,) * 3, 1.5, np.float32)
    src, crd = oracle.fill_ricker(T, S), oracle.fill_source_coords(S, n, n, n)
    oracle.run(u, m, src, crd, threads=8)
    wins = W.source_windows(crd, (n, n, n), half=20)
    assert len(wins) == 27
    nz = 0
    for w in wins:
        r = W.run_window(w, src)
        (ox, oy, oz), (wx, wy, wz) = w["off"], w["size"]
        got = u[:, ox + 4:ox + wx + 4, oy + 4:oy + wy + 4, oz + 4:oz + wz + 4]
        assert bits_equal(got, r[:, 4:-4, 4:-4, 4:-4])
        nz += int(np.count_nonzero(got))
    assert nz == int(np.count_nonzero(u))  # nothing outside the windows
    # coincident sources share one window and keep their p_src order
    crd64 = oracle.fill_source_coords(64, n, n, n)
    w64 = W.source_windows(crd64, (n, n, n), half=20)
    assert len(w64) == 27 and sorted(len(w["sources"]) for w in w64)[-1] == 38
