#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref/libref_openacc.so, i.e.
/root/reference/openacc.cpp compiled for the host by oracle/Makefile).

Run here (the build container), commit the outputs.  The GPU box has no /root/reference; the
tests only read the committed fixtures.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def stats(u):
    u64 = u.astype(np.float64)
    return {
        "sha256": sha(u),
        "max_abs": float(np.abs(u).max()),
        "l2": float(np.sqrt((u64 ** 2).sum())),
        "level_max_abs": [float(np.abs(u[i]).max()) for i in range(3)],
        "nonzero": int(np.count_nonzero(u)),
    }


def random_case(rng, shape, nsrc, T, with_halo_sources):
    """Seeded random fields; sources scattered over the box, some hugging / leaving the boundary."""
    nx, ny, nz = shape
    u = rng.uniform(-1, 1, (3, nx + 8, ny + 8, nz + 8)).astype(np.float32)
    m = rng.uniform(0.5, 3.0, (nx + 8, ny + 8, nz + 8)).astype(np.float32)
    src = rng.uniform(-50, 50, (T, nsrc)).astype(np.float32)
    h = np.float32(0.1)
    ext = np.array([nx - 1, ny - 1, nz - 1], np.float32) * h
    crd = (rng.uniform(0, 1, (nsrc, 3)).astype(np.float32) * ext).astype(np.float32)
    if with_halo_sources and nsrc >= 6:
        crd[0] = [0.0, 0.0, 0.0]                       # exactly on the first cell
        crd[1] = ext                                   # exactly on the last cell: +1 corners land in the halo
        crd[2] = [-0.05, 0.31, 0.52]                   # pos = -1 in x: only the rx=1 corners are in range
        crd[3] = ext + np.float32(0.04)                # pos = last cell, frac > 0: writes halo cells
        crd[4] = crd[5]                                # two sources in the same cell (ordering matters)
        crd[4, 2] += np.float32(0.003)
    return u, m, src, crd


def main():
    assert O.have_reference(), "build oracle/_ref first: make -C oracle"
    meta = {"generator": "tests/golden/make_golden.py", "source": "/root/reference/openacc.cpp (g++ -O3 -ffp-contract=off, oracle/stub/openacc.h)"}
    arrays = {}

    # ---- benchmark configs of main.cpp:279-356 (zero field, m = 1.5, Ricker, lattice sources)
    for name, (n, T, S) in {"bench64_s1": (64, 50, 1), "bench64_s64": (64, 50, 64),
                            "bench32_s27": (32, 50, 27), "bench128_T200_s64": (128, 200, 64),
                            "bench256_s1": (256, 50, 1)}.items():
        u = np.zeros((3, n + 8, n + 8, n + 8), np.float32)
        m = np.full((n + 8,) * 3, 1.5, np.float32)
        src, crd = O.fill_ricker(T, S), O.fill_source_coords(S, n, n, n)
        O.run(u, m, src, crd, impl="reference")
        meta[name] = {"n": n, "T": T, "S": S, **stats(u)}
        if n == 64 and S == 1:  # window around the source (pos = 15): the whole support is inside
            arrays["bench64_s1_window"] = u[:, 4 + 15 - 24 + 9:4 + 15 + 26, 4:4 + 42, 4:4 + 42].copy()
            meta[name]["window"] = [4 + 15 - 24 + 9, 4 + 15 + 26, 4, 46, 4, 46]
        print(name, meta[name]["max_abs"], meta[name]["l2"])

    # ---- dense parity field of main.cpp:525-570 (non-zero halos, h = 1, no sources)
    for n in (16, 32):
        u, m = O.fill_dense(n, n, n)
        O.run(u, m, time_M=49, h=1.0, impl="reference")
        meta[f"dense{n}"] = {"n": n, "T": 50, **stats(u)}
        if n == 16:
            arrays["dense16_u"] = u.copy()
        print(f"dense{n}", meta[f"dense{n}"]["max_abs"])

    # ---- seeded random fields, variable m, boundary-hugging and coincident sources
    rng = np.random.default_rng(1234)
    u, m, src, crd = random_case(rng, (16, 16, 16), 8, 14, True)
    arrays.update(rand16_u_in=u.copy(), rand16_m=m, rand16_src=src, rand16_crd=crd)
    O.run(u, m, src, crd, impl="reference")
    arrays["rand16_u_out"] = u.copy()
    meta["rand16"] = {"shape": [16, 16, 16], "T": 14, "S": 8, **stats(u)}

    # non-cubic, not a multiple of 4, ring phase time_m = 4, sub-range of sources
    u, m, src, crd = random_case(rng, (13, 10, 7), 6, 12, True)
    arrays.update(odd_u_in=u.copy(), odd_m=m, odd_src=src, odd_crd=crd)
    O.run(u, m, src, crd, impl="reference", time_m=4, time_M=11, p_src_m=1, p_src_M=4)
    arrays["odd_u_out"] = u.copy()
    meta["odd"] = {"shape": [13, 10, 7], "time_m": 4, "time_M": 11, "p_src_m": 1, "p_src_M": 4, **stats(u)}

    # ---- Ricker wavelet bits (main.cpp:290-298) and lattice source positions (openacc.cpp:125-131)
    arrays["ricker_T200"] = O.fill_ricker(200, 1)[:, 0].copy()
    pos_rows = []
    for n in (32, 64, 96, 128, 192, 256, 384, 512, 640, 768, 1024, 2048):
        c = O.fill_source_coords(27, n, n, n)
        for s in (0, 13, 26):
            pos, frac = O.source_pos(float(c[s, 0]), 0.0, 0.1)
            pos_rows.append([n, s, int(np.float32(c[s, 0]).view(np.uint32)), pos, int(np.float32(frac).view(np.uint32))])
    arrays["lattice_pos"] = np.array(pos_rows, np.int64)
    rc = rng.uniform(-3, 80, 64).astype(np.float32)
    ro = rng.uniform(-1, 1, 64).astype(np.float32)
    rh = rng.uniform(0.05, 2.0, 64).astype(np.float32)
    pf = [O.source_pos(float(a), float(b), float(c)) for a, b, c in zip(rc, ro, rh)]
    arrays.update(pos_coord=rc, pos_o=ro, pos_h=rh, pos_pos=np.array([p[0] for p in pf], np.int32),
                  pos_frac=np.array([p[1] for p in pf], np.float32))

    np.savez_compressed(os.path.join(OUT, "golden.npz"), **arrays)
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", os.path.getsize(os.path.join(OUT, "golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
