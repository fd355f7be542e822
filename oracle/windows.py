"""Cropped-window parity checks for fields that are too large for the oracle -- TEST INFRASTRUCTURE ONLY.

The reference's driver compares all three levels of u after a run (main.cpp:573-604).  At 512^3 .. 2048^3 the
oracle cannot run the whole grid in seconds, but the benchmark field (main.cpp:285-356: zero field, m = 1.5,
Ricker sources on a lattice) is LOCAL: after 50..200 steps its support (denormals included) stays within ~30
cells of a source, and the stencil is translation invariant.  So the oracle runs a small grid cut out around
every cluster of sources, with source coordinates chosen to reproduce the same (pos - offset, frac) BITS
(openacc.cpp:125-131), and the windows of the big run must equal it bit for bit; an order-independent checksum
of the whole field (fdtd_b200_plan_checksum) proves nothing else is non-zero anywhere.

Also here: a seeded dense random case with sources straddling every slab seam (full-grid oracle, small grids).
Used by tests/ and by bench.py's parity block (the checker, never the thing measured).
"""
from __future__ import annotations

import numpy as np

from . import oracle as O

HALO = 4


def _pos_frac(coord, h):
    pos, frac = O.source_pos(float(coord), 0.0, float(h))
    return pos, np.float32(frac)


def _shifted_coord(pos, frac, off, h):
    """A float32 coordinate c' with source_pos(c') == (pos - off, frac) bit for bit, or None."""
    if off == 0:
        return None
    want_bits = np.float32(frac).view(np.uint32)
    c = np.float32((np.float64(pos - off) + np.float64(frac)) * np.float64(np.float32(h)))
    cands = [c]
    lo = hi = c
    for _ in range(24):
        lo = np.nextafter(lo, np.float32(-np.inf), dtype=np.float32)
        hi = np.nextafter(hi, np.float32(np.inf), dtype=np.float32)
        cands += [lo, hi]
    for cand in cands:
        p, f = _pos_frac(cand, h)
        if p == pos - off and np.float32(f).view(np.uint32) == want_bits:
            return np.float32(cand)
    return None


def source_windows(coords, shape, h=0.1, half=46):
    """Cluster the sources and cut one oracle-sized grid around every cluster.

    Returns a list of windows: {"off": (ox, oy, oz) global unpadded offset of the crop's interior,
    "size": (wx, wy, wz) crop interior extents, "sources": [p, ...] ascending, "coords": float32 [k, 3] crop
    coordinates}.  `half` = cells kept on every side of the cluster (must exceed the support radius)."""
    coords = np.ascontiguousarray(coords, np.float32)
    S = coords.shape[0]
    pos = np.zeros((S, 3), np.int64)
    frac = np.zeros((S, 3), np.float32)
    for p in range(S):
        for a in range(3):
            pos[p, a], frac[p, a] = _pos_frac(coords[p, a], h)
    # union-find: sources whose windows would overlap belong to one cluster
    parent = list(range(S))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    for i in range(S):
        for j in range(i):
            if np.all(np.abs(pos[i] - pos[j]) <= 2 * half + 2):
                parent[find(i)] = find(j)
    clusters = {}
    for p in range(S):
        clusters.setdefault(find(p), []).append(p)
    out = []
    for members in clusters.values():
        members = sorted(members)
        lo = pos[members].min(axis=0) - half
        hi = pos[members].max(axis=0) + 1 + half + 1  # the +1 corner, exclusive
        off, size = [], []
        for a in range(3):
            n = shape[a]
            o = int(max(0, min(lo[a], n)))
            e = int(min(n, max(hi[a], 0)))
            if e - o < 8:  # a cluster outside the grid: keep a valid (empty-field) crop
                o, e = max(0, min(o, n - 8)), max(0, min(o, n - 8)) + 8
            off.append(o)
            size.append(e - o)
        # find coordinates that reproduce the same fractions on the shifted grid; nudge the offset if a rounding
        # of (c'/h) makes that impossible for some source
        for nudge in range(0, 12):
            trial = [max(0, o - nudge) if o > 0 else 0 for o in off]
            cc = np.array(coords[members], np.float32)
            ok = True
            for k, p in enumerate(members):
                for a in range(3):
                    if trial[a] == 0:
                        continue
                    c = _shifted_coord(int(pos[p, a]), frac[p, a], trial[a], h)
                    if c is None:
                        ok = False
                        break
                    cc[k, a] = c
                if not ok:
                    break
            if ok:
                size = [s + (o - t) for s, o, t in zip(size, off, trial)]
                off = trial
                break
        else:
            raise RuntimeError("no crop offset reproduces the source fractions")
        out.append({"off": tuple(off), "size": tuple(size), "sources": members, "coords": cc})
    return out


def run_window(win, src, m_value=1.5, threads=1, impl="port"):
    """The oracle on one crop of the benchmark field: zero u, constant m, the cluster's sources."""
    wx, wy, wz = win["size"]
    u = np.zeros((3, wx + 2 * HALO, wy + 2 * HALO, wz + 2 * HALO), np.float32)
    m = np.full(u.shape[1:], m_value, np.float32)
    s = np.ascontiguousarray(np.asarray(src, np.float32)[:, win["sources"]])
    O.run(u, m, s, win["coords"], impl=impl, threads=threads)
    return u


class WindowParity:
    """Accumulates the comparison of a resident (slab of a) field with oracle crops."""

    def __init__(self):
        self.bit_identical = True
        self.sq_err = 0.0
        self.sq_ref = 0.0
        self.max_abs_err = 0.0
        self.peak = 0.0
        self.cells = 0
        self.nonzero_in_windows = 0
        self.bit_sum_in_windows = 0

    def add(self, got, ref):
        self.cells += got.size
        if not np.array_equal(got.view(np.uint32), ref.view(np.uint32)):
            self.bit_identical = False
        d = got.astype(np.float64) - ref.astype(np.float64)
        self.sq_err += float(np.sum(d * d))
        self.sq_ref += float(np.sum(ref.astype(np.float64) ** 2))
        if got.size:
            self.max_abs_err = max(self.max_abs_err, float(np.abs(d).max()))
            self.peak = max(self.peak, float(np.abs(ref).max()))
        self.nonzero_in_windows += int(np.count_nonzero(got))
        self.bit_sum_in_windows += int(got.view(np.uint32).astype(np.uint64).sum())


def compare_windows(plan, x_offset, nx_local, windows, refs, acc=None):
    """Compare the part of every window this slab owns (interior planes [x_offset, x_offset + nx_local) of the
    global grid) with the oracle crops `refs`, all three ring levels.  plan: the product's Plan (download_window)."""
    acc = acc or WindowParity()
    for win, ref in zip(windows, refs):
        ox, oy, oz = win["off"]
        wx, wy, wz = win["size"]
        g0, g1 = max(ox, x_offset), min(ox + wx, x_offset + nx_local)  # global unpadded x range owned here
        if g1 <= g0:
            continue
        lw = (g0 - x_offset + HALO, g1 - x_offset + HALO, oy + HALO, oy + wy + HALO, oz + HALO, oz + wz + HALO)
        for lvl in range(3):
            got = plan.download_window(lvl, lw)
            acc.add(got, ref[lvl, g0 - ox + HALO:g1 - ox + HALO, HALO:wy + HALO, HALO:wz + HALO])
    return acc


# --------------------------------------------------------------------------- dense seam case
def dense_seam_case(seed, shape, T, S, nparts):
    """Seeded random field / model with ONE halo shell shared by the three levels (a Dirichlet boundary; what
    two-step passes need), sources strictly inside, some of them straddling / next to every slab seam."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    u = rng.uniform(-1, 1, (3, nx + 8, ny + 8, nz + 8)).astype(np.float32)
    inner = (slice(4, nx + 4), slice(4, ny + 4), slice(4, nz + 4))
    for lvl in (1, 2):
        keep = u[lvl][inner].copy()
        u[lvl] = u[0]
        u[lvl][inner] = keep
    m = rng.uniform(0.5, 3.0, (nx + 8, ny + 8, nz + 8)).astype(np.float32)
    src = rng.uniform(-20, 20, (T, S)).astype(np.float32)
    crd = (rng.uniform(0.03, 0.96, (S, 3)) * (np.array(shape, np.float32) - 1) * np.float32(0.1)).astype(np.float32)
    if S >= 3:
        crd[1] = crd[0]  # coincident sources: per-cell summation order
    base, i = nx // max(1, nparts), 2
    for k in range(1, nparts):
        for dx, fr in ((-1, 0.04), (0, 0.0), (-2, 0.03), (1, 0.02)):
            if i < S:
                crd[i, 0] = np.float32((k * base + dx) * 0.1) + np.float32(fr)
                i += 1
    return u, m, src, crd
