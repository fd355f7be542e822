"""CPU tests of the x-slab host logic with torch.distributed (gloo, world_size 2 and 3): partition,
slab views, per-slab source ownership (the product's host-only scatter table) and the ghost-plane
exchange pattern.  Each rank advances its slab with the ORACLE as the compute step (test only -- the
product has no CPU path) and the assembled result must be bit-identical to the single-domain oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT, bits_equal


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(seed, shape, T, S):
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    u = rng.uniform(-1, 1, (3, nx + 8, ny + 8, nz + 8)).astype(np.float32)
    m = rng.uniform(0.5, 3.0, (nx + 8, ny + 8, nz + 8)).astype(np.float32)
    src = rng.uniform(-20, 20, (T, S)).astype(np.float32)
    crd = (rng.uniform(-0.06, 1.06, (S, 3)) * (np.array(shape, np.float32) - 1) * np.float32(0.1)).astype(np.float32)
    return u, m, src, crd


def _worker(rank, world, port, shape, T, S, seam_sources, out_path):
    import importlib
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    pkg = importlib.import_module(PKG_NAME)
    from oracle import oracle as O

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nxg, ny, nz = shape
    u_g, m_g, src, crd = _case(77, shape, T, S)
    parts = pkg.partition(nxg, world)
    if seam_sources:  # put sources exactly on / next to every slab seam (SURVEY hard part e)
        for i, (off, nx) in enumerate(parts[1:]):
            crd[2 * i % S, 0] = np.float32((off - 1) * 0.1) + np.float32(0.03)   # base on the left slab, +1 corner on the right
            crd[(2 * i + 1) % S, 0] = np.float32(off * 0.1)
    off, nx = parts[rank]
    u = pkg.slab.slab_view(u_g, off, nx)
    m = pkg.slab.slab_view(m_g, off, nx)
    geom = pkg.Geometry(nx, ny, nz, off, nxg, 1e-3, 0.1, 0.1, 0.1, 0.0, 0.0, 0.0, -1)
    cells, n_int, cp, cw, base = pkg.slab_source_cells(geom, crd)
    mflat = m.reshape(-1)
    X0, X1 = 4, 4 + nx
    for time in range(T):
        t0, t1, t2 = time % 3, (time + 2) % 3, (time + 1) % 3
        # Section0 on the slab's own planes (the oracle, extents = local interior), no sources inside
        O.run(u, m, time_m=time, time_M=time, impl="port")
        # Section1 from the product's slab table: owned cells only, contributions in p_src order
        for X, Y, Z, first, count in cells:
            v = u[t2, X, Y, Z]
            for k in range(first, first + count):
                p = cp[k]
                v = np.float32(v + np.float32(np.float32(cw[k] * src[time, p]) / mflat[base[p]]))
            u[t2, X, Y, Z] = v
        # ghost-plane exchange of the new level: two boundary planes to each neighbour
        reqs, bufs = [], {}
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(u[t2, X0:X0 + 2].copy()), rank - 1))
            bufs["lo"] = torch.empty((2, ny + 8, nz + 8), dtype=torch.float32)
            reqs.append(dist.irecv(bufs["lo"], rank - 1))
        if rank < world - 1:
            reqs.append(dist.isend(torch.from_numpy(u[t2, X1 - 2:X1].copy()), rank + 1))
            bufs["hi"] = torch.empty((2, ny + 8, nz + 8), dtype=torch.float32)
            reqs.append(dist.irecv(bufs["hi"], rank + 1))
        for r in reqs:
            r.wait()
        if "lo" in bufs:
            u[t2, X0 - 2:X0] = bufs["lo"].numpy()
        if "hi" in bufs:
            u[t2, X1:X1 + 2] = bufs["hi"].numpy()
    gathered = [None] * world
    dist.all_gather_object(gathered, u)
    if rank == 0:
        out = np.zeros_like(u_g)
        pkg.slab.assemble(out, gathered, parts)
        ref = u_g.copy()
        O.run(ref, m_g, src, crd, impl="port")
        np.save(out_path, np.array([bits_equal(out, ref), float(np.abs(out - ref).max())]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,seam", [(2, (12, 8, 8), False), (2, (16, 8, 12), True), (3, (17, 8, 8), True)])
def test_slab_decomposition_matches_single_domain(pkg, oracle, tmp_path, world, shape, seam):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, _free_port(), shape, 7, 6, seam, out), nprocs=world, join=True)
    ok, err = np.load(out)
    assert ok == 1.0, f"slab run differs from the single-domain oracle (max abs {err})"


def test_partition_and_views(pkg):
    assert pkg.partition(1024, 8) == [(i * 128, 128) for i in range(8)]
    assert pkg.partition(10, 3) == [(0, 4), (4, 3), (7, 3)]
    with pytest.raises(ValueError):
        pkg.partition(5, 3)
    g = np.arange(3 * 18 * 9 * 9, dtype=np.float32).reshape(3, 18, 9, 9)
    parts = pkg.partition(10, 3)
    slabs = [pkg.slab.slab_view(g, off, nx) for off, nx in parts]
    assert [s.shape[1] for s in slabs] == [12, 11, 11]
    assert np.array_equal(slabs[1][:, 4], g[:, 8])          # local plane 4 = first owned plane = global padded 4 + 4
    assert np.array_equal(slabs[1][:, 2:4], slabs[0][:, 6:8])  # my lower ghosts = the neighbour's last owned planes
    out = np.zeros_like(g)
    pkg.slab.assemble(out, slabs, parts)
    assert np.array_equal(out, g)


def test_slab_source_ownership_partitions_the_global_table(pkg):
    """Every (cell, source) pair of the single-domain table is owned by exactly one slab."""
    rng = np.random.default_rng(5)
    nxg, ny, nz = 24, 10, 12
    crd = (rng.uniform(-0.06, 1.06, (40, 3)) * (np.array([nxg, ny, nz], np.float32) - 1) * np.float32(0.1)).astype(np.float32)
    crd[0, 0] = np.float32(7 * 0.1) + np.float32(0.05)   # straddles the seam between slabs [0,8) and [8,16)
    crd[1, 0] = np.float32(-0.04)                        # pos = -1: lower physical halo
    crd[2, 0] = np.float32((nxg - 1) * 0.1) + np.float32(0.02)  # +1 corner in the upper physical halo
    whole = pkg.Geometry(nxg, ny, nz, 0, nxg, 1e-3, 0.1, 0.1, 0.1, 0, 0, 0, -1)
    cells, n_int, cp, cw, base = pkg.slab_source_cells(whole, crd)
    want = {(int(X), int(Y), int(Z), int(cp[k]), float(cw[k])) for X, Y, Z, f, c in cells for k in range(f, f + c)}
    got = []
    for off, nx in pkg.partition(nxg, 3):
        g = pkg.Geometry(nx, ny, nz, off, nxg, 1e-3, 0.1, 0.1, 0.1, 0, 0, 0, -1)
        c2, ni2, cp2, cw2, b2 = pkg.slab_source_cells(g, crd)
        for X, Y, Z, f, c in c2:
            assert (4 <= X < 4 + nx) or (off == 0 and X == 3) or (off + nx == nxg and X == 4 + nx)
            got += [(int(X) + off, int(Y), int(Z), int(cp2[k]), float(cw2[k])) for k in range(f, f + c)]
        # fused cells come first and are sorted by plane
        assert all(4 <= X < 4 + nx and 4 <= Y < 4 + ny and 4 <= Z < 4 + nz for X, Y, Z, _, _ in c2[:ni2])
        assert list(c2[:ni2, 0]) == sorted(c2[:ni2, 0])
    assert len(got) == len(set(got)) == len(want) and set(got) == want


# ------------------------------------------------------------------ two-step launches: ghost-zone source cells, 4-plane exchange
def test_two_step_source_table_covers_the_neighbours_nearest_planes(pkg):
    """A two-step launch recomputes the neighbour slabs' two nearest planes of u^{n+1}, so its table holds the
    slab's interior cells plus the global table's cells on those planes -- same (source, weight) lists."""
    rng = np.random.default_rng(8)
    nxg, ny, nz = 30, 10, 12
    crd = (rng.uniform(0.03, 0.96, (60, 3)) * (np.array([nxg, ny, nz], np.float32) - 1) * np.float32(0.1)).astype(np.float32)
    for k, x in enumerate((9, 10, 8, 11, 19, 20, 18, 21)):  # on and next to the seams of [0,10) [10,20) [20,30)
        crd[k, 0] = np.float32(x * 0.1) + np.float32(0.03)
    whole = pkg.Geometry(nxg, ny, nz, 0, nxg, 1e-3, 0.1, 0.1, 0.1, 0, 0, 0, -1)
    cells, cp, cw, halo = pkg.slab_source_cells2(whole, crd)
    c1, n_int, cp1, cw1, _ = pkg.slab_source_cells(whole, crd)
    assert not halo and len(cells) == n_int == len(c1)  # one slab: exactly the interior cells
    glob = {}
    for X, Y, Z, f, c in cells:
        glob[(int(X), int(Y), int(Z))] = [(int(cp[k]), float(cw[k])) for k in range(f, f + c)]
    for r, (off, nx) in enumerate(pkg.partition(nxg, 3)):
        g = pkg.Geometry(nx, ny, nz, off, nxg, 1e-3, 0.1, 0.1, 0.1, 0, 0, 0, -1)
        c2, cp2, cw2, halo2 = pkg.slab_source_cells2(g, crd)
        assert not halo2
        lo = 4 - (2 if r > 0 else 0)
        hi = 4 + nx + (2 if r < 2 else 0)
        want = {k: v for k, v in glob.items() if lo <= k[0] - off < hi}
        got = {(int(X) + off, int(Y), int(Z)): [(int(cp2[k]), float(cw2[k])) for k in range(f, f + c)] for X, Y, Z, f, c in c2}
        assert got == want
        assert [tuple(c[:3]) for c in c2] == sorted(tuple(c[:3]) for c in c2)  # plane order: the kernel indexes by plane
    # a corner in a halo cell anywhere is seen by EVERY slab (they must all fall back to one step per launch)
    crd[5, 1] = np.float32(-0.04)
    for off, nx in pkg.partition(nxg, 3):
        g = pkg.Geometry(nx, ny, nz, off, nxg, 1e-3, 0.1, 0.1, 0.1, 0, 0, 0, -1)
        assert pkg.slab_source_cells2(g, crd)[3]


def _worker_two_step(rank, world, port, shape, T, S, out_path):
    """Each rank advances its slab in two-step passes with the ORACLE as the compute step (test only): step n on the
    slab's planes plus the neighbours' two nearest planes (sources from the product's two-step table), step n+1 on
    its own planes, then 2 planes of u^{n+1} and 4 planes of u^{n+2} go to each neighbour."""
    import importlib
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    pkg = importlib.import_module(PKG_NAME)
    from oracle import oracle as O

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nxg, ny, nz = shape
    rng = np.random.default_rng(123)
    u_g = rng.uniform(-1, 1, (3, nxg + 8, ny + 8, nz + 8)).astype(np.float32)
    m_g = rng.uniform(0.5, 3.0, (nxg + 8, ny + 8, nz + 8)).astype(np.float32)
    src = rng.uniform(-20, 20, (T, S)).astype(np.float32)
    crd = (rng.uniform(0.05, 0.95, (S, 3)) * (np.array(shape, np.float32) - 1) * np.float32(0.1)).astype(np.float32)
    parts = pkg.partition(nxg, world)
    for i, (off, _) in enumerate(parts[1:]):  # sources on, and one / two planes off, every seam
        for j, dx in enumerate((-1, 0, -2, 1)):
            crd[(4 * i + j) % S, 0] = np.float32((off + dx) * 0.1) + np.float32(0.03)
    off, nx = parts[rank]
    u = pkg.slab.slab_view(u_g, off, nx)
    m = pkg.slab.slab_view(m_g, off, nx)
    geom = pkg.Geometry(nx, ny, nz, off, nxg, 1e-3, 0.1, 0.1, 0.1, 0.0, 0.0, 0.0, -1)
    cells2, cp, cw, halo = pkg.slab_source_cells2(geom, crd)
    assert not halo
    # base corner of every source in LOCAL padded coordinates (the ghost planes hold the neighbours' m)
    base = {}
    for p in range(S):
        pos, _, _, _ = pkg.source_table(crd[p], (0, 0, 0), (0.1,) * 3, (0, 0, 0), (nxg - 1, ny - 1, nz - 1))
        base[p] = (int(pos[0]) - off + 4, int(pos[1]) + 4, int(pos[2]) + 4)
    X0, X1 = 4, 4 + nx
    lo2 = 2 if rank > 0 else 0
    hi2 = 2 if rank < world - 1 else 0

    def inject(level, time, planes):
        for X, Y, Z, first, count in cells2:
            if not (planes[0] <= X < planes[1]):
                continue
            v = u[level, X, Y, Z]
            for k in range(first, first + count):
                p = int(cp[k])
                v = np.float32(v + np.float32(np.float32(cw[k] * src[time, p]) / m[base[p]]))
            u[level, X, Y, Z] = v

    def exchange(level, depth):
        reqs, bufs = [], {}
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(u[level, X0:X0 + depth].copy()), rank - 1))
            bufs["lo"] = torch.empty((depth, ny + 8, nz + 8), dtype=torch.float32)
            reqs.append(dist.irecv(bufs["lo"], rank - 1))
        if rank < world - 1:
            reqs.append(dist.isend(torch.from_numpy(u[level, X1 - depth:X1].copy()), rank + 1))
            bufs["hi"] = torch.empty((depth, ny + 8, nz + 8), dtype=torch.float32)
            reqs.append(dist.irecv(bufs["hi"], rank + 1))
        for r in reqs:
            r.wait()
        if "lo" in bufs:
            u[level, X0 - depth:X0] = bufs["lo"].numpy()
        if "hi" in bufs:
            u[level, X1:X1 + depth] = bufs["hi"].numpy()

    time = 0
    while time < T:
        t2 = (time + 1) % 3
        if time + 1 < T:  # a two-step pass
            # step n on [X0-2, X1+2) towards neighbours: extents are unpadded, local x index = padded - 4
            O.run(u, m, time_m=time, time_M=time, impl="port", extents=(-lo2, nx - 1 + hi2, 0, ny - 1, 0, nz - 1))
            inject(t2, time, (X0 - lo2, X1 + hi2))
            O.run(u, m, time_m=time + 1, time_M=time + 1, impl="port")
            inject((time + 2) % 3, time + 1, (X0, X1))
            exchange(t2, 2)
            exchange((time + 2) % 3, 4)
            time += 2
        else:  # odd remainder: one step, 4 planes out (a pass may follow in a later run)
            O.run(u, m, time_m=time, time_M=time, impl="port")
            inject(t2, time, (X0, X1))
            exchange(t2, 4)
            time += 1
    gathered = [None] * world
    dist.all_gather_object(gathered, u)
    if rank == 0:
        out = np.zeros_like(u_g)
        pkg.slab.assemble(out, gathered, parts)
        ref = u_g.copy()
        O.run(ref, m_g, src, crd, impl="port")
        np.save(out_path, np.array([bits_equal(out, ref), float(np.abs(out - ref).max())]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,T", [(2, (16, 8, 12), 7), (3, (21, 8, 8), 6)])
def test_two_step_slab_passes_match_single_domain(pkg, oracle, tmp_path, world, shape, T):
    out = str(tmp_path / "res2.npy")
    mp.spawn(_worker_two_step, args=(world, _free_port(), shape, T, 9, out), nprocs=world, join=True)
    ok, err = np.load(out)
    assert ok == 1.0, f"two-step slab passes differ from the single-domain oracle (max abs {err})"
