"""x-slab domain decomposition of the FDTD grid (x = the reference's slowest axis, one plane is
(ny+8)*(nz+8) contiguous floats, so slabs need no packing).

Two ways to drive slabs:
  * ``SlabRun``    -- one process per GPU (torchrun); torch.distributed only carries the rendezvous:
                      the CUDA IPC blobs of the neighbours' arrays and the end-of-run barrier.  The halo
                      exchange itself is inside the stencil kernel (peer stores over NVLink + flags).
  * ``LocalSlabs`` -- one process driving several plans (several devices, or several slabs on one device
                      for tests) through fdtd_b200_run_slabs.
Host logic only; all compute goes to libfdtd_b200.so.
"""
from __future__ import annotations

import numpy as np

from . import host

HALO = host.HALO


def partition(nx_global: int, nparts: int):
    """[(x_offset, nx_local)] -- contiguous slabs, sizes differing by at most one plane."""
    if nparts < 1 or nx_global < 2 * nparts:
        raise ValueError("need at least 2 planes per slab")
    base, rem = divmod(nx_global, nparts)
    out, off = [], 0
    for r in range(nparts):
        n = base + (1 if r < rem else 0)
        out.append((off, n))
        off += n
    return out


def slab_view(global_arr: np.ndarray, x_offset: int, nx: int) -> np.ndarray:
    """The padded local array of a slab cut from the padded global array (ghost planes included).
    Works for u [3, nxp, nyp, nzp] and m [nxp, nyp, nzp]; local padded plane X = global padded X - x_offset."""
    if global_arr.ndim == 4:
        return np.ascontiguousarray(global_arr[:, x_offset:x_offset + nx + 2 * HALO])
    return np.ascontiguousarray(global_arr[x_offset:x_offset + nx + 2 * HALO])


def assemble(global_out: np.ndarray, slabs, parts):
    """Write the slabs' interior planes (and the physical halo planes of the two end slabs) into global_out."""
    last = len(parts) - 1
    for r, ((off, nx), s) in enumerate(zip(parts, slabs)):
        lo = 0 if r == 0 else HALO
        hi = nx + 2 * HALO if r == last else nx + HALO
        global_out[:, off + lo:off + hi] = s[:, lo:hi]
    return global_out


def merge_receivers(parts):
    """[(rec [rows, nrec], owned [nrec]) per slab] -> rec [rows, nrec]: each receiver belongs to exactly one slab."""
    rec = np.zeros_like(parts[0][0])
    seen = np.zeros(rec.shape[1], int)
    for r, owned in parts:
        rec[:, owned] = r[:, owned]
        seen += owned
    if rec.shape[1] and not np.all(seen == 1):
        raise ValueError("every receiver must be owned by exactly one slab")
    return rec


class SlabRun:
    """This rank's slab of a (nx_global, ny, nz) grid, neighbours attached through CUDA IPC."""

    def __init__(self, dist, nx_global, ny, nz, device, *, dt=1e-3, h=(0.1, 0.1, 0.1), o=(0.0, 0.0, 0.0)):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.parts = partition(nx_global, self.world)
        self.x_offset, self.nx = self.parts[self.rank]
        self.plan = host.Plan(self.nx, ny, nz, dt=dt, h=h, o=o, x_offset=self.x_offset, nx_global=nx_global,
                              deviceid=device)
        blobs = [None] * self.world
        dist.all_gather_object(blobs, self.plan.ipc_export())
        if self.rank > 0:
            self.plan.ipc_attach(0, blobs[self.rank - 1])
        if self.rank < self.world - 1:
            self.plan.ipc_attach(1, blobs[self.rank + 1])
        dist.barrier()

    def run(self, time_m: int, time_M: int):
        """All ranks advance the same steps.  The barrier makes sure every neighbour's last boundary
        planes have landed (and nobody refills a field a neighbour is still writing ghosts into)."""
        if self.plan.get_option("t_fuse") >= 2:
            # two-step passes only if every slab can run them (same pass schedule on all ranks)
            depths = [None] * self.world
            self.dist.all_gather_object(depths, self.plan.probe_fuse())
            self.plan.set_option("t_fuse_agreed", min(depths))
        t = self.plan.run(time_m, time_M)
        self.dist.barrier()
        return t

    def close(self):
        self.dist.barrier()
        self.plan.close()


class LocalSlabs:
    """Several slabs driven by this process (fdtd_b200_run_slabs)."""

    def __init__(self, nx_global, ny, nz, devices, *, dt=1e-3, h=(0.1, 0.1, 0.1), o=(0.0, 0.0, 0.0), options=None):
        self.parts = partition(nx_global, len(devices))
        self.plans = [host.Plan(nx, ny, nz, dt=dt, h=h, o=o, x_offset=off, nx_global=nx_global, deviceid=d)
                      for (off, nx), d in zip(self.parts, devices)]
        for p in self.plans:
            for k, v in (options or {}).items():
                p.set_option(k, v)
        for r, p in enumerate(self.plans):
            if r > 0:
                p.attach_local(0, self.plans[r - 1])
            if r < len(self.plans) - 1:
                p.attach_local(1, self.plans[r + 1])

    def upload(self, u_global, m_global):
        for (off, nx), p in zip(self.parts, self.plans):
            p.upload(slab_view(u_global, off, nx), slab_view(m_global, off, nx))

    def set_sources(self, src, coords, p_src_m=0, p_src_M=None):
        for p in self.plans:
            p.set_sources(src, coords, p_src_m, p_src_M)

    def fill(self, u_value=0.0, m_value=1.5):
        for p in self.plans:
            p.fill(u_value, m_value)

    def fill_dense(self):
        for p in self.plans:
            p.fill_dense()

    def run(self, time_m, time_M):
        return host.run_slabs(self.plans, time_m, time_M)

    def download(self, out):
        return assemble(out, [p.download() for p in self.plans], self.parts)

    def set_receivers(self, coords):
        for p in self.plans:
            p.set_receivers(coords)

    def receivers(self):
        """Traces of the last run, every column taken from the slab that owns the receiver."""
        return merge_receivers([p.receivers() for p in self.plans])

    def close(self):
        for p in reversed(self.plans):  # a slab may share the stream of the plan created before it
            p.close()
