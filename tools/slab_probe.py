#!/usr/bin/env python
"""One linked-slab workload under torchrun (tuning aid): Gpts/s per mode for a global grid nxg x ny x nz.
    torchrun --nproc-per-node 2 tools/slab_probe.py --nxg 256 --ny 1024 --nz 1024 --T 60 [--modes 0:2,1:1]"""
import argparse, importlib, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("accelerated-3d-acoustic-fdtd-kernel_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--nxg", type=int, default=256)
ap.add_argument("--ny", type=int, default=1024)
ap.add_argument("--nz", type=int, default=1024)
ap.add_argument("--T", type=int, default=60)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--modes", default="0:2,1:1")
ap.add_argument("--xchunk", type=int, default=0)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sr = pkg.SlabRun(dist, a.nxg, a.ny, a.nz, local)
src, crd = pkg.fill_ricker(a.T, 1), pkg.fill_source_coords(1, a.nxg, a.ny, a.nz)
sr.plan.set_sources(src, crd)
for mode in a.modes.split(","):
    ex, tf = [int(x) for x in mode.split(":")]
    sr.plan.set_option("exact", ex)
    sr.plan.set_option("t_fuse", tf)
    sr.plan.set_option("xchunk", a.xchunk)
    best = 0.0
    for _ in range(a.reps):
        sr.plan.fill(0.0, 1.5)
        dist.barrier()
        t = sr.run(0, a.T - 1)
        tt = torch.tensor([t.section0 + t.section1], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        best = max(best, a.nxg * a.ny * a.nz * (a.T - 5) / tt.item() / 1e9)
    if rank == 0:
        print(f"slabs {world} x {sr.nx}x{a.ny}x{a.nz} exact={ex} t_fuse={sr.plan.get_option('t_fuse_used')} tile "
              f"{sr.plan.get_option('tile_y_used')}x{sr.plan.get_option('tile_z_used')} xchunk {sr.plan.get_option('xchunk_used')} "
              f"edge {os.environ.get('FDTD_B200_SLAB_EDGE', 'auto')}: {best:8.1f} Gpts/s = {best / world:7.1f} per GPU", flush=True)
sr.close()
dist.destroy_process_group()
