#!/usr/bin/env python
"""One short run of a chosen kernel configuration (profiling target for ncu):
    python tools/prof_one.py --n 512 --tfuse 2 --exact 0 --tile 32x64 [--xchunk 0] [--steps 6]"""
import argparse, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("accelerated-3d-acoustic-fdtd-kernel_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--tfuse", type=int, default=1)
ap.add_argument("--exact", type=int, default=1)
ap.add_argument("--tile", default="")
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--xchunk", type=int, default=0)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--cluster", type=int, default=0)
ap.add_argument("--lean", type=int, default=1)
a = ap.parse_args()
n, T = a.n, a.steps + 5
src, crd = pkg.fill_ricker(T, 1), pkg.fill_source_coords(1, n, n, n)
with pkg.Plan(n, n, n, deviceid=0) as p:
    p.set_sources(src, crd)
    p.set_option("kernel", 2)
    p.set_option("t_fuse", a.tfuse)
    p.set_option("exact", a.exact)
    p.set_option("cluster", a.cluster)
    p.set_option("tb2_lean", a.lean)
    if a.tile:
        ty, tz = [int(x) for x in a.tile.split("x")]
        p.set_option("tile_y", ty)
        p.set_option("tile_z", tz)
    p.set_option("rows", a.rows)
    p.set_option("xchunk", a.xchunk)
    p.fill(0.0, 1.5)
    t = p.run(0, T - 1)
    g = n ** 3 * a.steps / (t.section0 + t.section1) / 1e9
    print(f"t_fuse={p.get_option('t_fuse_used')} exact={a.exact} tile={p.get_option('tile_y_used')}x{p.get_option('tile_z_used')} "
          f"xchunk={p.get_option('xchunk_used')}: {g:.1f} Gpts/s", flush=True)
