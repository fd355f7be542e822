#!/usr/bin/env python
"""Tile / stage / x-chunk sweep of the streaming kernel on one GPU (tuning aid, not a bench).

    python tools/sweep.py --n 512 --steps 30 > gpurun_out/sweep_512.txt
"""
import argparse
import importlib
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("accelerated-3d-acoustic-fdtd-kernel_b200")

TILES = [(8, 64, 2, 5), (8, 64, 1, 5), (8, 128, 1, 5), (8, 128, 2, 5), (16, 64, 2, 5), (16, 64, 1, 5), (16, 128, 2, 5),
         (16, 128, 1, 5), (32, 64, 2, 5), (32, 64, 1, 5), (16, 32, 2, 5), (16, 32, 1, 5), (8, 32, 1, 5), (32, 32, 2, 5),
         (32, 128, 2, 5), (32, 64, 4, 5), (32, 128, 4, 5), (14, 64, 1, 5), (14, 128, 1, 5), (28, 64, 2, 5),
         (8, 64, 2, 10), (8, 128, 2, 10), (16, 64, 2, 10), (8, 64, 1, 10), (16, 128, 2, 10)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--shape", type=str, default="", help="nx,ny,nz (overrides --n)")
    ap.add_argument("--xchunks", type=str, default="0")
    ap.add_argument("--exact", type=str, default="1,0")
    ap.add_argument("--tiles", type=str, default="")
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--nsrc", type=int, default=1)
    ap.add_argument("--global-nx", type=int, default=0, help="place the source lattice as if the slab were part of this global grid")
    ap.add_argument("--dense", action="store_true", help="dense sin field instead of the zero field")
    ap.add_argument("--tfuse", type=int, default=1, help="2: two-step passes (tiles are then the output tiles TYxTZ of stencil_tb2.cu)")
    ap.add_argument("--lean", type=int, default=1, help="two-step passes: 1 = stencil_tb2l.cu, 0 = stencil_tb2.cu")
    ap.add_argument("--cluster", type=int, default=0, help="1: two-step passes on 2-CTA clusters (stencil_tc2.cu)")
    a = ap.parse_args()
    n, T = a.n, a.steps + 5
    nx, ny, nz = [int(x) for x in a.shape.split(",")] if a.shape else (n, n, n)
    npts = nx * ny * nz
    tiles = TILES if not a.tiles else [tuple(int(x) for x in t.split("x")) for t in a.tiles.split(",")]
    tiles = [t + ((1, 5) if len(t) == 2 else (5,) if len(t) == 3 else ()) for t in tiles]
    src, crd = pkg.fill_ricker(T, a.nsrc), pkg.fill_source_coords(a.nsrc, a.global_nx or nx, ny, nz)
    with pkg.Plan(nx, ny, nz, deviceid=0) as p:
        p.set_sources(src, crd)
        rows = []
        if a.generic:
            for exact in (1, 0):
                p.fill(0.0, 1.5)
                p.set_option("kernel", 1)
                p.set_option("exact", exact)
                t = p.run(0, T - 1)
                g = npts * a.steps / (t.section0 + t.section1) / 1e9
                print(f"generic exact={exact}: {g:8.1f} Gpts/s  {16 * g / 6551.7:6.3f} of HBM", flush=True)
        for (ty, tz, st, ns), xc, exact in itertools.product(tiles, [int(x) for x in a.xchunks.split(",")],
                                                          [int(x) for x in a.exact.split(",")]):
            p.fill_dense() if a.dense else p.fill(0.0, 1.5)
            for k, v in (("kernel", 2), ("exact", exact), ("tile_y", ty), ("tile_z", tz), ("rows", st), ("stages", ns),
                         ("xchunk", xc), ("t_fuse", a.tfuse), ("cluster", a.cluster), ("tb2_lean", a.lean)):
                p.set_option(k, v)
            try:
                t = p.run(0, T - 1)
            except pkg.FdtdError as e:
                print(f"tile {ty}x{tz} r{st} s{ns} xc{xc} exact={exact}: {e}", flush=True)
                continue
            g = npts * a.steps / (t.section0 + t.section1) / 1e9
            rows.append((g, ty, tz, st, ns, p.get_option("xchunk_used"), exact))
            if p.get_option("t_fuse_used") != a.tfuse:
                print(f"tile {ty}x{tz}: t_fuse {a.tfuse} not honoured", flush=True)
            print(f"tile {ty:3d}x{tz:3d} r{st} s{ns:2d} xchunk {rows[-1][5]:4d} exact={exact}: {g:8.1f} Gpts/s  "
                  f"{16 * g / 6551.7:6.3f} of HBM  ({p.last_kernel_seconds * 1e6:8.1f} us/step)", flush=True)
    rows.sort(reverse=True)
    print("best:", rows[:5])


if __name__ == "__main__":
    main()
