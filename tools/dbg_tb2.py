#!/usr/bin/env python
"""Debug aid: one run of the two-step passes on an nx,ny,nz grid through the plan API.
usage: dbg_tb2.py nx,ny,nz T exact [tile_y tile_z xchunk]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("accelerated-3d-acoustic-fdtd-kernel_b200")
nx, ny, nz = [int(x) for x in sys.argv[1].split(",")]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
exact = int(sys.argv[3]) if len(sys.argv) > 3 else 1
src, crd = pkg.fill_ricker(T, 1), pkg.fill_source_coords(1, nx, ny, nz)
with pkg.Plan(nx, ny, nz, deviceid=0) as p:
    p.set_sources(src, crd)
    p.set_option("t_fuse", 2)
    p.set_option("kernel", 2)
    p.set_option("exact", exact)
    for k, v in zip(("tile_y", "tile_z", "xchunk"), [int(x) for x in sys.argv[4:7]]):
        p.set_option(k, v)
    p.fill(0.0, 1.5)
    try:
        t = p.run(0, T - 1)
        print("ok  ", sys.argv[1:], p.get_option("t_fuse_used"), p.get_option("tile_y_used"), p.get_option("tile_z_used"),
              p.get_option("xchunk_used"), t.section0, flush=True)
    except pkg.FdtdError as e:
        print("FAIL", sys.argv[1:], e, flush=True)
