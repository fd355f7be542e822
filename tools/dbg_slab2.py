#!/usr/bin/env python
"""Debug aid: two (or N) ranks over CUDA IPC with two-step launches against the oracle; prints which planes of which
level differ.  usage: dbg_slab2.py [world] [nx_per_rank] [T] [ny] [nz]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.multiprocessing as mp
    from test_slab_gpu import _ipc_worker, _free_port
    from test_tb2_gpu import fused_case
    from oracle import oracle as O
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    nxr = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 13
    ny = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    nz = int(sys.argv[5]) if len(sys.argv) > 5 else 128
    shape, S = (nxr * world, ny, nz), 6
    u, m, src, crd = fused_case(5, shape, T, S, seam_parts=world)
    ref = u.copy()
    O.run(ref, m, src, crd, impl="port", threads=8)
    out_path = "/tmp/dbg_slab2.npy"
    mp.spawn(_ipc_worker, args=(world, _free_port(), shape, T, S, out_path, 2), nprocs=world, join=True)
    out = np.load(out_path)
    bad = out.view(np.uint32) != ref.view(np.uint32)
    print("shape", shape, "T", T, "mismatching cells:", int(bad.sum()))
    for lvl in range(3):
        planes = np.nonzero(bad[lvl].any(axis=(1, 2)))[0]
        if len(planes):
            print(" level", lvl, "padded planes", planes.tolist()[:40], "cells per plane", [int(bad[lvl, x].sum()) for x in planes[:12]])
            x = planes[0]
            ys, zs = np.nonzero(bad[lvl, x])
            print("   first plane: y range", ys.min(), ys.max(), "z range", zs.min(), zs.max(),
                  "max abs err", float(np.abs(out[lvl, x] - ref[lvl, x]).max()))


if __name__ == "__main__":
    main()
