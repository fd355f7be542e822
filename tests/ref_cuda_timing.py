#!/usr/bin/env python
"""Time the REFERENCE's own CUDA kernels (cuda.cu, cuda_optimized.cu compiled unmodified for sm_100a into
oracle/_ref/libref_cuda.so) on this GPU, beside the new kernel, and report their error against the oracle.
Test infrastructure (it loads oracle/_ref), run by hand:  python tests/ref_cuda_timing.py --sizes 64,256,512
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

pkg = importlib.import_module("accelerated-3d-acoustic-fdtd-kernel_b200")


def call(fn, u, m, src, crd, n, T, S):
    def obj(a):
        sizes = (C.c_int * a.ndim)(*a.shape)
        d = O.Dataobj()
        d.data, d.size, d.nbytes, d._k = a.ctypes.data, C.cast(sizes, C.POINTER(C.c_int)), a.nbytes, sizes
        return d
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(O.Dataobj)] * 4 + [C.c_int] * 6 + [C.c_float] * 7 + [C.c_int] * 6 + [C.POINTER(O.Profiler)]
    t = O.Profiler(0, 0)
    mo, so, co, uo = obj(m), obj(src), obj(crd), obj(u)
    rc = fn(mo, so, co, uo, n - 1, 0, n - 1, 0, n - 1, 0, 1e-3, .1, .1, .1, 0, 0, 0, S - 1, 0, T - 1, 0, 0, 1, C.byref(t))
    assert rc == 0, rc
    return t.section0 + t.section1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64,256,512")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so"))
    mine = pkg.lib()
    T, S = 50, 1
    rows = []
    for n in [int(x) for x in a.sizes.split(",")]:
        m = np.full((n + 8,) * 3, 1.5, np.float32)
        src, crd = pkg.fill_ricker(T, S), pkg.fill_source_coords(S, n, n, n)
        gold = None
        if n <= 256:
            gold = np.zeros((3, n + 8, n + 8, n + 8), np.float32)
            O.run(gold, m, src, crd, impl="reference" if O.have_reference() else "port", threads=os.cpu_count())
        for name, fn in (("ref Kernel_CUDA (cuda.cu)", ref.Kernel_CUDA),
                         ("ref Kernel_CUDA_Optimized (cuda_optimized.cu)", ref.Kernel_CUDA_Optimized),
                         ("new Kernel_B200", mine.Kernel_B200)):
            best = None
            for _ in range(a.reps):
                u = np.zeros((3, n + 8, n + 8, n + 8), np.float32)
                dev = call(fn, u, m, src, crd, n, T, S)
                best = dev if best is None else min(best, dev)
            gpts = n ** 3 * (T - 5) / best / 1e9
            row = {"n": n, "impl": name, "device_s": best, "gpts": gpts, "hbm_frac_of_measured": gpts * 16 / 6551.7,
                   "max_abs": float(np.abs(u).max())}
            if gold is not None:
                row["rel_l2_vs_oracle"] = O.rel_l2(u, gold)
                row["bit_exact"] = bool(np.array_equal(u.view(np.uint32), gold.view(np.uint32)))
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
