#!/usr/bin/env python
"""Host<->device copy rates on this box (pinned vs pageable) and the phase breakdown of one Kernel_B200 call."""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("accelerated-3d-acoustic-fdtd-kernel_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nb = 3 * (n + 8) ** 3 * 4
d = torch.empty(nb // 4, dtype=torch.float32, device="cuda")
for name, h in (("pinned", torch.zeros(nb // 4, dtype=torch.float32).pin_memory()), ("pageable", torch.zeros(nb // 4, dtype=torch.float32))):
    for direction in ("H2D", "D2H"):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t = time.perf_counter()
            if direction == "H2D":
                d.copy_(h, non_blocking=True)
            else:
                h.copy_(d, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t)
        print(f"{name:9s} {direction}: {nb / best / 1e9:6.1f} GB/s ({best * 1e3:.1f} ms for {nb / 1e9:.2f} GB)")
del d
os.environ["FDTD_B200_TRACE"] = "1"
for pin in (True, False):
    u = torch.zeros((3, n + 8, n + 8, n + 8), dtype=torch.float32)
    m = torch.full((n + 8, n + 8, n + 8), 1.5, dtype=torch.float32)
    if pin:
        u, m = u.pin_memory(), m.pin_memory()
    u, m = u.numpy(), m.numpy()
    src, crd = pkg.fill_ricker(50, 1), pkg.fill_source_coords(1, n, n, n)
    for _ in range(2):
        t = time.perf_counter()
        rc = pkg.Kernel_B200(m, src, crd, u, n - 1, 0, n - 1, 0, n - 1, 0, 1e-3, .1, .1, .1, 0, 0, 0, 0, 0, 49, 0, 0, 1)
        print(f"pinned={pin} Kernel_B200 rc={rc} wall {1e3 * (time.perf_counter() - t):.1f} ms", flush=True)
