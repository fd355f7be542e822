#!/usr/bin/env python
"""bench.py -- headline benchmark of the 3D acoustic FDTD hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   (the reference's own CPU implementation)

Metric (BASELINE.json): Gpts/s = grid-point updates per second, and its fraction of the HBM roofline.
One "step" = one pass of the operator (reference Kernel_* semantics: T time steps of Section0 + Section1,
the first 5 untimed by the operator's own section timers) over one synthetic grid of the driver's
benchmark configuration (main.cpp:285-356: zero field, m = 1.5, Ricker wavelet, lattice sources).
  N = 1 : BASELINE configs[2], 512^3, T = 50, 1 source (the size the metric is quoted on).
  N > 1 : x-slab decomposition, one 512 x 512 x 512 slab per GPU (global (512 N) x 512 x 512), weak scaling.
`value` follows the reference's definition: points * timed steps / (section0 + section1) with the fields
resident in HBM; `e2e` is the same operator through the reference-facing C ABI (Kernel_B200) with host
buffers, host<->device copies inside the timed region.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "accelerated-3d-acoustic-fdtd-kernel_b200"
ALGO_BYTES_PER_POINT = 16.0  # read u[t0], u[t1], m + write u[t2], fp32


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"  # B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.rows, self._stop, self._t = gpu, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][2]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[3]) for r in self.rows)}


def cpu_reference(n, nsrc, timed_steps, reps=1):
    """The reference's own OpenACC source compiled for the host (oracle/_ref, OpenMP stand-in for
    -acc=multicore) -- or the oracle port when _ref is absent -- on a bounded sample of the workload."""
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    kind = "reference" if O.have_reference() else "port"
    T = 5 + timed_steps
    u = np.zeros((3, n + 8, n + 8, n + 8), np.float32)
    m = np.full((n + 8,) * 3, 1.5, np.float32)
    src, crd = O.fill_ricker(T, nsrc), O.fill_source_coords(nsrc, n, n, n)
    best = None
    for _ in range(reps):
        u[...] = 0
        s0, s1 = O.run(u, m, src, crd, impl=kind, threads=cores)
        dev = s0 + s1
        best = dev if best is None else min(best, dev)
    return {"value": n ** 3 * timed_steps / best / 1e9, "unit": "Gpts/s", "cores": cores, "kind": kind,
            "sample": f"{n}^3, {T} of 50 time steps ({timed_steps} timed), {nsrc} source(s), OpenMP x{cores}",
            "seconds": best}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, nsrc, timed = a.n, a.nsrc, 2
    vals = []
    for i in range(a.warmup + a.steps):
        r = cpu_reference(n, nsrc, timed)
        if i >= a.warmup:
            vals.append(r)
    tot_pts = sum(n ** 3 * timed for _ in vals)
    tot_s = sum(r["seconds"] for r in vals)
    v = tot_pts / tot_s / 1e9
    base = vals[-1]
    line = {
        "impl": "reference", "metric": "Gpts/s (grid-point updates/s) at 512^3", "value": v, "unit": "Gpts/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot_s / max(1, len(vals)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n}^3 grid, 50 timesteps, {nsrc} source, fp32 (reference CPU path: bounded sample)"},
        "cpu_baseline": {"value": v, "unit": "Gpts/s", "cores": base["cores"], "kind": base["kind"], "sample": base["sample"]},
        "e2e": {"value": v, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_b200_arm(a):
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module(PKG)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, T, S = a.n, a.timesteps, a.nsrc
    nxg, scaling = n * world, "weak"
    if a.workload == "1024-strong":      # BASELINE configs[3]
        n, nxg, T, S, scaling = 1024, 1024, 200, 1, "strong"
    elif a.workload == "2048-weak":      # BASELINE configs[4]
        n, nxg, T, S = 2048, 2048, 200, 64
    elif a.workload:
        raise SystemExit(f"unknown workload {a.workload}")
    timed_steps = T - min(5, T)
    if world == 1:
        from_slab = None
        plan = pkg.Plan(nxg, n, n, deviceid=local)
    else:
        # x-slab decomposition, one process per GPU; torch.distributed only carries the rendezvous
        from_slab = pkg.SlabRun(dist, nxg, n, n, local)
        plan = from_slab.plan
    nx_local = plan.shape[1] - 8
    # headline configuration: the fastest one that meets the reference's own tolerance (relative L2 < 1e-4,
    # README.md:33): contracted arithmetic (rel L2 ~1e-6 vs the oracle) and two time steps per pass.  --exact 1
    # --tfuse 1 is the bit-identical one-step configuration; both are reported (other_modes).
    for k, v in (("exact", a.exact), ("kernel", a.kernel), ("t_fuse", a.tfuse)):
        if v is not None:
            plan.set_option(k, v)
    src = pkg.fill_ricker(T, S)
    crd = pkg.fill_source_coords(S, nxg, n, n)
    plan.set_sources(src, crd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        plan.fill(0.0, 1.5)
        if from_slab is not None:
            dist.barrier()  # a neighbour's first step already writes ghost planes into this slab
            return from_slab.run(0, T - 1)
        return plan.run(0, T - 1)

    passes = []  # (wall, section0, section1) seconds of every timed operator pass on this rank

    def measure(steps, warmup):
        """W untimed + K timed operator passes; device seconds (section timers = CUDA events on the compute
        stream), wall seconds bracketed by barrier + synchronize, max over ranks."""
        for _ in range(warmup):
            one_step()
        barrier()
        dev_s, kern_s, launches = 0.0, 0.0, 0
        t0 = time.perf_counter()
        for _ in range(steps):
            tp = time.perf_counter()
            t = one_step()
            dev_s += t.section0 + t.section1
            kern_s += plan.last_kernel_seconds
            launches += plan.last_launches + 2  # + the two fill kernels
            passes.append((time.perf_counter() - tp, t.section0, t.section1))
        barrier()
        wall = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dev_s, wall, kern_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dev_s, wall, kern_s = tt.tolist()
        return dev_s, wall, kern_s, launches

    with ClockSampler(local) as clk:
        dev_s, wall, kern_s, launches = measure(a.steps, a.warmup)
    headline_passes = list(passes)
    pts_per_step = float(nxg) * n * n
    value = pts_per_step * timed_steps * a.steps / dev_s / 1e9
    peak, peak_kind = measured_peak()
    # last_kernel_seconds = stencil seconds per TIME STEP in the timed region; a two-step launch covers two of them
    t_fuse_used = plan.get_option("t_fuse_used")
    kern_step = kern_s / a.steps
    achieved = ALGO_BYTES_PER_POINT * float(nx_local) * n * n / kern_step / 1e9
    arith = "exact" if plan.get_option("exact") else "contracted"
    traffic = None
    try:  # ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel (profiles/README.md)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{nx_local}x{n}x{n}", {}).get(f"{arith}_t{t_fuse_used}")
    except Exception:  # noqa: BLE001
        pass

    line = {
        "metric": "Gpts/s (grid-point updates/s) at 512^3 and fraction of B200 HBM roofline",
        "value": value, "unit": "Gpts/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"{nxg}x{n}x{n} grid, {T} timesteps, {S} source, fp32"
                                + (f", {world} x-slabs of {nx_local}x{n}x{n}" if world > 1 else "")
                                + (" (BASELINE configs[2])" if (nxg, world) == (512, 1) else "")),
                   "timed_steps_per_pass": timed_steps, "arithmetic": arith,
                   "time_steps_per_launch": t_fuse_used,
                   "kernel": {1: "generic", 2: "tma"}[plan.get_option("kernel_used")] + ("_two_step" if t_fuse_used == 2 else ""),
                   "tile": [plan.get_option("tile_y_used"), plan.get_option("tile_z_used"), plan.get_option("rows_used"),
                            plan.get_option("xchunk_used")],
                   "l2": f"arrays ({16 * (nx_local + 8) * (n + 8) ** 2 / 1e9:.2f} GB per GPU) exceed the 126 MB L2; no flush needed"},
        "value_bracketed": pts_per_step * T * a.steps / wall / 1e9,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_kind": peak_kind, "bytes_per_point": ALGO_BYTES_PER_POINT,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_POINT * float(nx_local) * n * n * t_fuse_used,
                     "kernel_us": kern_step * t_fuse_used * 1e6},
        "clocks": clk.summary(),
        "gpu_launches": launches,
    }

    # ---- the other configurations on the same workload (fewer passes): bit-exact arithmetic, one step per launch
    if not a.no_modes:
        others = []
        headline = (1 if arith == "exact" else 0, plan.get_option("t_fuse"))
        for ex, tf in ((1, 1), (0, 1), (1, 2), (0, 2)):
            if (ex, tf) == headline:
                continue
            plan.set_option("exact", ex)
            plan.set_option("t_fuse", tf)
            d, _, k, _ = measure(max(2, a.steps // 4), 1)
            reps = max(2, a.steps // 4)
            others.append({"arithmetic": "exact" if ex else "contracted", "time_steps_per_launch": plan.get_option("t_fuse_used"),
                           "value": pts_per_step * timed_steps * reps / d / 1e9,
                           "roofline_frac": ALGO_BYTES_PER_POINT * float(nx_local) * n * n / (k / reps) / 1e9 / peak})
        line["other_modes"] = others
        plan.set_option("exact", headline[0])
        plan.set_option("t_fuse", headline[1])

    # ---- e2e: the reference-facing C ABI with host buffers (H2D + 50 steps + D2H inside the timed region)
    if world == 1 and rank == 0 and not a.no_e2e and not a.workload:
        volp = (n + 8) ** 3
        u_h = torch.zeros((3, n + 8, n + 8, n + 8), dtype=torch.float32).pin_memory().numpy()
        m_h = torch.full((n + 8, n + 8, n + 8), 1.5, dtype=torch.float32).pin_memory().numpy()
        # library defaults: bit-exact arithmetic, staged pipeline (H2D, x-skewed time loop and D2H overlapped)
        e2e_t = []
        for i in range(1 + a.e2e_reps):
            u_h[...] = 0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = pkg.Kernel_B200(m_h, src, crd, u_h, n - 1, 0, n - 1, 0, n - 1, 0, 1e-3, 0.1, 0.1, 0.1, 0.0, 0.0, 0.0,
                                 S - 1, 0, T - 1, 0, local, 1)
            dt = time.perf_counter() - t0
            if rc != 0:
                raise SystemExit(f"Kernel_B200 failed: cudaError {rc}")
            if i > 0:
                e2e_t.append(dt)
        assert abs(float(np.abs(u_h).max()) - 0.1168) < 1e-3  # the D2H result is read
        e2e_s = float(np.median(e2e_t))  # median: an occasional cudaMalloc/cudaFree hiccup of the box is not the path's speed
        line["e2e"] = {"value": pts_per_step * T / e2e_s / 1e9, "unit": "Gpts/s", "calls_timed": len(e2e_t),
                       "seconds_per_call_min_max": [min(e2e_t), max(e2e_t)],
                       "h2d_bytes_per_step": 4 * volp * 4 + src.nbytes + crd.nbytes,
                       "d2h_bytes_per_step": 3 * n * (n + 8) ** 2 * 4,  # x-halo planes never change and are not read back
                       "seconds_per_call": e2e_s, "api": "Kernel_B200 (reference ABI), pinned host buffers",
                       "arithmetic": "exact" if int(os.environ.get("FDTD_B200_EXACT", "1")) else "contracted",
                       "staging": ("pipelined: chunks of x planes (%s), time loop skewed along x, D2H of finished planes overlapped"
                                   % os.environ.get("FDTD_B200_STAGE_PLANES", "auto"))
                       if int(os.environ.get("FDTD_B200_STAGE_PLANES", "-1")) != 0 else "three phases (H2D, run, D2H)"}
        del u_h, m_h

    # ---- e2e at N > 1: every rank stages its own slab through the plan API (upload from pinned host memory, the
    # linked run, download), wall clock bracketed by barriers, max over ranks.  Slabs keep the three-phase order (the
    # skewed pipeline of the single-GPU path would have to interleave with the neighbours' ghost planes).
    if world > 1 and not a.no_e2e:
        nxl = nx_local + 8
        u_h = torch.zeros((3, nxl, n + 8, n + 8), dtype=torch.float32).pin_memory().numpy()
        m_h = torch.full((nxl, n + 8, n + 8), 1.5, dtype=torch.float32).pin_memory().numpy()
        e2e_t = []
        for i in range(1 + a.e2e_reps):
            u_h[...] = 0
            barrier()
            t0 = time.perf_counter()
            plan.upload(u_h, m_h)
            dist.barrier()  # a neighbour's first pass already writes ghost planes into this slab
            from_slab.run(0, T - 1)
            plan.download(u_h)
            barrier()
            if i > 0:
                e2e_t.append(time.perf_counter() - t0)
        tt = torch.tensor([float(np.median(e2e_t))], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
        line["e2e"] = {"value": pts_per_step * T / e2e_s / 1e9, "unit": "Gpts/s",
                       "h2d_bytes_per_step": world * (4 * nxl * (n + 8) ** 2 * 4) + src.nbytes + crd.nbytes,
                       "d2h_bytes_per_step": world * 3 * nxl * (n + 8) ** 2 * 4, "seconds_per_call": e2e_s,
                       "api": "SlabRun: plan.upload + linked run + plan.download per rank, pinned host buffers",
                       "arithmetic": arith, "time_steps_per_launch": t_fuse_used, "staging": "three phases (H2D, run, D2H)"}
        del u_h, m_h

    # ---- CPU baseline: the reference's OpenACC source on this box's host cores (bounded sample)
    if world == 1 and rank == 0 and not a.no_cpu and not a.workload:
        plan.close()
        cb = cpu_reference(n, S, 3)
        cb.pop("seconds")
        line["cpu_baseline"] = cb

    if rank == 0 and a.csv:
        # one row in the reference's benchmark.csv schema (main.cpp:201-249) for harnesses that bypass main.cpp
        # (SURVEY 8b): mean and POPULATION std over the timed passes, the driver's 36 flop / 64 B per point models
        # (main.cpp:129-146) over ALL T steps divided by the device time of the timed ones (main.cpp:404,430)
        def stat(v):
            v = np.asarray(v, np.float64)
            return float(v.mean()), float(v.std())

        tot, s0, s1 = (np.array([p[i] for p in headline_passes]) for i in range(3))
        dev = s0 + s1
        gf, gb = pts_per_step * T * 36 / dev / 1e9, pts_per_step * T * 64 / dev / 1e9
        pkg.write_benchmark_csv(a.csv, f"B200_{world}gpu_t{t_fuse_used}", stat(tot), stat(s0), stat(s1), stat(dev),
                                stat(np.maximum(0.0, tot - dev)), stat(gf), stat(gb), 148 * 128 * 2 * 1.965 * world, peak * world,
                                36.0 / 64.0, nxg, n, n, T, S)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="", help="named BASELINE config: 1024-strong (1024^3, T=200, 1 source, slabs of "
                    "1024/N planes) or 2048-weak (2048^3, T=200, 64 sources); default: one 512^3 slab per GPU")
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--timesteps", type=int, default=50)
    ap.add_argument("--nsrc", type=int, default=1)
    ap.add_argument("--exact", type=int, default=0, help="1 = bit-exact arithmetic (0 ulp vs the reference built for the host), "
                    "0 = contracted (FMA, rel L2 ~1e-6; tolerance 1e-4)")
    ap.add_argument("--tfuse", type=int, default=2, help="time steps per launch: 2 = two-step passes (temporal blocking), 1 = one")
    ap.add_argument("--csv", default="", help="append one row in the reference's benchmark.csv schema to this file")
    ap.add_argument("--no-modes", action="store_true", help="skip the short runs of the other arithmetic / t_fuse modes")
    ap.add_argument("--kernel", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-reps", type=int, default=5)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_b200_arm(a)


if __name__ == "__main__":
    sys.exit(main())
