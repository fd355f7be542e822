"""CPU tests of the drop-in boundary: libfdtd_b200.so loads, exports every symbol include/fdtd_b200.h
declares, its structs match the reference ABI, and the host-only helpers (driver input synthesis,
source table, benchmark.csv writer) are bit-exact against the oracle / golden fixtures.
No compute call is made here (there is no GPU in the build container)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bits_equal


def header_functions():
    src = open(os.path.join(ROOT, "include", "fdtd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:Kernel_|FDTD_|fdtd_b200_)\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    declared = header_functions()
    assert set(declared) == set(pkg.exported_symbols())
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    L = pkg.lib()
    for s in declared:
        assert getattr(L, s) is not None


def test_struct_layout_matches_reference_abi(pkg):
    # main.cpp:35-50 on LP64: 9 pointer-sized fields, two doubles
    assert C.sizeof(pkg.Dataobj) == 72 and C.sizeof(pkg.Profiler) == 16
    assert pkg.Dataobj.data.offset == 0 and pkg.Dataobj.size.offset == 8 and pkg.Dataobj.nbytes.offset == 16
    assert pkg.Profiler.section1.offset == 8
    assert pkg.HALO == 4 and pkg.WARMUP_STEPS == 5


def test_sass_is_blackwell_native(pkg):
    """The shipped cubin is sm_100a and the streaming kernel really uses TMA + mbarrier (UTMALDG / SYNCS)."""
    r = subprocess.run(["cuobjdump", "-sass", pkg.lib_path()], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    assert "UTMALDG" in r.stdout and "SYNCS" in r.stdout
    assert "STG.E.128" in r.stdout


def test_driver_input_synthesis_is_bit_exact(pkg, oracle, golden):
    _, arrs = golden
    assert bits_equal(pkg.fill_ricker(200, 1)[:, 0], arrs["ricker_T200"])
    assert bits_equal(pkg.fill_ricker(50, 7), oracle.fill_ricker(50, 7))
    for n in (32, 64, 96, 512, 768, 1024, 2048):
        for S in (1, 5, 27, 64):
            assert bits_equal(pkg.fill_source_coords(S, n, n, n), oracle.fill_source_coords(S, n, n, n))
    assert bits_equal(pkg.fill_source_coords(30, 40, 52, 64), oracle.fill_source_coords(30, 40, 52, 64))


def test_source_table_is_bit_exact(pkg, oracle, golden):
    _, arrs = golden
    # lattice positions recorded from the reference build (pos, frac bits), openacc.cpp:125-131
    for n, s, cbits, pos, fbits in arrs["lattice_pos"]:
        c = np.array([cbits], np.uint32).view(np.float32)[0]
        p, f, w, inr = pkg.source_table((c, c, c), (0, 0, 0), (0.1, 0.1, 0.1), (0, 0, 0), (n - 1,) * 3)
        assert p.tolist() == [pos] * 3 and f.view(np.uint32).tolist() == [fbits] * 3 and inr.all()
    for c, o, h, pos, frac in zip(arrs["pos_coord"], arrs["pos_o"], arrs["pos_h"], arrs["pos_pos"], arrs["pos_frac"]):
        p, f, w, inr = pkg.source_table((c, c, c), (o, o, o), (h, h, h), (0, 0, 0), (63, 63, 63))
        assert p[0] == pos and f.view(np.uint32)[0] == np.float32(frac).view(np.uint32)
        # weights: ((1e-2f*wx)*wy)*wz with w = r*p + (1-r)*(1-p), openacc.cpp:134
        one, fr = np.float32(1), np.float32(f[0])
        ax = [one - fr, fr]
        for rx in (0, 1):
            for ry in (0, 1):
                for rz in (0, 1):
                    ref = np.float32(np.float32(np.float32(np.float32(1.0e-2) * ax[rx]) * ax[ry]) * ax[rz])
                    assert w[rx * 4 + ry * 2 + rz].view(np.uint32) == ref.view(np.uint32)
                    ok = all(-1 <= r + pos <= 64 for r in (rx, ry, rz))
                    assert bool(inr[rx * 4 + ry * 2 + rz]) == ok


def test_source_table_drives_the_oracle_scatter(pkg, oracle):
    """Scatter built from the product's table == the oracle's Section1, including halo and coincident cells."""
    rng = np.random.default_rng(3)
    n, S = 12, 9
    crd = (rng.uniform(-0.08, 1.08, (S, 3)) * (n - 1) * 0.1).astype(np.float32)
    crd[4] = crd[3]
    m = rng.uniform(0.5, 3, (n + 8,) * 3).astype(np.float32)
    src = rng.uniform(-5, 5, (1, S)).astype(np.float32)
    u = np.zeros((3, n + 8, n + 8, n + 8), np.float32)
    # oracle: one step on a zero field with time_M = 0 -> u[1] holds exactly the scatter
    oracle.run(u, m, src, crd, impl="port", time_m=0, time_M=0)
    mine = np.zeros_like(u[1])
    for p in range(S):
        pos, frac, w, inr = pkg.source_table(crd[p], (0, 0, 0), (0.1,) * 3, (0, 0, 0), (n - 1,) * 3)
        mb = m[pos[0] + 4, pos[1] + 4, pos[2] + 4] if inr.any() else np.float32(1)
        for rx in (0, 1):
            for ry in (0, 1):
                for rz in (0, 1):
                    i = rx * 4 + ry * 2 + rz
                    if inr[i]:
                        r0 = np.float32(np.float32(w[i] * src[0, p]) / mb)
                        idx = (pos[0] + rx + 4, pos[1] + ry + 4, pos[2] + rz + 4)
                        mine[idx] = np.float32(mine[idx] + r0)
    assert bits_equal(mine, u[1])


def test_benchmark_csv_schema(pkg, tmp_path):
    f = str(tmp_path / "benchmark.csv")
    for method in ("B200_1gpu", "B200_8gpu_t4"):
        pkg.write_benchmark_csv(f, method, (1.5, 0.1), (0.4, 0.01), (0.0, 0.0), (0.4, 0.01), (1.1, 0.1), (1234.5, 6.7),
                                (890.1, 2.3), 80000.0, 8000.0, 0.5625, 512, 512, 512, 50, 1)
    lines = open(f).read().splitlines()
    assert lines[0] == ("Method,Total_Time(ms),Total_Std(ms),Section0_Time(ms),Section0_Std(ms),Section1_Time(ms),"
                        "Section1_Std(ms),Device_Time(ms),Device_Std(ms),Overhead(ms),Overhead_Std(ms),GFLOPS,GFLOPS_Std,"
                        "GBps,GBps_Std,Compute_Eff(%),Memory_Eff(%),AI,NX,NY,NZ,Timesteps,Sources,StencilOrder")  # main.cpp:222-225
    assert len(lines) == 3 and all(len(ln.split(",")) == 24 for ln in lines)
    row = lines[1].split(",")
    assert row[0] == "B200_1gpu" and row[1] == "1500" and row[11] == "1234.5" and row[-6:] == ["512", "512", "512", "50", "1", "4"]


def test_no_silent_fallback_without_gpu(pkg):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.FdtdError):
        pkg.Plan(8, 8, 8)
    u = np.zeros((3, 16, 16, 16), np.float32)
    m = np.ones((16, 16, 16), np.float32)
    rc = pkg.Kernel_B200(m, None, None, u, 7, 0, 7, 0, 7, 0, 1e-3, 1.0, 1.0, 1.0, 0, 0, 0, -1, 0, 3, 0, 0, 1)
    assert rc != 0 and not u.any()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, load or link it."""
    pk = os.path.join(ROOT, "accelerated-3d-acoustic-fdtd-kernel_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|liboracle|libref_|oracle/|oracle\.py|fdtd_oracle")
    for dirpath, _, files in os.walk(pk):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or fn == "Makefile":
                assert not pat.search(open(os.path.join(dirpath, fn)).read()), fn


def test_shipped_kernels_are_sm100a_tma_code():
    """Static evidence that the hot kernels in the shipped library are Blackwell-native: sm_100a SASS with TMA tensor loads
    (UTMALDG), the L2 tensor prefetch of the lean two-step kernel (UTMAPF), mbarrier operations (SYNCS) and 128-bit stores.
    cuobjdump needs no GPU."""
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    lib = os.path.join(ROOT, "accelerated-3d-acoustic-fdtd-kernel_b200", "libfdtd_b200.so")

    def sass(mangled):
        out = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, lib], capture_output=True, text=True, check=True).stdout
        assert "arch = sm_100a" in out
        return out

    lean = sass("_ZN4fdtd19stencil_tb2l_kernelILi20ELi34ELb0EEEvNS_7Tb2ArgsE")  # 16 x 128 tile, contracted: the bench headline
    for op in ("UTMALDG.4D", "UTMALDG.3D", "UTMAPF.L2", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "STG.E.128", "LDS.128", "STS.128", "ELECT"):
        assert op in lean, op
    one = sass("_ZN4fdtd18stencil_tma_kernelILi8ELi64ELi1ELi5ELb1ELi4EEEvNS_7TmaArgsE")  # 8 x 64 tile, exact: the bit-identical default
    for op in ("UTMALDG.4D", "UTMALDG.3D", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "STG.E.128", "LDS.128"):
        assert op in one, op
