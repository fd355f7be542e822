"""SURVEY 8(f) row 1 as a test: the reference's UNMODIFIED main.cpp, linked against libfdtd_b200.so as
`Kernel_CUDA_Optimized` (oracle/Makefile: _ref/fdtd_benchmark_b200; Kernel_OpenACC = the host build of openacc.cpp,
Kernel_CUDA = the reference's plain cuda.cu), runs end to end: its own correctness test (main.cpp:511-652) must
print an L2 error of exactly 0 for the drop-in at every size, and its benchmark.csv (main.cpp:201-249) must hold one
24-column row per grid size for the method.  The driver's OpenACC leg runs on the host cores (a few minutes)."""
import csv
import os
import re
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "oracle", "_ref", "fdtd_benchmark_b200")
HEADER = ("Method,Total_Time(ms),Total_Std(ms),Section0_Time(ms),Section0_Std(ms),Section1_Time(ms),Section1_Std(ms),"
          "Device_Time(ms),Device_Std(ms),Overhead(ms),Overhead_Std(ms),GFLOPS,GFLOPS_Std,GBps,GBps_Std,Compute_Eff(%),"
          "Memory_Eff(%),AI,NX,NY,NZ,Timesteps,Sources,StencilOrder").split(",")


def test_unmodified_reference_driver_end_to_end(tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/fdtd_benchmark_b200 not built (needs /root/reference at build time)")
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 8))
    r = subprocess.run([BIN], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=3000)
    out = r.stdout + r.stderr
    (tmp_path / "driver.log").write_text(out)
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        with open(os.path.join(keep, "main_driver_test.log"), "w") as f:
            f.write(out)
    assert r.returncode == 0, out[-3000:]
    # the driver's own correctness test (main.cpp:511-652): per "Test configuration" block it prints the L2 error of
    # Kernel_CUDA (main.cpp:599) and then of Kernel_CUDA_Optimized (main.cpp:637) against Kernel_OpenACC
    blocks = out.split("Test configuration:")[1:]
    assert len(blocks) >= 5, "correctness blocks missing"
    for blk in blocks:
        blk = blk.split("STEP 2")[0]
        errs = re.findall(r"L2 norm error:\s*([0-9.eE+-]+)", blk)
        assert len(errs) == 2, blk[:400]
        assert float(errs[1]) == 0.0, f"CUDA_Optimized (libfdtd_b200) L2 error {errs[1]} in block: {blk[:80]}"
        assert blk.count("PASS") >= 1
    rows = list(csv.reader(open(tmp_path / "benchmark.csv")))
    assert rows[0] == HEADER
    mine = [row for row in rows[1:] if row[0] == "CUDA_Optimized"]
    assert len(mine) == 10 and all(len(row) == 24 for row in mine)
    assert [int(row[18]) for row in mine] == [32, 64, 96, 128, 192, 256, 384, 512, 640, 768]
    if os.path.isdir(keep):
        with open(os.path.join(keep, "main_driver_test_benchmark.csv"), "w") as f:
            f.write(open(tmp_path / "benchmark.csv").read())
