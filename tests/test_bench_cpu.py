"""CPU checks of bench.py's contract: the reference arm runs without a GPU (the reference's own OpenACC source on
the host cores, or the oracle port) and prints ONE JSON line with the keys the driver reads; the B200 arm refuses
to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import torch

from conftest import ROOT


def test_reference_arm_prints_one_json_line(oracle):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "64", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gpts/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_b200_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return  # on a GPU box the arm runs; the GPU tiers cover it
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
