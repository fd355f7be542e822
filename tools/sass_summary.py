#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library: which mnemonics prove TMA / mbarrier / 128-bit stores.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt   (cuobjdump needs no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "accelerated-3d-acoustic-fdtd-kernel_b200", "libfdtd_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTMALDG", "UTMAPF", "SYNCS", "STG.E.128", "LDS.128", "LDS.64", "STS.128", "FFMA", "FADD", "FMUL", "MUFU.RCP", "LDG", "ATOM", "NANOSLEEP",
        "ACQBULK", "ERRBAR", "MEMBAR", "CCTL"]
arch = re.search(r"arch = (sm_\w+)", out)
print(f"# {os.path.relpath(lib, ROOT)}: {arch.group(1) if arch else '?'}; instruction counts per kernel (static SASS, cuobjdump -sass)")
print("kernel".ljust(78) + "".join(k.rjust(11) for k in ["insts"] + KEYS))
cur, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if cur and m:
        op = m.group(1)
        counts[cur]["insts"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for name, nice in zip(counts, demangle):
    c = counts[name]
    tot.update(c)
    nice = re.sub(r"\(fdtd::\w+\)$|void |fdtd::|\(int\)|\(bool\)", "", nice)
    print(nice[:77].ljust(78) + "".join(str(c[k]).rjust(11) for k in ["insts"] + KEYS))
print("TOTAL".ljust(78) + "".join(str(tot[k]).rjust(11) for k in ["insts"] + KEYS))
