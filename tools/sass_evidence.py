#!/usr/bin/env python
"""Regenerate profiles/<round>_sass_evidence.txt from the library that is benchmarked
(audio_fir_filter_b200/libfir_gpu.so), here, without a GPU: per kernel the counts of the SASS
mnemonics that prove which hardware paths it uses (DMMA = FP64 tensor MMA, UBLKCP = 1-D TMA
bulk copy, UTMALDG = tiled TMA, SYNCS = mbarrier, DFMA = FP64 FMA pipe), the ptxas resource
lines, and an excerpt of the default FIR kernel's unrolled tap tile.
    python tools/sass_evidence.py [--round r2]"""
import argparse
import collections
import hashlib
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio_fir_filter_b200", "libfir_gpu.so")
ap = argparse.ArgumentParser()
ap.add_argument("--round", default="r2")
a = ap.parse_args()

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
MNEMONICS = ["DMMA", "DFMA", "UBLKCP", "UTMALDG", "SYNCS", "LDS.128", "LDS.64", "LDG.E.128", "STG.E.128", "STS", "F2I", "BAR.SYNC",
             "BPT.TRAP"]
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        kernels[cur].append(m.group(1).strip())


def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", r) or n


out = []
sha = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
out.append(f"# SASS evidence of {os.path.relpath(LIB, ROOT)} (sha256 {sha}..., {os.path.getsize(LIB)} bytes), "
           f"cuobjdump -sass, {time.strftime('%Y-%m-%d')}")
out.append("# regenerate with: python tools/sass_evidence.py   (the library is rebuilt by __graft_entry__.build())")
out.append(f"# {'kernel':72s} " + " ".join(f"{m:>9s}" for m in MNEMONICS) + "   instrs")
for k, ins in kernels.items():
    counts = [sum(1 for i in ins if re.search(r"(^|\s)" + re.escape(m) + r"(\.|\s|$)", i)) for m in MNEMONICS]
    out.append(f"{demangle(k)[:74]:74s} " + " ".join(f"{c:9d}" for c in counts) + f"   {len(ins)}")

# the default FIR kernel: the steady-state cadence of its fully unrolled tap tile
default = next((k for k in kernels if "fir_dmma_kernel" in k), None)
if default:
    ins = kernels[default]
    idx = [i for i, s in enumerate(ins) if s.startswith("DMMA") or " DMMA" in s]
    # a window in the middle of the longest run of DMMAs
    mid = idx[len(idx) // 2]
    lo = max(0, mid - 22)
    out.append("")
    out.append(f"# {demangle(default)}: {len(idx)} DMMA in {len(ins)} instructions.  Excerpt from the middle of the unrolled")
    out.append("# 256-tap tile (T = 2 tiles of 64 outputs per warp): per 8-tap step ONE LDS.128 (A fragments of both MMAs of")
    out.append("# the newest tile), TWO LDS.64 (Toeplitz tap fragments, even / odd), 2T = 4 DMMA; the ring registers feed the")
    out.append("# older tile without touching shared memory.")
    for s in ins[lo:lo + 44]:
        out.append("    " + s)

# ptxas resources
r = subprocess.run(["make", "-C", os.path.join(ROOT, "audio_fir_filter_b200", "csrc"), "ptxas-info"], capture_output=True, text=True)
lines = (r.stdout + r.stderr).splitlines()
out.append("")
out.append("# ptxas -v (registers / spills / shared memory) of the same sources")
for i, l in enumerate(lines):
    if "Compiling entry function" in l:
        name = re.search(r"'(\S+)'", l).group(1)
        res = " ".join(x.strip() for x in lines[i + 1:i + 3]).replace("ptxas info    :", "")
        out.append(f"{demangle(name)[:70]:70s} {res}")
path = os.path.join(ROOT, "profiles", f"{a.round}_sass_evidence.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path, f"({len(kernels)} kernels)")
