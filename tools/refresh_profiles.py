#!/usr/bin/env python
"""Turn what a gpurun call brought back in gpurun_out/ into the tracked evidence under
profiles/ (run here, no GPU needed):

  gpurun_out/<rep>.ncu-rep  -> profiles/<name>_ncu.txt        (tools/ncu_summary.py)
                            -> profiles/fir_traffic.json      (dram read+write of that launch)
  gpurun_out/launches_*.csv -> profiles/<name>_launches*.{csv,txt}

usage: python tools/refresh_profiles.py --rep gpurun_out/x.ncu-rep --cfg 2 --round r2
(tools/sass_evidence.py regenerates the SASS evidence from the shipped library)
"""
import argparse
import collections
import csv
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--rep", default="")
ap.add_argument("--launches", default="")
ap.add_argument("--cfg", type=int, default=2)
ap.add_argument("--round", default="r2")
ap.add_argument("--name", default="fir_dmma", help="kernel tag in the output file name")
a = ap.parse_args()
P = os.path.join(ROOT, "profiles")

if a.rep:
    out = os.path.join(P, f"{a.round}_{a.name}_cfg{a.cfg}_ncu.txt")
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), a.rep], capture_output=True,
                         text=True).stdout
    open(out, "w").write(txt)
    mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    kern, tot = "", 0.0
    for line in txt.splitlines():
        p = line.split()
        if line.startswith("# kernel:"):
            kern = line.split("kernel:")[1].split("(")[0].strip()
        if p and p[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(p[1]) * mul[p[2]]
    tp = os.path.join(P, "fir_traffic.json")
    if a.name != "fir_dmma":
        print(out, tot)
        sys.exit(0)
    d = json.load(open(tp)) if os.path.exists(tp) else {}
    d[f"cfg{a.cfg}"] = tot
    d["source"] = (f"profiles/{os.path.basename(out)} (ncu --set full, one launch of {kern} on bench.py's config "
                   f"{a.cfg}, captured {time.strftime('%Y-%m-%d')}): dram__bytes_read.sum + dram__bytes_write.sum")
    d["captured"] = time.strftime("%Y-%m-%d")
    d["note"] = ("bench.py copies this constant into roofline.traffic and names this file in roofline.traffic_source; "
                 "it is NOT re-measured by a bench run (a bench number is never taken under a profiler)")
    json.dump(d, open(tp, "w"), indent=1)
    print(out, tot)

if a.launches:
    rows = [r for r in csv.reader(open(a.launches)) if len(r) > 10 and r[0].isdigit()]
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        name = r[4].split("(")[0]
        tot[name] += float(r[-1])
        cnt[name] += 1
    T = sum(tot.values())
    shutil.copy(a.launches, os.path.join(P, f"{a.round}_launches_bench_cfg{a.cfg}.csv"))
    with open(os.path.join(P, f"{a.round}_launches_bench_cfg{a.cfg}_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 "
                "--no-cpu\n# (cold-cache, serialised; compare shares)   kernel, launches, total ns, share\n")
        for n, t in tot.most_common():
            f.write(f"{n:70s} {cnt[n]:5d} {t:14.0f} {t / T:8.4f}\n")
    print(open(os.path.join(P, f"{a.round}_launches_bench_cfg{a.cfg}_summary.txt")).read())
