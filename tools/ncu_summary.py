#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the text committed under profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_fp64.avg.pct", "sm__pipe_fp64_cycles_active.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct",
        "lts__t_bytes.sum", "sm__throughput.avg.pct", "dram__cycles_active.avg.pct"]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"# kernel: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}  (source: {rep})")
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in KEYS) and "_op_" not in h and ".min" not in h and ".max" not in h:
            print(f"{h:95s} {v:>20s} {u}")
