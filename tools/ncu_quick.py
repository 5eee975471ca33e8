#!/usr/bin/env python3
"""Prints the handful of ncu metrics the kernel notes cite, plus the shared-memory instructions with excess wavefronts.
usage: python tools/ncu_quick.py prof.ncu-rep [ncells]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ncell = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg"]
d = {}
for h, u, v in zip(hdr, units, vals):
    if h in want:
        d[h] = (v, u)
        print(f"{h:90s} {v:>16s} {u}")
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.3:
        print(f"{h:90s} {v:>16s}")
if ncell:
    w = float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0])
    c = float(d["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"][0])
    print(f"per cell: smem wavefronts {w / ncell:.0f}, of them bank-conflict replays {c / ncell:.0f} ({100 * c / w:.0f} %)")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
tot_w = tot_e = 0
print("shared-memory instructions with excess wavefronts (index, executed, wavefronts, ideal, SASS):")
for i, r in enumerate(rows[2:]):
    w = int(r[idx["L1 Wavefronts Shared"]] or 0)
    e = int(r[idx["L1 Wavefronts Shared Excessive"]] or 0)
    tot_w += w
    tot_e += e
    if e * 10 > w and w > 0:
        print(f"  {i:5d} {r[idx['Instructions Executed']]:>9s} {w:>10d} {r[idx['L1 Wavefronts Shared Ideal']]:>10s}  {r[idx['Source']].strip()[:60]}")
print(f"source page: shared wavefronts {tot_w}, excessive {tot_e}")
