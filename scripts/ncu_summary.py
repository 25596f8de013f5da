#!/usr/bin/env python3
"""Summarise an .ncu-rep (read offline with `ncu -i`): headline metrics + dynamic SASS mix per tile.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [rows_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, data = r[0], r[1], r[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
d = data[0]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:82s} {d[i]:>18s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(src)))
h = None
byop = collections.Counter(); stall = collections.Counter(); tot = 0
nk = 0
for row in rr:
    if row and row[0] == "Address":
        h = row; nk += 1; continue
    if h is None or len(row) < len(h):
        continue
    if nk > 1:
        break
    try:
        e = int(row[h.index("Instructions Executed")]); s = int(row[h.index("Warp Stall Sampling (All Samples)")])
    except Exception:
        continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", row[h.index("Source")])
    op = m.group(2) if m else "?"
    byop[op] += e; stall[op] += s; tot += e
tiles = rows / 32
print(f"\nwarp-instructions executed: {tot}  = {tot / tiles:.1f} per 32-row tile (per thread-row)")
for op, c in byop.most_common(28):
    print(f"  {op:10s} {c / tiles:8.1f} /tile   stall samples {stall[op]}")
