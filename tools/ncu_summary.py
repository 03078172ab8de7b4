#!/usr/bin/env python3
"""Text summary of an .ncu-rep (runs on the GPU box, so that the multi-MB report need not travel):
key launch / issue / memory metrics per captured kernel, then the hottest SASS instructions by warp-stall samples
with their executed-instruction share and mean active threads.
usage: ncu_summary.py report.ncu-rep [topN] > summary.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print("== %s" % rep)
for r in rows[2:]:
    print("\n-- kernel: %s" % r[hdr.index("Kernel Name")])
    for w in WANT:
        if w in hdr:
            print("   %-92s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
for b in blocks:
    h = b["hdr"]
    ia, it, isamp, ithr = (h.index(n) for n in ("Instructions Executed", "Thread Instructions Executed", "# Samples", "Avg. Threads Executed"))
    tot = sum(int(r[ia]) for r in b["rows"]) or 1
    tots = sum(int(r[isamp]) for r in b["rows"]) or 1
    thr = sum(int(r[it]) for r in b["rows"])
    print("\n-- SASS of %s: %d instructions, %d warp instructions executed, %.2f active threads per instruction" % (
        b["name"][:60], len(b["rows"]), tot, thr / tot))
    print("   idx  exec%  samples%  threads  instruction")
    order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][isamp]))[:top]
    for i in sorted(order):
        r = b["rows"][i]
        print("   %4d %5.2f %8.2f %8s  %s" % (i, 100.0 * int(r[ia]) / tot, 100.0 * int(r[isamp]) / tots, r[ithr], r[1].strip()[:80]))
    # executed-instruction share and mean threads of the memory / test instructions (structure of the trip loop)
    for pat in ("LDG", "LDL", "STL", "LDS", "STS", "VOTE", "BRA"):
        sel = [r for r in b["rows"] if pat in r[1]]
        e = sum(int(r[ia]) for r in sel)
        t = sum(int(r[it]) for r in sel)
        if e:
            print("   %-5s %5.2f %% of executed warp instructions, %.1f threads" % (pat, 100.0 * e / tot, t / e))
