#!/usr/bin/env python3
"""profiles/traffic.json from one `ncu --set full` capture of the first launches of the traversal kernels:
real DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per ray of k_intersect / k_shadow, the ray counts taken
from the library's own queue-size log of the same run (tunable 6).
usage: traffic_from_ncu.py report.ncu-rep render.log "source note" > traffic.json"""
import csv
import io
import json
import re
import subprocess
import sys

rep, log, note = sys.argv[1], sys.argv[2], sys.argv[3]
counts = {"main": [], "shadow": []}
for line in open(log, errors="replace"):
    m = re.search(r"pbrt_b200 wave: iteration (\d+) main (\d+) shadow (\d+) probe (\d+)", line)
    if m and len(counts["main"]) <= int(m.group(1)):   # first wave only
        counts["main"].append(int(m.group(2)))
        counts["shadow"].append(int(m.group(3)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")


def to_bytes(v, unit):
    return float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


out = {}
seen = {"k_intersect": 0, "k_shadow": 0}
acc = {"k_intersect": [0.0, 0], "k_shadow": [0.0, 0]}
for r in rows[2:]:
    for k, q in (("k_intersect", "main"), ("k_shadow", "shadow")):
        if k + "<" in r[ik]:
            i = seen[k]
            seen[k] += 1
            if i < len(counts[q]) and counts[q][i] > 0:
                acc[k][0] += to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
                acc[k][1] += counts[q][i]
for k, (b, n) in acc.items():
    if n:
        out[k] = {"dram_bytes_per_ray": b / n, "rays": n, "launches": seen[k], "source": note}
print(json.dumps(out, indent=1))
