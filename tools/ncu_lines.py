#!/usr/bin/env python3
"""Per-source-line dynamic instruction and stall histogram from an ncu report + the matching .so.
usage: ncu_lines.py report.ncu-rep lib.so kernel_substring [topN [mangled_substring]]
(mangled_substring picks the template instance in the disassembly, e.g. k_path_fastILb0ELi0E)"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
b = [x for x in blocks if kname in x["name"]][0]
hdr = b["hdr"]
ia, it, isrc, ismp = (hdr.index(n) for n in ("Instructions Executed", "Thread Instructions Executed", "Source", "# Samples"))
ex = [(int(r[ia]), int(r[it]), r[isrc].strip(), int(r[ismp])) for r in b["rows"]]
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cubin)], capture_output=True, text=True).stdout.split("\n")
mangled = sys.argv[5] if len(sys.argv) > 5 else re.sub(r"[^A-Za-z0-9_]", "", kname)
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l][0]
end = [i for i, l in enumerate(dis) if l.startswith("//--------------------- .text.") and i > start][0]
curl, sass = None, []
for line in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', line)
    if m:
        curl = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        sass.append((curl, m.group(2)))
n = min(len(ex), len(sass))
mism = sum(1 for k in range(n) if ex[k][2].split()[:1] != sass[k][1].split()[:1])
tot_i = sum(e[0] for e in ex)
tot_t = sum(e[1] for e in ex)
tot_s = sum(e[3] for e in ex)
print("%s: sass %d/%d opcode mismatches %d | warp-inst %d avg threads %.2f samples %d" % (b["name"][:40], len(ex), len(sass), mism, tot_i, tot_t / tot_i, tot_s))
byline = collections.defaultdict(lambda: [0, 0, 0])
for k in range(n):
    a = byline[sass[k][0]]
    a[0] += ex[k][0]; a[1] += ex[k][1]; a[2] += ex[k][3]
print("  inst%  smpl%  thr  line")
for k, v in sorted(byline.items(), key=lambda kv: -kv[1][2])[:top]:
    print("  %5.1f  %5.1f  %4.1f  %s" % (100 * v[0] / tot_i, 100 * v[2] / tot_s, v[1] / max(1, v[0]), k))
