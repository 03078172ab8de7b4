#!/usr/bin/env python3
"""Times pbrt_b200_render_host (upload + render + download) on the bench workload, step by step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi
ffi = load_ffi()
out = os.path.join(ROOT, "gpurun_out"); os.makedirs(out, exist_ok=True)
sp = gen_scene(os.path.join(out, "e2e_c3.pbrt"), "c3", "--filename", os.path.join(out, "e2e.exr"))
sc = ffi.Scene(sp)
for i in range(5):
    t0 = time.perf_counter()
    outs, _, st = ffi.render_host(sc)
    t1 = time.perf_counter()
    print("step %d: wall %.1f ms  render %.1f  upload %.2f  download %.2f" % (i, 1e3 * (t1 - t0), st["render_ms"], st["upload_ms"], st["download_ms"]))
