#!/usr/bin/env python3
"""Times pbrt_b200_render_host (upload + render + download) on the bench workload, step by step, the way bench.py's
e2e leg calls it: a resident DeviceScene of the same scene alive next to it, frames into pinned host memory."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi
import torch
ffi = load_ffi()
import tempfile
out = tempfile.mkdtemp(prefix="pbrt_tool_")
sp = gen_scene(os.path.join(out, "e2e_c3.pbrt"), "c3", "--filename", os.path.join(out, "e2e.exr"))
sc = ffi.Scene(sp)
host_out = torch.empty((sc.height, sc.width, 4), dtype=torch.float32).pin_memory().numpy()
for resident in (False, True):
    ds = None
    if resident:
        ds = ffi.DeviceScene(sc)
        st = ds.render()
        print("resident scene rendered once: %.1f ms" % st["render_ms"])
    for i in range(5):
        t0 = time.perf_counter()
        outs, _, st = ffi.render_host(sc, [sc.full_window()], device=0, outs=[host_out])
        t1 = time.perf_counter()
        print("resident=%s step %d: wall %.1f ms  render %.1f  upload %.2f  download %.2f" % (
            resident, i, 1e3 * (t1 - t0), st["render_ms"], st["upload_ms"], st["download_ms"]))
    if ds:
        ds.close()

# the same steps spelled out through the resident-scene calls, to see which one a slow step spends its time in
ds = ffi.DeviceScene(sc)
ds.render()
for i in range(8):
    t0 = time.perf_counter()
    d2 = ffi.DeviceScene(sc)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    st = d2.render()
    t2 = time.perf_counter()
    img = d2.download(0)
    t3 = time.perf_counter()
    d2.close()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print("spelled out step %d: create %.2f ms  render %.2f (device %.2f)  download %.2f  destroy %.2f" % (
        i, 1e3 * (t1 - t0), 1e3 * (t2 - t1), st["render_ms"], 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
ds.close()
