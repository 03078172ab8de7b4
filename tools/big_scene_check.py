#!/usr/bin/env python3
"""BASELINE configs[3] / [4] class scenes (100k / 1M triangle blob): GPU vs oracle twin on a crop, then
full-size throughput with the kd-walk counters and the roofline arithmetic of SURVEY 8(d).
usage: big_scene_check.py c4|c5|inst [--full]   (inst: 27 instances of one 20 k-triangle object)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi  # noqa: E402

ffi = load_ffi()
name = sys.argv[1]
import tempfile
out = tempfile.mkdtemp(prefix="pbrt_tool_")   # (scene files are large: keep them out of gpurun_out/)
res, spp = ((1024, 1024), (8, 8)) if name in ("c4", "inst") else ((1920, 1080), (16, 16))
if "--full" not in sys.argv:
    res, spp = (res[0] // 4, res[1] // 4), (2, 2)
t0 = time.time()
sp = gen_scene(os.path.join(out, "big_%s.pbrt" % name), name, "--res", res[0], res[1], "--spp", spp[0], spp[1],
               "--filename", os.path.join(out, "big.exr"))
t1 = time.time()
sc = ffi.Scene(sp)
t2 = time.time()
print("%s: %d prims, %d kd nodes, %d prim refs; generate %.1fs parse+kd build %.1fs" % (name, sc.n_prims, sc.n_nodes,
                                                                                         sc.n_prim_refs, t1 - t0, t2 - t1), flush=True)
ds = ffi.DeviceScene(sc)
# parity on a crop: 32x32 pixels in the middle of the blob
w = ffi.Window(res[0] // 2 - 16, res[1] // 2 - 16, 32, 32)
ds.render([w])
g = ds.download(0)
t, _, _, stt = ffi.twin_render(sc, w)
print("crop parity: bit-identical to the oracle twin:", bool(np.array_equal(g.view(np.uint32), t.view(np.uint32))), flush=True)
lib = ffi.cuda()
lib.pbrt_b200_tunable(1, 1)
st = ds.render()
lib.pbrt_b200_tunable(1, 0)
ks = ds.kernel_stats()
for k in ("k_intersect", "k_shadow", "k_probe"):
    v = ks[k]
    if v["rays"]:
        b = (8 * v["interior_visits"] + 8 * v["leaf_visits"] + 4 * v["prim_refs_read"] + 48 * v["prim_tests"]) / v["rays"] + 64
        print("%-12s rays %.1fM  per ray: interior %.1f leaves %.1f refs %.1f tests %.1f -> %.0f algorithmic B/ray" % (
            k, v["rays"] / 1e6, v["interior_visits"] / v["rays"], v["leaf_visits"] / v["rays"], v["prim_refs_read"] / v["rays"],
            v["prim_tests"] / v["rays"], b))
lib.pbrt_b200_tunable(3, 1)
for _ in range(2):
    st = ds.render()
lib.pbrt_b200_tunable(3, 0)
ks2 = ds.kernel_stats()
rays = st["intersect_rays"] + st["shadow_rays"]
print("full frame %dx%d %d spp: %.1f ms, %.1f Mrays, %.1f Mrays/s" % (res[0], res[1], spp[0] * spp[1], st["render_ms"], rays / 1e6,
                                                                      rays / st["render_ms"] / 1e3))
print("  ".join("%s %.1f" % (k.replace("k_", ""), v["ms"]) for k, v in ks2.items() if v["ms"] > 0.05))
v, v2 = ks["k_intersect"], ks2["k_intersect"]
bytes_ = 8 * v["interior_visits"] + 8 * v["leaf_visits"] + 4 * v["prim_refs_read"] + 48 * v["prim_tests"] + 64 * v["rays"]
print("k_intersect roofline: %.0f GB/s algorithmic = %.3f of 6549 GB/s measured HBM peak" % (bytes_ / v2["ms"] / 1e6,
                                                                                          bytes_ / v2["ms"] / 1e6 / 6549.1))
