#!/usr/bin/env python3
"""BASELINE configs[2] (Cornell box + glass sphere, path, 1024x1024, 64 spp) with textured walls: a 1024x1024
image map behind a checkerboard / scale expression, marble, and an fbm bump map on every wall.  Checks a
crop against the oracle twin bit for bit, then times the full frame next to the untextured scene.
usage: textured_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi, write_exr  # noqa: E402

ffi = load_ffi()
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
tex = os.path.join(out, "wall_1024.exr")
yy, xx = np.mgrid[0:1024, 0:1024].astype(np.float32)
rng = np.random.RandomState(5)
img = np.stack([.5 + .5 * np.sin(xx / 9) * np.cos(yy / 13), .5 + .5 * np.sin((xx + yy) / 21), .5 + .5 * np.cos(xx / 5 - yy / 17)], -1)
write_exr(tex, (img * (.6 + .4 * rng.rand(1024, 1024, 1))).astype(np.float32))
lib = ffi.cuda()
res = {}
for name, extra in (("plain", []), ("textured", ["--textured", tex])):
    sp = gen_scene(os.path.join(out, "c3_%s.pbrt" % name), "c3", "--filename", os.path.join(out, "t.exr"), *extra)
    sc = ffi.Scene(sp)
    ds = ffi.DeviceScene(sc)
    w = ffi.Window(480, 600, 32, 32)
    ds.render([w])
    g = ds.download(0)
    t, _, _, _ = ffi.twin_render(sc, w)
    same = bool(np.array_equal(g.view(np.uint32), t.view(np.uint32)))
    lib.pbrt_b200_tunable(3, 1)
    for _ in range(3):
        st = ds.render()
    lib.pbrt_b200_tunable(3, 0)
    ks = ds.kernel_stats()
    rays = st["intersect_rays"] + st["shadow_rays"]
    res[name] = st["render_ms"]
    print("%-9s crop bit-identical to the twin: %s; full frame %.1f ms, %.1f Mrays, %.1f Mrays/s" % (
        name, same, st["render_ms"], rays / 1e6, rays / st["render_ms"] / 1e3))
    print("   " + "  ".join("%s %.1f" % (k.replace("k_", ""), v["ms"]) for k, v in ks.items() if v["ms"] > 0.05), flush=True)
    ds.close()
print("textured / plain frame time: %.3f" % (res["textured"] / res["plain"]))
