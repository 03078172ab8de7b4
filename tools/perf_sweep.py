#!/usr/bin/env python3
"""Render a config under several tunable settings and print per-kernel-class times.
Usage: python tools/perf_sweep.py c3 [gen args] -- key=value[,key=value] ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi  # noqa: E402

ffi = load_ffi()
argv = sys.argv[1:]
if "--" in argv:
    i = argv.index("--")
    gen, variants = argv[:i], argv[i + 1:]
else:
    gen, variants = argv, [""]
name = gen[0] if gen else "c3"
import tempfile
out = tempfile.mkdtemp(prefix="pbrt_tool_")   # (scene files are large: keep them out of gpurun_out/)
sp = gen_scene(os.path.join(out, "sweep_%s.pbrt" % name), name, "--filename", os.path.join(out, "sweep.exr"), *gen[1:])
sc = ffi.Scene(sp)
lib = ffi.cuda()
defaults = {k: lib.pbrt_b200_tunable(k, -1) for k in range(12)}
for var in variants:
    for k, v in defaults.items():
        if v >= 0:
            lib.pbrt_b200_tunable(k, v)
    for kv in filter(None, var.split(",")):
        k, v = kv.split("=")
        lib.pbrt_b200_tunable(int(k), int(v))
    ds = ffi.DeviceScene(sc)
    ds.render()
    lib.pbrt_b200_tunable(3, 1)
    best = None
    for _ in range(3):
        st = ds.render()
        if best is None or st["render_ms"] < best[0]["render_ms"]:
            best = (st, ds.kernel_stats())
    lib.pbrt_b200_tunable(3, 0)
    st, ks = best
    rays = st["intersect_rays"] + st["shadow_rays"]
    print("[%s] %.2f ms  %.1f Mrays/s | %s" % (var or "default", st["render_ms"], rays / st["render_ms"] / 1e3,
          "  ".join("%s %.1f" % (k.replace("k_", ""), v["ms"]) for k, v in ks.items() if v["ms"] > 0.05)), flush=True)
    ds.close()
