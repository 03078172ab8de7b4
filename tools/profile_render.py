#!/usr/bin/env python3
"""One render of a BASELINE config for profiling under ncu:  python tools/profile_render.py [c3|c4|c5] [gen args...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi  # noqa: E402

ffi = load_ffi()
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
import tempfile
out = tempfile.mkdtemp(prefix="pbrt_tool_")   # (scene files are large: keep them out of gpurun_out/)
sp = gen_scene(os.path.join(out, "prof_%s.pbrt" % name), name, "--filename", os.path.join(out, "prof.exr"), *sys.argv[2:])
sc = ffi.Scene(sp)
ds = ffi.DeviceScene(sc)
if os.environ.get("PBRT_PROFILE_WAVE"):   # (ncu's kernel replay saves and restores the wave arena: keep it small)
    ffi.cuda().pbrt_b200_tunable(0, int(os.environ["PBRT_PROFILE_WAVE"]))
if os.environ.get("PBRT_PROFILE_COUNTS"):
    ffi.cuda().pbrt_b200_tunable(6, 1)
st = ds.render()
rays = st["intersect_rays"] + st["shadow_rays"]
print("render %.2f ms, %.1f Mrays, %.1f Mrays/s, %d launches" % (st["render_ms"], rays / 1e6, rays / st["render_ms"] / 1e3,
                                                              st["kernel_launches"]))
