#!/usr/bin/env python3
"""Host time per frame outside the device's kernels, for a multi-GPU rank's share rendered on ONE GPU:
the tiles rank 0 of an 8-rank job owns, through pbrt_b200_render_tiles, in a frame loop -- wall clock per call
against the device time the call reports."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi
import tempfile
ffi = load_ffi()
out = tempfile.mkdtemp(prefix="pbrt_tool_")
sp = gen_scene(os.path.join(out, "gap_c3.pbrt"), "c3", "--filename", os.path.join(out, "gap.exr"))
sc = ffi.Scene(sp)
ds = ffi.DeviceScene(sc)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tiles = [t for t, o in ffi.frame_partition(sc.full_window(), world, 0) if o == 0]
print("%d tiles of rank 0 of %d" % (len(tiles), world))
for i in range(3):
    ds.render_tiles(tiles)
walls, devs = [], []
for i in range(20):
    t0 = time.perf_counter()
    st = ds.render_tiles(tiles)
    walls.append(1e3 * (time.perf_counter() - t0)); devs.append(st["render_ms"])
walls.sort(); devs.sort()
print("render_tiles: wall median %.3f ms (min %.3f), device median %.3f ms -> host gap %.3f ms per frame, %d launches" % (
    walls[10], walls[0], devs[10], walls[10] - devs[10], st["kernel_launches"]))
walls, devs = [], []
for i in range(20):
    t0 = time.perf_counter()
    st = ds.render_frame(None, 0, None)
    walls.append(1e3 * (time.perf_counter() - t0)); devs.append(st["render_ms"])
walls.sort(); devs.sort()
print("whole frame (render_frame, one rank): wall median %.3f ms, device median %.3f ms -> gap %.3f ms" % (walls[10], devs[10], walls[10] - devs[10]))
