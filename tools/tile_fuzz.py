#!/usr/bin/env python3
"""Tile mode on seeded fuzz scenes (tests/fuzz_scenes.py): the film window cut into four tiles and rendered through
pbrt_b200_render_tiles must equal the whole-frame render bit for bit (scenes with the bestcandidate sampler are
skipped: tile mode does not support it).  usage: tile_fuzz.py N [first_seed]"""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import load_ffi
from fuzz_scenes import scene as fuzz_scene
ffi = load_ffi()
n = int(sys.argv[1]); first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bad = skipped = done = 0
for seed in range(first, first + n):
    txt = fuzz_scene(random.Random(500009 + seed), seed)
    if '"bestcandidate"' in txt:
        skipped += 1
        continue
    sp = "/tmp/tile_fuzz_%d.pbrt" % seed
    open(sp, "w").write(txt)
    sc = ffi.Scene(sp)
    ds = ffi.DeviceScene(sc)
    ds.render()
    full = ds.download(0)
    w = sc.full_window()
    hx, hy = max(1, w.width // 2), max(1, w.height // 2)
    tiles = [t for t in (ffi.Window(w.x0, w.y0, hx, hy), ffi.Window(w.x0 + hx, w.y0, w.width - hx, hy),
                         ffi.Window(w.x0, w.y0 + hy, hx, w.height - hy), ffi.Window(w.x0 + hx, w.y0 + hy, w.width - hx, w.height - hy))
             if t.width > 0 and t.height > 0]
    try:
        ds.render_tiles(tiles)
    except RuntimeError as e:
        if "bestcandidate" in str(e):      # (the scene named no sampler: pbrt's default is bestcandidate)
            skipped += 1
            ds.close(); sc.close(); os.remove(sp)
            continue
        raise
    ok = True
    for i, t in enumerate(tiles):
        got = ds.download(i)
        want = full[t.y0 - w.y0:t.y0 - w.y0 + t.height, t.x0 - w.x0:t.x0 - w.x0 + t.width]
        if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
            ok = False
    done += 1
    if not ok:
        bad += 1
        print("seed %d: tiles differ from the frame" % seed, flush=True)
    ds.close(); sc.close(); os.remove(sp)
print("tile fuzz: %d scenes, %d skipped (bestcandidate), %d with differing tiles" % (done, skipped, bad))
