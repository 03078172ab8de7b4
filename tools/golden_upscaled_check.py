#!/usr/bin/env python3
"""GPU vs oracle twin, bit for bit, on the golden scenes re-rendered at a larger resolution (the committed
fixtures are 20-50 pixels wide).  usage: golden_upscaled_check.py [scale] [name-substring ...]"""
import glob
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import load_ffi  # noqa: E402

ffi = load_ffi()
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 6
subs = sys.argv[2:]
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
bad = 0
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.pbrt"))):
    name = os.path.basename(path)[:-5]
    if subs and not any(s in name for s in subs):
        continue
    text = open(path).read()
    text = re.sub(r'"integer xresolution" \[(\d+)\]', lambda m: '"integer xresolution" [%d]' % (int(m.group(1)) * scale), text)
    text = re.sub(r'"integer yresolution" \[(\d+)\]', lambda m: '"integer yresolution" [%d]' % (int(m.group(1)) * scale), text)
    sp = os.path.join(out, "up_%s.pbrt" % name)
    open(sp, "w").write(text)
    sc = ffi.Scene(sp)
    ds = ffi.DeviceScene(sc)
    ffi.cuda().pbrt_b200_tunable(1, 1)
    st = ds.render()
    ffi.cuda().pbrt_b200_tunable(1, 0)
    rgba, raw = ds.download(0, want_raw=True)
    rt, rawt, _, stt = ffi.twin_render(sc)
    same = np.array_equal(raw.view(np.uint32), rawt.view(np.uint32))
    counters = all(st[k] == stt[k] for k in ("camera_rays", "intersect_rays", "shadow_rays", "points_shaded", "prim_tests"))
    print("%-36s %4dx%-4d film identical: %s  counters identical: %s  (%d rays)" % (
        name, sc.width, sc.height, same, counters, st["intersect_rays"] + st["shadow_rays"]), flush=True)
    bad += (not same) or (not counters)
    ds.close()
print("mismatching scenes:", bad)
sys.exit(1 if bad else 0)
