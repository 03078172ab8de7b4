#!/usr/bin/env python3
"""One seeded fuzz scene (tests/fuzz_scenes.py) on the GPU against the twin: how many samples differ, by how much,
and whether the walk counters agree.  usage: fuzz_one.py SEED"""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import load_ffi
from fuzz_scenes import scene as fuzz_scene
ffi = load_ffi()
seed = int(sys.argv[1])
sp = "/tmp/fuzz_%d.pbrt" % seed
open(sp, "w").write(fuzz_scene(random.Random(500009 + seed), seed))
scene = ffi.Scene(sp)
ds = ffi.DeviceScene(scene)
lib = ffi.cuda()
lib.pbrt_b200_tunable(1, 1)
lib.pbrt_b200_sample_radiance(ds.h, 0, 1, None)
st = ds.render()
lib.pbrt_b200_tunable(1, 0)
rgba_t, raw_t, samples_t, st_t = ffi.twin_render(scene, want_samples=True)
ls = ds.sample_radiance(0, samples_t.shape[0])
a = ls[:, :3]; b = samples_t[:, 2:5]
diff = np.any(a.view(np.uint32) != b.view(np.uint32), axis=1)
print("lib %s: %d of %d samples differ" % (os.environ.get("PBRT_B200_LIB", "default"), int(diff.sum()), len(diff)))
for i in np.nonzero(diff)[0][:8]:
    print("  sample %d: gpu %s twin %s  ulp diff %s" % (i, a[i], b[i], np.abs(a[i].view(np.int32).astype(np.int64) - b[i].view(np.int32).astype(np.int64))))
for k in ("camera_rays", "intersect_rays", "shadow_rays", "points_shaded", "interior_visits", "leaf_visits", "prim_refs_read", "prim_tests",
          "nan_samples", "negative_samples", "infinite_samples"):
    print("  %-16s gpu %d twin %d %s" % (k, st[k], st_t[k], "" if st[k] == st_t[k] else "<-- differs"))
