#!/usr/bin/env python3
"""Developer check on a GPU box: GPU stages vs the oracle twin on small scenes, then a first
throughput number.  Usage: python tools/gpu_check.py [--big]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, image_error_stats, load_ffi  # noqa: E402

ffi = load_ffi()
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)


def check(name, *args):
    sp = gen_scene(os.path.join(out, "chk_%s.pbrt" % name), name, "--filename", os.path.join(out, "x.exr"), *args)
    sc = ffi.Scene(sp)
    ds = ffi.DeviceScene(sc)
    rays_t, xy_t = ffi.twin_camera_rays(sc)
    rays_g, xy_g = ds.camera_rays(ffi.twin_sample_count(scene, scene.full_window()))
    print(name, "camera rays identical:", bool((rays_t.view(np.uint32) == rays_g.view(np.uint32)).all()),
          bool((xy_t == xy_g).all()))
    pt, tt, b1t, b2t = ffi.twin_intersect(sc, rays_t)
    pg, tg, b1g, b2g = ds.intersect(rays_t)
    print(name, "closest hits identical:", bool((pt == pg).all()), bool((tt == tg).all()),
          bool((b1t == b1g).all() and (b2t == b2g).all()))
    ffi.cuda().pbrt_b200_tunable(1, 1)
    ffi.cuda().pbrt_b200_sample_radiance(ds.h, 0, 1, None)
    st = ds.render()
    rgba, raw = ds.download(0, want_raw=True)
    ns = ffi.twin_sample_count(sc, sc.full_window())
    sg = ds.sample_radiance(0, ns)
    rt, rawt, st_t, stt = ffi.twin_render(sc, want_samples=True)
    d = np.abs(sg[:, :3] - st_t[:, 2:5]).max(axis=1)
    print(name, "samples: n=%d identical=%.5f  >1e-4: %d  max %.4g" % (ns, (d == 0).mean(), (d > 1e-4).sum(), d.max()))
    print(name, "image:", image_error_stats(rgba, rt))
    keys = ["camera_rays", "intersect_rays", "shadow_rays", "points_shaded", "interior_visits", "leaf_visits",
            "prim_refs_read", "prim_tests", "nan_samples"]
    print(name, "gpu  stats:", {k: st[k] for k in keys}, "ms %.3f launches %d" % (st["render_ms"], st["kernel_launches"]))
    print(name, "twin stats:", {k: stt[k] for k in keys})
    ds.close()
    ffi.cuda().pbrt_b200_tunable(1, 0)


def perf(name, *args):
    sp = gen_scene(os.path.join(out, "perf_%s.pbrt" % name), name, "--filename", os.path.join(out, "x.exr"), *args)
    t0 = time.time()
    sc = ffi.Scene(sp)
    t1 = time.time()
    ds = ffi.DeviceScene(sc)
    for i in range(3):
        st = ds.render()
        rays = st["intersect_rays"] + st["shadow_rays"]
        print("%s %s: parse+kd %.2fs  render %.2f ms  rays %.1fM  -> %.1f Mrays/s  launches %d" % (
            name, " ".join(map(str, args)), t1 - t0, st["render_ms"], rays / 1e6, rays / st["render_ms"] / 1e3,
            st["kernel_launches"]), flush=True)
    ds.close()


if __name__ == "__main__":
    check("c1", "--res", "64", "64")
    check("c2", "--res", "64", "64", "--spp", "2", "2")
    check("c3", "--res", "64", "64", "--spp", "4", "4")
    check("c3", "--res", "32", "32", "--spp", "2", "2", "--filter", "mitchell")
    check("c4", "--res", "48", "48", "--spp", "2", "2", "--nphi", "64", "--ntheta", "32")
    perf("c3", "--res", "512", "512", "--spp", "4", "4")
    if "--big" in sys.argv:
        perf("c3")
        perf("c4", "--res", "512", "512", "--spp", "4", "4")
