#!/usr/bin/env python3
"""Differential fuzzing of the oracle twin against the COMPILED REFERENCE (build container only: needs
oracle/_ref).  Random scenes mixing every supported plugin -- shapes, materials, lights, cameras,
samplers, integrators, filters, instancing, constant textures -- are rendered by
oracle/_ref/pbrt_ref_keyed and by the glibc-float-libm twin; every per-sample radiance and the image
must agree bit for bit.  usage: fuzz_parity.py [n_scenes] [seed]"""
import ctypes as C
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import load_ffi, read_sample_dump  # noqa: E402

ffi = load_ffi()
assert ffi.have_reference(), "oracle/_ref missing"
lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libpbrt_twin_libm.so"))
lib.pbrt_twin_sample_count.restype = C.c_uint64
lib.pbrt_twin_sample_count.argtypes = [C.c_void_p, C.c_void_p]
lib.pbrt_twin_render.argtypes = [C.c_void_p] * 6


from fuzz_scenes import scene  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    td = tempfile.mkdtemp(prefix="pbrt_fuzz_")
    bad = 0
    for i in range(n):
        r = random.Random(seed * 100003 + i)
        sp = os.path.join(td, "fuzz_%d.pbrt" % i)
        open(sp, "w").write(scene(r, i))
        dump = os.path.join(td, "fuzz_%d.samples" % i)
        try:
            ffi.run_reference(sp, keyed=True, dump_samples=dump, cwd=td)
        except Exception as e:   # the reference itself died (e.g. Severe on a degenerate shape)
            print("scene %d: reference failed (%s); skipped" % (i, str(e)[:80]))
            continue
        ref_rgba, _ = ffi.read_raw(os.path.join(td, "fuzz_%d.exr.raw" % i))
        ref_samples = read_sample_dump(dump)
        cwd = os.getcwd()
        os.chdir(td)
        try:
            sc = ffi.Scene(sp)
        finally:
            os.chdir(cwd)
        w = sc.full_window()
        rgba = np.zeros((w.height, w.width, 4), np.float32)
        ns = lib.pbrt_twin_sample_count(sc.ptr, C.byref(w))
        samples = np.zeros((ns, 6), np.float32)
        st = ffi.Stats()
        rc = lib.pbrt_twin_render(sc.ptr, C.byref(w), rgba.ctypes.data_as(C.c_void_p), None, samples.ctypes.data_as(C.c_void_p), C.byref(st))
        ok = rc == 0 and samples.shape == ref_samples.shape and np.array_equal(samples.view(np.uint32), ref_samples.view(np.uint32)) and \
            np.array_equal(rgba.view(np.uint32), ref_rgba.view(np.uint32))
        if not ok:
            bad += 1
            nd = -1
            if samples.shape == ref_samples.shape:
                nd = int((samples.view(np.uint32) != ref_samples.view(np.uint32)).any(axis=1).sum())
            print("scene %d: MISMATCH (rc %d, samples %s vs %s, differing samples %d) -> %s" % (i, rc, samples.shape, ref_samples.shape, nd, sp))
        elif i % 10 == 0:
            print("scene %d ok (%d samples)" % (i, ns), flush=True)
    print("scenes: %d, mismatches: %d (kept in %s)" % (n, bad, td))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
