"""ctypes bindings of the CPU twin (oracle/_build/libpbrt_twin.so, built from oracle/twin.cpp).

TEST INFRASTRUCTURE: only tests/, bench.py's checker legs and __graft_entry__.smoke() use these; the product
package (pbrt-v1_b200/) does not know the twin exists.  parity_util.load_ffi() attaches the functions below to the
product's ctypes module so that a test can write ffi.twin_render(...) next to ffi.render_host(...).
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TWIN_LIB = os.environ.get("PBRT_TWIN_LIB", os.path.join(ROOT, "oracle", "_build", "libpbrt_twin.so"))
Stats = None   # set by attach(): the product binding's PbrtB200Stats mirror


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


_twin = None


def twin():
    global _twin
    if _twin is None:
        lib = C.CDLL(TWIN_LIB)
        lib.pbrt_twin_last_error.restype = C.c_char_p
        lib.pbrt_twin_sample_count.restype = C.c_uint64
        lib.pbrt_twin_sample_count.argtypes = [C.c_void_p, C.c_void_p]
        lib.pbrt_twin_render.argtypes = [C.c_void_p] * 6
        lib.pbrt_twin_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 4
        lib.pbrt_twin_intersect_p.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.pbrt_twin_camera_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        _twin = lib
    return _twin


def twin_render(scene, window=None, want_samples=False):
    w = window or scene.full_window()
    rgba = np.zeros((w.height, w.width, 4), np.float32)
    raw = np.zeros((w.height, w.width, 5), np.float32)
    ns = twin().pbrt_twin_sample_count(scene.ptr, C.byref(w))
    samples = np.zeros((ns, 6), np.float32) if want_samples else None
    st = Stats()
    rc = twin().pbrt_twin_render(scene.ptr, C.byref(w), _fp(rgba), _fp(raw),
                                 _fp(samples) if want_samples else None, C.byref(st))
    if rc != 0:
        raise RuntimeError("twin render failed: %s" % twin().pbrt_twin_last_error().decode())
    return rgba, raw, samples, st.as_dict()


def twin_intersect(scene, rays):
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]
    prim = np.zeros(n, np.int32)
    t = np.zeros(n, np.float32)
    b1 = np.zeros(n, np.float32)
    b2 = np.zeros(n, np.float32)
    rc = twin().pbrt_twin_intersect(scene.ptr, _fp(rays), n, _fp(prim), _fp(t), _fp(b1), _fp(b2))
    assert rc == 0
    return prim, t, b1, b2


def twin_intersect_p(scene, rays):
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]
    occ = np.zeros(n, np.uint8)
    rc = twin().pbrt_twin_intersect_p(scene.ptr, _fp(rays), n, _fp(occ))
    assert rc == 0
    return occ


def twin_camera_rays(scene, window=None, n=None):
    w = window or scene.full_window()
    total = twin().pbrt_twin_sample_count(scene.ptr, C.byref(w))
    n = total if n is None else min(n, total)
    rays = np.zeros((n, 8), np.float32)
    xy = np.zeros((n, 2), np.float32)
    rc = twin().pbrt_twin_camera_rays(scene.ptr, C.byref(w), n, _fp(rays), _fp(xy))
    assert rc == 0
    return rays, xy


def twin_render_tile(scene, frame, tile):
    """TEST ONLY: the oracle twin's restatement of tile semantics (pbrt_b200_render_tiles)."""
    lib = twin()
    lib.pbrt_twin_render_tile.argtypes = [C.c_void_p] * 6
    rgba = np.zeros((tile.height, tile.width, 4), np.float32)
    st = Stats()
    rc = lib.pbrt_twin_render_tile(scene.ptr, C.byref(frame), C.byref(tile), _fp(rgba), None, C.byref(st))
    if rc != 0:
        raise RuntimeError("twin tile render failed (%d)" % rc)
    return rgba, st.as_dict()


def twin_sample_count(scene, w):
    return int(twin().pbrt_twin_sample_count(scene.ptr, C.byref(w)))




def twin_intersect_treelets(scene, rays, paired):
    """The closest-hit walk over the reference's node array (paired = 0) or the device-side treelet order
    (csrc/pb_treelet.h, paired = 1); returns hits and the summed walk counters."""
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]
    prim = np.zeros(n, np.int32)
    t = np.zeros(n, np.float32)
    b1 = np.zeros(n, np.float32)
    b2 = np.zeros(n, np.float32)
    counters = np.zeros(4, np.uint64)
    lib = twin()
    lib.pbrt_twin_intersect_treelets.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 5 + [C.c_int]
    rc = lib.pbrt_twin_intersect_treelets(scene.ptr, _fp(rays), n, _fp(prim), _fp(t), _fp(b1), _fp(b2), _fp(counters), paired)
    assert rc == 0, rc
    return prim, t, b1, b2, counters


def attach(mod):
    """Make the twin's functions attributes of the product's ctypes module (tests only)."""
    global Stats
    Stats = mod.Stats
    for name, obj in list(globals().items()):
        if name.startswith("twin") and callable(obj):
            setattr(mod, name, obj)
    mod.TWIN_LIB = TWIN_LIB
    return mod
