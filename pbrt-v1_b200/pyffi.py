"""ctypes bindings used by tests/, bench.py and __graft_entry__.py.

Two shared libraries, both plain C ABIs (the bindings of the test-only CPU twin, oracle/_build/libpbrt_twin.so,
live in tests/twin_ffi.py; tests/parity_util.load_ffi() attaches them to this module for the tests' convenience):

  host/libpbrt_host.so    front end: .pbrt parser + api.h state machine + kd build;
                          `load_scene()` parses a scene file and returns the flattened
                          PbrtB200Scene (opaque pointer) that Scene::Render() would hand
                          to the GPU library.
  csrc/libpbrt_b200.so    the product: CUDA render hot path (include/pbrt_b200.h).

Nothing here computes anything; it only marshals numpy arrays.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_LIB = os.path.join(ROOT, "pbrt-v1_b200", "host", "libpbrt_host.so")
CUDA_LIB = os.environ.get("PBRT_B200_LIB", os.path.join(ROOT, "pbrt-v1_b200", "csrc", "libpbrt_b200.so"))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class Window(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "camera_rays", "intersect_rays", "shadow_rays", "points_shaded", "interior_visits",
        "leaf_visits", "prim_refs_read", "prim_tests", "nan_samples", "negative_samples",
        "infinite_samples")] + [("render_ms", C.c_double), ("upload_ms", C.c_double),
                                ("download_ms", C.c_double), ("kernel_launches", C.c_uint64),
                                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class KernelStats(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("launches", C.c_uint64), ("ms", C.c_double), ("rays", C.c_uint64),
                ("interior_visits", C.c_uint64), ("leaf_visits", C.c_uint64), ("prim_refs_read", C.c_uint64),
                ("prim_tests", C.c_uint64)]


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


_host = None


def host():
    global _host
    if _host is None:
        lib = C.CDLL(HOST_LIB)
        lib.pbrt_host_load_scene.restype = C.c_void_p
        lib.pbrt_host_load_scene.argtypes = [C.c_char_p]
        lib.pbrt_host_scene.restype = C.c_void_p
        lib.pbrt_host_scene.argtypes = [C.c_void_p]
        lib.pbrt_host_free_scene.argtypes = [C.c_void_p]
        lib.pbrt_host_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        lib.pbrt_host_render_file.argtypes = [C.c_char_p]
        lib.pbrt_host_write_exr.argtypes = [C.c_char_p, C.c_void_p] + [C.c_int] * 6
        _host = lib
    return _host


class Scene:
    """A parsed + flattened scene (owns the host-side arrays)."""

    def __init__(self, path):
        self.handle = host().pbrt_host_load_scene(os.fsencode(path))
        if not self.handle:
            raise RuntimeError("failed to load scene %s" % path)
        self.ptr = host().pbrt_host_scene(self.handle)
        info = np.zeros(16, np.int64)
        host().pbrt_host_scene_info(self.handle, _fp(info))
        (self.xres, self.yres, self.x0, self.y0, self.width, self.height, self.spp, self.n_nodes,
         self.n_prim_refs, self.n_prims, self.n_lights, self.n_materials, self.integrator) = [int(v) for v in info[:13]]

    def full_window(self):
        return Window(self.x0, self.y0, self.width, self.height)

    def close(self):
        if self.handle:
            host().pbrt_host_free_scene(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------- CUDA library (the product)
_cuda = None


def cuda():
    """Load libpbrt_b200.so.  Raises if it is missing: there is no fallback."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB):
            raise RuntimeError("CUDA library %s not built (run __graft_entry__.build())" % CUDA_LIB)
        lib = C.CDLL(CUDA_LIB)
        lib.pbrt_b200_last_error.restype = C.c_char_p
        lib.pbrt_b200_scene_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.pbrt_b200_scene_destroy.argtypes = [C.c_void_p]
        lib.pbrt_b200_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.pbrt_b200_film_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.pbrt_b200_film_device_ptr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.pbrt_b200_render_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        lib.pbrt_b200_tunable.restype = C.c_int64
        lib.pbrt_b200_tunable.argtypes = [C.c_int, C.c_int64]
        lib.pbrt_b200_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 4
        lib.pbrt_b200_intersect_p.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.pbrt_b200_camera_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        lib.pbrt_b200_sample_radiance.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
        lib.pbrt_b200_kernel_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.pbrt_b200_render_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.pbrt_b200_comm_unique_id.argtypes = [C.c_void_p]
        lib.pbrt_b200_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.pbrt_b200_comm_destroy.argtypes = [C.c_void_p]
        lib.pbrt_b200_render_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.pbrt_b200_frame_device_ptr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pbrt_b200_render_host_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _cuda = lib
    return _cuda


def _check(rc):
    if rc != 0:
        raise RuntimeError("pbrt_b200 error %d: %s" % (rc, cuda().pbrt_b200_last_error().decode()))


class DeviceScene:
    """Device-resident scene (pbrt_b200_scene_create / _destroy)."""

    def __init__(self, scene, device=0):
        self.scene = scene
        h = C.c_void_p()
        _check(cuda().pbrt_b200_scene_create(scene.ptr, device, C.byref(h)))
        self.h = h

    def render(self, windows=None):
        ws = windows or [self.scene.full_window()]
        arr = (Window * len(ws))(*ws)
        st = Stats()
        _check(cuda().pbrt_b200_render(self.h, arr, len(ws), C.byref(st)))
        self.windows = ws
        return st.as_dict()

    def render_tiles(self, tiles):
        """pbrt_b200_render_tiles: the windows are tiles of the scene's film window (bit-identical to the frame)."""
        arr = (Window * len(tiles))(*tiles)
        st = Stats()
        _check(cuda().pbrt_b200_render_tiles(self.h, arr, len(tiles), C.byref(st)))
        self.windows = list(tiles)
        return st.as_dict()

    def render_frame(self, comm=None, tile_edge=0, out=None):
        """pbrt_b200_render_frame: this rank's share of the whole film window + the one exchange; rank 0 may pass a
        float32 [h, w, 4] array (e.g. pinned) to receive the frame."""
        st = Stats()
        _check(cuda().pbrt_b200_render_frame(self.h, comm.c if comm else None, tile_edge,
                                             out.ctypes.data if out is not None else None, C.byref(st)))
        return st.as_dict()

    def frame_device_ptr(self):
        p = C.c_void_p()
        n = C.c_uint64()
        _check(cuda().pbrt_b200_frame_device_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def download(self, w=0, want_raw=False):
        win = self.windows[w]
        rgba = np.zeros((win.height, win.width, 4), np.float32)
        raw = np.zeros((win.height, win.width, 5), np.float32) if want_raw else None
        _check(cuda().pbrt_b200_film_download(self.h, w, _fp(rgba), _fp(raw) if want_raw else None))
        return (rgba, raw) if want_raw else rgba

    def kernel_stats(self):
        arr = (KernelStats * 16)()
        n = C.c_int()
        _check(cuda().pbrt_b200_kernel_stats(self.h, arr, 16, C.byref(n)))
        out = {}
        for i in range(n.value):
            k = arr[i]
            out[k.name.decode()] = {f: getattr(k, f) for f, _ in KernelStats._fields_ if f != "name"}
        return out

    def film_device_ptr(self, w=0):
        p = C.c_void_p()
        n = C.c_uint64()
        _check(cuda().pbrt_b200_film_device_ptr(self.h, w, C.byref(p), C.byref(n)))
        return p.value, n.value

    def sample_radiance(self, w, n):
        """First n per-sample records (L.rgb, alpha) in reference order.  With the bestcandidate
        sampler the library keeps one record per table entry; skipped entries (alpha == -1) are dropped."""
        lib = cuda()
        lib.pbrt_b200_sample_slots.restype = C.c_uint64
        lib.pbrt_b200_sample_slots.argtypes = [C.c_void_p, C.c_int]
        slots = int(lib.pbrt_b200_sample_slots(self.h, w))
        if slots <= n:
            out = np.zeros((n, 4), np.float32)
            _check(lib.pbrt_b200_sample_radiance(self.h, w, n, _fp(out)))
            return out
        out = np.zeros((slots, 4), np.float32)
        _check(lib.pbrt_b200_sample_radiance(self.h, w, slots, _fp(out)))
        if (out[:, 3] == -1).any():
            out = out[out[:, 3] != -1]
        return np.ascontiguousarray(out[:n])

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        b1 = np.zeros(n, np.float32)
        b2 = np.zeros(n, np.float32)
        _check(cuda().pbrt_b200_intersect(self.h, _fp(rays), n, _fp(prim), _fp(t), _fp(b1), _fp(b2)))
        return prim, t, b1, b2

    def intersect_p(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        occ = np.zeros(n, np.uint8)
        _check(cuda().pbrt_b200_intersect_p(self.h, _fp(rays), n, _fp(occ)))
        return occ

    def camera_rays(self, n, window=None):
        """The first n camera rays of a window (pbrt_b200_camera_rays, a test entry point); the caller knows the
        window's sample count."""
        w = window or self.scene.full_window()
        rays = np.zeros((n, 8), np.float32)
        xy = np.zeros((n, 2), np.float32)
        _check(cuda().pbrt_b200_camera_rays(self.h, C.byref(w), n, _fp(rays), _fp(xy)))
        return rays, xy

    def close(self):
        if self.h:
            cuda().pbrt_b200_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """pbrt_b200_comm_*: this rank among `world` ranks (one GPU each).  `unique_id` = the 128 bytes rank 0 got from
    Comm.unique_id(), broadcast by the caller (torch.distributed, MPI, a file ...)."""

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        _check(cuda().pbrt_b200_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, unique_id, rank, world, device):
        c = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(cuda().pbrt_b200_comm_create(buf, rank, world, device, C.byref(c)))
        self.c = c
        self.rank, self.world = rank, world

    def close(self):
        if self.c:
            cuda().pbrt_b200_comm_destroy(self.c)
            self.c = None


def frame_partition(window, world, tile_edge=0):
    """pbrt_b200_frame_partition: [(Window, owner rank)] of the multi-GPU partition of `window` (no GPU needed)."""
    lib = cuda()
    lib.pbrt_b200_frame_partition.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    n = lib.pbrt_b200_frame_partition(C.byref(window), world, tile_edge, None, None, 0)
    tiles = (Window * n)()
    owners = (C.c_int32 * n)()
    lib.pbrt_b200_frame_partition(C.byref(window), world, tile_edge, tiles, owners, n)
    return [(Window(t.x0, t.y0, t.width, t.height), int(o)) for t, o in zip(tiles, owners)]


def render_host_multi(scene, devices, out=None):
    """pbrt_b200_render_host_multi: HOST buffers both sides, one thread per device inside the library."""
    w = scene.full_window()
    if out is None:
        out = np.zeros((w.height, w.width, 4), np.float32)
    dev = (C.c_int * len(devices))(*devices)
    st = Stats()
    _check(cuda().pbrt_b200_render_host_multi(scene.ptr, dev, len(devices), out.ctypes.data, C.byref(st)))
    return out, st.as_dict()


def render_host(scene, windows=None, device=0, want_raw=False, outs=None):
    """pbrt_b200_render_host: HOST buffers in, HOST buffers out (the call Scene::Render makes).
    `outs`: optional preallocated float32 [h, w, 4] arrays (e.g. pinned memory) to receive the frames."""
    ws = windows or [scene.full_window()]
    arr = (Window * len(ws))(*ws)
    if outs is None:
        outs = [np.zeros((w.height, w.width, 4), np.float32) for w in ws]
    assert len(outs) == len(ws) and all(o.dtype == np.float32 and o.flags.c_contiguous for o in outs)
    raws = [np.zeros((w.height, w.width, 5), np.float32) for w in ws] if want_raw else None
    optr = (C.c_void_p * len(ws))(*[o.ctypes.data for o in outs])
    rptr = (C.c_void_p * len(ws))(*[r.ctypes.data for r in raws]) if want_raw else None
    st = Stats()
    _check(cuda().pbrt_b200_render_host(scene.ptr, device, arr, len(ws), optr, rptr, C.byref(st)))
    return outs, raws, st.as_dict()


# ----------------------------------------------------------------- compiled reference (oracle/_ref)
def read_raw(path):
    """Reads the raw float dump oracle/ref_glue.cpp writes in place of an EXR."""
    with open(path, "rb") as f:
        assert f.read(8) == b"PBRTRAW1"
        hdr = np.frombuffer(f.read(24), np.int32)
        xres, yres = int(hdr[0]), int(hdr[1])
        px = np.frombuffer(f.read(), np.float32).reshape(yres, xres, 4)
    return px.copy(), [int(v) for v in hdr]


def have_reference():
    return os.path.exists(os.path.join(REF_DIR, "pbrt_ref_keyed"))


def run_reference(scene_path, keyed=True, dump_samples=None, cwd=None):
    """Runs the compiled reference on a scene file; returns (stdout+stderr, seconds of the sample loop)."""
    import subprocess
    exe = os.path.join(REF_DIR, "pbrt_ref_keyed" if keyed else "pbrt_ref_exact")
    env = dict(os.environ)
    env["PBRT_SEARCHPATH"] = os.path.join(REF_DIR, "plugins_keyed" if keyed else "plugins")
    if dump_samples:
        env["PBRT_REF_DUMP_SAMPLES"] = dump_samples
    p = subprocess.run([exe, scene_path], env=env, cwd=cwd, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, check=True)
    out = p.stdout.decode(errors="replace")
    secs = None
    for line in out.splitlines():
        if "PBRT_REF_RENDER_SECONDS" in line:
            secs = float(line.split("PBRT_REF_RENDER_SECONDS")[1].split()[0])
    return out, secs
