"""Shared helpers for the parity tests: scene generation, reference runs, error statistics."""
import importlib.util
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_ffi():
    spec = importlib.util.spec_from_file_location("pbrt_b200_pyffi", os.path.join(ROOT, "pbrt-v1_b200", "pyffi.py"))
    mod = sys.modules.get("pbrt_b200_pyffi")
    if mod is None:
        mod = importlib.util.module_from_spec(spec)
        sys.modules["pbrt_b200_pyffi"] = mod
        spec.loader.exec_module(mod)
    if not hasattr(mod, "twin"):
        # the CPU twin's bindings are test infrastructure and live here, not in the product package
        tspec = importlib.util.spec_from_file_location("pbrt_twin_ffi", os.path.join(ROOT, "tests", "twin_ffi.py"))
        tmod = importlib.util.module_from_spec(tspec)
        sys.modules["pbrt_twin_ffi"] = tmod
        tspec.loader.exec_module(tmod)
        tmod.attach(mod)
    return mod


def gen_scene(out_path, name, *args):
    """Runs scenes/gen_scenes.py; `args` are its command-line flags."""
    cmd = [sys.executable, os.path.join(ROOT, "scenes", "gen_scenes.py"), name, "-o", out_path] + [str(a) for a in args]
    subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL)
    return out_path


def read_sample_dump(path):
    """oracle/ref_glue.cpp PBRT_REF_DUMP_SAMPLES format: imageX, imageY, L.rgb, alpha per sample."""
    return np.fromfile(path, np.float32).reshape(-1, 6)


def image_error_stats(a, b):
    """Per-pixel L-infinity over RGBA between two float images plus the tolerance fractions."""
    d = np.abs(a.astype(np.float64) - b.astype(np.float64)).max(axis=-1)
    ref = np.abs(b.astype(np.float64)).max(axis=-1)
    tol = np.maximum(ref * 2.0 ** -10, 1e-3)
    return {
        "linf": float(d.max()),
        "mean_abs": float(d.mean()),
        "frac_identical": float((d == 0).mean()),
        "frac_within_tol": float((d <= tol).mean()),
        "frac_gt_1e-2": float((d > 1e-2).mean()),
        "mean_rel": float(abs(a[..., :3].mean() - b[..., :3].mean()) / max(abs(b[..., :3].mean()), 1e-30)),
    }


def write_exr(path, rgb):
    """Uncompressed scan-line OpenEXR file, HALF channels B, G, R, from a float array [h, w, 3] (texture
    fixtures; what oracle/ref_glue.cpp's and the host front end's ReadImage both accept)."""
    import struct
    rgb = np.asarray(rgb, np.float32)
    h, w, _ = rgb.shape
    half = rgb.astype(np.float16)

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data
    chlist = b"".join(n + b"\0" + struct.pack("<iBBBBii", 1, 0, 0, 0, 0, 1, 1) for n in (b"B", b"G", b"R")) + b"\0"
    hdr = struct.pack("<II", 20000630, 2)
    hdr += attr("channels", "chlist", chlist)
    hdr += attr("compression", "compression", b"\0")
    hdr += attr("dataWindow", "box2i", struct.pack("<4i", 0, 0, w - 1, h - 1))
    hdr += attr("displayWindow", "box2i", struct.pack("<4i", 0, 0, w - 1, h - 1))
    hdr += attr("lineOrder", "lineOrder", b"\0")
    hdr += attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    hdr += attr("screenWindowCenter", "v2f", struct.pack("<2f", 0.0, 0.0))
    hdr += attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    hdr += b"\0"
    row_bytes = 8 + 3 * 2 * w
    base = len(hdr) + 8 * h
    out = [hdr, b"".join(struct.pack("<Q", base + y * row_bytes) for y in range(h))]
    for y in range(h):
        out.append(struct.pack("<ii", y, 3 * 2 * w))
        for c in (2, 1, 0):
            out.append(half[y, :, c].tobytes())
    with open(path, "wb") as f:
        f.write(b"".join(out))


# ---------------------------------------------------------------------------------------------------
# Running the compiled reference (oracle/_ref) at the BASELINE configurations' real sizes.
# The reference is single-threaded; its own way to use several cores is independent `cropwindow`
# renders pasted by data-window offset (film/image.cpp:79-85,148-156, tools/exrassemble.cpp:53-66).
def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def split_scene_text(text):
    """(header, world) of a generated scene: everything before `WorldBegin`, and the rest."""
    i = text.index("WorldBegin")
    return text[:i], text[i:]


def header_with_crop(header, xres, yres, win):
    """The header with the Film's cropwindow set so that the film covers exactly the pixel window
    win = (x0, y0, w, h): ImageFilm takes Ceil2Int(res * crop) (film/image.cpp:79-85), so the bounds are
    placed half a pixel below the wanted integer."""
    x0, y0, w, h = win
    crop = ((x0 - 0.5) / xres, (x0 + w - 0.5) / xres, (y0 - 0.5) / yres, (y0 + h - 0.5) / yres)
    out = []
    done = False
    for ln in header.splitlines():
        if ln.startswith("Film "):
            j = ln.find(' "float cropwindow"')
            if j >= 0:
                ln = ln[:j]
            ln += ' "float cropwindow" [%s]' % " ".join("%.9g" % c for c in crop)
            done = True
        out.append(ln)
    assert done, "scene header has no Film line"
    return "\n".join(out) + "\n"


def row_strips(win, n):
    """Splits the pixel window (x0, y0, w, h) into at most n horizontal strips."""
    x0, y0, w, h = win
    n = max(1, min(n, h))
    b = [y0 + (i * h) // n for i in range(n + 1)]
    return [(x0, b[i], w, b[i + 1] - b[i]) for i in range(n)]


def run_reference_windows(ffi, scene_path, xres, yres, windows, workdir, keyed=True, dump_samples=False,
                          dump_kd=None, tag="ref"):
    """Renders each pixel window of `scene_path` with its own process of the compiled reference, all
    concurrently (the scene's world block is shared, each process gets a header with its crop window).
    Returns a list of dicts per window: rgba [h, w, 4] float32 (what WriteRGBAImage received), samples
    [n, 6] or None, intersect / intersect_p call counts, loop seconds, kd (nodes, leaves, refs, prims),
    the reference's own statistics text; plus the wall time of the whole batch."""
    import time
    exe = os.path.join(ffi.REF_DIR, "pbrt_ref_keyed" if keyed else "pbrt_ref_exact")
    env0 = dict(os.environ)
    env0["PBRT_SEARCHPATH"] = os.path.join(ffi.REF_DIR, "plugins_keyed" if keyed else "plugins")
    header, world = split_scene_text(open(scene_path).read())
    world_path = os.path.join(workdir, tag + "_world.pbrt")
    with open(world_path, "w") as f:
        f.write(world)
    procs = []
    t0 = time.perf_counter()
    for i, win in enumerate(windows):
        exr = os.path.join(workdir, "%s_%d.exr" % (tag, i))
        hp = os.path.join(workdir, "%s_%d_head.pbrt" % (tag, i))
        hd = header_with_crop(header, xres, yres, win)
        lines = []
        for ln in hd.splitlines():
            if ln.startswith("Film "):
                j = ln.find('"string filename" ["')
                k = ln.find('"]', j)
                ln = ln[:j] + '"string filename" ["%s"]' % exr + ln[k + 2:]
            lines.append(ln)
        with open(hp, "w") as f:
            f.write("\n".join(lines) + "\n")
        env = dict(env0)
        dump = None
        if dump_samples:
            dump = os.path.join(workdir, "%s_%d.samples" % (tag, i))
            env["PBRT_REF_DUMP_SAMPLES"] = dump
        if dump_kd and i == 0:
            if os.path.exists(dump_kd):
                os.remove(dump_kd)
            env["PBRT_REF_DUMP_KD"] = dump_kd
        p = subprocess.Popen([exe, hp, world_path], env=env, cwd=workdir, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT)
        procs.append((p, exr, dump, win))
    results = []
    for p, exr, dump, win in procs:
        out = p.communicate()[0].decode(errors="replace")
        if p.returncode != 0:
            raise RuntimeError("reference failed on window %s:\n%s" % (win, out[-2000:]))
        r = {"window": win, "intersect": 0, "intersect_p": 0, "seconds": None, "kd": None, "samples": None}
        for ln in out.replace("\r", "\n").splitlines():
            if "PBRT_REF_RAYS" in ln:
                f = ln.split()
                r["intersect"] = int(f[f.index("INTERSECT") + 1]); r["intersect_p"] = int(f[f.index("INTERSECTP") + 1])
            elif "PBRT_REF_RENDER_SECONDS" in ln:
                r["seconds"] = float(ln.split("PBRT_REF_RENDER_SECONDS")[1].split()[0])
            elif "PBRT_REF_KD" in ln and r["kd"] is None:
                f = ln.split()
                r["kd"] = tuple(int(f[f.index(k) + 1]) for k in ("NODES", "LEAVES", "REFS", "PRIMS"))
        i = out.find("Statistics:")
        r["statistics"] = out[i:] if i >= 0 else ""
        rgba, hdr = ffi.read_raw(exr + ".raw")
        assert (hdr[4], hdr[5], hdr[0], hdr[1]) == tuple(win), (hdr, win)
        r["rgba"] = rgba
        if dump:
            r["samples"] = read_sample_dump(dump)
            os.remove(dump)
        os.remove(exr + ".raw")
        results.append(r)
    wall = time.perf_counter() - t0
    return results, wall


def reference_statistic(text, name):
    """A counter of the reference's post-render statistics (core/util.cpp:228-276): it prints %.0f below
    1e4, then 'k' / 'M' / 'B' with 1 or 3 decimals.  Returns (value, tolerance of the printed form)."""
    for ln in text.splitlines():
        if name in ln:
            tok = ln.split()[-1] if "(" not in ln else ln.split("(")[0].split()[-1]
            tok = tok.split(":")[0]
            mult, tol = 1.0, 0.5
            if tok[-1] in "kMB":
                mult = {"k": 1e3, "M": 1e6, "B": 1e9}[tok[-1]]
                tol = mult * (0.05 if tok[-1] == "k" else 0.0005) * 1.01
                tok = tok[:-1]
            return float(tok) * mult, tol
    return None, None


def read_kd_dump(path):
    """PBRT_REF_DUMP_KD file of the first KdTreeAccel built: (bounds[6], node words [n, 2] uint32, refs uint32)."""
    d = np.fromfile(path, np.uint32)
    assert d[0] == 0x4b445452
    nn, nr = int(d[1]), int(d[2])
    return d[4:10].view(np.float32).copy(), d[10:10 + 2 * nn].reshape(nn, 2).copy(), d[10 + 2 * nn:10 + 2 * nn + nr].copy(), int(d[3])


def host_kd_arrays(scene):
    """The kd-tree arrays of the flattened scene the host front end built (head of PbrtB200Scene)."""
    import ctypes as C

    class Head(C.Structure):
        _fields_ = [("abi", C.c_uint32), ("nodes", C.c_void_p), ("n_nodes", C.c_uint32), ("n_top", C.c_uint32),
                    ("refs", C.c_void_p), ("n_refs", C.c_uint32), ("bounds", C.c_float * 6)]
    h = Head.from_address(scene.ptr)
    nodes = np.ctypeslib.as_array(C.cast(h.nodes, C.POINTER(C.c_uint32)), shape=(h.n_nodes, 2)).copy()
    refs = (np.ctypeslib.as_array(C.cast(h.refs, C.POINTER(C.c_uint32)), shape=(h.n_refs,)).copy()
            if h.n_refs else np.zeros(0, np.uint32))
    return np.array(list(h.bounds), np.float32), nodes, refs, int(h.n_top)


def error_histogram(img, ref, light_max=None, spp=None):
    """The per-pixel error summary SURVEY.md 8(c) asks for, over RGBA of the fp32 film."""
    d = np.abs(img.astype(np.float64) - ref.astype(np.float64)).max(axis=-1)
    mag = np.abs(ref.astype(np.float64)).max(axis=-1)
    tol = np.maximum(mag * 2.0 ** -10, 1e-3)
    out = {
        "pixels": int(d.size),
        "linf": float(d.max()),
        "p99": float(np.percentile(d, 99)), "p99_9": float(np.percentile(d, 99.9)),
        "frac_identical": float((d == 0).mean()),
        "frac_within_tol": float((d <= tol).mean()),
        "frac_gt_1e-3": float((d > 1e-3).mean()), "frac_gt_1e-2": float((d > 1e-2).mean()),
        "mean_abs": float(d.mean()),
        "mean_rel": float(abs(img[..., :3].mean(dtype=np.float64) - ref[..., :3].mean(dtype=np.float64)) /
                          max(abs(ref[..., :3].mean(dtype=np.float64)), 1e-30)),
    }
    if light_max is not None and spp:
        out["outlier_bound"] = 4.0 * light_max / spp
        out["outliers_over_bound"] = int((d > out["outlier_bound"]).sum())
    return out
