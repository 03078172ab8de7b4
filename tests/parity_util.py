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
