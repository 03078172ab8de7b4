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
