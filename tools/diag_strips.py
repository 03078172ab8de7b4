#!/usr/bin/env python3
"""Diagnostic: full-frame C3 render vs the same frame rendered as row-strip windows (must be bit-identical)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi, row_strips
import tempfile
ffi = load_ffi()
td = tempfile.mkdtemp()
res = int(os.environ.get("DIAG_RES", "1024")); spp = int(os.environ.get("DIAG_SPP", "8"))
sp = gen_scene(os.path.join(td, "c3.pbrt"), "c3", "--res", res, res, "--spp", spp, spp, "--filename", os.path.join(td, "c3.exr"))
sc = ffi.Scene(sp)
lib = ffi.cuda()
ds = ffi.DeviceScene(sc)
ds.render(); full = ds.download(0, want_raw=True)
ds.render(); full2 = ds.download(0, want_raw=True)
print("full frame deterministic:", np.array_equal(full[0].view(np.uint32), full2[0].view(np.uint32)))
for wave in (-1, 1 << 22, 1 << 26):
    if wave > 0:
        lib.pbrt_b200_tunable(0, wave)
    for n in (2, 16):
        strips = row_strips((0, 0, res, res), n)
        ds.render([ffi.Window(*w) for w in strips])
        got = [ds.download(i, want_raw=True) for i in range(n)]
        rgba = np.concatenate([g[0] for g in got], axis=0); raw = np.concatenate([g[1] for g in got], axis=0)
        d = (rgba.view(np.uint32) != full[0].view(np.uint32)).any(axis=-1)
        dr = (raw.view(np.uint32) != full[1].view(np.uint32)).any(axis=-1)
        rows = np.nonzero(d.any(axis=1))[0]
        print("wave %d strips %d: rgba pixels differing %d, raw differing %d, rows %s, max abs diff %.3g, weight diff px %d" % (
            wave, n, int(d.sum()), int(dr.sum()), rows[:12].tolist(), float(np.abs(rgba - full[0]).max()),
            int((raw[..., 4] != full[1][..., 4]).sum())))
        if d.any():
            ys, xs = np.nonzero(d)
            print("   first differing pixels:", list(zip(ys[:6].tolist(), xs[:6].tolist())), "cols range", xs.min(), xs.max())
