#!/usr/bin/env python3
"""Trace-driven model of the warp trip loop of pb_kd_warp.cuh (CPU only).

Every ray's step string (I interior, L leaf reached, R mailboxed reference, T reference fetched and tested, P pop)
comes from the oracle twin (pbrt_twin_trace); warps of 32 lanes with lane refill replay them under a scheduling
policy, and the model adds up warp instructions with the per-step costs read off the SASS of the shipped kernel
(profiles/r02e_*: interior step 58, leaf-list step 33, test 97, pop 30, loop top 30).  It knows nothing about memory
latency: it ranks policies by issued instructions only.

usage: simt_sim2.py [c3|c4|c5] [gen args]      (SIMT_RES / SIMT_SPP / SIMT_ANY=1 for shadow-like any-hit rays)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
RES = os.environ.get("SIMT_RES", "64")
SPP = os.environ.get("SIMT_SPP", "2")
ANY = int(os.environ.get("SIMT_ANY", "0"))
cache = "/tmp/simt2_%s_%s_%s_%d.npz" % (name, RES, SPP, ANY)
if os.path.exists(cache):
    z = np.load(cache)
    codes, offs = z["codes"], z["offs"]
else:
    ffi = load_ffi()
    sp = gen_scene("/tmp/simt2_%s.pbrt" % name, name, "--res", RES, RES, "--spp", SPP, SPP, "--filename", "/tmp/simt.exr", *sys.argv[2:])
    sc = ffi.Scene(sp)
    rays, _ = ffi.twin_camera_rays(sc)
    prim, t, _, _ = ffi.twin_intersect(sc, rays)
    ok = prim >= 0
    p = rays[ok, :3] + rays[ok, 3:6] * t[ok, None]
    rng = np.random.default_rng(1)
    target = rng.uniform(5, 550, p.shape).astype(np.float32)
    d = target - p
    dist = np.linalg.norm(d, axis=1, keepdims=True)
    d /= dist
    sec = np.zeros((p.shape[0], 8), np.float32)
    sec[:, :3] = p; sec[:, 3:6] = d; sec[:, 6] = 1e-3
    sec[:, 7] = dist[:, 0] * (1 - 1e-3) if ANY else np.inf
    n = sec.shape[0]
    lib = ffi.twin()
    lib.pbrt_twin_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    cap = n * 600
    codes = np.zeros(cap, np.uint8)
    offs = np.zeros(n + 1, np.uint64)
    assert lib.pbrt_twin_trace(sc.ptr, sec.ctypes.data, n, ANY, codes.ctypes.data, cap, offs.ctypes.data) == 0
    codes = codes[:int(offs[n])]
    np.savez(cache, codes=codes, offs=offs)
n = len(offs) - 1
traces = [codes[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)]
traces = [t for t in traces if t]
n = len(traces)
tot = {c: sum(t.count(c.encode()) for t in traces) for c in "ILRTP"}
print("%s: %d rays; per ray: %s" % (name, n, "  ".join("%s %.2f" % (k, v / n) for k, v in tot.items())))
empty = sum(t.count(b"LP") for t in traces)
print("empty leaves per ray %.2f of %.2f" % (empty / n, tot["L"] / n))

I_, L_, R_, T_, P_ = (ord(c) for c in "ILRTP")


def simulate(DS=2, LS=2, inline_empty_pop=False, adaptive=0, cI=58, cR=33, cT=97, cP=30, cTop=30, cRefill=110,
             vote=8, refill_thr=16, merge_r=False, max_rays=None):
    """returns (warp instructions per ray, trips per ray, mean lanes in I / R / T steps)"""
    total = 0
    trips = 0
    lanesI = stepsI = lanesR = stepsR = lanesT = stepsT = 0
    nq = n if max_rays is None else min(n, max_rays)
    nwarps = 8   # rays are dealt to warps from one queue, like the persistent grid's atomic counter
    nxt = 0
    # per-warp state
    W = [dict(tr=[None] * 32, pos=[0] * 32, leaf=[False] * 32, pend=[False] * 32, done=False) for _ in range(nwarps)]
    live = nwarps
    while live:
        for w in W:
            if w["done"]:
                continue
            tr, pos, leaf, pend = w["tr"], w["pos"], w["leaf"], w["pend"]
            idle = [l for l in range(32) if tr[l] is None]
            total += cTop
            if nxt < nq and (len(idle) >= refill_thr or len(idle) == 32):
                total += cRefill
                for l in idle:
                    if nxt < nq:
                        tr[l] = traces[nxt]; pos[l] = 0; leaf[l] = False; pend[l] = False
                        nxt += 1
            act = [l for l in range(32) if tr[l] is not None]
            if not act:
                if nxt >= nq:
                    w["done"] = True
                    live -= 1
                continue
            trips += 1
            # I phase
            step = 0
            while True:
                desc = [l for l in act if not leaf[l]]
                if not desc:
                    total += vote
                    break
                if adaptive:
                    if step >= 1 and len(desc) < adaptive or step >= 12:
                        total += vote
                        break
                elif step >= DS:
                    break
                total += cI
                stepsI += 1; lanesI += len(desc)
                for l in desc:
                    c = tr[l][pos[l]]
                    pos[l] += 1
                    if c == L_:
                        leaf[l] = True
                        t = tr[l]
                        if inline_empty_pop and pos[l] + 1 < len(t) and t[pos[l]] == P_:
                            pos[l] += 1
                            leaf[l] = False
                        elif merge_r and pos[l] < len(t):
                            c2 = t[pos[l]]
                            if c2 == R_:
                                pos[l] += 1
                            elif c2 == T_:
                                pend[l] = True
                step += 1
            # R phase
            for step in range(LS):
                walk = [l for l in act if leaf[l] and not pend[l] and pos[l] < len(tr[l]) and tr[l][pos[l]] in (R_, T_)]
                if not walk:
                    total += vote
                    break
                total += cR
                stepsR += 1; lanesR += len(walk)
                for l in walk:
                    if tr[l][pos[l]] == R_:
                        pos[l] += 1
                    else:
                        pend[l] = True
            # T phase
            pl = [l for l in act if pend[l]]
            if pl:
                total += cT + (12 if merge_r else 0)
                stepsT += 1; lanesT += len(pl)
                for l in pl:
                    pos[l] += 1
                    pend[l] = False
                    if merge_r and pos[l] < len(tr[l]):
                        c2 = tr[l][pos[l]]
                        if c2 == R_:
                            pos[l] += 1
                        elif c2 == T_:
                            pend[l] = True
            else:
                total += 4
            # pop phase
            popped = False
            for l in act:
                if leaf[l] and not pend[l]:
                    t = tr[l]
                    if pos[l] >= len(t):
                        tr[l] = None; popped = True
                    elif t[pos[l]] == P_:
                        pos[l] += 1
                        leaf[l] = False
                        popped = True
                        if pos[l] >= len(t):
                            tr[l] = None
            total += cP if popped else 8
    return total / nq, trips * 32.0 / nq / 24, lanesI / max(stepsI, 1), lanesR / max(stepsR, 1), lanesT / max(stepsT, 1)


MAXR = int(os.environ.get("SIMT_RAYS", "12000"))
base = None
for label, kw in [
    ("shipped r01: DS=2 LS=2", dict()),
    ("DS=3", dict(DS=3)),
    ("DS=4", dict(DS=4)),
    ("DS=6", dict(DS=6)),
    ("DS=4 + inline empty-leaf pop", dict(DS=4, inline_empty_pop=True)),
    ("DS=6 + inline empty-leaf pop", dict(DS=6, inline_empty_pop=True)),
    ("adaptive >=8 lanes", dict(adaptive=8)),
    ("adaptive >=8 + inline pop", dict(adaptive=8, inline_empty_pop=True)),
    ("adaptive >=12 + inline pop", dict(adaptive=12, inline_empty_pop=True)),
    ("adaptive >=6 + inline pop", dict(adaptive=6, inline_empty_pop=True)),
    ("DS=4 + inline pop + merged R", dict(DS=4, inline_empty_pop=True, merge_r=True, cI=64)),
    ("DS=4 + inline pop + merged R, LS=1", dict(DS=4, LS=1, inline_empty_pop=True, merge_r=True, cI=64)),
    ("adaptive >=8 + inline pop + merged R, LS=1", dict(adaptive=8, LS=1, inline_empty_pop=True, merge_r=True, cI=64)),
    ("DS=4 + inline pop, lean steps (cI 38, cR 26, cT 90, cP 24, top 24)", dict(DS=4, inline_empty_pop=True, cI=38, cR=26, cT=90, cP=24, cTop=24)),
    ("adaptive 8 + inline pop + merged R LS=1, lean", dict(adaptive=8, LS=1, inline_empty_pop=True, merge_r=True, cI=44, cR=26, cT=90, cP=24, cTop=24)),
]:
    r = simulate(max_rays=MAXR, **kw)
    if base is None:
        base = r[0]
    print("%-70s %7.1f warp-instr/ray (%.3f)  lanes I %.1f R %.1f T %.1f" % (label, r[0], r[0] / base, r[2], r[3], r[4]), flush=True)
