#!/usr/bin/env python3
"""Offline SIMT-efficiency model of kd-walk scheduling strategies on secondary rays (CPU only).
Builds bounce-1-like rays (camera-ray hit points, direction towards a random point of the room),
traces each through the oracle twin recording its step string, then replays warps of 32."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import gen_scene, load_ffi  # noqa: E402

ffi = load_ffi()
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
RES = os.environ.get("SIMT_RES", "96")
SPP = os.environ.get("SIMT_SPP", "2")
sp = gen_scene("/tmp/simt_%s.pbrt" % name, name, "--res", RES, RES, "--spp", SPP, SPP, "--filename", "/tmp/simt.exr", *sys.argv[2:])
sc = ffi.Scene(sp)
rays, _ = ffi.twin_camera_rays(sc)
prim, t, _, _ = ffi.twin_intersect(sc, rays)
ok = prim >= 0
p = rays[ok, :3] + rays[ok, 3:6] * t[ok, None]
rng = np.random.default_rng(1)
target = rng.uniform(5, 550, p.shape).astype(np.float32)
d = target - p
d /= np.linalg.norm(d, axis=1, keepdims=True)
sec = np.zeros((p.shape[0], 8), np.float32)
sec[:, :3] = p; sec[:, 3:6] = d; sec[:, 6] = 1e-3; sec[:, 7] = np.inf
n = sec.shape[0]
lib = ffi.twin()
lib.pbrt_twin_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
cap = n * 400
codes = np.zeros(cap, np.uint8)
offs = np.zeros(n + 1, np.uint64)
assert lib.pbrt_twin_trace(sc.ptr, sec.ctypes.data, n, 0, codes.ctypes.data, cap, offs.ctypes.data) == 0
traces = [codes[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)]
COST = {ord('I'): 40, ord('L'): 12, ord('R'): 25, ord('T'): 25 + 90, ord('P'): 25}
useful = sum(sum(COST[c] for c in tr) for tr in traces)
print("rays %d  mean steps %.1f  useful thread-instr/ray %.0f" % (n, np.mean([len(t) for t in traces]), useful / n))


def sim_steploop(order):
    """current kernel: per trip, cheap steps until each lane holds a test, then the test together"""
    total = 0
    for w in range(0, len(order) - 31, 32):
        lanes = [traces[i] for i in order[w:w + 32]]
        pos = [0] * 32
        while True:
            maxcheap = 0
            anytest = False
            alive = False
            for l in range(32):
                tr = lanes[l]
                c = 0
                while pos[l] < len(tr) and tr[pos[l]] != ord('T'):
                    c += COST[tr[pos[l]]] if tr[pos[l]] != ord('R') else 25
                    pos[l] += 1
                if pos[l] < len(tr):
                    anytest = True
                    c += 25
                    pos[l] += 1
                    alive = True
                maxcheap = max(maxcheap, c)
            total += maxcheap + (90 if anytest else 0)
            if not alive:
                break
    return total


def sim_whilewhile(order):
    total = 0
    for w in range(0, len(order) - 31, 32):
        lanes = [traces[i].split(b'P') for i in order[w:w + 32]]   # segments: I* L (R|T)*
        nseg = max(len(s) for s in lanes)
        for k in range(nseg):
            mi = mt = 0
            for s in lanes:
                if k < len(s):
                    seg = s[k]
                    mi = max(mi, seg.count(b'I'))
                    mt = max(mt, seg.count(b'T') * 115 + seg.count(b'R') * 25)
            total += mi * 40 + 12 + mt + 25
    return total


idx = np.arange(n)
octant = (d[:, 0] > 0).astype(int) | ((d[:, 1] > 0).astype(int) << 1) | ((d[:, 2] > 0).astype(int) << 2)
orders = {"pixel order": idx}
for chunk in (256, 1024, 4096):
    o = []
    for s in range(0, n, chunk):
        seg = idx[s:s + chunk]
        o.extend(seg[np.argsort(octant[seg], kind="stable")])
    orders["octant within %d" % chunk] = np.array(o)
cell = (np.clip(p / 555.0 * 4, 0, 3.999).astype(int) * np.array([1, 4, 16])).sum(axis=1)
orders["global sort (octant, cell4^3)"] = np.lexsort((cell, octant))
orders["global sort (cell4^3, octant)"] = np.lexsort((octant, cell))
def morton3(v, bits):
    out = np.zeros(v.shape[0], np.int64)
    for b in range(bits):
        for a in range(3):
            out |= ((v[:, a] >> b) & 1).astype(np.int64) << (3 * b + a)
    return out


for ob, db in ((3, 2), (4, 3), (5, 4), (6, 5)):
    oc = np.clip(p / 555.0 * (1 << ob), 0, (1 << ob) - 1).astype(np.int64)
    dc = np.clip((d * 0.5 + 0.5) * (1 << db), 0, (1 << db) - 1).astype(np.int64)
    key = (morton3(dc, db) << (3 * ob)) | morton3(oc, ob)
    orders["sort dir%db/axis then origin%db/axis" % (db, ob)] = np.argsort(key, kind="stable")
    key2 = (morton3(oc, ob) << (3 * db)) | morton3(dc, db)
    orders["sort origin%db/axis then dir%db/axis" % (ob, db)] = np.argsort(key2, kind="stable")
for k, o in orders.items():
    nw = (len(o) // 32) * 32
    u = sum(sum(COST[c] for c in traces[i]) for i in o[:nw])
    a, b = sim_steploop(o), sim_whilewhile(o)
    print("%-32s step-loop eff %.3f   while-while eff %.3f" % (k, u / (32.0 * a), u / (32.0 * b)))


def sim_onestep(order):
    """each trip every live lane executes exactly one step; the warp pays for every kind present"""
    total = 0
    for w in range(0, len(order) - 31, 32):
        lanes = [traces[i] for i in order[w:w + 32]]
        L = max(len(t) for t in lanes)
        for k in range(L):
            kinds = set(t[k] for t in lanes if k < len(t))
            total += sum(COST[c] for c in kinds)
    return total


def sim_kcheap(order, K):
    """per trip: up to K cheap steps per lane (stop early at a test), then the test for lanes that hold one"""
    total = 0
    for w in range(0, len(order) - 31, 32):
        lanes = [traces[i] for i in order[w:w + 32]]
        pos = [0] * 32
        while True:
            maxcheap = 0
            anytest = False
            alive = False
            for l in range(32):
                tr = lanes[l]
                c = 0
                k = 0
                while pos[l] < len(tr) and tr[pos[l]] != ord('T') and k < K:
                    c += COST[tr[pos[l]]]
                    pos[l] += 1
                    k += 1
                if pos[l] < len(tr) and tr[pos[l]] == ord('T'):
                    anytest = True
                    c += 25
                    pos[l] += 1
                if pos[l] < len(tr):
                    alive = True
                maxcheap = max(maxcheap, c)
            total += maxcheap + (90 if anytest else 0)
            if not alive:
                break
    return total


o = idx
nw = (len(o) // 32) * 32
u = sum(sum(COST[c] for c in traces[i]) for i in o[:nw])
print("one-step-per-trip eff %.3f" % (u / (32.0 * sim_onestep(o))))
for K in (1, 2, 3, 4, 6, 8):
    print("K=%d cheap steps then test: eff %.3f" % (K, u / (32.0 * sim_kcheap(o, K))))
