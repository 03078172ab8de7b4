// pb_kd.h -- kd-tree walk, the device twin of KdTreeAccel::Intersect
// (accelerators/kdtree.cpp:313-403) and IntersectP (:404-488), over the packed
// 8-byte nodes of include/pbrt_b200.h.
//
// Same tree, same child ordering, same near/far/both decisions, same
// `ray.maxt < tmin` early-out, so primitives are tested in the reference's
// order and equal-t ties resolve the same way (later test wins,
// shapes/trianglemesh.cpp:245).  The reference's mailboxes
// (kdtree.cpp:373-377) only skip repeated tests of a primitive that straddles
// several leaves; a small per-ray ring of recently tested primitive ids plays
// that role here.
#ifndef PB_KD_H
#define PB_KD_H
#include "pb_shapes.h"

struct KdCounters {
  uint32_t interior, leaves, refs, tests;
};

// core/geometry.cpp:51-68
PB_HD bool bbox_intersect_p(const float *bmin, const float *bmax, const Ray &ray, float *hitt0,
                            float *hitt1) {
  float t0 = ray.mint, t1 = ray.maxt;
  for (int i = 0; i < 3; ++i) {
    float invRayDir = 1.f / ray.d[i];
    float tNear = (bmin[i] - ray.o[i]) * invRayDir;
    float tFar = (bmax[i] - ray.o[i]) * invRayDir;
    if (tNear > tFar) { float tmp = tNear; tNear = tFar; tFar = tmp; }
    t0 = tNear > t0 ? tNear : t0;
    t1 = tFar < t1 ? tFar : t1;
    if (t0 > t1) return false;
  }
  *hitt0 = t0; *hitt1 = t1;
  return true;
}

#ifndef PB_MAILBOX
#define PB_MAILBOX 2   // measured 2 > 4 > 8 on B200: the ring costs more than the few re-tests it avoids
#endif
// Scenes without object instances filter repeats with a direct-mapped table of PB_MB_SLOTS recently tested
// primitives per ray instead (slot = multiplicative hash of the primitive number; an entry is only ever compared
// for equality, so a collision costs a repeated test, never a skipped one).  oracle/twin.cpp's
// pbrt_twin_mailbox_study on C4's bounce rays: tests per closest-hit ray 14.5 with the 2-entry ring, 8.75 / 8.30 with
// 8 / 16 slots, 8.21 with the reference's exact per-primitive mailboxes (kdtree.cpp:373-377); C3 13.0 / 9.8 / 9.0 /
// 8.8.  The warp walk keeps the table in shared memory (pb_kd_warp.cuh), this per-thread walk in a local array.
// 0 = the ring everywhere.
#ifndef PB_MB_SLOTS
#define PB_MB_SLOTS 16          // closest-hit walks (Scene::Intersect, MIS probes)
#endif
#ifndef PB_MB_SLOTS_ANY
#define PB_MB_SLOTS_ANY PB_MB_SLOTS   // any-hit walks (Scene::IntersectP)
#endif
template <int SLOTS>
PB_HD uint32_t pb_mb_slot(uint32_t prim) {
  static_assert(SLOTS == 4 || SLOTS == 8 || SLOTS == 16 || SLOTS == 32, "mailbox table: 4, 8, 16 or 32 slots (0: the ring)");
  return (prim * 0x9E3779B1u) >> (SLOTS == 32 ? 27 : SLOTS == 16 ? 28 : SLOTS == 8 ? 29 : 30);
}

// The scene's top-level tree / an instanced object's tree.
PB_HD KdTreeView kd_top_tree(const SceneView &sv) {
  KdTreeView t;
  t.nodes = sv.nodes; t.primRefs = sv.primRefs; t.nNodes = sv.nNodes; t.primBase = 0; t.paired = sv.pairedNodes;
  for (int i = 0; i < 3; ++i) { t.bmin[i] = sv.bmin[i]; t.bmax[i] = sv.bmax[i]; }
  return t;
}
PB_HD KdTreeView kd_instance_tree(const SceneView &sv, const PbrtB200Instance &in) {
  KdTreeView t;
  t.nodes = sv.nodes + in.node_base; t.primRefs = sv.primRefs + in.ref_base;
  t.nNodes = in.n_nodes; t.primBase = in.prim_base; t.paired = 0;
  for (int i = 0; i < 3; ++i) { t.bmin[i] = in.bounds[i]; t.bmax[i] = in.bounds[3 + i]; }
  return t;
}

// QK = the primitive kinds the scene holds, so that traversal kernels for simpler scenes compile the
// rest out: 0 triangles only, 1 triangles + spheres, 2 every quadric, 3 + object instances.
#define PB_QK_TRI 0
#define PB_QK_SPHERE 1
#define PB_QK_ALL 2
#define PB_QK_INSTANCES 3

template <bool ANY, bool COUNT, int QK, bool TABLE = false>
PB_HD bool kd_tree_intersect(const SceneView &sv, const KdTreeView &tree, Ray ray, Hit *hit, KdCounters *cnt);

// One primitive test: GeometricPrimitive::Intersect (core/primitive.cpp:113-122)
// shrinks ray.maxt on an accepted hit.  With QK == PB_QK_INSTANCES a primitive may be an
// InstancePrimitive (core/primitive.cpp:66-85): the ray goes into instance space, the instanced
// object's own kd-tree (or single primitive) is intersected, r.maxt = ray.maxt.
template <bool ANY, int QK = PB_QK_INSTANCES, bool COUNT = false>
PB_HD bool kd_test_prim(const SceneView &sv, uint32_t prim, Ray &ray, Hit *hit, KdCounters *cnt = nullptr) {
  TriVerts tv = load_prim(sv, prim);
  float t, b1 = 0.f, b2 = 0.f;
  bool ok;
  const uint32_t kind = tv.flags & PBRT_B200_PF_KIND_MASK;
  if (QK == PB_QK_TRI || kind == PBRT_B200_PRIM_TRIANGLE) {
    ok = tri_test(tv.p1, tv.p2, tv.p3, ray, &t, &b1, &b2);
  } else if (QK != PB_QK_INSTANCES || kind == PBRT_B200_PRIM_QUADRIC) {
    V3 phit; float phi;
    if (QK == PB_QK_SPHERE) ok = sphere_test(sv.quadrics[sv.aux[prim].owner], ray, &t, &phit, &phi, false);
    else ok = quadric_test(sv.quadrics[sv.aux[prim].owner], ray, &t, &phit, &phi, false);
  } else if constexpr (QK == PB_QK_INSTANCES) {
    const uint32_t instId = sv.aux[prim].owner;
    const PbrtB200Instance &in = sv.instances[instId];
    Ray r2;
    r2.o = xf_point(in.w2i, ray.o);
    r2.d = xf_vector(in.w2i, ray.d);
    r2.mint = ray.mint; r2.maxt = ray.maxt;
    Hit h2;
    h2.prim = -1; h2.t = 0.f; h2.b1 = h2.b2 = 0.f; h2.inst = -1;
    bool got;
    if (in.n_nodes == 0) got = kd_test_prim<ANY, PB_QK_ALL>(sv, in.prim_base, r2, &h2);
    else got = kd_tree_intersect<ANY, COUNT, PB_QK_ALL, true>(sv, kd_instance_tree(sv, in), r2, &h2, cnt);   // (table per object walk)
    if (!got) return false;
    if (!ANY) {
      *hit = h2;
      hit->inst = (int32_t)instId;
      ray.maxt = h2.t;
    }
    return true;
  } else
    return false;
  if (!ok) return false;
  if (!ANY) {
    hit->prim = (int32_t)prim; hit->t = t; hit->b1 = b1; hit->b2 = b2; hit->inst = -1;
    ray.maxt = t;
  }
  return true;
}

// Returns true when something was hit.  ANY=false: closest hit in *hit.
//
// Loop shape ("while-while"): descend through interior nodes until a leaf is
// reached, then test the leaf's primitives, then pop.  On a GPU this lets the
// lanes of a warp run the short interior step together and the long triangle
// tests together instead of interleaving them; per ray the sequence of nodes
// visited and primitives tested is exactly that of the reference's single loop
// (kdtree.cpp:328-401).  The reference re-checks `ray.maxt < tmin` at the top of
// every iteration; neither value changes while descending, so checking after
// each pop is equivalent.
template <bool ANY, bool COUNT, int QK, bool TABLE>
PB_HD bool kd_tree_intersect(const SceneView &sv, const KdTreeView &tree, Ray ray, Hit *hit, KdCounters *cnt) {
  if (!ANY) hit->prim = -1;
  float tmin, tmax;
  if (tree.nNodes == 0) return false;
  if (!bbox_intersect_p(tree.bmin, tree.bmax, ray, &tmin, &tmax)) return false;
  V3 invDir = v3(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
  uint32_t todoNode[PB_KD_STACK];
  float todoTmin[PB_KD_STACK], todoTmax[PB_KD_STACK];
  int todoPos = 0;
  uint32_t mailbox[PB_MAILBOX];
#pragma unroll
  for (int i = 0; i < PB_MAILBOX; ++i) mailbox[i] = 0xffffffffu;
  int mbPos = 0;
  constexpr int MBS = !TABLE ? 0 : (ANY ? PB_MB_SLOTS_ANY : PB_MB_SLOTS);
  constexpr bool HASHED = MBS != 0;
  uint32_t tested[HASHED ? MBS : 1];
  if (HASHED)
    for (int i = 0; i < MBS; ++i) tested[i] = 0xffffffffu;
  bool anyHit = false;
  uint32_t node = 0;
  for (;;) {
    uint32_t w0, w1;
    // ---- descend to a leaf
    for (;;) {
#ifdef __CUDA_ARCH__
      uint2 nd = __ldg((const uint2 *)(tree.nodes + node));
      w0 = nd.x; w1 = nd.y;
#else
      w0 = tree.nodes[node].word0; w1 = tree.nodes[node].word1;
#endif
      if ((w0 & 3u) == 3u) break;
      if (COUNT) ++cnt->interior;
      int axis = (int)(w0 & 3u);
#ifdef __CUDA_ARCH__
      float split = __uint_as_float(w0);   // the axis bits stay in the mantissa (kdtree.cpp:66-72)
#else
      float split; { uint32_t b = w0; memcpy(&split, &b, 4); }
#endif
      float o = ray.o[axis], d = ray.d[axis];
      float tplane = (split - o) * invDir[axis];
      bool belowFirst = (o < split) || (o == split && d >= 0);
      const uint32_t below = tree.paired ? w1 : node + 1, above = tree.paired ? w1 + 1 : w1;   // (pb_treelet.h)
      uint32_t firstChild = belowFirst ? below : above;
      uint32_t secondChild = belowFirst ? above : below;
      if (tplane > tmax || tplane <= 0)
        node = firstChild;
      else if (tplane < tmin)
        node = secondChild;
      else {
        todoNode[todoPos] = secondChild;
        todoTmin[todoPos] = tplane;
        todoTmax[todoPos] = tmax;
        ++todoPos;
        node = firstChild;
        tmax = tplane;
      }
    }
    // ---- leaf
    if (COUNT) ++cnt->leaves;
    uint32_t nPrims = w0 >> 2;
    for (uint32_t i = 0; i < nPrims; ++i) {
      uint32_t prim = tree.primBase + ((nPrims == 1) ? w1 : tree.primRefs[w1 + i]);
      if (COUNT) ++cnt->refs;
      if (HASHED) {
        const uint32_t slot = pb_mb_slot<HASHED ? MBS : 4>(prim);
        if (tested[slot] == prim) continue;
        tested[slot] = prim;
      } else {
        bool seen = false;
#pragma unroll
        for (int k = 0; k < PB_MAILBOX; ++k) seen |= (mailbox[k] == prim);
        if (seen) continue;
#pragma unroll
        for (int k = 0; k < PB_MAILBOX; ++k)
          if (k == mbPos) mailbox[k] = prim;
        mbPos = (mbPos + 1) & (PB_MAILBOX - 1);
      }
      if (COUNT) ++cnt->tests;
      if (kd_test_prim<ANY, QK, COUNT>(sv, prim, ray, hit, cnt)) {
        if (ANY) return true;
        anyHit = true;
      }
    }
    // ---- next node from the todo list
    if (todoPos == 0) break;
    --todoPos;
    node = todoNode[todoPos];
    tmin = todoTmin[todoPos];
    tmax = todoTmax[todoPos];
    if (!ANY && ray.maxt < tmin) break;
  }
  return anyHit;
}
// KdTreeAccel::Intersect / IntersectP on the scene's top-level tree.
template <bool ANY, bool COUNT>
PB_HD bool kd_intersect(const SceneView &sv, Ray ray, Hit *hit, KdCounters *cnt) {
  if (!ANY) { hit->prim = -1; hit->inst = -1; }
  if (sv.nInstances != 0) return kd_tree_intersect<ANY, COUNT, PB_QK_INSTANCES>(sv, kd_top_tree(sv), ray, hit, cnt);
  return kd_tree_intersect<ANY, COUNT, PB_QK_ALL, true>(sv, kd_top_tree(sv), ray, hit, cnt);
}
#endif
