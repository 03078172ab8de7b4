// pb_kd_warp.cuh -- warp-scheduled kd-tree walk used by the traversal kernels.
//
// Per ray this performs exactly the sequence of node visits and primitive
// tests of pb_kd.h's kd_intersect() (== KdTreeAccel::Intersect / IntersectP,
// accelerators/kdtree.cpp:313-488): same tree, same near/far order, same
// `maxt < tmin` cut, same mailbox ring.  What differs is how a warp's 32 rays
// are scheduled, because a 1-ray-per-thread transcription of the reference loop
// measured 3.7 active lanes out of 32 on a path-traced Cornell box (ncu,
// profiles/r01a_*): rays sit in leaves of different sizes at different times.
//
//   * Each lane is a small state machine: DESCEND (interior steps), LEAF
//     (walk the leaf's primitive list), DONE.
//   * One trip of the warp loop = every live lane first advances through cheap
//     steps (interior nodes, leaf bookkeeping, pops) until it holds ONE
//     primitive that needs testing, then all lanes run the expensive
//     ray/primitive test together.  The triangle test therefore executes with
//     every lane that still has work, instead of only the lanes that happen to
//     be in a leaf in the same iteration.
//   * Persistent threads with lane refill: when enough lanes have finished
//     their ray, they claim fresh rays from the queue's atomic work counter
//     (warp-aggregated), so short rays do not wait for the longest one.
#ifndef PB_KD_WARP_CUH
#define PB_KD_WARP_CUH
#include "pb_kd.h"

#ifndef PB_REFILL_THRESHOLD
#define PB_REFILL_THRESHOLD 16  // refill when at least this many lanes are idle (measured: 1 << 4 << 8 < 16)
#endif
#ifndef PB_DESCEND_STEPS
#define PB_DESCEND_STEPS 2
#endif
#ifndef PB_LEAF_STEPS
#define PB_LEAF_STEPS 2
#endif
#ifndef PB_GENERAL_ADAPTIVE
#define PB_GENERAL_ADAPTIVE 1   // instance-capable walk: adaptive interior phase and early refills like the lean walk (the
#endif                          // 27-placement scene: 916 -> 781 ms per frame, profiles/r02p_inst_variants.log)
#ifndef PB_GENERAL_REFILL
#define PB_GENERAL_REFILL 8
#endif
#ifndef PB_GENERAL_DESCEND_LANES
#define PB_GENERAL_DESCEND_LANES 8
#endif
#ifndef PB_GENERAL_DESCEND_MAX
#define PB_GENERAL_DESCEND_MAX 8
#endif
#ifndef PB_SMEM_NODES
#define PB_SMEM_NODES 0   // > 0: the first PB_SMEM_NODES nodes of the top-level tree are staged in shared memory
#endif                    //      by one TMA bulk copy per CTA (experiment)

#if PB_SMEM_NODES > 0
__device__ __forceinline__ uint32_t pb_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// One thread arms an mbarrier with the byte count and issues cp.async.bulk (TMA, 1-D); everyone waits on it.
__device__ __forceinline__ void pb_stage_nodes(const PbrtB200KdNode *src, uint2 *dst, unsigned long long *bar, uint32_t nNodes) {
  const uint32_t bytes = nNodes * 8u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pb_smem_addr(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (bytes == 0) return;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pb_smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pb_smem_addr(dst)), "l"(src), "r"(bytes), "r"(pb_smem_addr(bar)) : "memory");
  }
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PB_WAIT_NODES:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@!p bra PB_WAIT_NODES;\n"
      "}\n" ::"r"(pb_smem_addr(bar)) : "memory");
}
#endif

// Source: Ray load(uint32_t i).   Sink: void done(uint32_t i, const Ray &ray, const Hit &hit, bool anyHit).
// DS: interior steps per trip (PB_DESCEND_STEPS for small trees; trees with > ~100 k nodes visit 3-4 interior
// nodes per primitive test and run 3-4 % faster with 4: profiles/r02c_sweep.log, r02d_sweep.log)
template <bool ANY, bool COUNT, int QK, int DS, class Source, class Sink>
__device__ __forceinline__ void kd_warp_traverse(const SceneView &sv, uint32_t n, uint32_t *work,
                                                 const Source &src, const Sink &sink, KdCounters &cnt) {
  const uint32_t lane = threadIdx.x & 31u;
  // todo list of 8-byte entries (node, tmax): the popped tmin is the tmax that is current at the pop (see the lean
  // walk below); entering an instanced object parks the outer tmax in saved[], because the object's walk runs with
  // its own interval
  uint2 todo[PB_KD_STACK];
  // lane state
  bool active = false;        // holds an unfinished ray
  bool inLeaf = false;
  bool exhausted = false;     // the queue has no more rays (warp-uniform)
  uint32_t idx = 0;
  Ray ray;
  ray.o = v3(0, 0, 0); ray.d = v3(0, 0, 0); ray.mint = 0.f; ray.maxt = 0.f;
  V3 invDir = v3(0, 0, 0);
  float tmin = 0.f, tmax = 0.f;
  uint32_t node = 0;
  int todoPos = 0;
  uint32_t leafW1 = 0, leafI = 0, leafN = 0;
  uint32_t mailbox[PB_MAILBOX];
  int mbPos = 0;
  Hit hit;
  hit.prim = -1; hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.inst = -1;
  bool anyHit = false;
  // Object instances (QK == PB_QK_INSTANCES kernels only): while a lane walks an instanced object's
  // tree its ray is the instance-space ray and node / reference / primitive numbers are relative to
  // these offsets; the world-space ray, the interrupted leaf walk and the outer mailbox ring wait
  // in `saved`.  instBase = todo depth at entry (-1: not inside an instance).
  constexpr bool INSTV = QK == PB_QK_INSTANCES;
  uint32_t nodeBase = 0, refBase = 0, primBase = 0;
  int instBase = -1;
  int32_t instId = -1;
  float saved[INSTV ? 10 : 1];
  // repeats inside an instanced object's tree are filtered by the hashed table of pb_kd.h (cleared at every entry);
  // the top-level tree, which holds the placements and a few loose primitives, keeps the 2-entry ring
  constexpr int MBS = INSTV ? (ANY ? PB_MB_SLOTS_ANY : PB_MB_SLOTS) : 0;
  __shared__ uint32_t testedAll[MBS ? MBS : 1][MBS ? PB_BLOCK : 1];
  uint32_t *const tested = &testedAll[0][MBS ? threadIdx.x : 0];

#if PB_SMEM_NODES > 0
  __shared__ __align__(128) uint2 topNodes[PB_SMEM_NODES];
  __shared__ __align__(8) unsigned long long topBar;
  const uint32_t nTop = (sv.nNodes < (uint32_t)PB_SMEM_NODES ? sv.nNodes : (uint32_t)PB_SMEM_NODES) & ~1u;   // 16-byte granules
  pb_stage_nodes(sv.nodes, topNodes, &topBar, nTop);
#endif
  for (;;) {
    // ---- refill idle lanes
    uint32_t idleMask = __ballot_sync(0xffffffffu, !active);
    if (!exhausted && (uint32_t)__popc(idleMask) >= (idleMask == 0xffffffffu ? 1u : (uint32_t)PB_GENERAL_REFILL)) {
      uint32_t base = 0;
      uint32_t nIdle = (uint32_t)__popc(idleMask);
      if (lane == 0) base = atomicAdd(work, nIdle);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + nIdle >= n) exhausted = true;
      if (!active) {
        uint32_t i = base + (uint32_t)__popc(idleMask & ((1u << lane) - 1u));
        if (i < n) {
          idx = i;
          ray = src.load(i);
          hit.prim = -1;
          anyHit = false;
          todoPos = 0;
          inLeaf = false;
          node = 0;
          if (INSTV) { nodeBase = refBase = primBase = 0; instBase = -1; instId = -1; }
#pragma unroll
          for (int k = 0; k < PB_MAILBOX; ++k) mailbox[k] = 0xffffffffu;
          mbPos = 0;
          active = sv.nNodes != 0 && bbox_intersect_p(sv.bmin, sv.bmax, ray, &tmin, &tmax);
          invDir = v3(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);

          if (!active) sink.done(idx, ray, hit, false);   // missed the tree's bounds
        }
      }
    }
    if (__ballot_sync(0xffffffffu, active) == 0) {
      if (exhausted) break;
      continue;
    }
    // ---- bookkeeping, in two uniform phases so that lanes doing the same kind of step run together:
    //      A. descend: up to PB_DESCEND_STEPS interior nodes (entering the leaf when one is reached)
    //      B. leaf: up to PB_LEAF_STEPS list entries / pops, until the lane holds one primitive to test
    // (bounded, so a lane deep in interior nodes never holds back 31 lanes that are ready to test;
    // offline SIMT model: tools/simt_sim.py)
    bool pending = false;
    uint32_t prim = 0;
#pragma unroll 1
    for (int step = 0; step < (PB_GENERAL_ADAPTIVE ? PB_GENERAL_DESCEND_MAX : DS); ++step) {
      bool descending = active && !inLeaf;
      const uint32_t dm = __ballot_sync(0xffffffffu, descending);
      if (dm == 0) break;
      // (PB_GENERAL_ADAPTIVE: like the lean walk, the interior phase goes on while enough lanes still descend)
      if (PB_GENERAL_ADAPTIVE && step != 0 && (uint32_t)__popc(dm) < (uint32_t)PB_GENERAL_DESCEND_LANES) break;
      if (descending) {
#if PB_SMEM_NODES > 0
        uint2 nd;
        if (node < nTop && !(INSTV && nodeBase != 0u)) nd = topNodes[node];
        else nd = __ldg((const uint2 *)(sv.nodes + (INSTV ? nodeBase : 0u) + node));
#else
        uint2 nd = __ldg((const uint2 *)(sv.nodes + (INSTV ? nodeBase : 0u) + node));
#endif
        if ((nd.x & 3u) == 3u) {
          if (COUNT) ++cnt.leaves;
          inLeaf = true;
          leafN = nd.x >> 2; leafW1 = nd.y; leafI = 0;
        } else {
          if (COUNT) ++cnt.interior;
          int axis = (int)(nd.x & 3u);
          float split = __uint_as_float(nd.x);   // axis bits stay in the mantissa (kdtree.cpp:66-72)
          float o = ray.o[axis], d = ray.d[axis];
          float tplane = (split - o) * invDir[axis];
          bool belowFirst = (o < split) || (o == split && d >= 0);
          uint32_t firstChild = belowFirst ? node + 1 : nd.y;
          uint32_t secondChild = belowFirst ? nd.y : node + 1;
          if (tplane > tmax || tplane <= 0)
            node = firstChild;
          else if (tplane < tmin)
            node = secondChild;
          else {
            todo[todoPos] = make_uint2(secondChild, __float_as_uint(tmax));
            ++todoPos;
            node = firstChild;
            tmax = tplane;
          }
        }
      }
    }
#pragma unroll 1
    for (int step = 0; step < PB_LEAF_STEPS; ++step) {
      bool walking = active && inLeaf && !pending && leafI < leafN;
      if (__ballot_sync(0xffffffffu, walking) == 0) break;
      if (walking) {
        prim = (INSTV ? primBase : 0u) + ((leafN == 1) ? leafW1 : __ldg(sv.primRefs + (INSTV ? refBase : 0u) + leafW1 + leafI));
        ++leafI;
        if (COUNT) ++cnt.refs;
        if (MBS != 0 && instBase >= 0) {
          uint32_t *const entry = tested + pb_mb_slot<MBS ? MBS : 4>(prim) * PB_BLOCK;
          if (*entry != prim) {
            *entry = prim;
            pending = true;
          }
        } else {
          bool seen = false;
#pragma unroll
          for (int k = 0; k < PB_MAILBOX; ++k) seen |= (mailbox[k] == prim);
          if (!seen) {
#pragma unroll
            for (int k = 0; k < PB_MAILBOX; ++k)
              if (k == mbPos) mailbox[k] = prim;
            mbPos = (mbPos + 1) & (PB_MAILBOX - 1);
            pending = true;
          }
        }
      }
    }
    // ---- the expensive part, all lanes with work together
    if (pending) {
      if (COUNT) ++cnt.tests;
      bool entered = false;
      if (INSTV && instBase < 0 &&
          (__float_as_uint(__ldg(&sv.prims[3u * prim + 2].b)) & PBRT_B200_PF_KIND_MASK) == PBRT_B200_PRIM_INSTANCE) {
        // InstancePrimitive::Intersect (core/primitive.cpp:66-85): Ray ray = WorldToInstance(r), then
        // the object's KdTreeAccel::Intersect as part of this lane's state machine
        const uint32_t id = sv.aux[prim].owner;
        const PbrtB200Instance &in = sv.instances[id];
        entered = true;
        Ray r2;
        r2.o = xf_point(in.w2i, ray.o);
        r2.d = xf_vector(in.w2i, ray.d);
        r2.mint = ray.mint; r2.maxt = ray.maxt;
        if (in.n_nodes == 0) {
          // the object is one intersectable primitive (no accelerator, api.cpp:441)
          Hit h2;
          h2.prim = -1; h2.t = 0.f; h2.b1 = h2.b2 = 0.f; h2.inst = -1;
          if (kd_test_prim<ANY, PB_QK_ALL>(sv, in.prim_base, r2, &h2)) {
            anyHit = true;
            if (!ANY) { hit = h2; hit.inst = (int32_t)id; ray.maxt = h2.t; }
            else { active = false; sink.done(idx, ray, hit, true); }
          }
        } else {
          float t0, t1;
          if (bbox_intersect_p(in.bounds, in.bounds + 3, r2, &t0, &t1)) {
            saved[0] = ray.o.x; saved[1] = ray.o.y; saved[2] = ray.o.z;
            saved[3] = ray.d.x; saved[4] = ray.d.y; saved[5] = ray.d.z;
            saved[6] = __uint_as_float(leafW1); saved[7] = __uint_as_float(leafI); saved[8] = __uint_as_float(leafN);
            saved[9] = tmax;
            if (MBS != 0) {
#pragma unroll
              for (int k = 0; k < MBS; ++k) tested[k * PB_BLOCK] = 0xffffffffu;
            } else {
              // (no table: the object's walk would need the ring; not a configuration that is built)
            }
            ray = r2;
            invDir = v3(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
            tmin = t0; tmax = t1;
            nodeBase = in.node_base; refBase = in.ref_base; primBase = in.prim_base;
            instBase = todoPos;
            instId = (int32_t)id;
            node = 0;
            inLeaf = false;
          }
        }
      }
      if (!entered) {
        if (kd_test_prim<ANY, (INSTV ? PB_QK_ALL : QK)>(sv, prim, ray, &hit)) {
          anyHit = true;
          if (INSTV && !ANY && instBase >= 0) hit.inst = instId;
          if (ANY) {
            active = false;
            sink.done(idx, ray, hit, true);
          }
        }
      }
    }
    // ---- leaf finished (its last test is in): next node from the todo list (kdtree.cpp:392-399), so the
    //      lane descends again in the next trip instead of spending a trip on the pop
    if (active && inLeaf && leafI >= leafN) {
      inLeaf = false;
      bool finished = false;   // this tree's walk is over
      if (todoPos == (INSTV && instBase >= 0 ? instBase : 0)) finished = true;
      else {
        --todoPos;
        const uint2 e = todo[todoPos];
        node = e.x;
        tmin = tmax;
        tmax = __uint_as_float(e.y);
        if (!ANY && ray.maxt < tmin) finished = true;   // kdtree.cpp:330
      }
      if (INSTV && finished && instBase >= 0) {
        // the instanced object's walk is over: r.maxt = ray.maxt (primitive.cpp:71), back to the
        // world-space ray and the leaf that held the instance
        todoPos = instBase;
        ray.o = v3(saved[0], saved[1], saved[2]);
        ray.d = v3(saved[3], saved[4], saved[5]);
        invDir = v3(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
        leafW1 = __float_as_uint(saved[6]); leafI = __float_as_uint(saved[7]); leafN = __float_as_uint(saved[8]);
        tmax = saved[9];
        nodeBase = refBase = primBase = 0;
        instBase = -1;
        inLeaf = true;
      } else if (finished) {
        active = false;
        sink.done(idx, ray, hit, anyHit);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// The walk of scenes without object instances (every BASELINE configuration), written for issue slots: round-2
// ncu of the general loop above on C5 (profiles/r02e_ncu_trav_*) shows 333 warp instructions per trip at 12 / 6 / 8
// active lanes in the interior / leaf-list / test steps, 58 of them per interior step (the per-axis operands were
// picked by branches, the lane flags lived in byte registers).  Same per-ray visit and test sequence; differences:
//   * lane state is one word (idle / descending / in a leaf / finished), the per-axis operands of an interior step
//     come from predicated selects, the 2-entry mailbox ring is a shift register;
//   * the todo list holds 8-byte entries (node, tmax) -- the popped tmin is the current tmax, see below;
//   * repeats are filtered by a hashed table of recently tested primitives in shared memory (pb_kd.h);
//   * (PB_LEAN_INLINE_POP, off) an EMPTY leaf -- 43 % of the leaves C4's bounce rays reach -- can be popped inside
//     the interior step that finds it, so that the lane keeps descending in the same trip.  The instruction-count
//     model liked it (-7 %), the GPU did not: the pop's local-memory loads sit in the divergent leaf branch of every
//     interior step and their latency is exposed there (C4 closest hit 117.5 ms with, 108.2 ms without;
//     profiles/r02g_sweep_walk_variants.log);
//   * the interior phase runs while at least PB_DESCEND_LANES lanes are still descending (at most PB_DESCEND_MAX
//     steps) instead of a fixed step count (tools/simt_sim2.py: -17 % issued instructions on C4's bounce rays);
//   * hit.t is not carried: it always equals ray.maxt once something was hit.
#ifndef PB_DESCEND_LANES
#define PB_DESCEND_LANES 8    // (profiles/r02h_sweep_walk_tuning.log: 8 < 12 < 16 on C3, C4 and C5)
#endif
// Idle lanes fetch new rays once this many of the warp's lanes are idle.  Small trees (C3: 61 nodes) like full
// refills, big ones earlier ones: C4 247.7 -> 237.1 ms and C5 1592 -> 1535 ms with 8 instead of 16, C3 127.8 -> 130.1.
#ifndef PB_REFILL_SMALL
#define PB_REFILL_SMALL PB_REFILL_THRESHOLD
#endif
#ifndef PB_REFILL_BIG
#define PB_REFILL_BIG 8
#endif
#ifndef PB_BIG_TREE_NODES
#define PB_BIG_TREE_NODES 100000u
#endif
#ifndef PB_DESCEND_MAX
#define PB_DESCEND_MAX 8
#endif
#ifndef PB_LEAN_INLINE_POP
#define PB_LEAN_INLINE_POP 0   // pop an empty leaf inside the interior step that finds it (measured: slower, see below)
#endif
#ifndef PB_LEAN_LEAF_STEPS
#define PB_LEAN_LEAF_STEPS 2
#endif
// o[axis], invDir[axis] with two compares and four selects
__device__ __forceinline__ void pb_pick_axis(uint32_t axis, const V3 &o, const V3 &inv, float *oa, float *ia) {
  asm("{\n\t.reg .pred p0, p1;\n\t"
      "setp.eq.u32 p0, %2, 0;\n\t"
      "setp.eq.u32 p1, %2, 1;\n\t"
      "selp.f32 %0, %4, %5, p1;\n\t"
      "selp.f32 %1, %7, %8, p1;\n\t"
      "selp.f32 %0, %3, %0, p0;\n\t"
      "selp.f32 %1, %6, %1, p0;\n\t}"
      : "=&f"(*oa), "=&f"(*ia)
      : "r"(axis), "f"(o.x), "f"(o.y), "f"(o.z), "f"(inv.x), "f"(inv.y), "f"(inv.z));
}
enum { PB_LS_IDLE = 0, PB_LS_DESCEND = 1, PB_LS_LEAF = 2, PB_LS_FINISHED = 3, PB_LS_QUADRIC = 4 };
#ifndef PB_QUADRIC_LANES
#define PB_QUADRIC_LANES 1   // > 1: quadric tests are parked until this many lanes of the warp hold one (measured: slower)
#endif

template <bool ANY, bool COUNT, int QK, class Source, class Sink>
__device__ __forceinline__ void kd_warp_traverse_lean(const SceneView &sv, uint32_t n, uint32_t *work,
                                                      const Source &src, const Sink &sink, KdCounters &cnt) {
  static_assert(QK != PB_QK_INSTANCES, "object instances use kd_warp_traverse");
  static_assert(PB_MAILBOX == 2, "the shift-register mailbox holds two entries");
  // per-lane table of recently tested primitives (pb_kd.h): column threadIdx.x, so a warp's accesses never conflict
  constexpr int MBS = ANY ? PB_MB_SLOTS_ANY : PB_MB_SLOTS;
  __shared__ uint32_t testedAll[MBS ? MBS : 1][MBS ? PB_BLOCK : 1];
  uint32_t *const tested = &testedAll[0][MBS ? threadIdx.x : 0];
  const uint32_t lane = threadIdx.x & 31u;
  // todo list, 8 bytes per entry: (node, tmax).  The reference's entry also holds tmin (kdtree.cpp:318-322), but the
  // interval that is current when an entry is popped always ends where that entry's begins -- a push sets
  // tmax = tplane = the new entry's tmin, single-child steps leave tmax alone, and a pop restores the tmax that was
  // current at the push, which is the tmin of the entry below -- so the popped tmin is the value tmax already holds.
  uint2 todo[PB_KD_STACK];
  uint32_t state = PB_LS_IDLE;
  bool exhausted = false;     // the queue has no more rays (warp-uniform)
  uint32_t idx = 0;
  Ray ray;
  ray.o = v3(0, 0, 0); ray.d = v3(0, 0, 0); ray.mint = 0.f; ray.maxt = 0.f;
  V3 invDir = v3(0, 0, 0);
  uint32_t dirNeg = 0;        // bit a: !(ray.d[a] >= 0)
  float tmin = 0.f, tmax = 0.f;
  uint32_t node = 0;
  int todoPos = 0;
  uint32_t leafW1 = 0, leafI = 0, leafN = 0;
  uint32_t mb0 = 0xffffffffu, mb1 = 0xffffffffu;   // the two most recently tested primitives
  Hit hit;
  hit.prim = -1; hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.inst = -1;
  uint32_t parkedPrim = 0;    // PB_LS_QUADRIC: the quadric waiting to be tested
  const uint2 *nodes = (const uint2 *)sv.nodes;
  const uint32_t refillAt = sv.nNodes > PB_BIG_TREE_NODES ? (uint32_t)PB_REFILL_BIG : (uint32_t)PB_REFILL_SMALL;
  for (;;) {
    // ---- refill idle lanes
    const uint32_t idleMask = __ballot_sync(0xffffffffu, state == PB_LS_IDLE);
    if (!exhausted && (uint32_t)__popc(idleMask) >= (idleMask == 0xffffffffu ? 1u : refillAt)) {
      uint32_t base = 0;
      const uint32_t nIdle = (uint32_t)__popc(idleMask);
      if (lane == 0) base = atomicAdd(work, nIdle);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + nIdle >= n) exhausted = true;
      if (state == PB_LS_IDLE) {
        const uint32_t i = base + (uint32_t)__popc(idleMask & ((1u << lane) - 1u));
        if (i < n) {
          idx = i;
          ray = src.load(i);
          hit.prim = -1;
          todoPos = 0;
          node = 0;
          mb0 = mb1 = 0xffffffffu;
          if (MBS) {
#pragma unroll
            for (int k = 0; k < MBS; ++k) tested[k * PB_BLOCK] = 0xffffffffu;
          }
          const bool in = sv.nNodes != 0 && bbox_intersect_p(sv.bmin, sv.bmax, ray, &tmin, &tmax);
          invDir = v3(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
          dirNeg = (ray.d.x >= 0 ? 0u : 1u) | (ray.d.y >= 0 ? 0u : 2u) | (ray.d.z >= 0 ? 0u : 4u);
          if (in) state = PB_LS_DESCEND;
          else sink.done(idx, ray, hit, false);   // missed the tree's bounds
        }
      }
    }
    if (idleMask == 0xffffffffu && __ballot_sync(0xffffffffu, state != PB_LS_IDLE) == 0) {
      if (exhausted) break;
      continue;
    }
    // ---- A. interior steps (while the queue drains, the bar follows the number of lanes that still hold a ray)
    const uint32_t descendLanes = exhausted ? min((uint32_t)PB_DESCEND_LANES, (32u - (uint32_t)__popc(idleMask)) >> 1)
                                            : (uint32_t)PB_DESCEND_LANES;
#pragma unroll 1
    for (int step = 0; step < PB_DESCEND_MAX; ++step) {
      const bool descending = state == PB_LS_DESCEND;
      const uint32_t m = __ballot_sync(0xffffffffu, descending);
      if (m == 0 || (step != 0 && (uint32_t)__popc(m) < descendLanes)) break;
      if (descending) {
        const uint2 nd = __ldg(nodes + node);
        const uint32_t axis = nd.x & 3u;
        if (axis == 3u) {
          if (COUNT) ++cnt.leaves;
          leafN = nd.x >> 2; leafW1 = nd.y; leafI = 0;
          if (leafN != 0) state = PB_LS_LEAF;
          else if (!PB_LEAN_INLINE_POP) state = PB_LS_LEAF;
          else if (todoPos == 0) state = PB_LS_FINISHED;
          else {
            // empty leaf: next node from the todo list right away (kdtree.cpp:392-399, :330)
            --todoPos;
            const uint2 e = todo[todoPos];
            node = e.x;
            tmin = tmax;
            tmax = __uint_as_float(e.y);
            if (!ANY && ray.maxt < tmin) state = PB_LS_FINISHED;
          }
        } else {
          if (COUNT) ++cnt.interior;
          const float split = __uint_as_float(nd.x);   // axis bits stay in the mantissa (kdtree.cpp:66-72)
          float o, inv;
          pb_pick_axis(axis, ray.o, invDir, &o, &inv);
          const float tplane = (split - o) * inv;
          const bool belowFirst = (o < split) || (o == split && !((dirNeg >> axis) & 1u));
          const uint32_t below = sv.pairedNodes ? nd.y : node + 1u;        // (pb_treelet.h)
          const uint32_t above = sv.pairedNodes ? nd.y + 1u : nd.y;
          const uint32_t firstChild = belowFirst ? below : above;
          const uint32_t secondChild = belowFirst ? above : below;
          if (tplane > tmax || tplane <= 0)
            node = firstChild;
          else if (tplane < tmin)
            node = secondChild;
          else {
            todo[todoPos] = make_uint2(secondChild, __float_as_uint(tmax));
            ++todoPos;
            node = firstChild;
            tmax = tplane;
          }
        }
      }
    }
    // ---- B. leaf-list steps until the lane holds one primitive to test
    bool pending = false;
    uint32_t prim = 0;
#pragma unroll 1
    for (int step = 0; step < PB_LEAN_LEAF_STEPS; ++step) {
      const bool walking = state == PB_LS_LEAF && !pending && leafI < leafN;
      if (__ballot_sync(0xffffffffu, walking) == 0) break;
      if (walking) {
        prim = (leafN == 1) ? leafW1 : __ldg(sv.primRefs + leafW1 + leafI);
        ++leafI;
        if (COUNT) ++cnt.refs;
        if (MBS) {
          uint32_t *const entry = tested + pb_mb_slot<MBS ? MBS : 4>(prim) * PB_BLOCK;
          if (*entry != prim) {
            *entry = prim;
            pending = true;
          }
        } else if (prim != mb0 && prim != mb1) {
          mb1 = mb0; mb0 = prim;
          pending = true;
        }
      }
    }
    // ---- C. the expensive part, all lanes with work together.  (PB_QUADRIC_LANES > 1, an experiment: a lane whose
    //      primitive is not a triangle parks in PB_LS_QUADRIC and the parked lanes are tested together once that
    //      many wait or nothing else in the warp can advance -- the quadric test is several hundred instructions and
    //      on C3's bounce rays it runs with one or two active lanes.  A parked lane resumes later with its own
    //      sequence of tests unchanged, but its idle trips cost more than the fuller quadric tests return: C3 frame
    //      128.3 / 128.6 / 130.2 / 134.6 ms with 1 / 2 / 4 / 8 lanes, profiles/r02i_sweep_quadric_parking.log.)
    if (pending) {
      if (COUNT) ++cnt.tests;
      if (QK == PB_QK_TRI || PB_QUADRIC_LANES <= 1) {
        if (kd_test_prim<ANY, QK>(sv, prim, ray, &hit)) {
          if (ANY) { hit.prim = 0; state = PB_LS_FINISHED; }
        }
      } else {
        const TriVerts tv = load_prim(sv, prim);
        if ((tv.flags & PBRT_B200_PF_KIND_MASK) == PBRT_B200_PRIM_TRIANGLE) {
          float t, b1, b2;
          if (tri_test(tv.p1, tv.p2, tv.p3, ray, &t, &b1, &b2)) {
            if (ANY) { hit.prim = 0; state = PB_LS_FINISHED; }
            else { hit.prim = (int32_t)prim; hit.b1 = b1; hit.b2 = b2; ray.maxt = t; }
          }
        } else {
          parkedPrim = prim;
          state = PB_LS_QUADRIC;
        }
      }
    }
    if (QK != PB_QK_TRI && PB_QUADRIC_LANES > 1) {
      const uint32_t qm = __ballot_sync(0xffffffffu, state == PB_LS_QUADRIC);
      if (qm != 0u) {
        bool run = (uint32_t)__popc(qm) >= (uint32_t)PB_QUADRIC_LANES || exhausted;
        if (!run) run = __ballot_sync(0xffffffffu, state == PB_LS_DESCEND || (state == PB_LS_LEAF && leafI < leafN)) == 0u;
        if (run && state == PB_LS_QUADRIC) {
          state = PB_LS_LEAF;
          if (kd_test_prim<ANY, QK>(sv, parkedPrim, ray, &hit)) {
            if (ANY) { hit.prim = 0; state = PB_LS_FINISHED; }
          }
        }
      }
    }
    // ---- D. leaf finished: next node from the todo list (kdtree.cpp:392-399)
    if (state == PB_LS_LEAF && leafI >= leafN) {
      if (todoPos == 0) state = PB_LS_FINISHED;
      else {
        --todoPos;
        const uint2 e = todo[todoPos];
        node = e.x;
        tmin = tmax;
        tmax = __uint_as_float(e.y);
        state = (!ANY && ray.maxt < tmin) ? PB_LS_FINISHED : PB_LS_DESCEND;   // kdtree.cpp:330
      }
    }
    if (state == PB_LS_FINISHED) {
      state = PB_LS_IDLE;
      hit.t = ray.maxt;
      sink.done(idx, ray, hit, hit.prim >= 0);
    }
  }
}
#endif
