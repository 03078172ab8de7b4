// kdtree.cpp -- host-side SAH kd-tree construction.
//
// Restates KdTreeAccel's constructor and buildTree (accelerators/kdtree.cpp:
// 141-312) so that, for the same primitive list, the tree is node-for-node the
// reference's tree (same split axes/positions, same leaf contents, same node
// numbering: below child at node+1, above child index stored).  That makes
// the per-ray kd counters of the GPU walk comparable with an instrumented
// reference.  The output is the packed 8-byte node array of
// include/pbrt_b200.h instead of the reference's pointer-carrying nodes.
#include "core.h"
#include <thread>
#include <atomic>
#include <climits>
#include <chrono>
static double host_seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void Primitive::FullyRefine(vector<Reference<Primitive> > &refined) const {
  // core/primitive.cpp:40-53: LIFO todo list => a mesh's triangles come out
  // in reverse index order.
  vector<Reference<Primitive> > todo;
  todo.push_back(const_cast<Primitive *>(this));
  while (todo.size()) {
    Reference<Primitive> prim = todo.back();
    todo.pop_back();
    if (prim->CanIntersect()) refined.push_back(prim);
    else prim->Refine(todo);
  }
}

void GeometricPrimitive::Refine(vector<Reference<Primitive> > &refined) const {
  vector<Reference<Shape> > r;
  shape->Refine(r);
  for (u_int i = 0; i < r.size(); ++i)
    refined.push_back(new GeometricPrimitive(r[i], material, areaLight));
}

// One end of a primitive's bounding interval along the axis being examined.  Events are ordered by
// position; where an interval opens at the very position another closes, the opening comes first (so a
// primitive that merely touches the plane from above is counted on both sides of it, like the reference's
// edge order, accelerators/kdtree.cpp:89-105).  Events that compare equal keep whatever order std::sort
// leaves them in -- the build below depends on that being the same algorithm the reference ran (libstdc++'s
// introsort on the same input sequence), see DESIGN.md "kd-tree build".
struct KdTreeAccel::BoundEdge {
  float pos;
  int prim;
  unsigned char closes;   // 0: the interval opens here, 1: it closes here
};
namespace {
typedef KdTreeAccel::BoundEdge PlaneEvent;
struct EventOrder {
  bool operator()(const PlaneEvent &a, const PlaneEvent &b) const {
    return a.pos == b.pos ? a.closes < b.closes : a.pos < b.pos;
  }
};
}  // namespace

// ---------------------------------------------------------------------------
// Build.  The recursion is the reference's (accelerators/kdtree.cpp:172-312); what is added is that
// a subtree is built into its own node / reference arrays with numbers relative to the subtree, so
// that the two children of a big node can be built by different threads and spliced afterwards:
// node numbering is depth first (below child = parent + 1, above child = after the whole below
// subtree) and leaves append their lists in visiting order, so concatenating
// [node] + below + above (and the reference lists likewise) and shifting the stored indices
// reproduces exactly the arrays the sequential build writes.
namespace {
struct Subtree {
  vector<PbrtB200KdNode> nodes;
  vector<uint32_t> refs;
};
struct BuildParams {
  int isectCost, traversalCost, maxPrims;
  float emptyBonus;
  const vector<BBox> *allPrimBounds;
  int sortThreads;   // threads a big node's event sort may use (sort_events)
};
void init_leaf(Subtree &st, int nodeNum, const int *primNums, int np) {
  PbrtB200KdNode &n = st.nodes[nodeNum];
  n.word0 = ((uint32_t)np << 2) | 3u;          // kdtree.cpp:53-54
  if (np == 0) n.word1 = 0;
  else if (np == 1) n.word1 = (uint32_t)primNums[0];
  else {
    n.word1 = (uint32_t)st.refs.size();
    for (int i = 0; i < np; ++i) st.refs.push_back((uint32_t)primNums[i]);
  }
}
}  // namespace

// The split decision of one node (kdtree.cpp:189-299).  Returns false when the node becomes a leaf;
// otherwise axis / tsplit and the two primitive lists in prims0[0..n0) / prims1[0..n1).
static bool choose_split(const BuildParams &bp, const BBox &nodeBounds, const int *primNums, int nPrims,
                         KdTreeAccel::BoundEdge *edges[3], int *prims0, int *prims1, int *badRefines,
                         int *bestAxisOut, float *tsplitOut, int *n0Out, int *n1Out);

namespace { struct Piece; }
struct KdBuilder {
  typedef KdTreeAccel::BoundEdge BoundEdge;
  static void buildTree(const BuildParams &bp, Subtree &st, int nodeNum, const BBox &nodeBounds,
                        int *primNums, int nPrims, int depth, BoundEdge *edges[3], int *prims0,
                        int *prims1, int badRefines);
  static void buildParallel(const BuildParams &bp, Piece &pc, const BBox &nodeBounds, vector<int> &primNums,
                            int depth, int badRefines, int forkPrims);
};
// Surface-area heuristic along one axis: walks the sorted events keeping the primitive counts on either side
// of the candidate plane and returns the cheapest plane strictly inside the node (index into `ev`, -1 if none).
// cost = traversal + intersection * (1 - bonus) * (P(below) * nBelow + P(above) * nAbove), where P is the ratio
// of the child's surface area to the node's and the bonus applies when one side is empty.
namespace {
struct AxisSweep {
  const BuildParams &bp;
  const BBox &box;
  Vector extent;
  float invNodeArea;
  int best;
  float bestCost;
  AxisSweep(const BuildParams &p, const BBox &b) : bp(p), box(b), extent(b.pMax - b.pMin), best(-1), bestCost(INFINITY) {
    float nodeArea = (2.f * (extent.x * extent.y + extent.x * extent.z + extent.y * extent.z));
    invNodeArea = 1.f / nodeArea;
  }
  void run(int axis, const PlaneEvent *ev, int nEvents, int nPrims) {
    const int u = (axis + 1) % 3 < (axis + 2) % 3 ? (axis + 1) % 3 : (axis + 2) % 3;   // the two other axes,
    const int v = 3 - axis - u;                                                         // lower one first
    const float lo = box.pMin[axis], hi = box.pMax[axis];
    int below = 0, above = nPrims;
    for (int i = 0; i < nEvents; ++i) {
      const PlaneEvent &e = ev[i];
      if (e.closes) --above;
      if (e.pos > lo && e.pos < hi) {
        float areaBelow = 2 * (extent[u] * extent[v] + (e.pos - lo) * (extent[u] + extent[v]));
        float areaAbove = 2 * (extent[u] * extent[v] + (hi - e.pos) * (extent[u] + extent[v]));
        float pBelow = areaBelow * invNodeArea, pAbove = areaAbove * invNodeArea;
        float bonus = (above == 0 || below == 0) ? bp.emptyBonus : 0.f;
        float cost = bp.traversalCost + bp.isectCost * (1.f - bonus) * (pBelow * below + pAbove * above);
        if (cost < bestCost) { bestCost = cost; best = i; }
      }
      if (!e.closes) ++below;
    }
  }
};
}  // namespace
// std::sort's result, computed by several threads.  The order std::sort leaves EQUAL events in decides which
// primitives land in which child (and in which order), so the sort of a big node cannot be replaced by another
// algorithm -- but it can be run concurrently: libstdc++'s std::sort is
//     __introsort_loop(first, last, 2 * lg(n));  __final_insertion_sort(first, last);
// and the introsort loop (bits/stl_algo.h) partitions [first, last) around a median-of-three pivot, recurses into the
// upper part and continues with the lower one.  The two parts are disjoint ranges and nothing else is shared, so
// handing the upper part of a large range to another thread changes no comparison and no swap.  The functions
// called are the library's own (same pivot rule, same heap-sort fallback at depth 0), hence the same bytes out; the
// final insertion sort then runs once over the whole range as in std::sort.
// (tests/test_oracle.py::test_parallel_kd_build_equals_sequential compares the trees.)
#if defined(__GLIBCXX__)
#include <future>
namespace {
template <class It, class Cmp>
void introsort_loop_parallel(It first, It last, long depthLimit, Cmp cmp, int forks) {
  std::vector<std::future<void> > pending;
  while (last - first > 16) {   // _S_threshold
    if (depthLimit == 0) {
      std::__partial_sort(first, last, last, cmp);
      break;
    }
    --depthLimit;
    It cut = std::__unguarded_partition_pivot(first, last, cmp);
    if (forks > 0 && last - cut > 32768) {
      --forks;
      const int f = forks;
      pending.push_back(std::async(std::launch::async, [=]() { introsort_loop_parallel(cut, last, depthLimit, cmp, f); }));
    } else
      std::__introsort_loop(cut, last, depthLimit, cmp);
    last = cut;
  }
  for (size_t i = 0; i < pending.size(); ++i) pending[i].get();
}
}  // namespace
static void sort_events(PlaneEvent *first, PlaneEvent *last, int threads) {
  if (threads <= 1 || last - first < 131072) {
    std::sort(first, last, EventOrder());
    return;
  }
  auto cmp = __gnu_cxx::__ops::__iter_comp_iter(EventOrder());
  int forks = 0;
  while ((1 << forks) < threads) ++forks;
  introsort_loop_parallel(first, last, std::__lg(last - first) * 2, cmp, forks + 1);
  std::__final_insertion_sort(first, last, cmp);
}
#else
static void sort_events(PlaneEvent *first, PlaneEvent *last, int) { std::sort(first, last, EventOrder()); }
#endif

static bool choose_split(const BuildParams &bp, const BBox &nodeBounds, const int *primNums, int nPrims,
                         KdTreeAccel::BoundEdge *edges[3], int *prims0, int *prims1, int *badRefines,
                         int *bestAxisOut, float *tsplitOut, int *n0Out, int *n1Out) {
  const vector<BBox> &primBounds = *bp.allPrimBounds;
  AxisSweep sweep(bp, nodeBounds);
  // the longest axis first; if no plane of an axis falls inside the node, the next axis (cyclically) is tried
  const Vector &ext = sweep.extent;
  const int firstAxis = (ext.x > ext.y && ext.x > ext.z) ? 0 : (ext.y > ext.z ? 1 : 2);
  int splitAxis = -1;
  for (int attempt = 0; attempt < 3 && splitAxis < 0; ++attempt) {
    const int axis = (firstAxis + attempt) % 3;
    PlaneEvent *ev = edges[axis];
    for (int i = 0; i < nPrims; ++i) {
      const BBox &b = primBounds[primNums[i]];
      ev[2 * i].pos = b.pMin[axis]; ev[2 * i].prim = primNums[i]; ev[2 * i].closes = 0;
      ev[2 * i + 1].pos = b.pMax[axis]; ev[2 * i + 1].prim = primNums[i]; ev[2 * i + 1].closes = 1;
    }
    sort_events(ev, ev + 2 * nPrims, bp.sortThreads);
    sweep.run(axis, ev, 2 * nPrims, nPrims);
    if (sweep.best >= 0) splitAxis = axis;
  }
  // give up (leaf) when splitting is not worth it: repeatedly worse than not splitting, or much worse on a
  // small node, or no plane found at all
  const float leafCost = bp.isectCost * float(nPrims);
  if (sweep.bestCost > leafCost) ++*badRefines;
  if ((sweep.bestCost > 4.f * leafCost && nPrims < 16) || splitAxis < 0 || *badRefines == 3) return false;
  // primitives whose interval opens before the chosen event go below, those that close after it go above
  const PlaneEvent *ev = edges[splitAxis];
  int nBelow = 0, nAbove = 0;
  for (int i = 0; i < sweep.best; ++i)
    if (!ev[i].closes) prims0[nBelow++] = ev[i].prim;
  for (int i = sweep.best + 1; i < 2 * nPrims; ++i)
    if (ev[i].closes) prims1[nAbove++] = ev[i].prim;
  *bestAxisOut = splitAxis;
  *tsplitOut = ev[sweep.best].pos;
  *n0Out = nBelow; *n1Out = nAbove;
  return true;
}

// interior node: split float with the axis in its two low mantissa bits
// (kdtree.cpp:66-72: `split = s; flags &= ~3; flags |= axis;`)
static uint32_t interior_word0(float tsplit, int axis) {
  uint32_t bits;
  memcpy(&bits, &tsplit, 4);
  return (bits & ~3u) | (uint32_t)axis;
}

void KdBuilder::buildTree(const BuildParams &bp, Subtree &st, int nodeNum, const BBox &nodeBounds,
                          int *primNums, int nPrims, int depth, BoundEdge *edges[3], int *prims0,
                          int *prims1, int badRefines) {
  if ((size_t)nodeNum != st.nodes.size()) Severe("kd-tree build: node numbering out of sync");
  PbrtB200KdNode blank = {0u, 0u};
  st.nodes.push_back(blank);
  if (nPrims <= bp.maxPrims || depth == 0) {
    init_leaf(st, nodeNum, primNums, nPrims);
    return;
  }
  int bestAxis, n0, n1;
  float tsplit;
  if (!choose_split(bp, nodeBounds, primNums, nPrims, edges, prims0, prims1, &badRefines, &bestAxis, &tsplit,
                    &n0, &n1)) {
    init_leaf(st, nodeNum, primNums, nPrims);
    return;
  }
  st.nodes[nodeNum].word0 = interior_word0(tsplit, bestAxis);
  BBox bounds0 = nodeBounds, bounds1 = nodeBounds;
  bounds0.pMax[bestAxis] = bounds1.pMin[bestAxis] = tsplit;
  buildTree(bp, st, nodeNum + 1, bounds0, prims0, n0, depth - 1, edges, prims0, prims1 + nPrims, badRefines);
  st.nodes[nodeNum].word1 = (uint32_t)st.nodes.size();
  buildTree(bp, st, (int)st.nodes[nodeNum].word1, bounds1, prims1, n1, depth - 1, edges, prims0,
            prims1 + nPrims, badRefines);
}

#include <future>
#include <memory>
// A big node: decide its split here and build the two children concurrently.  The result is a tree of pieces -- a
// forked node holds only its own split, a piece built sequentially holds its whole subtree with subtree-relative
// numbers -- which assemble_pieces() lays out afterwards: node numbering is depth first (a node, its whole below
// subtree, its above subtree) and leaves append their reference lists in the same order, so every piece's position
// in the final arrays follows from the sizes of the pieces before it, and the pieces are copied (indices shifted)
// by all threads at once.  Same arrays as the sequential build, bit for bit.
namespace {
struct Piece {
  Subtree st;                       // a sequentially built subtree (or a single leaf); unused for a fork
  bool fork = false;
  uint32_t word0 = 0;               // fork: the node's split word
  std::unique_ptr<Piece> below, above;
  size_t nodeBase = 0, refBase = 0, nNodes = 0, nRefs = 0;
};
void measure_pieces(Piece &pc, size_t nodeBase, size_t refBase, vector<Piece *> *leaves) {
  pc.nodeBase = nodeBase; pc.refBase = refBase;
  if (!pc.fork) {
    pc.nNodes = pc.st.nodes.size(); pc.nRefs = pc.st.refs.size();
    leaves->push_back(&pc);
    return;
  }
  measure_pieces(*pc.below, nodeBase + 1, refBase, leaves);
  measure_pieces(*pc.above, nodeBase + 1 + pc.below->nNodes, refBase + pc.below->nRefs, leaves);
  pc.nNodes = 1 + pc.below->nNodes + pc.above->nNodes;
  pc.nRefs = pc.below->nRefs + pc.above->nRefs;
}
void write_forks(const Piece &pc, vector<PbrtB200KdNode> &nodes) {
  if (!pc.fork) return;
  nodes[pc.nodeBase].word0 = pc.word0;
  nodes[pc.nodeBase].word1 = (uint32_t)pc.above->nodeBase;
  write_forks(*pc.below, nodes);
  write_forks(*pc.above, nodes);
}
void assemble_pieces(Piece &root, int threads, vector<PbrtB200KdNode> *nodes, vector<uint32_t> *refs) {
  vector<Piece *> leaves;
  measure_pieces(root, 0, 0, &leaves);
  PbrtB200KdNode blank = {0u, 0u};
  nodes->assign(root.nNodes, blank);
  refs->assign(root.nRefs, 0u);
  write_forks(root, *nodes);
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= leaves.size()) return;
      const Piece &pc = *leaves[i];
      const uint32_t nodeBase = (uint32_t)pc.nodeBase, refBase = (uint32_t)pc.refBase;
      PbrtB200KdNode *dst = nodes->data() + pc.nodeBase;
      for (size_t k = 0; k < pc.st.nodes.size(); ++k) {
        PbrtB200KdNode n = pc.st.nodes[k];
        if ((n.word0 & 3u) != 3u) n.word1 += nodeBase;            // interior: above-child index
        else if ((n.word0 >> 2) > 1u) n.word1 += refBase;         // leaf list offset
        dst[k] = n;
      }
      if (!pc.st.refs.empty()) memcpy(refs->data() + pc.refBase, pc.st.refs.data(), pc.st.refs.size() * sizeof(uint32_t));
    }
  };
  vector<std::thread> pool;
  for (int t = 1; t < std::max(1, threads); ++t) pool.emplace_back(worker);
  worker();
  for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
}
}  // namespace
void KdBuilder::buildParallel(const BuildParams &bp, Piece &pc, const BBox &nodeBounds, vector<int> &primNums,
                              int depth, int badRefines, int forkPrims) {
  const int nPrims = (int)primNums.size();
  const size_t np = (size_t)nPrims;
  Subtree &st = pc.st;
  if (nPrims < forkPrims || nPrims <= bp.maxPrims || depth == 0) {
    // sequential build of this subtree, the reference's recursion with its scratch arrays
    vector<BoundEdge> e0(2 * np + 1), e1(2 * np + 1), e2(2 * np + 1);
    BoundEdge *edges[3] = {&e0[0], &e1[0], &e2[0]};
    vector<int> prims0(np + 1), prims1((size_t)(depth + 1) * np + 1);
    buildTree(bp, st, 0, nodeBounds, nPrims ? &primNums[0] : NULL, nPrims, depth, edges, &prims0[0], &prims1[0],
              badRefines);
    return;
  }
  vector<int> below(np + 1), above(np + 1);
  int bestAxis, n0, n1;
  float tsplit;
  bool split;
  {
    vector<BoundEdge> e0(2 * np + 1), e1(2 * np + 1), e2(2 * np + 1);
    BoundEdge *edges[3] = {&e0[0], &e1[0], &e2[0]};
    split = choose_split(bp, nodeBounds, &primNums[0], nPrims, edges, &below[0], &above[0], &badRefines, &bestAxis,
                         &tsplit, &n0, &n1);
  }
  if (!split) {
    PbrtB200KdNode blank = {0u, 0u};
    st.nodes.push_back(blank);
    init_leaf(st, 0, &primNums[0], nPrims);
    return;
  }
  vector<int>().swap(primNums);   // the parent's list is not needed any more
  below.resize(n0); above.resize(n1);
  pc.fork = true;
  pc.word0 = interior_word0(tsplit, bestAxis);
  pc.below.reset(new Piece);
  pc.above.reset(new Piece);
  BBox bounds0 = nodeBounds, bounds1 = nodeBounds;
  bounds0.pMax[bestAxis] = bounds1.pMin[bestAxis] = tsplit;
  Piece *pa = pc.above.get();
  std::future<void> other = std::async(std::launch::async, [&, pa]() {
    buildParallel(bp, *pa, bounds1, above, depth - 1, badRefines, forkPrims);
  });
  buildParallel(bp, *pc.below, bounds0, below, depth - 1, badRefines, forkPrims);
  other.get();
}

KdTreeAccel::KdTreeAccel(const vector<Reference<Primitive> > &p, int icost, int tcost,
                         float ebonus, int maxp, int maxDepth)
    : isectCost(icost), traversalCost(tcost), maxPrims(maxp), emptyBonus(ebonus) {
  const bool timing = getenv("PBRT_HOST_TIMING") != NULL;
  const double t0 = host_seconds();
  for (u_int i = 0; i < p.size(); ++i) p[i]->FullyRefine(prims);
  if (maxDepth <= 0) maxDepth = Round2Int(8 + 1.3f * Log2Int(float(prims.size())));
  vector<BBox> primBounds;
  primBounds.reserve(prims.size());
  for (u_int i = 0; i < prims.size(); ++i) {
    BBox b = prims[i]->WorldBound();
    bounds = Union(bounds, b);
    primBounds.push_back(b);
  }
  const double t1 = host_seconds();
  size_t np = prims.size();
  vector<int> primNums(np);
  for (u_int i = 0; i < np; ++i) primNums[i] = i;
  // PBRT_KD_THREADS=1 forces the sequential build; default: one thread per core
  int threads = (int)std::thread::hardware_concurrency();
  if (const char *e = getenv("PBRT_KD_THREADS")) threads = atoi(e);
  threads = std::max(1, threads);
  BuildParams bp = {isectCost, traversalCost, maxPrims, emptyBonus, &primBounds, threads};
  // nodes with at least forkPrims primitives fork their children onto threads: about 16 sequential pieces per
  // thread, so that pieces of different cost still balance
  const int forkPrims = threads > 1 ? std::max(4096, (int)(np / (16 * (size_t)threads))) : INT_MAX;
  Piece root;
  KdBuilder::buildParallel(bp, root, bounds, primNums, maxDepth, 0, forkPrims);
  const double t2 = host_seconds();
  assemble_pieces(root, threads, &nodes, &primRefs);
  if (timing) fprintf(stderr, "kd-tree: pieces built in %.3f s, assembled in %.3f s\n", t2 - t1, host_seconds() - t2);
  if (timing)
    fprintf(stderr, "kd-tree: %zu primitives, refine + bounds %.3f s, build %.3f s (%d threads), %zu nodes\n", np,
            t1 - t0, host_seconds() - t1, threads, nodes.size());
}

// Test hook (tests/test_oracle.py): sort_events() on `threads` threads against std::sort on the same input --
// n events at pseudo-random positions drawn from `distinct` values, so that runs of equal keys are long.
// Returns the number of positions where the two results differ (0 = same permutation).
extern "C" long pbrt_host_test_sort_events(long n, long distinct, unsigned seed, int threads) {
  vector<PlaneEvent> a((size_t)n), b;
  unsigned s = seed * 2654435761u + 12345u;
  for (long i = 0; i < n; ++i) {
    s = s * 1664525u + 1013904223u;
    a[(size_t)i].pos = (float)((s >> 8) % (unsigned)distinct) * 0.25f;
    s = s * 1664525u + 1013904223u;
    a[(size_t)i].closes = (unsigned char)((s >> 16) & 1u);
    a[(size_t)i].prim = (int)i;
  }
  b = a;
  std::sort(a.begin(), a.end(), EventOrder());
  sort_events(&b[0], &b[0] + n, threads);
  long diff = 0;
  for (long i = 0; i < n; ++i) diff += a[(size_t)i].prim != b[(size_t)i].prim;
  return diff;
}
