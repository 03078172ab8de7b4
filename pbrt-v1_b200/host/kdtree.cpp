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
// Shift a subtree's stored indices and append it.
void splice(Subtree &dst, const Subtree &src) {
  uint32_t nodeBase = (uint32_t)dst.nodes.size(), refBase = (uint32_t)dst.refs.size();
  for (size_t i = 0; i < src.nodes.size(); ++i) {
    PbrtB200KdNode n = src.nodes[i];
    if ((n.word0 & 3u) != 3u) n.word1 += nodeBase;            // interior: above-child index
    else if ((n.word0 >> 2) > 1u) n.word1 += refBase;         // leaf list offset
    dst.nodes.push_back(n);
  }
  dst.refs.insert(dst.refs.end(), src.refs.begin(), src.refs.end());
}
}  // namespace

// The split decision of one node (kdtree.cpp:189-299).  Returns false when the node becomes a leaf;
// otherwise axis / tsplit and the two primitive lists in prims0[0..n0) / prims1[0..n1).
static bool choose_split(const BuildParams &bp, const BBox &nodeBounds, const int *primNums, int nPrims,
                         KdTreeAccel::BoundEdge *edges[3], int *prims0, int *prims1, int *badRefines,
                         int *bestAxisOut, float *tsplitOut, int *n0Out, int *n1Out);

struct KdBuilder {
  typedef KdTreeAccel::BoundEdge BoundEdge;
  static void buildTree(const BuildParams &bp, Subtree &st, int nodeNum, const BBox &nodeBounds,
                        int *primNums, int nPrims, int depth, BoundEdge *edges[3], int *prims0,
                        int *prims1, int badRefines);
  static void buildParallel(const BuildParams &bp, Subtree &st, const BBox &nodeBounds, vector<int> &primNums,
                            int depth, int badRefines, int forkLevels);
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
    const bool tm = nPrims > 900000 && getenv("PBRT_HOST_TIMING");
    double ta = tm ? host_seconds() : 0;
    std::sort(ev, ev + 2 * nPrims, EventOrder());
    double tb = tm ? host_seconds() : 0;
    sweep.run(axis, ev, 2 * nPrims, nPrims);
    if (tm) fprintf(stderr, "  root axis %d: sort %.3f s, sweep %.3f s\n", axis, tb - ta, host_seconds() - tb);
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
// A big node: decide its split here, build the two children concurrently, splice.
void KdBuilder::buildParallel(const BuildParams &bp, Subtree &st, const BBox &nodeBounds, vector<int> &primNums,
                              int depth, int badRefines, int forkLevels) {
  const int nPrims = (int)primNums.size();
  const size_t np = (size_t)nPrims;
  if (forkLevels <= 0 || nPrims < 4096 || nPrims <= bp.maxPrims || depth == 0) {
    // sequential build of this subtree, the reference's recursion with its scratch arrays
    vector<BoundEdge> e0(2 * np + 1), e1(2 * np + 1), e2(2 * np + 1);
    BoundEdge *edges[3] = {&e0[0], &e1[0], &e2[0]};
    vector<int> prims0(np + 1), prims1((size_t)(depth + 1) * np + 1);
    buildTree(bp, st, 0, nodeBounds, nPrims ? &primNums[0] : NULL, nPrims, depth, edges, &prims0[0], &prims1[0],
              badRefines);
    return;
  }
  PbrtB200KdNode blank = {0u, 0u};
  st.nodes.push_back(blank);
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
    init_leaf(st, 0, &primNums[0], nPrims);
    return;
  }
  vector<int>().swap(primNums);   // the parent's list is not needed any more
  below.resize(n0); above.resize(n1);
  st.nodes[0].word0 = interior_word0(tsplit, bestAxis);
  BBox bounds0 = nodeBounds, bounds1 = nodeBounds;
  bounds0.pMax[bestAxis] = bounds1.pMin[bestAxis] = tsplit;
  Subtree sub0, sub1;
  std::future<void> other = std::async(std::launch::async, [&]() {
    buildParallel(bp, sub1, bounds1, above, depth - 1, badRefines, forkLevels - 1);
  });
  buildParallel(bp, sub0, bounds0, below, depth - 1, badRefines, forkLevels - 1);
  other.get();
  splice(st, sub0);
  st.nodes[0].word1 = (uint32_t)st.nodes.size();
  splice(st, sub1);
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
  BuildParams bp = {isectCost, traversalCost, maxPrims, emptyBonus, &primBounds};
  // PBRT_KD_THREADS=1 forces the sequential build; default: as many fork levels as the cores warrant
  int threads = (int)std::thread::hardware_concurrency();
  if (const char *e = getenv("PBRT_KD_THREADS")) threads = atoi(e);
  int forkLevels = 0;
  while ((1 << forkLevels) < threads) ++forkLevels;
  if (threads > 1) forkLevels += 2;   // children differ in size: over-decompose a little
  Subtree st;
  KdBuilder::buildParallel(bp, st, bounds, primNums, maxDepth, 0, forkLevels);
  nodes.swap(st.nodes);
  primRefs.swap(st.refs);
  if (timing)
    fprintf(stderr, "kd-tree: %zu primitives, refine + bounds %.3f s, build %.3f s (%d threads), %zu nodes\n", np,
            t1 - t0, host_seconds() - t1, threads, nodes.size());
}
