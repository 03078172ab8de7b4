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

struct KdTreeAccel::BoundEdge {
  BoundEdge() {}
  BoundEdge(float tt, int pn, bool starting) : t(tt), primNum(pn), type(starting ? START : END) {}
  bool operator<(const BoundEdge &e) const {
    if (t == e.t) return (int)type < (int)e.type;
    else return t < e.t;
  }
  float t;
  int primNum;
  enum { START, END } type;
};

KdTreeAccel::KdTreeAccel(const vector<Reference<Primitive> > &p, int icost, int tcost,
                         float ebonus, int maxp, int maxDepth)
    : isectCost(icost), traversalCost(tcost), maxPrims(maxp), emptyBonus(ebonus) {
  for (u_int i = 0; i < p.size(); ++i) p[i]->FullyRefine(prims);
  if (maxDepth <= 0) maxDepth = Round2Int(8 + 1.3f * Log2Int(float(prims.size())));
  vector<BBox> primBounds;
  primBounds.reserve(prims.size());
  for (u_int i = 0; i < prims.size(); ++i) {
    BBox b = prims[i]->WorldBound();
    bounds = Union(bounds, b);
    primBounds.push_back(b);
  }
  size_t np = prims.size();
  BoundEdge *edges[3];
  for (int i = 0; i < 3; ++i) edges[i] = new BoundEdge[2 * np + 1];
  int *prims0 = new int[np + 1];
  int *prims1 = new int[(size_t)(maxDepth + 1) * np + 1];
  int *primNums = new int[np + 1];
  for (u_int i = 0; i < np; ++i) primNums[i] = i;
  buildTree(0, bounds, primBounds, primNums, (int)np, maxDepth, edges, prims0, prims1, 0);
  delete[] primNums;
  for (int i = 0; i < 3; ++i) delete[] edges[i];
  delete[] prims0;
  delete[] prims1;
}

void KdTreeAccel::initLeaf(int nodeNum, int *primNums, int np) {
  PbrtB200KdNode &n = nodes[nodeNum];
  n.word0 = ((uint32_t)np << 2) | 3u;          // kdtree.cpp:53-54
  if (np == 0) n.word1 = 0;
  else if (np == 1) n.word1 = (uint32_t)primNums[0];
  else {
    n.word1 = (uint32_t)primRefs.size();
    for (int i = 0; i < np; ++i) primRefs.push_back((uint32_t)primNums[i]);
  }
}

void KdTreeAccel::buildTree(int nodeNum, const BBox &nodeBounds,
                            const vector<BBox> &allPrimBounds, int *primNums, int nPrims,
                            int depth, BoundEdge *edges[3], int *prims0, int *prims1,
                            int badRefines) {
  if ((size_t)nodeNum != nodes.size()) Severe("kd-tree build: node numbering out of sync");
  PbrtB200KdNode blank = {0u, 0u};
  nodes.push_back(blank);
  if (nPrims <= maxPrims || depth == 0) {
    initLeaf(nodeNum, primNums, nPrims);
    return;
  }
  // choose split: surface-area heuristic over all bounding-box edges
  int bestAxis = -1, bestOffset = -1;
  float bestCost = INFINITY;
  float oldCost = isectCost * float(nPrims);
  Vector d = nodeBounds.pMax - nodeBounds.pMin;
  float totalSA = (2.f * (d.x * d.y + d.x * d.z + d.y * d.z));
  float invTotalSA = 1.f / totalSA;
  int axis;
  if (d.x > d.y && d.x > d.z) axis = 0;
  else axis = (d.y > d.z) ? 1 : 2;
  int retries = 0;
retrySplit:
  for (int i = 0; i < nPrims; ++i) {
    int pn = primNums[i];
    const BBox &bbox = allPrimBounds[pn];
    edges[axis][2 * i] = BoundEdge(bbox.pMin[axis], pn, true);
    edges[axis][2 * i + 1] = BoundEdge(bbox.pMax[axis], pn, false);
  }
  std::sort(&edges[axis][0], &edges[axis][2 * nPrims]);
  int nBelow = 0, nAbove = nPrims;
  for (int i = 0; i < 2 * nPrims; ++i) {
    if (edges[axis][i].type == BoundEdge::END) --nAbove;
    float edget = edges[axis][i].t;
    if (edget > nodeBounds.pMin[axis] && edget < nodeBounds.pMax[axis]) {
      static const int otherAxis[3][2] = {{1, 2}, {0, 2}, {0, 1}};
      int otherAxis0 = otherAxis[axis][0];
      int otherAxis1 = otherAxis[axis][1];
      float belowSA = 2 * (d[otherAxis0] * d[otherAxis1] +
                           (edget - nodeBounds.pMin[axis]) * (d[otherAxis0] + d[otherAxis1]));
      float aboveSA = 2 * (d[otherAxis0] * d[otherAxis1] +
                           (nodeBounds.pMax[axis] - edget) * (d[otherAxis0] + d[otherAxis1]));
      float pBelow = belowSA * invTotalSA;
      float pAbove = aboveSA * invTotalSA;
      float eb = (nAbove == 0 || nBelow == 0) ? emptyBonus : 0.f;
      float cost = traversalCost + isectCost * (1.f - eb) * (pBelow * nBelow + pAbove * nAbove);
      if (cost < bestCost) {
        bestCost = cost;
        bestAxis = axis;
        bestOffset = i;
      }
    }
    if (edges[axis][i].type == BoundEdge::START) ++nBelow;
  }
  if (bestAxis == -1 && retries < 2) {
    ++retries;
    axis = (axis + 1) % 3;
    goto retrySplit;
  }
  if (bestCost > oldCost) ++badRefines;
  if ((bestCost > 4.f * oldCost && nPrims < 16) || bestAxis == -1 || badRefines == 3) {
    initLeaf(nodeNum, primNums, nPrims);
    return;
  }
  int n0 = 0, n1 = 0;
  for (int i = 0; i < bestOffset; ++i)
    if (edges[bestAxis][i].type == BoundEdge::START) prims0[n0++] = edges[bestAxis][i].primNum;
  for (int i = bestOffset + 1; i < 2 * nPrims; ++i)
    if (edges[bestAxis][i].type == BoundEdge::END) prims1[n1++] = edges[bestAxis][i].primNum;
  float tsplit = edges[bestAxis][bestOffset].t;
  // interior node: split float with the axis in its two low mantissa bits
  // (kdtree.cpp:66-72: `split = s; flags &= ~3; flags |= axis;`)
  {
    uint32_t bits;
    memcpy(&bits, &tsplit, 4);
    nodes[nodeNum].word0 = (bits & ~3u) | (uint32_t)bestAxis;
  }
  BBox bounds0 = nodeBounds, bounds1 = nodeBounds;
  bounds0.pMax[bestAxis] = bounds1.pMin[bestAxis] = tsplit;
  buildTree(nodeNum + 1, bounds0, allPrimBounds, prims0, n0, depth - 1, edges, prims0,
            prims1 + nPrims, badRefines);
  nodes[nodeNum].word1 = (uint32_t)nodes.size();
  buildTree((int)nodes[nodeNum].word1, bounds1, allPrimBounds, prims1, n1, depth - 1, edges,
            prims0, prims1 + nPrims, badRefines);
}
