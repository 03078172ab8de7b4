// pb_treelet.h -- device-side node order of the top-level kd-tree (host code, used by pbrt_b200_scene_create).
//
// The ABI carries the reference's array: depth first, the below child right after its parent, the above child's
// index stored (accelerators/kdtree.cpp:172-312).  The walk (pb_kd_warp.cuh) fetches one 8-byte node per interior
// step and waits for it -- round-2 ncu on C4 / C5: that wait is the largest stall (17-24 % of the samples), L1 hit
// rate 47-66 %, because a descent in depth-first order changes cache line at every step that takes the above
// child.  The device copy is therefore re-ordered; the tree, every node's content and the sequence of nodes a ray
// visits are unchanged, only where a node lies:
//   * siblings are adjacent, so an interior node stores the index of its below child and the above child is the
//     next entry (word1 = below, above = below + 1);
//   * three levels of sibling pairs (a pair, the pairs under its two nodes, the pairs under those: at most 7 pairs
//     = 14 nodes = 112 bytes) form a treelet that never straddles a 128-byte line, so three consecutive steps of a
//     descent read one line;
//   * treelets are written breadth first, which keeps the top of the tree -- shared by every ray -- contiguous.
// Lines are packed first-fit in that order (a treelet that does not fit the rest of the current line starts the
// next one); unused entries are empty leaves nothing points to.
//
// MEASURED (profiles/r02g_sweep_walk_variants.log, rows "[default]" = treelets, "[9=0]" = depth first): no gain on
// any configuration -- C4 248.6 vs 247.4 ms, C5 1587.8 vs 1591.1 ms per frame.  The wait after a node fetch is the
// latency of the slowest of a warp's ~14 concurrent fetches into an L2-resident tree, which a better line layout
// does not shorten; the L1 misses come from the per-lane todo lists and primitive data.  Kept behind tunable 9
// (default off) with its CPU test (tests/test_oracle.py::test_treelet_node_order_changes_no_hit_and_no_counter).
#ifndef PB_TREELET_H
#define PB_TREELET_H
#include <stdint.h>
#include <vector>
#include "../../include/pbrt_b200.h"

static void pb_treelet_order(const PbrtB200KdNode *nodes, uint32_t nNodes, std::vector<PbrtB200KdNode> *out) {
  const uint32_t NONE = 0xffffffffu;
  const uint32_t LINE = 16;   // nodes per 128-byte line
  std::vector<uint32_t> newIndex(nNodes, NONE);
  // queue of treelet roots: sibling pairs by old index (the tree's root has no sibling)
  std::vector<uint32_t> qa, qb;
  qa.reserve(nNodes / 8 + 16); qb.reserve(nNodes / 8 + 16);
  qa.push_back(0); qb.push_back(NONE);
  uint32_t cursor = 0;
  uint32_t pa[8], pb[8];
  auto interior = [&](uint32_t n) { return n != NONE && (nodes[n].word0 & 3u) != 3u; };
  for (size_t head = 0; head < qa.size(); ++head) {
    int nPairs = 1, levelStart = 0, levelEnd = 1;
    pa[0] = qa[head]; pb[0] = qb[head];
    for (int depth = 1; depth < 3; ++depth) {
      for (int i = levelStart; i < levelEnd; ++i) {
        const uint32_t two[2] = {pa[i], pb[i]};
        for (int k = 0; k < 2; ++k)
          if (interior(two[k])) { pa[nPairs] = two[k] + 1u; pb[nPairs] = nodes[two[k]].word1; ++nPairs; }
      }
      levelStart = levelEnd; levelEnd = nPairs;
    }
    for (int i = levelStart; i < levelEnd; ++i) {
      const uint32_t two[2] = {pa[i], pb[i]};
      for (int k = 0; k < 2; ++k)
        if (interior(two[k])) { qa.push_back(two[k] + 1u); qb.push_back(nodes[two[k]].word1); }
    }
    const uint32_t size = 2u * (uint32_t)nPairs;
    if (cursor % LINE + size > LINE) cursor = (cursor / LINE + 1u) * LINE;
    for (int i = 0; i < nPairs; ++i) {
      newIndex[pa[i]] = cursor + 2u * (uint32_t)i;
      if (pb[i] != NONE) newIndex[pb[i]] = cursor + 2u * (uint32_t)i + 1u;
    }
    cursor += size;
  }
  PbrtB200KdNode emptyLeaf = {3u, 0u};
  out->assign(cursor, emptyLeaf);
  for (uint32_t i = 0; i < nNodes; ++i) {
    if (newIndex[i] == NONE) continue;   // (unreachable entry: cannot happen for a tree built by the front end)
    PbrtB200KdNode n = nodes[i];
    if ((n.word0 & 3u) != 3u) n.word1 = newIndex[i + 1u];
    (*out)[newIndex[i]] = n;
  }
}
#endif
