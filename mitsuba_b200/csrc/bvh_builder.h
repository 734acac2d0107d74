// Host-side BVH2 construction (binned SAH, multi-threaded).  Replaces GenericKDTree::buildInternal
// (include/mitsuba/render/gkdtree.h:958-1263) for the device: any structure is admissible as long as a
// query returns the same argmin-t TriAccel hit.
#pragma once
#include <stdint.h>
#include <vector>
#include "b2_types.h"

namespace b2 {

struct PrimBox {
    float lo[3], hi[3];
};

struct BVHResult {
    std::vector<BVHNode> nodes;      // BFS order: the head of the array is the top of the tree (TMA staging)
    std::vector<uint32_t> leafPrims; // leaf-ordered prim ids
    int32_t rootRef = -1;
    int depth = 0;
    // wide tree over the same leaves (only when requested): 8-wide nodes with quantised child boxes, BFS order (the internal children of a
    // node are consecutive); its leaves index the SAME leaf-ordered triangle array (leafPrims is then emitted in the wide tree's order)
    std::vector<BVH8Node> nodes8;
    int depth8 = 0;
};

// boxes: one per candidate prim (ids[i] is its global prim id).  maxLeaf <= 7 (<= 3 with `wide`), maxDepth <= B2_STACK_DEPTH.
// wide: also collapse the binary tree into the 8-wide compressed tree (BVHResult::nodes8).
void buildBVH(const std::vector<PrimBox> &boxes, const std::vector<uint32_t> &ids, int maxLeaf, int maxDepth, int threads, BVHResult &out, bool wide = false);

} // namespace b2
