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
};

// boxes: one per candidate prim (ids[i] is its global prim id).  maxLeaf <= 7, maxDepth <= B2_STACK_DEPTH.
void buildBVH(const std::vector<PrimBox> &boxes, const std::vector<uint32_t> &ids, int maxLeaf, int maxDepth, int threads, BVHResult &out);

} // namespace b2
