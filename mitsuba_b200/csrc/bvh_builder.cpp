// Binned-SAH BVH2 builder (16 bins, top-down, subtrees built by worker threads), BFS re-layout.
#include "bvh_builder.h"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <future>
#include <limits>
#include <queue>
#include <thread>

namespace b2 {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int i = 0; i < 3; ++i) { lo[i] = std::numeric_limits<float>::infinity(); hi[i] = -std::numeric_limits<float>::infinity(); }
    }
    void grow(const float *l, const float *h) {
        for (int i = 0; i < 3; ++i) { lo[i] = std::min(lo[i], l[i]); hi[i] = std::max(hi[i], h[i]); }
    }
    void growPt(const float *p) {
        for (int i = 0; i < 3; ++i) { lo[i] = std::min(lo[i], p[i]); hi[i] = std::max(hi[i], p[i]); }
    }
    float area() const {
        float d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        if (d[0] < 0) return 0;
        return 2.0f * (d[0] * d[1] + d[1] * d[2] + d[0] * d[2]);
    }
};

struct TmpNode {
    Box box;
    int left = -1, right = -1; // children (TmpNode indices) or -1 for leaf
    uint32_t start = 0, count = 0;
    int depth = 0;
};

struct Builder {
    const std::vector<PrimBox> &boxes;
    std::vector<uint32_t> order; // permutation of [0, n) partitioned in place
    std::vector<float> cx, cy, cz;
    int maxLeaf, maxDepth;
    std::atomic<int> threadsLeft;

    Builder(const std::vector<PrimBox> &b, int ml, int md, int threads) : boxes(b), maxLeaf(ml), maxDepth(md), threadsLeft(threads - 1) {
        size_t n = b.size();
        order.resize(n);
        cx.resize(n); cy.resize(n); cz.resize(n);
        for (size_t i = 0; i < n; ++i) {
            order[i] = (uint32_t) i;
            cx[i] = 0.5f * (b[i].lo[0] + b[i].hi[0]);
            cy[i] = 0.5f * (b[i].lo[1] + b[i].hi[1]);
            cz[i] = 0.5f * (b[i].lo[2] + b[i].hi[2]);
        }
    }
    float centroid(uint32_t p, int axis) const { return axis == 0 ? cx[p] : (axis == 1 ? cy[p] : cz[p]); }

    // builds the subtree over order[start, start+count) into `nodes` (local vector), returns local root index
    int build(std::vector<TmpNode> &nodes, uint32_t start, uint32_t count, int depth) {
        int me = (int) nodes.size();
        nodes.emplace_back();
        Box box, cbox;
        box.reset(); cbox.reset();
        for (uint32_t i = start; i < start + count; ++i) {
            uint32_t p = order[i];
            box.grow(boxes[p].lo, boxes[p].hi);
            float c[3] = {cx[p], cy[p], cz[p]};
            cbox.growPt(c);
        }
        nodes[me].box = box;
        nodes[me].depth = depth;
        nodes[me].start = start;
        nodes[me].count = count;
        if (count == 1) return me;
        const bool canLeaf = (int) count <= maxLeaf;
        // depth cap: if the remaining levels can only just hold a balanced tree, split at the object median
        int remaining = maxDepth - depth;
        int need = 0;
        { uint32_t leaves = (count + (uint32_t) maxLeaf - 1) / (uint32_t) maxLeaf; while ((1u << need) < leaves) ++need; }
        bool forceMedian = need + 1 >= remaining;
        if (canLeaf && (forceMedian || remaining <= 1)) return me;
        int axis = 0;
        float ext[3] = {cbox.hi[0] - cbox.lo[0], cbox.hi[1] - cbox.lo[1], cbox.hi[2] - cbox.lo[2]};
        if (ext[1] > ext[axis]) axis = 1;
        if (ext[2] > ext[axis]) axis = 2;
        uint32_t mid = start + count / 2;
        bool done = false;
        if (!forceMedian && ext[axis] > 0) {
            const int NB = 16;
            float bestCost = std::numeric_limits<float>::infinity();
            int bestAxis = -1, bestBin = -1;
            for (int a = 0; a < 3; ++a) {
                if (!(ext[a] > 0)) continue;
                Box bb[NB];
                uint32_t bc[NB];
                for (int k = 0; k < NB; ++k) { bb[k].reset(); bc[k] = 0; }
                float scale = NB / ext[a];
                for (uint32_t i = start; i < start + count; ++i) {
                    uint32_t p = order[i];
                    int k = (int) ((centroid(p, a) - cbox.lo[a]) * scale);
                    k = std::min(std::max(k, 0), NB - 1);
                    bb[k].grow(boxes[p].lo, boxes[p].hi);
                    bc[k]++;
                }
                float rightArea[NB];
                uint32_t rightCount[NB];
                Box acc;
                acc.reset();
                uint32_t cnt = 0;
                for (int k = NB - 1; k > 0; --k) {
                    if (bc[k]) acc.grow(bb[k].lo, bb[k].hi);
                    cnt += bc[k];
                    rightArea[k] = acc.area();
                    rightCount[k] = cnt;
                }
                acc.reset();
                cnt = 0;
                for (int k = 0; k < NB - 1; ++k) {
                    if (bc[k]) acc.grow(bb[k].lo, bb[k].hi);
                    cnt += bc[k];
                    if (cnt == 0 || rightCount[k + 1] == 0) continue;
                    float cost = acc.area() * cnt + rightArea[k + 1] * rightCount[k + 1];
                    if (cost < bestCost) { bestCost = cost; bestAxis = a; bestBin = k; }
                }
            }
            if (bestAxis >= 0) {
                float leafCost = box.area() * count;
                float splitCost = 1.0f * box.area() + bestCost; // traversal cost 1, intersection cost 1
                if (splitCost < leafCost || !canLeaf) {
                    float scale = NB / ext[bestAxis];
                    auto it = std::partition(order.begin() + start, order.begin() + start + count, [&](uint32_t p) {
                        int k = (int) ((centroid(p, bestAxis) - cbox.lo[bestAxis]) * scale);
                        k = std::min(std::max(k, 0), NB - 1);
                        return k <= bestBin;
                    });
                    mid = (uint32_t) (it - order.begin());
                    done = mid > start && mid < start + count;
                }
            }
        }
        if (!done && canLeaf) return me;
        if (!done) {
            std::nth_element(order.begin() + start, order.begin() + start + count / 2, order.begin() + start + count,
                             [&](uint32_t a, uint32_t b) { return centroid(a, axis) < centroid(b, axis); });
            mid = start + count / 2;
        }
        uint32_t lc = mid - start, rc = count - lc;
        // large subtrees: build the left child on another thread
        int l, r;
        if (lc > 200000 && threadsLeft.fetch_sub(1) > 0) {
            std::vector<TmpNode> sub;
            auto fut = std::async(std::launch::async, [&]() { return build(sub, start, lc, depth + 1); });
            r = build(nodes, mid, rc, depth + 1);
            int subRoot = fut.get();
            threadsLeft.fetch_add(1);
            int offset = (int) nodes.size();
            for (auto &t : sub) {
                if (t.left >= 0) { t.left += offset; t.right += offset; }
                nodes.push_back(t);
            }
            l = subRoot + offset;
        } else {
            if (lc > 200000) threadsLeft.fetch_add(1);
            l = build(nodes, start, lc, depth + 1);
            r = build(nodes, mid, rc, depth + 1);
        }
        nodes[me].left = l;
        nodes[me].right = r;
        return me;
    }
};

inline void padBox(const Box &b, float tiny, float *lo, float *hi) {
    for (int i = 0; i < 3; ++i) {
        lo[i] = b.lo[i] - (std::fabs(b.lo[i]) * 4e-7f + tiny);
        hi[i] = b.hi[i] + (std::fabs(b.hi[i]) * 4e-7f + tiny);
    }
}

} // namespace

void buildBVH(const std::vector<PrimBox> &boxes, const std::vector<uint32_t> &ids, int maxLeaf, int maxDepth, int threads, BVHResult &out) {
    out.nodes.clear();
    out.leafPrims.clear();
    out.depth = 0;
    const uint32_t n = (uint32_t) boxes.size();
    if (n == 0) { out.rootRef = -1; return; } // leaf with count 0
    Builder B(boxes, maxLeaf, maxDepth, std::max(1, threads));
    std::vector<TmpNode> tmp;
    tmp.reserve(2 * (size_t) n / std::max(1, maxLeaf) + 16);
    int root = B.build(tmp, 0, n, 0);
    // scene scale for the padding
    float diag = 0;
    for (int i = 0; i < 3; ++i) diag = std::max(diag, tmp[root].box.hi[i] - tmp[root].box.lo[i]);
    const float tiny = diag * 1e-7f + 1e-30f;
    // leaf prim order = permutation order (leaves cover disjoint contiguous ranges)
    out.leafPrims.resize(n);
    for (uint32_t i = 0; i < n; ++i) out.leafPrims[i] = ids[B.order[i]];
    auto leafRef = [&](const TmpNode &t) -> int32_t { return (int32_t) ~((uint32_t) t.start | ((uint32_t) t.count << 28)); };
    if (tmp[root].left < 0) { out.rootRef = leafRef(tmp[root]); out.depth = 1; return; }
    // BFS relayout of inner nodes
    std::vector<int> innerIndex(tmp.size(), -1);
    std::vector<int> bfs;
    bfs.reserve(tmp.size());
    {
        std::queue<int> q;
        q.push(root);
        while (!q.empty()) {
            int t = q.front(); q.pop();
            innerIndex[t] = (int) bfs.size();
            bfs.push_back(t);
            if (tmp[tmp[t].left].left >= 0) q.push(tmp[t].left);
            if (tmp[tmp[t].right].left >= 0) q.push(tmp[t].right);
        }
    }
    out.nodes.resize(bfs.size());
    int maxd = 0;
    for (size_t k = 0; k < bfs.size(); ++k) {
        const TmpNode &t = tmp[bfs[k]];
        const TmpNode &L = tmp[t.left], &R = tmp[t.right];
        BVHNode &nd = out.nodes[k];
        padBox(L.box, tiny, nd.lmin, nd.lmax);
        padBox(R.box, tiny, nd.rmin, nd.rmax);
        nd.left = L.left >= 0 ? innerIndex[t.left] : leafRef(L);
        nd.right = R.left >= 0 ? innerIndex[t.right] : leafRef(R);
        nd.pad0 = nd.pad1 = 0;
        maxd = std::max(maxd, t.depth + 2);
    }
    out.rootRef = 0;
    out.depth = maxd;
}

} // namespace b2
