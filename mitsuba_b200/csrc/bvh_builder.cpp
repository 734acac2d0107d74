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

// ---- 8-wide collapse ------------------------------------------------------------------------------------------------------------
// Every wide node takes a binary node and opens its largest (surface area) internal child again and again until eight children are
// reached; children go to the slots whose octant direction fits their position best (greedy over the 8 x 8 scores); boxes are quantised
// to 8 bits per axis relative to the node's padded box.  Wide nodes are emitted breadth first so that the internal children of a node are
// consecutive; the triangles are emitted node by node (every binary leaf stays contiguous), which defines the leaf order of BOTH trees.
struct WideBuild {
    const std::vector<TmpNode> &tmp;
    float tiny;
    std::vector<BVH8Node> nodes;
    std::vector<uint32_t> leafStart;   // per TmpNode leaf: its first position in the new triangle order
    std::vector<uint32_t> newOrder;    // positions of the old permutation in emission order
    int depth = 0;

    WideBuild(const std::vector<TmpNode> &t, float tn) : tmp(t), tiny(tn), leafStart(t.size(), 0xFFFFFFFFu) {}

    void run(int root) {
        struct Item { int tnode; uint32_t index; int depth; };
        std::queue<Item> q;
        nodes.emplace_back();
        q.push({root, 0u, 1});
        while (!q.empty()) {
            const Item it = q.front(); q.pop();
            depth = std::max(depth, it.depth);
            int ch[8], n = 0;
            ch[n++] = tmp[it.tnode].left; ch[n++] = tmp[it.tnode].right;
            while (n < 8) {
                int best = -1;
                float bestArea = -1.0f;
                for (int k = 0; k < n; ++k)
                    if (tmp[ch[k]].left >= 0 && tmp[ch[k]].box.area() > bestArea) { bestArea = tmp[ch[k]].box.area(); best = k; }
                if (best < 0) break;
                const int t = ch[best];
                ch[best] = tmp[t].left; ch[n++] = tmp[t].right;
            }
            // padded node box and quantisation grid
            Box nb; nb.reset();
            float clo[8][3], chi[8][3];
            for (int k = 0; k < n; ++k) { padBox(tmp[ch[k]].box, tiny, clo[k], chi[k]); nb.grow(clo[k], chi[k]); }
            BVH8Node nd;
            memset(&nd, 0, sizeof(nd));
            double scale[3];
            for (int a = 0; a < 3; ++a) {
                nd.p[a] = nb.lo[a];
                const double ext = (double) nb.hi[a] - (double) nb.lo[a];
                int e = ext > 0 ? (int) std::ceil(std::log2(ext / 255.0)) : -100;
                e = std::max(-100, std::min(100, e));
                while (std::ldexp(255.0, e) < ext && e < 100) ++e;
                nd.e[a] = (int8_t) e;
                scale[a] = std::ldexp(1.0, e);
            }
            // slot assignment: greedy on dot(child centre - node centre, octant direction of the slot)
            int slotOf[8], used = 0;
            bool done[8] = {false, false, false, false, false, false, false, false};
            float score[8][8];
            for (int k = 0; k < n; ++k)
                for (int s = 0; s < 8; ++s) {
                    float v = 0;
                    for (int a = 0; a < 3; ++a) {
                        const float c = 0.5f * (tmp[ch[k]].box.lo[a] + tmp[ch[k]].box.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
                        v += ((s >> a) & 1) ? c : -c;
                    }
                    score[k][s] = v;
                }
            for (int round = 0; round < n; ++round) {
                int bk = -1, bs = -1;
                float bv = -std::numeric_limits<float>::infinity();
                for (int k = 0; k < n; ++k) {
                    if (done[k]) continue;
                    for (int s = 0; s < 8; ++s)
                        if (!((used >> s) & 1) && score[k][s] > bv) { bv = score[k][s]; bk = k; bs = s; }
                }
                done[bk] = true; used |= 1 << bs; slotOf[bk] = bs;
            }
            int childAt[8];
            for (int s = 0; s < 8; ++s) childAt[s] = -1;
            for (int k = 0; k < n; ++k) childAt[slotOf[k]] = k;
            nd.childBase = (uint32_t) nodes.size();
            nd.triBase = (uint32_t) newOrder.size();
            uint32_t triOff = 0;
            for (int s = 0; s < 8; ++s) {
                const int k = childAt[s];
                if (k < 0) continue; // empty slot: qlo = qhi = 0 and no flag -- its test result is masked out
                const TmpNode &c = tmp[ch[k]];
                for (int a = 0; a < 3; ++a) {
                    int lo = (int) std::floor(((double) clo[k][a] - (double) nd.p[a]) / scale[a]);
                    int hi = (int) std::ceil(((double) chi[k][a] - (double) nd.p[a]) / scale[a]);
                    lo = std::max(0, std::min(255, lo)); hi = std::max(0, std::min(255, hi));
                    // the device decodes p + q * 2^e in float: keep the decoded box around the padded child box
                    while (lo > 0 && (float) ((double) nd.p[a] + lo * scale[a]) > clo[k][a]) --lo;
                    while (hi < 255 && (float) ((double) nd.p[a] + hi * scale[a]) < chi[k][a]) ++hi;
                    nd.qlo[a][s] = (uint8_t) lo; nd.qhi[a][s] = (uint8_t) hi;
                }
                if (c.left >= 0) {
                    nd.imask |= (uint8_t) (1u << s);
                    const uint32_t idx = (uint32_t) nodes.size();
                    nodes.emplace_back();
                    q.push({ch[k], idx, it.depth + 1});
                } else {
                    nd.meta[s] = (uint8_t) ((c.count << 5) | triOff);
                    leafStart[ch[k]] = (uint32_t) newOrder.size();
                    for (uint32_t i = 0; i < c.count; ++i) newOrder.push_back(c.start + i);
                    triOff += c.count;
                }
            }
            nodes[it.index] = nd;
        }
    }
};

void buildBVH(const std::vector<PrimBox> &boxes, const std::vector<uint32_t> &ids, int maxLeaf, int maxDepth, int threads, BVHResult &out, bool wide) {
    out.nodes.clear();
    out.leafPrims.clear();
    out.nodes8.clear();
    out.depth = 0; out.depth8 = 0;
    if (wide) maxLeaf = std::min(maxLeaf, 3); // a leaf child of the wide node holds at most 3 triangles
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
    // leaf prim order = permutation order (leaves cover disjoint contiguous ranges); with the wide tree: its emission order
    out.leafPrims.resize(n);
    std::vector<uint32_t> leafStart;
    if (wide && tmp[root].left >= 0) {
        WideBuild W(tmp, tiny);
        W.run(root);
        out.nodes8.swap(W.nodes);
        out.depth8 = W.depth;
        leafStart.swap(W.leafStart);
        for (uint32_t i = 0; i < n; ++i) out.leafPrims[i] = ids[B.order[W.newOrder[i]]];
    } else {
        for (uint32_t i = 0; i < n; ++i) out.leafPrims[i] = ids[B.order[i]];
    }
    auto leafRef = [&](const TmpNode &t) -> int32_t {
        const uint32_t start = leafStart.empty() ? t.start : leafStart[&t - tmp.data()];
        return (int32_t) ~(start | ((uint32_t) t.count << 28));
    };
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

// ---- host-side self test of the wide tree (no device): structure + conservativeness -----------------------------------------------
// Builds the trees over n random triangles' boxes and checks (1) every leaf position is referenced by exactly one leaf child, and the
// binary tree's leaves reference the same positions; (2) walking the wide tree with the device's arithmetic (b2_trace.cuh
// traverseQueue8: t = q * (2^e * idir) + (p * idir - o * idir), near / far planes by direction sign, tmin <= tmax * (1 + 3e-7)) reaches
// every primitive whose own box a random ray hits.  Returns 0 when all checks pass, else the number of the failed check.
extern "C" int b2_bvh_selftest(uint32_t n, uint32_t seed, uint32_t nRays) {
    using namespace b2;
    std::vector<PrimBox> boxes(n);
    std::vector<uint32_t> ids(n);
    uint32_t st = seed * 747796405u + 2891336453u;
    auto rnd = [&]() { st = st * 747796405u + 2891336453u; uint32_t w = ((st >> ((st >> 28u) + 4u)) ^ st) * 277803737u; return (float) (((w >> 22u) ^ w) >> 8) * (1.0f / 16777216.0f); };
    for (uint32_t i = 0; i < n; ++i) {
        ids[i] = i;
        float c[3] = {rnd() * 10 - 5, rnd() * 10 - 5, rnd() * 2 - 1}, r = 0.01f + 0.2f * rnd() * rnd();
        for (int a = 0; a < 3; ++a) { boxes[i].lo[a] = c[a] - r * rnd(); boxes[i].hi[a] = c[a] + r * rnd(); }
    }
    BVHResult res;
    buildBVH(boxes, ids, 4, B2_STACK_DEPTH - 2, 2, res, true);
    if (n > 3 && res.nodes8.empty()) return 1;
    if (res.nodes8.empty()) return 0; // a single leaf: no wide tree
    if (res.depth8 > B2_STACK8_DEPTH - 1) return 2;
    // (1) leaf coverage
    std::vector<int> seen(n, 0);
    for (const BVH8Node &nd : res.nodes8)
        for (int s = 0; s < 8; ++s) {
            if ((nd.imask >> s) & 1) continue;
            const uint32_t cnt = nd.meta[s] >> 5, off = nd.meta[s] & 31u;
            for (uint32_t k = 0; k < cnt; ++k) { if (nd.triBase + off + k >= n) return 3; ++seen[nd.triBase + off + k]; }
        }
    for (uint32_t i = 0; i < n; ++i) if (seen[i] != 1) return 4;
    std::vector<int> seen2(n, 0);
    auto leafMark = [&](int32_t ref) { const uint32_t bits = ~(uint32_t) ref, start = bits & 0x0FFFFFFFu, cnt = bits >> 28; for (uint32_t k = 0; k < cnt; ++k) ++seen2[start + k]; };
    for (const BVHNode &nd : res.nodes) { if (nd.left < 0) leafMark(nd.left); if (nd.right < 0) leafMark(nd.right); }
    for (uint32_t i = 0; i < n; ++i) if (seen2[i] != 1) return 5;
    { std::vector<int> perm(n, 0); for (uint32_t i = 0; i < n; ++i) { if (res.leafPrims[i] >= n) return 6; ++perm[res.leafPrims[i]]; } for (uint32_t i = 0; i < n; ++i) if (perm[i] != 1) return 7; }
    // (2) traversal emulation
    for (uint32_t r = 0; r < nRays; ++r) {
        float o[3] = {rnd() * 14 - 7, rnd() * 14 - 7, rnd() * 6 - 3}, d[3] = {rnd() * 2 - 1, rnd() * 2 - 1, rnd() * 2 - 1};
        if (r % 7 == 0) d[r % 3] = 0.0f; // axis-parallel components
        const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (!(len > 1e-3f)) continue;
        for (int a = 0; a < 3; ++a) d[a] /= len;
        float idir[3], ood[3];
        for (int a = 0; a < 3; ++a) { const float da = std::fabs(d[a]) > 1e-20f ? d[a] : std::copysign(1e-20f, d[a]); idir[a] = 1.0f / da; ood[a] = o[a] * idir[a]; }
        const float mint = 0.0f, maxt = 1e30f;
        std::vector<char> reached(n, 0);
        std::vector<uint32_t> stack; stack.push_back(0);
        while (!stack.empty()) {
            const BVH8Node &nd = res.nodes8[stack.back()]; stack.pop_back();
            float adj[3], b[3];
            for (int a = 0; a < 3; ++a) { adj[a] = std::ldexp(1.0f, nd.e[a]) * idir[a]; b[a] = std::fmaf(nd.p[a], idir[a], -ood[a]); }
            uint32_t nInner = 0;
            for (int s = 0; s < 8; ++s) {
                float tmin = mint, tmax = maxt;
                for (int a = 0; a < 3; ++a) {
                    const float qn = idir[a] >= 0 ? nd.qlo[a][s] : nd.qhi[a][s], qf = idir[a] >= 0 ? nd.qhi[a][s] : nd.qlo[a][s];
                    tmin = std::max(tmin, std::fmaf(qn, adj[a], b[a])); tmax = std::min(tmax, std::fmaf(qf, adj[a], b[a]));
                }
                const bool inner = (nd.imask >> s) & 1;
                if (tmin <= tmax * 1.0000003f) {
                    if (inner) stack.push_back(nd.childBase + nInner);
                    else for (uint32_t k = 0; k < (uint32_t) (nd.meta[s] >> 5); ++k) reached[nd.triBase + (nd.meta[s] & 31u) + k] = 1;
                }
                if (inner) ++nInner;
            }
        }
        for (uint32_t i = 0; i < n; ++i) { // exact slab test of the primitive's own box in double
            const PrimBox &pb = boxes[res.leafPrims[i]];
            double t0 = 0, t1 = 1e30;
            bool hit = true;
            for (int a = 0; a < 3 && hit; ++a) {
                if (d[a] == 0) { if (o[a] < pb.lo[a] || o[a] > pb.hi[a]) hit = false; continue; }
                double ta = ((double) pb.lo[a] - o[a]) / d[a], tb = ((double) pb.hi[a] - o[a]) / d[a];
                if (ta > tb) std::swap(ta, tb);
                t0 = std::max(t0, ta); t1 = std::min(t1, tb);
                if (t0 > t1) hit = false;
            }
            if (hit && !reached[i]) return 8;
        }
    }
    return 0;
}
