// Binned-SAH BVH2 builder (16 bins, top-down, subtrees built by worker threads), BFS re-layout.
#include "bvh_builder.h"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <chrono>
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

// Worker threads the top of the tree can borrow: a node with many primitives splits its reductions (bounds, bins) and its partition over
// them; further down whole subtrees go to their own thread.  All reductions are min / max / integer counts and the partition is stable,
// so the tree does not depend on the number of threads (b2_bvh_selftest compares 1 against many).
struct Workers {
    std::atomic<int> idle;
    explicit Workers(int threads) : idle(threads - 1) {}
    int borrow(int want) { // up to `want` extra threads
        int got = 0;
        while (got < want) {
            int cur = idle.load();
            if (cur <= 0) break;
            if (idle.compare_exchange_weak(cur, cur - 1)) ++got;
        }
        return got;
    }
    void giveBack(int n) { idle.fetch_add(n); }
};

// run f(chunkIndex, begin, end) over [0, count) cut into `parts` contiguous chunks, parts - 1 of them on new threads
template <typename F> void forChunks(uint32_t count, int parts, F f) {
    if (parts <= 1) { f(0, 0u, count); return; }
    std::vector<std::thread> th;
    th.reserve(parts > 1 ? parts - 1 : 0);
    for (int c = 1; c < parts; ++c)
        th.emplace_back([=]() { f(c, (uint32_t) ((uint64_t) count * c / parts), (uint32_t) ((uint64_t) count * (c + 1) / parts)); });
    f(0, 0u, (uint32_t) ((uint64_t) count / parts));
    for (auto &t : th) t.join();
}

struct Builder {
    // the primitives travel with the permutation (box + id, 28 bytes): every pass over a node streams its range instead of gathering
    // boxes through an index (the gathers were DRAM-latency bound: 57 ns per primitive and level on 2.5 M triangles)
    struct Ref { float lo[3], hi[3]; uint32_t id; };
    const std::vector<PrimBox> &boxes;
    std::vector<Ref> refs;      // partitioned in place
    std::vector<Ref> scratch;   // the right-hand side of a partition, same index range as the node (disjoint between subtrees)
    std::vector<uint32_t> order; // refs[i].id after the build: permutation of [0, n)
    int maxLeaf, maxDepth;
    Workers workers;
    int totalThreads;
    static const uint32_t kParallelNode = 1u << 17; // nodes with at least this many primitives are worth splitting over threads
    static const int NB = 16;

    Builder(const std::vector<PrimBox> &b, int ml, int md, int threads) : boxes(b), maxLeaf(ml), maxDepth(md), workers(threads), totalThreads(threads) {
        const size_t n = b.size();
        refs.resize(n); scratch.resize(n);
        const int parts = n >= kParallelNode ? 1 + workers.borrow(std::min<int>(threads - 1, (int) (n / kParallelNode))) : 1;
        forChunks((uint32_t) n, parts, [&](int, uint32_t lo, uint32_t hi) {
            for (uint32_t i = lo; i < hi; ++i) {
                memcpy(refs[i].lo, b[i].lo, 12); memcpy(refs[i].hi, b[i].hi, 12);
                refs[i].id = i;
            }
        });
        workers.giveBack(parts - 1);
    }
    static float centroid(const Ref &r, int axis) { return 0.5f * (r.lo[axis] + r.hi[axis]); }
    void finish() { // the permutation the callers read
        order.resize(refs.size());
        for (size_t i = 0; i < refs.size(); ++i) order[i] = refs[i].id;
        std::vector<Ref>().swap(scratch);
    }
    static int binOf(float c, float lo, float scale) { return std::min(std::max((int) ((c - lo) * scale), 0), NB - 1); }

    struct Bins { Box bb[3][NB]; uint32_t bc[3][NB]; };

    // builds the subtree over order[start, start+count) into `nodes` (local vector), returns local root index
    int build(std::vector<TmpNode> &nodes, uint32_t start, uint32_t count, int depth) {
        int me = (int) nodes.size();
        nodes.emplace_back();
        // threads for the passes over this node's primitives
        int parts = 1;
        if (count >= kParallelNode) parts += workers.borrow((int) (((uint64_t) totalThreads * count + boxes.size() - 1) / boxes.size()) - 1); // a share in proportion to the node's size
        Box box, cbox;
        box.reset(); cbox.reset();
        {
            Box pb1[2];
            std::vector<Box> pbN;
            if (parts > 1) pbN.resize(2 * (size_t) parts);
            Box *pb = parts > 1 ? pbN.data() : pb1, *pc = pb + parts; // no heap traffic for the millions of small nodes
            forChunks(count, parts, [&](int c, uint32_t lo, uint32_t hi) {
                Box b1, c1;
                b1.reset(); c1.reset();
                for (uint32_t i = start + lo; i < start + hi; ++i) {
                    const Ref &p = refs[i];
                    b1.grow(p.lo, p.hi);
                    const float c3[3] = {centroid(p, 0), centroid(p, 1), centroid(p, 2)};
                    c1.growPt(c3);
                }
                pb[c] = b1; pc[c] = c1;
            });
            for (int c = 0; c < parts; ++c) { box.grow(pb[c].lo, pb[c].hi); cbox.grow(pc[c].lo, pc[c].hi); }
        }
        nodes[me].box = box;
        nodes[me].depth = depth;
        nodes[me].start = start;
        nodes[me].count = count;
        if (count == 1) { workers.giveBack(parts - 1); return me; }
        const bool canLeaf = (int) count <= maxLeaf;
        // depth cap: if the remaining levels can only just hold a balanced tree, split at the object median
        int remaining = maxDepth - depth;
        int need = 0;
        { uint32_t leaves = (count + (uint32_t) maxLeaf - 1) / (uint32_t) maxLeaf; while ((1u << need) < leaves) ++need; }
        bool forceMedian = need + 1 >= remaining;
        if (canLeaf && (forceMedian || remaining <= 1)) { workers.giveBack(parts - 1); return me; }
        int axis = 0;
        float ext[3] = {cbox.hi[0] - cbox.lo[0], cbox.hi[1] - cbox.lo[1], cbox.hi[2] - cbox.lo[2]};
        if (ext[1] > ext[axis]) axis = 1;
        if (ext[2] > ext[axis]) axis = 2;
        uint32_t mid = start + count / 2;
        bool done = false;
        if (!forceMedian && ext[axis] > 0) {
            float bestCost = std::numeric_limits<float>::infinity();
            int bestAxis = -1, bestBin = -1;
            float scale[3];
            for (int a = 0; a < 3; ++a) scale[a] = ext[a] > 0 ? NB / ext[a] : 0.0f;
            // one pass fills the bins of all three axes; chunks keep their own bins, merged afterwards (min / max / counts: exact)
            Bins bins1;
            std::vector<Bins> binsN;
            if (parts > 1) binsN.resize(parts);
            Bins *pbins = parts > 1 ? binsN.data() : &bins1;
            forChunks(count, parts, [&](int c, uint32_t lo, uint32_t hi) {
                Bins &bn = pbins[c];
                for (int a = 0; a < 3; ++a) for (int k = 0; k < NB; ++k) { bn.bb[a][k].reset(); bn.bc[a][k] = 0; }
                for (uint32_t i = start + lo; i < start + hi; ++i) {
                    const Ref &p = refs[i];
                    for (int a = 0; a < 3; ++a) {
                        if (!(ext[a] > 0)) continue;
                        const int k = binOf(centroid(p, a), cbox.lo[a], scale[a]);
                        bn.bb[a][k].grow(p.lo, p.hi);
                        bn.bc[a][k]++;
                    }
                }
            });
            Bins &bins = pbins[0];
            for (int c = 1; c < parts; ++c)
                for (int a = 0; a < 3; ++a) for (int k = 0; k < NB; ++k) {
                    if (pbins[c].bc[a][k]) bins.bb[a][k].grow(pbins[c].bb[a][k].lo, pbins[c].bb[a][k].hi);
                    bins.bc[a][k] += pbins[c].bc[a][k];
                }
            for (int a = 0; a < 3; ++a) {
                if (!(ext[a] > 0)) continue;
                const Box *bb = bins.bb[a];
                const uint32_t *bc = bins.bc[a];
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
                    const float sc = scale[bestAxis], clo = cbox.lo[bestAxis];
                    auto goesLeft = [&](const Ref &p) { return binOf(centroid(p, bestAxis), clo, sc) <= bestBin; };
                    // stable partition: left-hand primitives keep their order in place, right-hand ones pass through `scratch`
                    uint32_t totalLeft = 0;
                    if (parts == 1) {
                        uint32_t l = start, r = 0;
                        for (uint32_t i = start; i < start + count; ++i) {
                            const Ref p = refs[i];
                            if (goesLeft(p)) refs[l++] = p; else scratch[start + r++] = p;
                        }
                        memcpy(refs.data() + l, scratch.data() + start, (size_t) r * sizeof(Ref));
                        totalLeft = l - start;
                    } else {
                        std::vector<uint32_t> nLeft(parts);
                        forChunks(count, parts, [&](int c, uint32_t lo, uint32_t hi) {
                            uint32_t n = 0;
                            for (uint32_t i = start + lo; i < start + hi; ++i) n += goesLeft(refs[i]) ? 1u : 0u;
                            nLeft[c] = n;
                        });
                        for (int c = 0; c < parts; ++c) totalLeft += nLeft[c];
                        // chunk c writes its left-hand elements at leftBase[c] and its right-hand ones at totalLeft + rightBase[c] of `scratch`
                        std::vector<uint32_t> leftBase(parts), rightBase(parts);
                        uint32_t lb = 0, rb = 0;
                        for (int c = 0; c < parts; ++c) {
                            leftBase[c] = lb; rightBase[c] = rb;
                            const uint32_t lo = (uint32_t) ((uint64_t) count * c / parts), hi = (uint32_t) ((uint64_t) count * (c + 1) / parts);
                            lb += nLeft[c]; rb += (hi - lo) - nLeft[c];
                        }
                        forChunks(count, parts, [&](int c, uint32_t lo, uint32_t hi) {
                            uint32_t l = start + leftBase[c], r = start + totalLeft + rightBase[c];
                            for (uint32_t i = start + lo; i < start + hi; ++i) {
                                const Ref &p = refs[i];
                                if (goesLeft(p)) scratch[l++] = p; else scratch[r++] = p;
                            }
                        });
                        forChunks(count, parts, [&](int, uint32_t lo, uint32_t hi) {
                            memcpy(refs.data() + start + lo, scratch.data() + start + lo, (size_t) (hi - lo) * sizeof(Ref));
                        });
                    }
                    mid = start + totalLeft;
                    done = mid > start && mid < start + count;
                }
            }
        }
        workers.giveBack(parts - 1);
        if (!done && canLeaf) return me;
        if (!done) {
            std::nth_element(refs.begin() + start, refs.begin() + start + count / 2, refs.begin() + start + count,
                             [&](const Ref &a, const Ref &b) { const float ca = centroid(a, axis), cb = centroid(b, axis); return ca < cb || (ca == cb && a.id < b.id); });
            mid = start + count / 2;
        }
        uint32_t lc = mid - start, rc = count - lc;
        // large subtrees: build the left child on another thread
        int l, r;
        if (lc > 50000 && rc > 50000 && workers.borrow(1) == 1) {
            std::vector<TmpNode> sub;
            std::thread th([&]() { l = build(sub, start, lc, depth + 1); });
            r = build(nodes, mid, rc, depth + 1);
            th.join();
            workers.giveBack(1);
            int offset = (int) nodes.size();
            for (auto &t : sub) {
                if (t.left >= 0) { t.left += offset; t.right += offset; }
                nodes.push_back(t);
            }
            l += offset;
        } else {
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

    // what the structural pass decides per wide node; the quantisation pass fills the node from it
    struct Plan { int child[8]; uint32_t childBase, triBase; }; // child[s]: TmpNode in slot s or -1

    // Pass 1 (serial, breadth first: it assigns the consecutive child indices and the triangle order): collapse, slot assignment, numbering.
    // Pass 2 (parallel over the nodes): node box, quantisation grid, quantised child boxes, flags.
    void run(int root, int threads = 1) {
        struct Item { int tnode; uint32_t index; int depth; };
        std::queue<Item> q;
        std::vector<Plan> plans;
        plans.emplace_back();
        newOrder.reserve(tmp.size());
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
            Box nb; nb.reset();
            for (int k = 0; k < n; ++k) { float clo[3], chi[3]; padBox(tmp[ch[k]].box, tiny, clo, chi); nb.grow(clo, chi); }
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
            Plan pl;
            for (int s = 0; s < 8; ++s) pl.child[s] = -1;
            for (int k = 0; k < n; ++k) pl.child[slotOf[k]] = ch[k];
            pl.childBase = (uint32_t) plans.size();
            pl.triBase = (uint32_t) newOrder.size();
            for (int s = 0; s < 8; ++s) {
                if (pl.child[s] < 0) continue;
                const TmpNode &c = tmp[pl.child[s]];
                if (c.left >= 0) {
                    const uint32_t idx = (uint32_t) plans.size();
                    plans.emplace_back();
                    q.push({pl.child[s], idx, it.depth + 1});
                } else {
                    leafStart[pl.child[s]] = (uint32_t) newOrder.size();
                    for (uint32_t i = 0; i < c.count; ++i) newOrder.push_back(c.start + i);
                }
            }
            plans[it.index] = pl;
        }
        nodes.resize(plans.size());
        const int parts = plans.size() >= 4096 ? std::max(1, threads) : 1;
        forChunks((uint32_t) plans.size(), parts, [&](int, uint32_t lo, uint32_t hi) { for (uint32_t i = lo; i < hi; ++i) quantise(plans[i], nodes[i]); });
    }

    void quantise(const Plan &pl, BVH8Node &out) const {
        Box nb; nb.reset();
        float clo[8][3], chi[8][3];
        for (int s = 0; s < 8; ++s) if (pl.child[s] >= 0) { padBox(tmp[pl.child[s]].box, tiny, clo[s], chi[s]); nb.grow(clo[s], chi[s]); }
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
        nd.childBase = pl.childBase;
        nd.triBase = pl.triBase;
        uint32_t triOff = 0;
        for (int s = 0; s < 8; ++s) {
            if (pl.child[s] < 0) continue; // empty slot: qlo = qhi = 0 and no flag -- its test result is masked out
            const TmpNode &c = tmp[pl.child[s]];
            for (int a = 0; a < 3; ++a) {
                int lo = (int) std::floor(((double) clo[s][a] - (double) nd.p[a]) / scale[a]);
                int hi = (int) std::ceil(((double) chi[s][a] - (double) nd.p[a]) / scale[a]);
                lo = std::max(0, std::min(255, lo)); hi = std::max(0, std::min(255, hi));
                // the device decodes p + q * 2^e in float: keep the decoded box around the padded child box
                while (lo > 0 && (float) ((double) nd.p[a] + lo * scale[a]) > clo[s][a]) --lo;
                while (hi < 255 && (float) ((double) nd.p[a] + hi * scale[a]) < chi[s][a]) ++hi;
                nd.qlo[a][s] = (uint8_t) lo; nd.qhi[a][s] = (uint8_t) hi;
            }
            if (c.left >= 0) nd.imask |= (uint8_t) (1u << s);
            else { nd.meta[s] = (uint8_t) ((c.count << 5) | triOff); triOff += c.count; }
        }
        out = nd;
    }
};

namespace {
// B2_COMMIT_TIMING=1: phase times of the build on stderr (see b2_scene_commit)
struct BuildClock {
    bool on = getenv("B2_COMMIT_TIMING") != nullptr;
    std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[b2 commit]   bvh: %-22s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
        last = now;
    }
};
}
void buildBVH(const std::vector<PrimBox> &boxes, const std::vector<uint32_t> &ids, int maxLeaf, int maxDepth, int threads, BVHResult &out, bool wide) {
    BuildClock clk;
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
    clk.mark("centroids");
    int root = B.build(tmp, 0, n, 0);
    B.finish();
    clk.mark("binned SAH (binary)");
    // scene scale for the padding
    float diag = 0;
    for (int i = 0; i < 3; ++i) diag = std::max(diag, tmp[root].box.hi[i] - tmp[root].box.lo[i]);
    const float tiny = diag * 1e-7f + 1e-30f;
    // leaf prim order = permutation order (leaves cover disjoint contiguous ranges); with the wide tree: its emission order
    out.leafPrims.resize(n);
    std::vector<uint32_t> leafStart;
    if (wide && tmp[root].left >= 0) {
        WideBuild W(tmp, tiny);
        W.run(root, std::max(1, threads));
        out.nodes8.swap(W.nodes);
        out.depth8 = W.depth;
        leafStart.swap(W.leafStart);
        for (uint32_t i = 0; i < n; ++i) out.leafPrims[i] = ids[B.order[W.newOrder[i]]];
        clk.mark("8-wide collapse");
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
    for (size_t k = 0; k < bfs.size(); ++k) maxd = std::max(maxd, tmp[bfs[k]].depth + 2);
    forChunks((uint32_t) bfs.size(), bfs.size() >= 4096 ? std::max(1, threads) : 1, [&](int, uint32_t klo, uint32_t khi) {
        for (uint32_t k = klo; k < khi; ++k) {
            const TmpNode &t = tmp[bfs[k]];
            const TmpNode &L = tmp[t.left], &R = tmp[t.right];
            BVHNode &nd = out.nodes[k];
            padBox(L.box, tiny, nd.lmin, nd.lmax);
            padBox(R.box, tiny, nd.rmin, nd.rmax);
            nd.left = L.left >= 0 ? innerIndex[t.left] : leafRef(L);
            nd.right = R.left >= 0 ? innerIndex[t.right] : leafRef(R);
            nd.pad0 = nd.pad1 = 0;
        }
    });
    out.rootRef = 0;
    out.depth = maxd;
    clk.mark("binary relayout");
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

// Host-only timing entry (no device): builds the trees over n boxes (lo xyz, hi xyz per primitive) with `threads` threads; returns the
// number of binary inner nodes, *wideNodes = the number of 8-wide nodes.  Used by scripts/bvh_build_bench.py.
extern "C" int b2_bvh_build_only(const float *boxes6, uint32_t n, int threads, int wide, uint32_t *wideNodes) {
    using namespace b2;
    std::vector<PrimBox> boxes(n);
    std::vector<uint32_t> ids(n);
    for (uint32_t i = 0; i < n; ++i) {
        for (int a = 0; a < 3; ++a) { boxes[i].lo[a] = boxes6[6 * (size_t) i + a]; boxes[i].hi[a] = boxes6[6 * (size_t) i + 3 + a]; }
        ids[i] = i;
    }
    BVHResult res;
    buildBVH(boxes, ids, 4, 26, threads, res, wide != 0);
    if (wideNodes) *wideNodes = (uint32_t) res.nodes8.size();
    return (int) res.nodes.size();
}

// Host-only check that the build does not depend on the number of threads: the same n pseudo-random boxes built with threadsA and threadsB
// must give byte-identical binary nodes, wide nodes and leaf order.  0 = identical.
extern "C" int b2_bvh_thread_invariance(uint32_t n, uint32_t seed, int threadsA, int threadsB) {
    using namespace b2;
    std::vector<PrimBox> boxes(n);
    std::vector<uint32_t> ids(n);
    uint32_t st = seed * 747796405u + 2891336453u;
    auto rnd = [&]() { st = st * 747796405u + 2891336453u; uint32_t w = ((st >> ((st >> 28u) + 4u)) ^ st) * 277803737u; return (float) (((w >> 22u) ^ w) >> 8) * (1.0f / 16777216.0f); };
    for (uint32_t i = 0; i < n; ++i) {
        ids[i] = i;
        // clustered sizes and a sprinkling of exact duplicates (ties in every comparison the builder makes)
        float c[3] = {rnd() * 100 - 50, rnd() * 100 - 50, rnd() * 10 - 5}, r = 0.01f + 0.3f * rnd() * rnd();
        if (i > 0 && (i % 97) == 0) { boxes[i] = boxes[i - 1]; continue; }
        for (int a = 0; a < 3; ++a) { boxes[i].lo[a] = c[a] - r * rnd(); boxes[i].hi[a] = c[a] + r * rnd(); }
    }
    BVHResult a, b;
    buildBVH(boxes, ids, 4, B2_STACK_DEPTH - 2, threadsA, a, true);
    buildBVH(boxes, ids, 4, B2_STACK_DEPTH - 2, threadsB, b, true);
    if (a.nodes.size() != b.nodes.size() || a.nodes8.size() != b.nodes8.size() || a.leafPrims != b.leafPrims) return 1;
    if (a.rootRef != b.rootRef || a.depth != b.depth || a.depth8 != b.depth8) return 2;
    if (!a.nodes.empty() && memcmp(a.nodes.data(), b.nodes.data(), a.nodes.size() * sizeof(BVHNode))) return 3;
    if (!a.nodes8.empty() && memcmp(a.nodes8.data(), b.nodes8.data(), a.nodes8.size() * sizeof(BVH8Node))) return 4;
    return 0;
}
