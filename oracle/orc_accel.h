/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
 * Ray queries: TriAccel (include/mitsuba/render/triaccel.h), the three ShapeKDTree::rayIntersect
 * entry points with their adaptive epsilon (src/librender/skdtree.cpp:112-226), Havran traversal
 * (include/mitsuba/render/sahkdtree3.h:178-308) and a brute-force any-order loop used to validate
 * the tree.
 *
 * The kd-tree *builder* here is a plain exact-sweep SAH with the reference's cost constants
 * (gkdtree.h:732-744: traversal 15, query 20, emptyBonus 0.9, stopPrims 6, maxBadRefines 3,
 * maxDepth 8+1.3*log2(N) capped at 48, :986-988) but WITHOUT min-max binning, perfect splits /
 * clipping and retraction (gkdtree.h:958-1263,1792-2400): those change tree quality, not the set
 * of (t,u,v,prim) a query returns -- the returned hit is argmin_t over TriAccel tests, which is
 * tree independent except for exact-t ties (SURVEY.md Appendix A "Tie-breaking").  The mailbox
 * (sahkdtree3.h:138-152) is an optimisation with no effect on results and is omitted. */
#pragma once
#include "orc_math.h"
#include <vector>
#include <algorithm>

namespace orc {

struct Ray { /* include/mitsuba/core/ray.h:44-96 */
    V3 o; float mint; V3 d; float maxt; V3 dRcp;
    Ray() : mint(kEpsilon), maxt(kInf) {}
    Ray(const V3 &o_, const V3 &d_) : o(o_), mint(kEpsilon), d(d_), maxt(kInf) { setD(d_); }
    Ray(const V3 &o_, const V3 &d_, float mn, float mx) : o(o_), mint(mn), d(d_), maxt(mx) { setD(d_); }
    void setD(const V3 &dd) { d = dd; dRcp = V3(1.0f / dd.x, 1.0f / dd.y, 1.0f / dd.z); }
    V3 operator()(float t) const { return o + d * t; }
};

struct AABB {
    V3 min, max;
    AABB() : min(kInf), max(-kInf) {}
    void expandBy(const V3 &p) {
        for (int i = 0; i < 3; ++i) { min[i] = std::min(min[i], p[i]); max[i] = std::max(max[i], p[i]); }
    }
    float surfaceArea() const { V3 d = max - min; return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z); }
    /* include/mitsuba/core/aabb.h:308-338 */
    bool rayIntersect(const Ray &ray, float &nearT, float &farT) const {
        nearT = -kInf; farT = kInf;
        for (int i = 0; i < 3; i++) {
            const float origin = ray.o[i], minVal = min[i], maxVal = max[i];
            if (ray.d[i] == 0) {
                if (origin < minVal || origin > maxVal) return false;
            } else {
                float t1 = (minVal - origin) * ray.dRcp[i];
                float t2 = (maxVal - origin) * ray.dRcp[i];
                if (t1 > t2) std::swap(t1, t2);
                nearT = std::max(t1, nearT);
                farT = std::min(t2, farT);
                if (!(nearT <= farT)) return false;
            }
        }
        return true;
    }
};

/* include/mitsuba/render/triaccel.h:37-59 (layout), 61-94 (load), 96-158 (rayIntersect) */
struct TriAccel {
    uint32_t k; float n_u, n_v, n_d;
    float a_u, a_v, b_nu, b_nv;
    float c_nu, c_nv; uint32_t shapeIndex, primIndex;
    int load(const V3 &A, const V3 &B, const V3 &C) {
        static const int waldModulo[4] = {1, 2, 0, 1};
        V3 b = C - A, c = B - A, N = cross(c, b);
        k = 0;
        for (int j = 0; j < 3; j++)
            if (std::abs(N[j]) > std::abs(N[k])) k = j;
        uint32_t u = waldModulo[k], v = waldModulo[k + 1];
        const float n_k = N[k], denom = b[u] * c[v] - b[v] * c[u];
        if (denom == 0) { k = 3; return 1; }
        n_u = N[u] / n_k;
        n_v = N[v] / n_k;
        n_d = dot(A, N) / n_k;
        b_nu = b[u] / denom;
        b_nv = -b[v] / denom;
        a_u = A[u];
        a_v = A[v];
        c_nu = c[v] / denom;
        c_nv = -c[u] / denom;
        return 0;
    }
    bool rayIntersect(const Ray &ray, float mint, float maxt, float &u, float &v, float &t) const {
        float o_u, o_v, o_k, d_u, d_v, d_k;
        switch (k) {
            case 0: o_u = ray.o[1]; o_v = ray.o[2]; o_k = ray.o[0]; d_u = ray.d[1]; d_v = ray.d[2]; d_k = ray.d[0]; break;
            case 1: o_u = ray.o[2]; o_v = ray.o[0]; o_k = ray.o[1]; d_u = ray.d[2]; d_v = ray.d[0]; d_k = ray.d[1]; break;
            case 2: o_u = ray.o[0]; o_v = ray.o[1]; o_k = ray.o[2]; d_u = ray.d[0]; d_v = ray.d[1]; d_k = ray.d[2]; break;
            default: return false;
        }
        t = (n_d - o_u * n_u - o_v * n_v - o_k) / (d_u * n_u + d_v * n_v + d_k);
        if (t < mint || t > maxt) return false;
        const float hu = o_u + t * d_u - a_u;
        const float hv = o_v + t * d_v - a_v;
        u = hv * b_nu + hu * b_nv;
        v = hu * c_nu + hv * c_nv;
        return u >= 0 && v >= 0 && u + v <= 1.0f;
    }
};

struct Hit { float t, u, v; uint32_t prim; }; /* skdtree.h:237-241 IntersectionCache + t */

struct KDNode { /* semantic stand-in for gkdtree.h:452-601 (leaf flag, axis, split, children adjacent) */
    bool leaf; int axis; float split; uint32_t left; uint32_t primStart, primEnd;
};

struct Accel {
    std::vector<TriAccel> tri;
    std::vector<AABB> triBox;
    std::vector<KDNode> nodes;
    std::vector<uint32_t> indices;
    AABB aabb;          /* enlarged, gkdtree.h:1213-1220 */
    bool useTree = true;
    int maxDepth = 0;
    /* build statistics */
    uint64_t nLeaves = 0, nInner = 0;

    static const int kStackSize = 48; /* MTS_KD_MAXDEPTH gkdtree.h:37 */

    void build(bool tree) {
        useTree = tree;
        aabb = AABB();
        for (auto &b : triBox) { aabb.expandBy(b.min); aabb.expandBy(b.max); }
        AABB tight = aabb;
        if (tri.empty()) { nodes.clear(); return; }
        if (tree) {
            uint32_t n = (uint32_t) tri.size();
            uint32_t lg = 0; { uint32_t v = n; while (v >>= 1) ++lg; }
            maxDepth = std::min((int) (8 + 1.3f * lg), kStackSize); /* gkdtree.h:986-988 */
            std::vector<uint32_t> all(n);
            for (uint32_t i = 0; i < n; ++i) all[i] = i;
            nodes.clear(); indices.clear();
            nodes.push_back(KDNode());
            buildRec(0, all, tight, 0, 0);
        }
        /* gkdtree.h:1217-1220 -- note max uses the already-updated min */
        const float eps = 1e-3f;
        aabb.min = aabb.min - ((aabb.max - aabb.min) * eps + V3(eps));
        aabb.max = aabb.max + ((aabb.max - aabb.min) * eps + V3(eps));
    }

    struct Ev { float pos; int type; /* 0=end,1=planar,2=start */ };

    void makeLeaf(uint32_t ni, const std::vector<uint32_t> &prims) {
        KDNode &nd = nodes[ni];
        nd.leaf = true; nd.axis = 0; nd.split = 0; nd.left = 0;
        nd.primStart = (uint32_t) indices.size();
        indices.insert(indices.end(), prims.begin(), prims.end());
        nd.primEnd = (uint32_t) indices.size();
        ++nLeaves;
    }

    void buildRec(uint32_t ni, const std::vector<uint32_t> &prims, const AABB &box, int depth, int badRefines) {
        const float traversalCost = 15, queryCost = 20, emptyBonus = 0.9f;
        const uint32_t stopPrims = 6; const int maxBadRefines = 3;
        uint32_t n = (uint32_t) prims.size();
        if (n <= stopPrims || depth >= maxDepth) { makeLeaf(ni, prims); return; }
        float leafCost = queryCost * n;
        V3 ext = box.max - box.min;
        float invSA = 1.0f / (ext.x * ext.y + ext.y * ext.z + ext.x * ext.z);
        float bestCost = kInf, bestPos = 0; int bestAxis = -1;
        std::vector<Ev> ev;
        for (int axis = 0; axis < 3; ++axis) {
            if (!(ext[axis] > 0)) continue;
            ev.clear();
            for (uint32_t p : prims) {
                float lo = std::max(triBox[p].min[axis], box.min[axis]);
                float hi = std::min(triBox[p].max[axis], box.max[axis]);
                if (lo == hi) ev.push_back({lo, 1});
                else { ev.push_back({lo, 2}); ev.push_back({hi, 0}); }
            }
            std::sort(ev.begin(), ev.end(), [](const Ev &a, const Ev &b) { return a.pos < b.pos || (a.pos == b.pos && a.type < b.type); });
            int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
            float t0 = ext[a1] * ext[a2] * invSA, t1 = (ext[a1] + ext[a2]) * invSA; /* sahkdtree3.h:44-58 */
            uint32_t nL = 0, nR = n;
            size_t i = 0;
            while (i < ev.size()) {
                float pos = ev[i].pos;
                uint32_t pEnd = 0, pPlanar = 0, pStart = 0;
                while (i < ev.size() && ev[i].pos == pos && ev[i].type == 0) { ++pEnd; ++i; }
                while (i < ev.size() && ev[i].pos == pos && ev[i].type == 1) { ++pPlanar; ++i; }
                while (i < ev.size() && ev[i].pos == pos && ev[i].type == 2) { ++pStart; ++i; }
                nR -= pPlanar + pEnd;
                if (pos > box.min[axis] && pos < box.max[axis]) {
                    float pL = t0 + t1 * (pos - box.min[axis]), pR = t0 + t1 * (box.max[axis] - pos);
                    /* planar prims are placed on both sides here (see file header) */
                    uint32_t cl = nL + pPlanar, cr = nR + pPlanar;
                    float cost = traversalCost + queryCost * (pL * cl + pR * cr);
                    if (cl == 0 || cr == 0) cost *= emptyBonus;
                    if (cost < bestCost) { bestCost = cost; bestPos = pos; bestAxis = axis; }
                }
                nL += pStart + pPlanar;
            }
        }
        if (bestAxis < 0) { makeLeaf(ni, prims); return; }
        if (bestCost >= leafCost) { /* gkdtree.h bad-refine rule */
            if ((bestCost > 4 * leafCost && n < 16) || badRefines >= maxBadRefines) { makeLeaf(ni, prims); return; }
            ++badRefines;
        }
        std::vector<uint32_t> L, R;
        for (uint32_t p : prims) {
            float lo = std::max(triBox[p].min[bestAxis], box.min[bestAxis]);
            float hi = std::min(triBox[p].max[bestAxis], box.max[bestAxis]);
            if (lo == hi && lo == bestPos) { L.push_back(p); R.push_back(p); continue; }
            if (lo < bestPos) L.push_back(p);
            if (hi > bestPos) R.push_back(p);
        }
        if (L.size() == n && R.size() == n) { makeLeaf(ni, prims); return; }
        uint32_t li = (uint32_t) nodes.size();
        nodes.push_back(KDNode()); nodes.push_back(KDNode());
        nodes[ni].leaf = false; nodes[ni].axis = bestAxis; nodes[ni].split = bestPos; nodes[ni].left = li;
        ++nInner;
        AABB lb = box, rb = box;
        lb.max[bestAxis] = bestPos; rb.min[bestAxis] = bestPos;
        buildRec(li, L, lb, depth + 1, badRefines);
        buildRec(li + 1, R, rb, depth + 1, badRefines);
    }

    /* sahkdtree3.h:178-308 rayIntersectHavran<shadowRay> */
    struct StackEntry { int node; float t; uint32_t prev; V3 p; };
    template <bool shadowRay> bool havran(const Ray &ray, float mint, float maxt, Hit &hit,
                                          uint64_t *nodeVisits = nullptr, uint64_t *primTests = nullptr) const {
        StackEntry stack[kStackSize + 2];
        uint32_t enPt = 0;
        stack[enPt].t = mint; stack[enPt].p = ray(mint);
        uint32_t exPt = 1;
        stack[exPt].t = maxt; stack[exPt].p = ray(maxt); stack[exPt].node = -1;
        bool found = false;
        int curr = 0;
        while (curr != -1) {
            while (!nodes[curr].leaf) {
                if (nodeVisits) ++*nodeVisits;
                const float splitVal = nodes[curr].split;
                const int axis = nodes[curr].axis;
                int farChild;
                if (stack[enPt].p[axis] <= splitVal) {
                    if (stack[exPt].p[axis] <= splitVal) { curr = (int) nodes[curr].left; continue; }
                    if (stack[enPt].p[axis] == splitVal) { curr = (int) nodes[curr].left + 1; continue; }
                    curr = (int) nodes[curr].left;
                    farChild = curr + 1;
                } else {
                    if (splitVal < stack[exPt].p[axis]) { curr = (int) nodes[curr].left + 1; continue; }
                    farChild = (int) nodes[curr].left;
                    curr = farChild + 1;
                }
                float distToSplit = (splitVal - ray.o[axis]) * ray.dRcp[axis];
                const uint32_t tmp = exPt++;
                if (exPt == enPt) ++exPt;
                stack[exPt].prev = tmp;
                stack[exPt].t = distToSplit;
                stack[exPt].node = farChild;
                stack[exPt].p = ray(distToSplit);
                stack[exPt].p[axis] = splitVal;
            }
            for (uint32_t e = nodes[curr].primStart, last = nodes[curr].primEnd; e != last; ++e) {
                const uint32_t primIdx = indices[e];
                float tu, tv, tt;
                if (primTests) ++*primTests;
                if (tri[primIdx].rayIntersect(ray, mint, maxt, tu, tv, tt)) {
                    if (shadowRay) return true;
                    hit.t = tt; hit.u = tu; hit.v = tv; hit.prim = primIdx;
                    maxt = tt;
                    found = true;
                }
            }
            if (stack[exPt].t > maxt) break;
            enPt = exPt;
            curr = stack[exPt].node;
            exPt = stack[enPt].prev;
        }
        return found;
    }
    /* order-independent reference: test every triangle, keep the closest (ties: last wins, like
     * triaccel.h:147-148 `t > maxt` rejection) */
    template <bool shadowRay> bool brute(const Ray &ray, float mint, float maxt, Hit &hit) const {
        bool found = false;
        for (uint32_t i = 0; i < tri.size(); ++i) {
            float tu, tv, tt;
            if (tri[i].rayIntersect(ray, mint, maxt, tu, tv, tt)) {
                if (shadowRay) return true;
                hit.t = tt; hit.u = tu; hit.v = tv; hit.prim = i; maxt = tt; found = true;
            }
        }
        return found;
    }
    template <bool shadowRay> bool query(const Ray &ray, float mint, float maxt, Hit &hit,
                                         uint64_t *nv = nullptr, uint64_t *pt = nullptr) const {
        if (tri.empty()) return false;
        return useTree ? havran<shadowRay>(ray, mint, maxt, hit, nv, pt) : brute<shadowRay>(ray, mint, maxt, hit);
    }

    /* src/librender/skdtree.cpp:112-142 closest hit */
    bool rayIntersect(const Ray &ray, Hit &hit, uint64_t *nv = nullptr, uint64_t *pt = nullptr) const {
        float mint, maxt;
        hit.t = kInf;
        if (aabb.rayIntersect(ray, mint, maxt)) {
            float rayMinT = ray.mint;
            if (rayMinT == kEpsilon)
                rayMinT *= std::max(std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z)), kEpsilon);
            if (rayMinT > mint) mint = rayMinT;
            if (ray.maxt < maxt) maxt = ray.maxt;
            if (maxt > mint) return query<false>(ray, mint, maxt, hit, nv, pt);
        }
        return false;
    }
    /* src/librender/skdtree.cpp:207-226 occlusion (epsilon scale has no inner max with Epsilon) */
    bool rayOccluded(const Ray &ray, uint64_t *nv = nullptr, uint64_t *pt = nullptr) const {
        float mint, maxt;
        Hit h;
        if (aabb.rayIntersect(ray, mint, maxt)) {
            float rayMinT = ray.mint;
            if (rayMinT == kEpsilon)
                rayMinT *= std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z));
            if (rayMinT > mint) mint = rayMinT;
            if (ray.maxt < maxt) maxt = ray.maxt;
            if (maxt > mint) return query<true>(ray, mint, maxt, h, nv, pt);
        }
        return false;
    }
};

} // namespace orc
