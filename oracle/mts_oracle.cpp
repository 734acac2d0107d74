/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header; parity status: image level "unpinned").
 *
 * CPU restatement of the Mitsuba 0.6 `path` integrator hot path:
 *   a1  MIPathTracer::Li                    src/integrators/path/path.cpp:119-300
 *   a2  SamplingIntegrator::renderBlock     src/librender/integrator.cpp:140-188
 *   a3-a5 ray queries                       orc_accel.h
 *   a6  fillIntersectionRecord<true>        include/mitsuba/render/skdtree.h:343-428
 *   a7-a8 BSDFs / microfacet                orc_bsdf.h
 *   a9  Scene::sampleEmitterDirect          src/librender/scene.cpp:828-852,949-952
 *   a10 AreaLight / Shape / TriMesh / Triangle sampling
 *                                           src/emitters/area.cpp:104-183, src/librender/shape.cpp:102-126,
 *                                           src/librender/trimesh.cpp:358-424, src/libcore/triangle.cpp:24-62
 *   a11 samplers                            orc_sampler.h
 *   a12 perspective sensor                  src/sensors/perspective.cpp:271-298
 *   a13 ImageBlock::put + filter table      include/mitsuba/render/imageblock.h:124-204, src/libcore/rfilter.cpp:37-57
 *   a14 film merge / develop                include/mitsuba/render/imageblock.h:103-107, src/libcore/fmtconv.cpp:955-990
 * Exposed as a flat C API for ctypes (oracle/oracle_api.py).                                            */
#include "orc_math.h"
#include "orc_sampler.h"
#include "orc_accel.h"
#include "orc_bsdf.h"
#include "orc_medium.h"
#include "orc_envmap.h"
#include <thread>
#include <atomic>
#include <mutex>
#include <memory>
#include <cstdio>

using namespace orc;

extern "C" {
typedef struct OrcRenderParams {
    int32_t spp;
    int32_t sampler;        /* 0 sobol, 1 independent (SFMT, thread-order dependent), 2 counter-based TEA */
    uint64_t seed;          /* sobol: `scramble` property; others: seed */
    int32_t maxDepth, rrDepth, strictNormals, hideEmitters; /* integrator.cpp:190-225 */
    int32_t rfilter;        /* 0 box, 1 gaussian */
    float rfilterParam;     /* box: radius property (0.5); gaussian: stddev (0.5) */
    int32_t sampleLo, sampleHi; /* render sample indices [lo,hi) of every pixel (multi-GPU sharding mirror); hi<=0 -> spp */
    int32_t threads;        /* 0 -> hardware_concurrency */
    int32_t blockSize;      /* scene.cpp:24 default 32 */
    int32_t integrator;     /* 0 path (path.cpp), 1 volpath (volpath.cpp) */
} OrcRenderParams;
typedef struct OrcStats {
    uint64_t samples, rays, shadowRays, pathLengthSum, nodeVisits, primTests, badSamples, dimOverflow;
} OrcStats;
}

namespace {

struct Mesh {
    std::vector<V3> P, N; std::vector<float> UV; /* UV: 2 per vertex */
    std::vector<uint32_t> idx; /* 3 per triangle */
    std::vector<V3> dpdu, dpdv; /* per triangle, only if UVs exist (trimesh.cpp:683-735) */
    std::vector<uint8_t> hasTangent;
    int bsdf = -1; int emitter = -1;
    int interior = -1, exterior = -1; /* shape.h:427-435 media ids (-1 = vacuum) */
    int group = -1;                   /* >= 0: member of that ShapeGroup (src/shapes/shapegroup.cpp), in object space */
    bool isMediumTransition() const { return interior >= 0 || exterior >= 0; }
    uint32_t primOffset = 0;
    /* area sampling: trimesh.cpp:388-403 + pmf.h */
    std::vector<float> cdf; float surfaceArea = -1, invSurfaceArea = 0;
    uint32_t nTri() const { return (uint32_t) (idx.size() / 3); }
};

/* include/mitsuba/core/pmf.h:60-187 */
struct Discrete {
    std::vector<float> cdf{0.0f}; float sum = 0, normalization = 0;
    void append(float v) { cdf.push_back(cdf.back() + v); }
    float normalize() {
        sum = cdf.back();
        if (sum > 0) {
            normalization = 1.0f / sum;
            for (size_t i = 1; i < cdf.size(); ++i) cdf[i] *= normalization;
            cdf.back() = 1.0f;
        } else normalization = 0.0f;
        return sum;
    }
    float operator[](size_t i) const { return cdf[i + 1] - cdf[i]; }
    size_t sample(float v) const {
        auto entry = std::lower_bound(cdf.begin(), cdf.end(), v);
        size_t index = std::min(cdf.size() - 2, (size_t) std::max((ptrdiff_t) 0, (ptrdiff_t) (entry - cdf.begin()) - 1));
        while ((*this)[index] == 0 && index < cdf.size() - 1) ++index;
        return index;
    }
    size_t sampleReuse(float &v) const {
        size_t index = sample(v);
        v = (v - cdf[index]) / (cdf[index + 1] - cdf[index]);
        return index;
    }
    size_t sampleReuse(float &v, float &pdf) const {
        size_t index = sample(v);
        pdf = (*this)[index];
        v = (v - cdf[index]) / (cdf[index + 1] - cdf[index]);
        return index;
    }
};

struct Emitter { V3 radiance; float samplingWeight; int mesh; /* -1: environment emitter: `constant` (src/emitters/constant.cpp) ... */
                 int envmap = -1; /* ... or, when >= 0, `envmap` (src/emitters/envmap.cpp, orc_envmap.h): index into Scene::envmaps */ };

struct Intersection { /* include/mitsuba/render/shape.h:131-171 (fields `path` reads) */
    float t = kInf; V3 p; Frame geoFrame, shFrame; V3 wi; V3 dpdu, dpdv; int mesh = -1; uint32_t prim = 0; int instance = -1;
    TexCtx tex; /* uv, hasUVPartials, dudx.. (shape.h:147-165) in the form the BSDFs' texture look-up takes them */
    bool isValid() const { return t != kInf; }
};
struct RayDiff { bool has = false; V3 rxO, ryO, rxD, ryD; }; /* include/mitsuba/core/ray.h:120-176 */

struct DRec { /* include/mitsuba/render/common.h:85-123,241-255 */
    V3 p, n, ref, refN, d; float pdf = 0, dist = 0; int emitter = -1; bool solidAngle = true;
};

struct Scene {
    std::vector<OrcBsdf> bsdfs;
    std::vector<Texture> textures; /* `bitmap` textures (orc_texture.h) referenced by OrcBsdf::texture */
    std::vector<Mesh> meshes;
    std::vector<Emitter> emitters;
    std::vector<EnvMap> envmaps;
    std::vector<OrcMedium> media;
    /* instancing: src/shapes/{shapegroup,instance}.cpp */
    struct Instance { int group; float M[16], Minv[16]; };
    std::vector<Accel> groups;          /* one kd-tree per ShapeGroup (object space) */
    std::vector<Instance> instances;
    AABB topAABB;                       /* top-level kd-tree box: world triangles + instance boxes, enlarged */
    int envEmitter = -1;                 /* Scene::getEnvironmentEmitter */
    V3 bsCenter; float bsRadius = 0;     /* constant.cpp:67-70 m_sceneBSphere */
    std::vector<std::vector<float>> mediaData; /* owned copies of the density grids */
    std::vector<uint32_t> primMesh; /* prim -> mesh (m_shapeMap) */
    Accel accel;
    Discrete emitterPDF;
    SobolTables sobol;
    /* camera */
    float camToWorld[16], sampleToCamera[16]; float nearClip = 1e-2f, farClip = 1e4f; int W = 0, H = 0;
    float apertureRadius = 0, focusDistance = 0; /* > 0: `thinlens` sensor (src/sensors/thinlens.cpp) */
    V3 camOrigin;

    void commit(bool tree) {
        accel.tri.clear(); accel.triBox.clear(); primMesh.clear();
        for (auto &g : groups) { g.tri.clear(); g.triBox.clear(); }
        for (size_t mi = 0; mi < meshes.size(); ++mi) {
            Mesh &m = meshes[mi];
            Accel &accel = m.group >= 0 ? groups[m.group] : this->accel;
            m.primOffset = (uint32_t) accel.tri.size();
            for (uint32_t j = 0; j < m.nTri(); ++j) {
                const V3 &v0 = m.P[m.idx[3 * j]], &v1 = m.P[m.idx[3 * j + 1]], &v2 = m.P[m.idx[3 * j + 2]];
                TriAccel ta; memset(&ta, 0, sizeof(ta));
                ta.load(v0, v1, v2);                 /* skdtree.cpp:88-94 */
                ta.shapeIndex = (uint32_t) mi; ta.primIndex = j;
                accel.tri.push_back(ta);
                AABB b; b.expandBy(v0); b.expandBy(v1); b.expandBy(v2);
                accel.triBox.push_back(b);
                if (m.group < 0) primMesh.push_back((uint32_t) mi);
            }
            /* trimesh.cpp:683-735 computeUVTangents (called unconditionally at :385) */
            m.dpdu.clear(); m.dpdv.clear(); m.hasTangent.clear();
            if (!m.UV.empty()) {
                m.dpdu.resize(m.nTri()); m.dpdv.resize(m.nTri()); m.hasTangent.assign(m.nTri(), 1);
                for (uint32_t j = 0; j < m.nTri(); ++j) {
                    uint32_t i0 = m.idx[3 * j], i1 = m.idx[3 * j + 1], i2 = m.idx[3 * j + 2];
                    V3 dP1 = m.P[i1] - m.P[i0], dP2 = m.P[i2] - m.P[i0];
                    float du1 = m.UV[2 * i1] - m.UV[2 * i0], dv1 = m.UV[2 * i1 + 1] - m.UV[2 * i0 + 1];
                    float du2 = m.UV[2 * i2] - m.UV[2 * i0], dv2 = m.UV[2 * i2 + 1] - m.UV[2 * i0 + 1];
                    V3 n = cross(dP1, dP2);
                    float length = n.length();
                    if (length == 0) { m.dpdu[j] = V3(0.0f); m.dpdv[j] = V3(0.0f); continue; } /* memset-zero tangent, :695 */
                    float determinant = du1 * dv2 - dv1 * du2;
                    if (determinant == 0) {
                        V3 a, b; coordinateSystem(n / length, a, b); m.dpdu[j] = a; m.dpdv[j] = b;
                    } else {
                        float invDet = 1.0f / determinant;
                        m.dpdu[j] = (dv2 * dP1 - dv1 * dP2) * invDet;
                        m.dpdv[j] = (-du2 * dP1 + du1 * dP2) * invDet;
                    }
                }
            }
            /* trimesh.cpp:388-403 prepareSamplingTable (only emitters need it) */
            if (m.emitter >= 0) {
                Discrete dd;
                for (uint32_t j = 0; j < m.nTri(); ++j) {
                    V3 sideA = m.P[m.idx[3 * j + 1]] - m.P[m.idx[3 * j]], sideB = m.P[m.idx[3 * j + 2]] - m.P[m.idx[3 * j]];
                    dd.append(0.5f * cross(sideA, sideB).length()); /* triangle.cpp:64-70 */
                }
                m.surfaceArea = dd.normalize();
                m.invSurfaceArea = 1.0f / m.surfaceArea;
                m.cdf = dd.cdf;
            }
        }
        accel.build(tree);
        for (auto &g : groups) g.build(tree);
        /* top-level box: the kd-tree over world triangles and Instance::getAABB boxes (instance.cpp:80-96), enlarged like every tree */
        topAABB = accel.aabb;
        if (!instances.empty()) {
            AABB b;
            for (auto &tb : accel.triBox) { b.expandBy(tb.min); b.expandBy(tb.max); }
            for (auto &in : instances) {
                const AABB &ga = groups[in.group].aabb;
                if (groups[in.group].tri.empty()) continue;
                for (int c = 0; c < 8; ++c) {
                    V3 corner((c & 1) ? ga.max.x : ga.min.x, (c & 2) ? ga.max.y : ga.min.y, (c & 4) ? ga.max.z : ga.min.z);
                    b.expandBy(xfPoint(in.M, corner));
                }
            }
            const float eps = 1e-3f;
            b.min = b.min - ((b.max - b.min) * eps + V3(eps));
            b.max = b.max + ((b.max - b.min) * eps + V3(eps));
            topAABB = b;
        }
        /* order of Scene::m_emitters: emitters that are direct children of the scene (`constant`) are appended by Scene::addChild
           (scene.cpp:510-516); the area emitters of shapes only join in Scene::initialize -> addShape (scene.cpp:322-335, :570-571),
           i.e. behind them and in shape order, whatever the document order (pinned by tests/golden/path_ref_ext.npz) */
        {
            std::vector<int> remap(emitters.size(), -1);
            std::vector<Emitter> sorted;
            for (size_t e = 0; e < emitters.size(); ++e) if (emitters[e].mesh < 0) { remap[e] = (int) sorted.size(); sorted.push_back(emitters[e]); }
            for (auto &m : meshes) if (m.emitter >= 0) { remap[m.emitter] = (int) sorted.size(); sorted.push_back(emitters[m.emitter]); }
            for (auto &m : meshes) if (m.emitter >= 0) m.emitter = remap[m.emitter];
            emitters.swap(sorted);
        }
        /* scene.cpp:375-380 */
        emitterPDF = Discrete();
        for (auto &e : emitters) emitterPDF.append(e.samplingWeight);
        if (!emitters.empty()) emitterPDF.normalize();
        camOrigin = V3(camToWorld[3], camToWorld[7], camToWorld[11]); /* trafo.transformAffine(Point(0)) */
        /* scene.cpp:386-399: m_aabb = kd-tree box expanded by the sensor position; constant.cpp:67-70: its bounding sphere
           (aabb.cpp:44-47), radius * 1.5 */
        envEmitter = -1;
        for (size_t e = 0; e < emitters.size(); ++e) if (emitters[e].mesh < 0) envEmitter = (int) e;
        {
            AABB b = topAABB;
            if (accel.tri.empty() && instances.empty()) { b.min = V3(0.0f); b.max = V3(0.0f); }
            b.expandBy(camOrigin);
            bsCenter = (b.max + b.min) * 0.5f;
            bsRadius = std::max(kEpsilon, (bsCenter - b.max).length() * 1.5f);
        }
    }

    /* skdtree.h:343-428 fillIntersectionRecord<true> for triangle meshes */
    /* transform.h:108-124 (affine: w == 1), :175-183, :203-211 */
    static V3 xfPoint(const float *M, const V3 &p) { /* transform.h:108-124: homogeneous, divides unless w == 1 exactly -- which happens with
        the reference's own float Gauss-Jordan inverse of an affine matrix (its last row can come out as (0, 3e-8, 0, 0.99999994)) */
        const V3 r(M[0] * p.x + M[1] * p.y + M[2] * p.z + M[3], M[4] * p.x + M[5] * p.y + M[6] * p.z + M[7], M[8] * p.x + M[9] * p.y + M[10] * p.z + M[11]);
        const float w = M[12] * p.x + M[13] * p.y + M[14] * p.z + M[15];
        return w == 1.0f ? r : r / w;
    }
    static V3 xfVector(const float *M, const V3 &v) {
        return V3(M[0] * v.x + M[1] * v.y + M[2] * v.z, M[4] * v.x + M[5] * v.y + M[6] * v.z, M[8] * v.x + M[9] * v.y + M[10] * v.z);
    }
    static V3 xfNormal(const float *Minv, const V3 &n) { /* inverse transpose */
        return V3(Minv[0] * n.x + Minv[4] * n.y + Minv[8] * n.z, Minv[1] * n.x + Minv[5] * n.y + Minv[9] * n.z, Minv[2] * n.x + Minv[6] * n.y + Minv[10] * n.z);
    }
    static Ray xfRay(const float *Minv, const Ray &r) { /* transform.h:262-278 */
        return Ray(xfPoint(Minv, r.o), xfVector(Minv, r.d), r.mint, r.maxt);
    }
    void fill(const Ray &wray, const Hit &h, Intersection &its, int inst = -1) const {
        const Accel &accel = inst >= 0 ? groups[instances[inst].group] : this->accel;
        const Ray ray = inst >= 0 ? xfRay(instances[inst].Minv, wray) : wray;
        const uint32_t mi = accel.tri[h.prim].shapeIndex, pi = accel.tri[h.prim].primIndex;
        const Mesh &m = meshes[mi];
        const V3 b(1 - h.u - h.v, h.u, h.v);
        const uint32_t i0 = m.idx[3 * pi], i1 = m.idx[3 * pi + 1], i2 = m.idx[3 * pi + 2];
        const V3 &p0 = m.P[i0], &p1 = m.P[i1], &p2 = m.P[i2];
        its.t = h.t;
        its.p = p0 * b.x + p1 * b.y + p2 * b.z;
        V3 side1(p1 - p0), side2(p2 - p0);
        V3 faceNormal(cross(side1, side2));
        float length = faceNormal.length();
        if (!faceNormal.isZero()) faceNormal /= length;
        if (!m.dpdu.empty()) { its.dpdu = m.dpdu[pi]; its.dpdv = m.dpdv[pi]; } else { its.dpdu = side1; its.dpdv = side2; }
        its.tex = TexCtx();
        its.tex.textures = textures.data();
        if (!m.UV.empty()) { /* skdtree.h:398-405 */
            its.tex.u = m.UV[2 * i0] * b.x + m.UV[2 * i1] * b.y + m.UV[2 * i2] * b.z;
            its.tex.v = m.UV[2 * i0 + 1] * b.x + m.UV[2 * i1 + 1] * b.y + m.UV[2 * i2 + 1] * b.z;
        } else { its.tex.u = b.y; its.tex.v = b.z; }
        if (!m.N.empty()) {
            const V3 &n0 = m.N[i0], &n1 = m.N[i1], &n2 = m.N[i2];
            its.shFrame.n = normalize(n0 * b.x + n1 * b.y + n2 * b.z);
            if (dot(faceNormal, its.shFrame.n) < 0) faceNormal = -faceNormal;
        } else its.shFrame.n = faceNormal;
        its.geoFrame = Frame(faceNormal);
        its.mesh = (int) mi; its.prim = pi; its.instance = inst;
        if (inst >= 0) { /* Instance::fillIntersectionRecord, instance.cpp:149-162 (nested record: skdtree.h fillIntersectionRecord<false>) */
            const Instance &in = instances[inst];
            its.p = ray(its.t);
            its.shFrame.n = normalize(xfNormal(in.Minv, its.shFrame.n));
            its.geoFrame = Frame(normalize(xfNormal(in.Minv, its.geoFrame.n)));
            its.dpdu = xfVector(in.M, its.dpdu);
            its.dpdv = xfVector(in.M, its.dpdv);
            its.p = xfPoint(in.M, its.p);
        }
        computeShadingFrame(its.shFrame.n, its.dpdu, its.shFrame);
        its.wi = its.shFrame.toLocal(-wray.d);
    }
    /* top-level queries with instances: skdtree.cpp:112-142 / :207-226 over world triangles and Instance::rayIntersect
       (instance.cpp:115-133 -> the nested tree's plain query, skdtree.h:430-458) */
    bool topClosest(const Ray &ray, Hit &best, int &inst, OrcStats &st) const {
        float mint, maxt;
        best.t = kInf; inst = -1;
        if (!topAABB.rayIntersect(ray, mint, maxt)) return false;
        float rayMinT = ray.mint;
        if (rayMinT == kEpsilon) rayMinT *= std::max(std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z)), kEpsilon);
        if (rayMinT > mint) mint = rayMinT;
        if (ray.maxt < maxt) maxt = ray.maxt;
        if (!(maxt > mint)) return false;
        bool found = false;
        Hit h;
        if (!accel.tri.empty() && accel.query<false>(ray, mint, maxt, h, &st.nodeVisits, &st.primTests)) { best = h; maxt = h.t; found = true; }
        for (size_t i = 0; i < instances.size(); ++i) {
            const Accel &g = groups[instances[i].group];
            if (g.tri.empty()) continue;
            const Ray r2 = xfRay(instances[i].Minv, ray);
            float m0, m1;
            if (!g.aabb.rayIntersect(r2, m0, m1)) continue;
            if (mint > m0) m0 = mint;
            if (maxt < m1) m1 = maxt;
            if (!(m1 > m0)) continue;
            if (g.query<false>(r2, m0, m1, h, &st.nodeVisits, &st.primTests)) { best = h; maxt = h.t; inst = (int) i; found = true; }
        }
        return found;
    }
    bool topOccluded(const Ray &ray, OrcStats &st) const {
        float mint, maxt;
        if (!topAABB.rayIntersect(ray, mint, maxt)) return false;
        float rayMinT = ray.mint;
        if (rayMinT == kEpsilon) rayMinT *= std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z));
        if (rayMinT > mint) mint = rayMinT;
        if (ray.maxt < maxt) maxt = ray.maxt;
        if (!(maxt > mint)) return false;
        Hit h;
        if (!accel.tri.empty() && accel.query<true>(ray, mint, maxt, h, &st.nodeVisits, &st.primTests)) return true;
        for (size_t i = 0; i < instances.size(); ++i) {
            const Accel &g = groups[instances[i].group];
            if (g.tri.empty()) continue;
            const Ray r2 = xfRay(instances[i].Minv, ray);
            float m0, m1;
            if (!g.aabb.rayIntersect(r2, m0, m1)) continue;
            if (mint > m0) m0 = mint;
            if (maxt < m1) m1 = maxt;
            if (!(m1 > m0)) continue;
            if (g.query<true>(r2, m0, m1, h, &st.nodeVisits, &st.primTests)) return true;
        }
        return false;
    }

    /* area.cpp:104-109 */
    Spectrum emitterEval(int e, const Intersection &its, const V3 &d) const {
        if (dot(its.shFrame.n, d) <= 0) return Spectrum(0.0f);
        return emitters[e].radiance;
    }
    /* scene.cpp:828-852 with testVisibility=true; returns value, fills dRec; `occl` counts shadow rays */
    Spectrum sampleEmitterDirect(DRec &dRec, float sx, float sy, OrcStats &st, bool testVisibility = true, float *emPdfOut = nullptr) const {
        float emPdf;
        size_t index = emitterPDF.sampleReuse(sx, emPdf);
        const Emitter &em = emitters[index];
        if (em.mesh < 0 && em.envmap >= 0) { /* envmap.cpp:516-543 sampleDirect, then scene.cpp:838-851 */
            const EnvMap &env = envmaps[em.envmap];
            V3 value, d; float pdf;
            env.sampleDirection(sx, sy, d, value, pdf);
            const V3 dw = EnvMap::xfVector(env.toWorld, d);
            float nearT, farT;
            dRec.pdf = 0.0f;
            if (value.isZero() || pdf == 0 || !bsphereIntersect(dRec.ref, dw, nearT, farT) || nearT >= 0 || farT <= 0) return Spectrum(0.0f);
            dRec.pdf = pdf;
            dRec.p = dRec.ref + dw * farT;
            dRec.n = normalize(bsCenter - dRec.p);
            dRec.dist = farT; dRec.d = dw; dRec.solidAngle = true;
            value = value / pdf;
            if (!testVisibility) { dRec.emitter = (int) index; if (emPdfOut) *emPdfOut = emPdf; return value; }
            Ray ray(dRec.ref, dRec.d, kEpsilon, dRec.dist * (1 - kShadowEpsilon));
            ++st.shadowRays;
            if (instances.empty() ? accel.rayOccluded(ray, &st.nodeVisits, &st.primTests) : topOccluded(ray, st)) return Spectrum(0.0f);
            dRec.emitter = (int) index;
            dRec.pdf *= emPdf;
            value /= emPdf;
            return value;
        }
        if (em.mesh < 0) { /* constant.cpp:171-208 sampleDirect */
            V3 d; float pdf;
            if (!dRec.refN.isZero()) {
                d = squareToCosineHemisphere(sx, sy);
                pdf = squareToCosineHemispherePdf(d);
                d = Frame(dRec.refN).toWorld(d);
            } else {
                d = squareToUniformSphere(sx, sy);
                pdf = kInvFourPi;
            }
            dRec.pdf = 0.0f;
            float nearT, farT;
            if (!bsphereIntersect(dRec.ref, d, nearT, farT)) return Spectrum(0.0f);
            if (!(nearT < 0 && farT > 0)) return Spectrum(0.0f);
            dRec.p = dRec.ref + d * farT;
            dRec.n = normalize(bsCenter - dRec.p);
            dRec.solidAngle = true; dRec.d = d; dRec.dist = farT; dRec.pdf = pdf;
            Spectrum value = em.radiance / pdf;
            if (!dRec.refN.isZero() && dot(dRec.d, dRec.refN) <= 0) value = Spectrum(0.0f); /* roundoff moved it to the back side; pdf stays != 0 */
            if (!testVisibility) { dRec.emitter = (int) index; if (emPdfOut) *emPdfOut = emPdf; return value; }
            Ray ray(dRec.ref, dRec.d, kEpsilon, dRec.dist * (1 - kShadowEpsilon)); /* scene.cpp:838-843: not on a surface, but the same ray */
            ++st.shadowRays;
            if (instances.empty() ? accel.rayOccluded(ray, &st.nodeVisits, &st.primTests) : topOccluded(ray, st)) return Spectrum(0.0f);
            dRec.emitter = (int) index;
            dRec.pdf *= emPdf;
            value /= emPdf;
            return value;
        }
        const Mesh &m = meshes[em.mesh];
        /* area.cpp:158-173 -> shape.cpp:102-115 -> trimesh.cpp:412-424 -> triangle.cpp:24-62 */
        {
            Discrete dd; dd.cdf = m.cdf;
            size_t ti = dd.sampleReuse(sy);
            const V3 &p0 = m.P[m.idx[3 * ti]], &p1 = m.P[m.idx[3 * ti + 1]], &p2 = m.P[m.idx[3 * ti + 2]];
            float bx, by;
            squareToUniformTriangle(sx, sy, bx, by);
            V3 sideA = p1 - p0, sideB = p2 - p0;
            dRec.p = p0 + (sideA * bx) + (sideB * by);
            if (!m.N.empty()) {
                const V3 &n0 = m.N[m.idx[3 * ti]], &n1 = m.N[m.idx[3 * ti + 1]], &n2 = m.N[m.idx[3 * ti + 2]];
                dRec.n = normalize(n0 * (1.0f - bx - by) + n1 * bx + n2 * by);
            } else dRec.n = normalize(cross(sideA, sideB));
            dRec.pdf = m.invSurfaceArea;
        }
        dRec.d = dRec.p - dRec.ref;
        float distSquared = dRec.d.lengthSquared();
        dRec.dist = std::sqrt(distSquared);
        dRec.d /= dRec.dist;
        float dp = absDot(dRec.d, dRec.n);
        dRec.pdf *= dp != 0 ? (distSquared / dp) : 0.0f;
        dRec.solidAngle = true;
        Spectrum value;
        if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) value = em.radiance / dRec.pdf;
        else { dRec.pdf = 0.0f; value = Spectrum(0.0f); }
        if (!testVisibility) { /* sampleAttenuatedEmitterDirect (scene.cpp:854-898): the caller applies transmittance / emPdf */
            dRec.emitter = (int) index;
            if (emPdfOut) *emPdfOut = emPdf;
            return value;
        }
        if (dRec.pdf != 0) {
            Ray ray(dRec.ref, dRec.d, kEpsilon, dRec.dist * (1 - kShadowEpsilon));
            ++st.shadowRays;
            if (instances.empty() ? accel.rayOccluded(ray, &st.nodeVisits, &st.primTests) : topOccluded(ray, st)) return Spectrum(0.0f);
            dRec.emitter = (int) index;
            dRec.pdf *= emPdf;
            value /= emPdf;
            return value;
        }
        return Spectrum(0.0f);
    }
    /* Scene::evalEnvironment (scene.h:727-730): constant.cpp:151-153 or envmap.cpp:380-410 (filtered when the ray carries differentials) */
    Spectrum evalEnvironment(const Ray &ray, const RayDiff *rd = nullptr) const {
        const Emitter &em = emitters[envEmitter];
        if (em.envmap < 0) return em.radiance;
        return envmaps[em.envmap].evalEnvironment(ray.d, rd && rd->has ? &rd->rxD : nullptr, rd && rd->has ? &rd->ryD : nullptr);
    }
    /* scene.cpp:949-952; scene.h:848-850; area.cpp:175-183; shape.cpp:117-126; trimesh.cpp:358-360 */
    /* bsphere.h:88-95 + util.cpp:447-485 */
    bool bsphereIntersect(const V3 &ro, const V3 &rd, float &x0, float &x1) const {
        V3 o = ro - bsCenter;
        float a = rd.lengthSquared(), b = 2 * dot(o, rd), c = o.lengthSquared() - bsRadius * bsRadius;
        if (a == 0) { if (b != 0) { x0 = x1 = -c / b; return true; } return false; }
        float discrim = b * b - 4.0f * a * c;
        if (discrim < 0) return false;
        float temp, sqrtDiscrim = std::sqrt(discrim);
        if (b < 0) temp = -0.5f * (b - sqrtDiscrim); else temp = -0.5f * (b + sqrtDiscrim);
        x0 = temp / a; x1 = c / temp;
        if (x0 > x1) std::swap(x0, x1);
        return true;
    }
    float pdfEmitterDirect(const DRec &dRec) const {
        const Emitter &em = emitters[dRec.emitter];
        if (em.mesh < 0 && em.envmap >= 0) { /* envmap.cpp:545-556, measure == ESolidAngle */
            const EnvMap &env = envmaps[em.envmap];
            return env.pdfDirection(EnvMap::xfVector(env.toLocal, dRec.d)) * (em.samplingWeight * emitterPDF.normalization);
        }
        if (em.mesh < 0) { /* constant.cpp:210-224, measure == ESolidAngle */
            float pdfSA = !dRec.refN.isZero() ? kInvPi * std::max(0.0f, dot(dRec.d, dRec.refN)) : kInvFourPi;
            return pdfSA * (em.samplingWeight * emitterPDF.normalization);
        }
        float pdfDirect = 0.0f;
        if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0) {
            float pdfPos = meshes[em.mesh].invSurfaceArea;
            pdfDirect = pdfPos * (dRec.dist * dRec.dist) / absDot(dRec.d, dRec.n);
        }
        return pdfDirect * (em.samplingWeight * emitterPDF.normalization);
    }

    /* Intersection::computePartials, src/librender/intersection.cpp:23-85; called through Intersection::getBSDF(ray)
       (records.inl:69-75) when the BSDF uses ray differentials */
    static void computePartials(Intersection &its, const Ray &ray, const RayDiff &rd) {
        TexCtx &t = its.tex;
        if (t.hasUVPartials || !rd.has) return;
        t.hasUVPartials = true;
        if (its.dpdu.isZero() && its.dpdv.isZero()) { t.dudx = t.dvdx = t.dudy = t.dvdy = 0.0f; return; }
        const V3 &n = its.geoFrame.n;
        const float pp = dot(n, its.p), pox = dot(n, rd.rxO), poy = dot(n, rd.ryO), prx = dot(n, rd.rxD), pry = dot(n, rd.ryD);
        if (prx == 0 || pry == 0) { t.dudx = t.dvdx = t.dudy = t.dvdy = 0.0f; return; }
        const float tx = (pp - pox) / prx, ty = (pp - poy) / pry;
        const float absX = std::abs(n.x), absY = std::abs(n.y), absZ = std::abs(n.z);
        int axes[2];
        if (absX > absY && absX > absZ) { axes[0] = 1; axes[1] = 2; }
        else if (absY > absZ) { axes[0] = 0; axes[1] = 2; }
        else { axes[0] = 0; axes[1] = 1; }
        float A[2][2], Bx[2], By[2], x[2];
        A[0][0] = its.dpdu[axes[0]]; A[0][1] = its.dpdv[axes[0]];
        A[1][0] = its.dpdu[axes[1]]; A[1][1] = its.dpdv[axes[1]];
        const V3 px = rd.rxO + rd.rxD * tx, py = rd.ryO + rd.ryD * ty;
        Bx[0] = px[axes[0]] - its.p[axes[0]]; Bx[1] = px[axes[1]] - its.p[axes[1]];
        By[0] = py[axes[0]] - its.p[axes[0]]; By[1] = py[axes[1]] - its.p[axes[1]];
        if (solveLinearSystem2x2(A, Bx, x)) { t.dudx = x[0]; t.dvdx = x[1]; } else { t.dudx = 1; t.dvdx = 0; }
        if (solveLinearSystem2x2(A, By, x)) { t.dudy = x[0]; t.dvdy = x[1]; } else { t.dudy = 1; /* sic: `dudy = 0; dudy = 1;`, dvdy stays as it was (0 here) */ }
    }
    static bool solveLinearSystem2x2(const float a[2][2], const float b[2], float x[2]) { /* src/libcore/util.cpp:527-539 */
        const float det = a[0][0] * a[1][1] - a[0][1] * a[1][0];
        if (std::abs(det) <= 2.93873587705571876e-39f) return false; /* RCPOVERFLOW_FLT */
        const float inverse = 1.0f / det;
        x[0] = (a[1][1] * b[0] - a[0][1] * b[1]) * inverse;
        x[1] = (a[0][0] * b[1] - a[1][0] * b[0]) * inverse;
        return true;
    }

    static float miWeight(float pdfA, float pdfB) { pdfA *= pdfA; pdfB *= pdfB; return pdfA / (pdfA + pdfB); }

    bool rayIntersect(const Ray &ray, Intersection &its, OrcStats &st) const {
        Hit h; ++st.rays;
        its.t = kInf;
        if (!instances.empty()) {
            int inst;
            if (topClosest(ray, h, inst, st)) { fill(ray, h, its, inst); return true; }
            return false;
        }
        if (accel.rayIntersect(ray, h, &st.nodeVisits, &st.primTests)) { fill(ray, h, its); return true; }
        return false;
    }

    /* path.cpp:119-294.  `alpha` mirrors RadianceQueryRecord::rayIntersect (records.inl:117-144). */
    Spectrum Li(const Ray &r, Sampler *sampler, const OrcRenderParams &rp, float &alpha, OrcStats &st, const RayDiff *sensorDiff = nullptr) const {
        BsdfSet bs{bsdfs.data(), (int) bsdfs.size()};
        Intersection its;
        Ray ray(r);
        RayDiff rayDiff; /* RayDifferential ray(r), path.cpp:122; plain Rays assigned later carry no differentials (ray.h:150-157) */
        if (sensorDiff) rayDiff = *sensorDiff;
        Spectrum Li(0.0f);
        bool scattered = false;
        int depth = 1;                          /* newQuery: depth = 1 (integrator.h:221-227) */
        bool emittedRadiance = true;            /* ERadiance has EEmittedRadiance until path.cpp:274 */
        rayIntersect(ray, its, st);
        alpha = its.isValid() ? 1.0f : 0.0f;
        ray.mint = kEpsilon;
        Spectrum throughput(1.0f);
        float eta = 1.0f;
        const int maxDepth = rp.maxDepth, rrDepth = rp.rrDepth;
        while (depth <= maxDepth || maxDepth < 0) {
            if (!its.isValid()) { /* path.cpp:136-143 */
                if (envEmitter >= 0 && emittedRadiance && (!rp.hideEmitters || scattered)) Li += throughput * evalEnvironment(ray, &rayDiff);
                break;
            }
            const Mesh &mesh = meshes[its.mesh];
            const int bsdf = mesh.bsdf;
            if (mesh.emitter >= 0 && emittedRadiance && (!rp.hideEmitters || scattered))
                Li += throughput * emitterEval(mesh.emitter, its, -ray.d);
            if ((depth >= maxDepth && maxDepth > 0) ||
                (rp.strictNormals && dot(ray.d, its.geoFrame.n) * Frame::cosTheta(its.wi) >= 0))
                break;
            const uint32_t btype = bs.type(bsdf);
            if (bs.usesRayDifferentials(bsdf)) computePartials(its, ray, rayDiff); /* its.getBSDF(ray), path.cpp:162 */
            /* DirectSamplingRecord(its): records.inl:156-164 */
            DRec dRec;
            dRec.ref = its.p; dRec.refN = V3(0.0f);
            if ((btype & (ETransmission | EBackSide)) == 0) dRec.refN = its.shFrame.n;
            if (!emitters.empty() && (btype & ESmooth)) {
                float sx, sy; sampler->next2D(sx, sy);
                Spectrum value = sampleEmitterDirect(dRec, sx, sy, st);
                if (!value.isZero()) {
                    BRec bRec; bRec.wi = its.wi; bRec.wo = its.shFrame.toLocal(dRec.d); bRec.sampler = sampler; bRec.its = &its.tex;
                    const Spectrum bsdfVal = bs.eval(bsdf, bRec);
                    if (!bsdfVal.isZero() && (!rp.strictNormals || dot(its.geoFrame.n, dRec.d) * Frame::cosTheta(bRec.wo) > 0)) {
                        float bsdfPdf = bs.pdf(bsdf, bRec); /* area emitter: onSurface && solid angle */
                        float weight = miWeight(dRec.pdf, bsdfPdf);
                        Li += throughput * value * bsdfVal * weight;
                    }
                }
            } else if (emitters.empty() && (btype & ESmooth)) {
                /* scene without emitters: reference would still draw the 2D sample; sampleReuse on an
                   empty PMF is undefined there -- not a supported configuration */
            }
            float bsdfPdf;
            BRec bRec; bRec.wi = its.wi; bRec.sampler = sampler; bRec.its = &its.tex;
            float sx, sy; sampler->next2D(sx, sy);
            Spectrum bsdfWeight = bs.sample(bsdf, bRec, bsdfPdf, sx, sy);
            if (bsdfWeight.isZero()) break;
            scattered |= bRec.sampledType != ENull;
            const V3 wo = its.shFrame.toWorld(bRec.wo);
            float woDotGeoN = dot(its.geoFrame.n, wo);
            if (rp.strictNormals && woDotGeoN * Frame::cosTheta(bRec.wo) <= 0) break;
            bool hitEmitter = false;
            Spectrum value;
            ray = Ray(its.p, wo);
            rayDiff.has = false;
            if (rayIntersect(ray, its, st)) {
                const Mesh &m2 = meshes[its.mesh];
                if (m2.emitter >= 0) {
                    value = emitterEval(m2.emitter, its, -ray.d);
                    /* dRec.setQuery(ray, its): records.inl:171-179 */
                    dRec.p = its.p; dRec.n = its.shFrame.n; dRec.solidAngle = true; dRec.emitter = m2.emitter;
                    dRec.d = ray.d; dRec.dist = its.t;
                    hitEmitter = true;
                }
            } else { /* path.cpp:239-252 */
                if (envEmitter < 0) break;
                if (rp.hideEmitters && !scattered) break;
                value = evalEnvironment(ray);
                /* fillDirectSamplingRecord (constant.cpp:239-253): the ray starts inside the bounding sphere */
                float nearT, farT;
                if (!bsphereIntersect(ray.o, ray.d, nearT, farT) || nearT > 0 || farT < 0) break;
                dRec.p = ray(farT); dRec.n = normalize(bsCenter - dRec.p); dRec.solidAngle = true; dRec.emitter = envEmitter;
                dRec.d = ray.d; dRec.dist = farT;
                hitEmitter = true;
            }
            throughput *= bsdfWeight;
            eta *= bRec.eta;
            if (hitEmitter) {
                const float lumPdf = (!(bRec.sampledType & EDelta)) ? pdfEmitterDirect(dRec) : 0;
                Li += throughput * value * miWeight(bsdfPdf, lumPdf);
            }
            if (!its.isValid()) break;
            emittedRadiance = false;            /* rRec.type = ERadianceNoEmission */
            if (depth++ >= rrDepth) {
                float q = std::min(throughput.max() * eta * eta, 0.95f);
                if (sampler->next1D() >= q) break;
                throughput /= q;
            }
        }
        st.pathLengthSum += (uint64_t) depth;
        return Li;
    }


    /* ---------------------------------------------------------------------------------------------------------
     * volpath: src/integrators/path/volpath.cpp:84-366 (Li), :368-426 (rayIntersectAndLookForEmitter),
     * Scene::evalTransmittance scene.cpp:619-679, Scene::sampleAttenuatedEmitterDirect scene.cpp:854-898
     * --------------------------------------------------------------------------------------------------------- */
    int targetMedium(const Mesh &m, const V3 &geoN, const V3 &d) const { /* records.inl:81-86 */
        return dot(d, geoN) > 0 ? m.exterior : m.interior;
    }
    /* ShapeKDTree::rayIntersect(ray, t, shape, n, uv), skdtree.cpp:144-204: closest hit, epsilon scale without the
       inner max, unflipped face normal */
    bool rayIntersectNormal(const Ray &ray, float &t, int &mesh, V3 &n, OrcStats &st) const {
        float mint, maxt;
        t = kInf;
        ++st.shadowRays;
        if (accel.aabb.rayIntersect(ray, mint, maxt)) {
            float rayMinT = ray.mint;
            if (rayMinT == kEpsilon) rayMinT *= std::max(std::max(std::abs(ray.o.x), std::abs(ray.o.y)), std::abs(ray.o.z));
            if (rayMinT > mint) mint = rayMinT;
            if (ray.maxt < maxt) maxt = ray.maxt;
            if (maxt > mint) {
                Hit h;
                if (accel.query<false>(ray, mint, maxt, h, &st.nodeVisits, &st.primTests)) {
                    t = h.t;
                    const uint32_t mi = accel.tri[h.prim].shapeIndex, pi = accel.tri[h.prim].primIndex;
                    const Mesh &m = meshes[mi];
                    const V3 &p0 = m.P[m.idx[3 * pi]], &p1 = m.P[m.idx[3 * pi + 1]], &p2 = m.P[m.idx[3 * pi + 2]];
                    n = normalize(cross(p1 - p0, p2 - p0));
                    mesh = (int) mi;
                    return true;
                }
            }
        }
        return false;
    }
    Spectrum evalTransmittance(const V3 &p1, bool p1OnSurface, const V3 &p2, bool p2OnSurface, int medium, int &interactions,
                               Sampler *sampler, OrcStats &st) const {
        BsdfSet bs{bsdfs.data(), (int) bsdfs.size()};
        V3 d = p2 - p1;
        float remaining = d.length();
        d /= remaining;
        float lengthFactor = p2OnSurface ? (1 - kShadowEpsilon) : 1;
        Ray ray(p1, d, p1OnSurface ? kEpsilon : 0, remaining * lengthFactor);
        Spectrum transmittance(1.0f);
        int maxInteractions = interactions;
        interactions = 0;
        while (remaining > 0) {
            float t; int mesh = -1; V3 n;
            bool surface = rayIntersectNormal(ray, t, mesh, n, st);
            if (surface && (interactions == maxInteractions || !(bs.type(meshes[mesh].bsdf) & ENull))) return Spectrum(0.0f);
            if (medium >= 0)
                transmittance *= MediumEval(media[medium]).evalTransmittance(Ray(ray.o, ray.d, 0, std::min(t, remaining)), sampler);
            if (!surface || transmittance.isZero()) break;
            /* null BSDF: eval(bRec, EDiscrete) with typeMask = ENull is 1 (null.cpp:45-47) */
            const Mesh &m = meshes[mesh];
            if (m.isMediumTransition()) {
                if (medium != targetMedium(m, n, -d)) return Spectrum(0.0f); /* mediumInconsistencies */
                medium = targetMedium(m, n, d);
            }
            if (++interactions > 100) break;
            ray.o = ray(t);
            remaining -= t;
            ray.maxt = remaining * lengthFactor;
            ray.mint = kEpsilon;
        }
        return transmittance;
    }
    void rayIntersectAndLookForEmitter(Sampler *sampler, int medium, int maxInteractions, Ray ray, Intersection &_its, DRec &dRec,
                                       Spectrum &value, OrcStats &st) const {
        BsdfSet bs{bsdfs.data(), (int) bsdfs.size()};
        Intersection its2, *its = &_its;
        Spectrum transmittance(1.0f);
        bool surface = false;
        int interactions = 0;
        while (true) {
            surface = rayIntersect(ray, *its, st);
            if (medium >= 0)
                transmittance *= MediumEval(media[medium]).evalTransmittance(Ray(ray.o, ray.d, 0, its->t), sampler);
            if (surface && (interactions == maxInteractions || !(bs.type(meshes[its->mesh].bsdf) & ENull) || meshes[its->mesh].emitter >= 0)) break;
            if (!surface) break;
            if (transmittance.isZero()) return;
            const Mesh &m = meshes[its->mesh];
            if (m.isMediumTransition()) medium = targetMedium(m, its->geoFrame.n, ray.d);
            /* transmittance *= bsdf->eval(bRec, EDiscrete) = 1 for the null BSDF */
            ray.o = ray(its->t);
            ray.mint = kEpsilon;
            its = &its2;
            if (++interactions > 100) return;
        }
        if (surface) {
            const Mesh &m = meshes[its->mesh];
            if (m.emitter >= 0) {
                /* dRec.setQuery(ray, its): records.inl:171-179 */
                dRec.p = its->p; dRec.n = its->shFrame.n; dRec.solidAngle = true; dRec.emitter = m.emitter;
                dRec.d = ray.d; dRec.dist = its->t;
                value = transmittance * emitterEval(m.emitter, *its, -ray.d);
            }
        } else if (envEmitter >= 0) { /* volpath.cpp:418-424: env->fillDirectSamplingRecord(dRec, ray) && evalEnvironment */
            float nearT, farT;
            if (bsphereIntersect(ray.o, ray.d, nearT, farT) && !(nearT > 0) && !(farT < 0)) {
                dRec.p = ray(farT); dRec.n = normalize(bsCenter - dRec.p); dRec.solidAngle = true; dRec.emitter = envEmitter;
                dRec.d = ray.d; dRec.dist = farT;
                value = transmittance * evalEnvironment(ray);
            }
        }
    }
    Spectrum LiVol(const Ray &r, Sampler *sampler, const OrcRenderParams &rp, float &alpha, OrcStats &st, const RayDiff *sensorDiff = nullptr) const {
        BsdfSet bs{bsdfs.data(), (int) bsdfs.size()};
        Intersection its;
        Ray ray(r);
        RayDiff rayDiff; /* RayDifferential ray(r), volpath.cpp:89: only the sensor ray carries differentials (they reach envmap's filtered look-up) */
        if (sensorDiff) rayDiff = *sensorDiff;
        Spectrum Li(0.0f);
        float eta = 1.0f;
        int depth = 1;
        int medium = -1;                        /* sensor medium: vacuum (a camera inside a medium is not in scope) */
        bool emittedRadiance = true;            /* EEmittedRadiance bit of rRec.type */
        rayIntersect(ray, its, st);
        alpha = its.isValid() ? 1.0f : 0.0f;    /* Simpson-integrated opacity of index-matched boundaries not restated */
        Spectrum throughput(1.0f);
        bool scattered = false;
        const int maxDepth = rp.maxDepth, rrDepth = rp.rrDepth;
        while (depth <= maxDepth || maxDepth < 0) {
            MediumSamplingRecord mRec;
            if (medium >= 0 && MediumEval(media[medium]).sampleDistance(Ray(ray.o, ray.d, 0, its.t), mRec, sampler)) {
                MediumEval me(media[medium]);
                if (depth >= maxDepth && maxDepth != -1) break;
                throughput *= mRec.sigmaS * mRec.transmittance / mRec.pdfSuccess;
                /* luminaire sampling */
                DRec dRec; dRec.ref = mRec.p; dRec.refN = V3(0.0f);
                if (!emitters.empty()) {
                    int interactions = maxDepth - depth - 1;
                    float sx, sy; sampler->next2D(sx, sy);
                    float emPdf = 1;
                    Spectrum value = sampleEmitterDirect(dRec, sx, sy, st, false, &emPdf);
                    if (dRec.pdf != 0) {
                        value *= evalTransmittance(dRec.ref, false, dRec.p, true, medium, interactions, sampler, st) / emPdf;
                        dRec.pdf *= emPdf;
                    } else value = Spectrum(0.0f);
                    if (!value.isZero()) {
                        float phaseVal = me.phaseEval(-ray.d, dRec.d);
                        if (phaseVal != 0) {
                            float phasePdf = phaseVal; /* PhaseFunction::pdf = eval (phase.cpp:21-23); area emitter: on surface, solid angle */
                            const float weight = miWeight(dRec.pdf, phasePdf);
                            Li += throughput * value * phaseVal * weight;
                        }
                    }
                }
                /* phase function sampling */
                float phasePdf; V3 wo;
                float phaseVal = me.phaseSample(-ray.d, wo, phasePdf, sampler);
                if (phaseVal == 0) break;
                throughput *= phaseVal;
                ray = Ray(mRec.p, wo, 0, kInf);
                rayDiff.has = false;
                Spectrum value(0.0f);
                rayIntersectAndLookForEmitter(sampler, medium, maxDepth - depth - 1, ray, its, dRec, value, st);
                if (!value.isZero()) {
                    const float emitterPdf = pdfEmitterDirect(dRec);
                    Li += throughput * value * miWeight(phasePdf, emitterPdf);
                }
                emittedRadiance = false;
            } else {
                if (medium >= 0) throughput *= mRec.transmittance / mRec.pdfFailure;
                if (!its.isValid()) { /* volpath.cpp:190-202 */
                    if (envEmitter >= 0 && emittedRadiance && (!rp.hideEmitters || scattered)) {
                        Spectrum value = throughput * evalEnvironment(ray, &rayDiff);
                        if (medium >= 0) value *= MediumEval(media[medium]).evalTransmittance(ray, sampler);
                        Li += value;
                    }
                    break;
                }
                const Mesh &mesh = meshes[its.mesh];
                const int bsdf = mesh.bsdf;
                if (mesh.emitter >= 0 && emittedRadiance && (!rp.hideEmitters || scattered))
                    Li += throughput * emitterEval(mesh.emitter, its, -ray.d);
                if (depth >= maxDepth && maxDepth != -1) break;
                float wiDotGeoN = -dot(its.geoFrame.n, ray.d), wiDotShN = Frame::cosTheta(its.wi);
                if (wiDotGeoN * wiDotShN < 0 && rp.strictNormals) break;
                const uint32_t btype = bs.type(bsdf);
                DRec dRec; dRec.ref = its.p; dRec.refN = V3(0.0f);
                if ((btype & (ETransmission | EBackSide)) == 0) dRec.refN = its.shFrame.n;
                if (!emitters.empty() && (btype & ESmooth)) {
                    int interactions = maxDepth - depth - 1;
                    float sx, sy; sampler->next2D(sx, sy);
                    float emPdf = 1;
                    Spectrum value = sampleEmitterDirect(dRec, sx, sy, st, false, &emPdf);
                    if (dRec.pdf != 0) {
                        int med = medium;
                        if (mesh.isMediumTransition()) med = targetMedium(mesh, its.geoFrame.n, dRec.d);
                        value *= evalTransmittance(its.p, true, dRec.p, true, med, interactions, sampler, st) / emPdf;
                        dRec.pdf *= emPdf;
                    } else value = Spectrum(0.0f);
                    if (!value.isZero()) {
                        BRec bRec; bRec.wi = its.wi; bRec.wo = its.shFrame.toLocal(dRec.d); bRec.sampler = sampler;
                        const Spectrum bsdfVal = bs.eval(bsdf, bRec);
                        float woDotGeoN = dot(its.geoFrame.n, dRec.d);
                        if (!bsdfVal.isZero() && (!rp.strictNormals || woDotGeoN * Frame::cosTheta(bRec.wo) > 0)) {
                            float bsdfPdf = bs.pdf(bsdf, bRec);
                            const float weight = miWeight(dRec.pdf, bsdfPdf);
                            Li += throughput * value * bsdfVal * weight;
                        }
                    }
                }
                /* BSDF sampling */
                BRec bRec; bRec.wi = its.wi; bRec.sampler = sampler;
                float bsdfPdf;
                float sx, sy; sampler->next2D(sx, sy);
                Spectrum bsdfWeight = bs.sample(bsdf, bRec, bsdfPdf, sx, sy);
                if (bsdfWeight.isZero()) break;
                const V3 wo = its.shFrame.toWorld(bRec.wo);
                float woDotGeoN = dot(its.geoFrame.n, wo);
                if (woDotGeoN * Frame::cosTheta(bRec.wo) <= 0 && rp.strictNormals) break;
                ray = Ray(its.p, wo);
                rayDiff.has = false;
                throughput *= bsdfWeight;
                eta *= bRec.eta;
                if (mesh.isMediumTransition()) medium = targetMedium(mesh, its.geoFrame.n, ray.d);
                if (bRec.sampledType == ENull) {
                    emittedRadiance = !scattered; /* rRec.type = scattered ? ERadianceNoEmission : ERadiance */
                    rayIntersect(ray, its, st);
                    depth++;
                    continue;
                }
                Spectrum value(0.0f);
                rayIntersectAndLookForEmitter(sampler, medium, maxDepth - depth - 1, ray, its, dRec, value, st);
                if (!value.isZero()) {
                    const float emitterPdf = (!(bRec.sampledType & EDelta)) ? pdfEmitterDirect(dRec) : 0;
                    Li += throughput * value * miWeight(bsdfPdf, emitterPdf);
                }
                emittedRadiance = false;
            }
            if (depth++ >= rrDepth) {
                float q = std::min(throughput.max() * eta * eta, 0.95f);
                if (sampler->next1D() >= q) break;
                throughput /= q;
            }
            scattered = true;
            /* not in the reference (it raises an error, sobol.cpp:223-225): a path that ran past the Sobol' table ends here */
            if (sampler->exhausted()) break;
        }
        st.pathLengthSum += (uint64_t) depth;
        return Li;
    }

    V3 sampleToCameraPoint(float px, float py, float pz) const { /* Transform::operator()(Point), transform.h:108-125 */
        const float *M = sampleToCamera;
        float x = M[0] * px + M[1] * py + M[2] * pz + M[3];
        float y = M[4] * px + M[5] * py + M[6] * pz + M[7];
        float z = M[8] * px + M[9] * py + M[10] * pz + M[11];
        float w = M[12] * px + M[13] * py + M[14] * pz + M[15];
        V3 r(x, y, z);
        if (w != 1.0f) r = r / w;
        return r;
    }
    V3 camVectorToWorld(const V3 &d) const {
        const float *T = camToWorld;
        return V3(T[0] * d.x + T[1] * d.y + T[2] * d.z, T[4] * d.x + T[5] * d.y + T[6] * d.z, T[8] * d.x + T[9] * d.y + T[10] * d.z);
    }
    /* perspective.cpp:271-298 / thinlens.cpp:324-361 sampleRayDifferential; `diff` (optional) receives the offset rays, already
       scaled by diffScale = 1/sqrt(sampleCount) as SamplingIntegrator::renderBlock does (integrator.cpp:144-145,181; ray.h:163-168) */
    Ray sampleRay(float sxp, float syp, float apx = 0.5f, float apy = 0.5f, RayDiff *diff = nullptr, float diffScale = 1.0f) const {
        const float invResX = 1.0f / (float) W, invResY = 1.0f / (float) H; /* m_invResolution, sensor.cpp:104-107 */
        const V3 nearP = sampleToCameraPoint(sxp * invResX, syp * invResY, 0.0f);
        V3 dx, dy; /* m_dx, m_dy: perspective.cpp:160-163 */
        if (diff) {
            const V3 zero = sampleToCameraPoint(0.0f, 0.0f, 0.0f);
            dx = sampleToCameraPoint(invResX, 0.0f, 0.0f) - zero;
            dy = sampleToCameraPoint(0.0f, invResY, 0.0f) - zero;
        }
        auto finish = [&](const V3 &o, const V3 &d, const V3 &rxd, const V3 &ryd) {
            if (!diff) return;
            diff->has = true;
            diff->rxO = o; diff->ryO = o; /* rxOrigin = ryOrigin = ray.o */
            diff->rxD = d + (rxd - d) * diffScale;
            diff->ryD = d + (ryd - d) * diffScale;
        };
        if (apertureRadius > 0) { /* thinlens.cpp:327-350 */
            float tx, ty;
            squareToUniformDiskConcentric(apx, apy, tx, ty);
            tx *= apertureRadius; ty *= apertureRadius;
            V3 apertureP(tx, ty, 0.0f);
            V3 focusP = nearP * (focusDistance / nearP.z);
            V3 d = normalize(focusP - apertureP);
            float invZ = 1.0f / d.z;
            const float *T = camToWorld;
            V3 ow(T[0] * tx + T[1] * ty + T[2] * 0.0f + T[3], T[4] * tx + T[5] * ty + T[6] * 0.0f + T[7], T[8] * tx + T[9] * ty + T[10] * 0.0f + T[11]);
            V3 dw(T[0] * d.x + T[1] * d.y + T[2] * d.z, T[4] * d.x + T[5] * d.y + T[6] * d.z, T[8] * d.x + T[9] * d.y + T[10] * d.z);
            if (diff) { /* thinlens.cpp:341-357 */
                const float fDist = focusDistance / nearP.z;
                const V3 focusPx = (nearP + dx) * fDist, focusPy = (nearP + dy) * fDist;
                finish(ow, dw, camVectorToWorld(normalize(focusPx - apertureP)), camVectorToWorld(normalize(focusPy - apertureP)));
            }
            return Ray(ow, dw, nearClip * invZ, farClip * invZ);
        }
        V3 d = normalize(nearP);
        float invZ = 1.0f / d.z;
        const V3 dw = camVectorToWorld(d);
        if (diff) finish(camOrigin, dw, camVectorToWorld(normalize(nearP + dx)), camVectorToWorld(normalize(nearP + dy))); /* perspective.cpp:291-295 */
        return Ray(camOrigin, dw, nearClip * invZ, farClip * invZ);
    }
};

/* src/libcore/rfilter.cpp:37-57 + rfilters/{box,gaussian}.cpp */
struct RFilter {
    float radius, values[32], scaleFactor; int borderSize;
    RFilter(int kind, float param) {
        float stddev = param;
        if (kind == 0) radius = param + 1e-5f;          /* box.cpp:41 */
        else radius = 4 * stddev;                        /* gaussian.cpp:38 */
        float sum = 0.0f;
        for (int i = 0; i < 31; ++i) {
            float x = (radius * i) / 31, value;
            if (kind == 0) value = std::abs(x) <= radius ? 1.0f : 0.0f;
            else {
                float alpha = -1.0f / (2.0f * stddev * stddev);
                value = std::max(0.0f, fastexp(alpha * x * x) - fastexp(alpha * radius * radius));
            }
            values[i] = value; sum += value;
        }
        values[31] = 0.0f;
        scaleFactor = 31 / radius;
        borderSize = (int) std::ceil(radius - 0.5f);
        sum *= 2 * radius / 31;
        float normalization = 1.0f / sum;
        for (int i = 0; i < 31; ++i) values[i] *= normalization;
    }
    float evalDiscretized(float x) const { return values[std::min((int) std::abs(x * scaleFactor), 31)]; }
};

/* imageblock.h: a block of (size + 2*border)^2 x 5 floats */
struct ImageBlock {
    int ox, oy, sx, sy, border; std::vector<float> data; const RFilter *f;
    int bw() const { return sx + 2 * border; }
    int bh() const { return sy + 2 * border; }
    ImageBlock(int ox_, int oy_, int sx_, int sy_, const RFilter *f_) : ox(ox_), oy(oy_), sx(sx_), sy(sy_), border(f_->borderSize), f(f_) {
        data.assign((size_t) bw() * bh() * 5, 0.0f);
    }
    /* imageblock.h:124-204 */
    bool put(float px, float py, const Spectrum &spec, float alpha) {
        float value[5] = {spec.x, spec.y, spec.z, alpha, 1.0f};
        for (int i = 0; i < 5; ++i)
            if (!std::isfinite(value[i]) || value[i] < 0) return false;
        const float filterRadius = f->radius;
        const int sizeX = bw(), sizeY = bh();
        const float posx = px - 0.5f - (ox - border), posy = py - 0.5f - (oy - border);
        const int minx = std::max((int) std::ceil(posx - filterRadius), 0), miny = std::max((int) std::ceil(posy - filterRadius), 0),
                  maxx = std::min((int) std::floor(posx + filterRadius), sizeX - 1), maxy = std::min((int) std::floor(posy + filterRadius), sizeY - 1);
        float wX[64], wY[64]; /* m_weightsX/Y: 2*ceil(radius)+1 entries; radius <= 31 here */
        for (int x = minx, i = 0; x <= maxx; ++x) wX[i++] = f->evalDiscretized(x - posx);
        for (int y = miny, i = 0; y <= maxy; ++y) wY[i++] = f->evalDiscretized(y - posy);
        for (int y = miny, yr = 0; y <= maxy; ++y, ++yr) {
            const float weightY = wY[yr];
            float *dest = data.data() + ((size_t) y * sizeX + minx) * 5;
            for (int x = minx, xr = 0; x <= maxx; ++x, ++xr) {
                const float weight = wX[xr] * weightY;
                for (int k = 0; k < 5; ++k) *dest++ += weight * value[k];
            }
        }
        return true;
    }
};

} // namespace

extern "C" {

void *orc_scene_new() { return new Scene(); }
void orc_scene_free(void *s) { delete (Scene *) s; }
void orc_set_sobol_tables(void *s, const uint32_t *m32, const uint64_t *vdc, const uint64_t *inv) {
    Scene *sc = (Scene *) s; sc->sobol.m32 = m32; sc->sobol.vdc = vdc; sc->sobol.inv = inv;
}
int orc_add_bsdf(void *s, const OrcBsdf *b) { Scene *sc = (Scene *) s; sc->bsdfs.push_back(*b); return (int) sc->bsdfs.size() - 1; }
int orc_add_mesh(void *s, const float *P, const float *N, const float *UV, uint32_t nV, const uint32_t *idx, uint32_t nT,
                 int bsdf, const float *radiance, float samplingWeight) {
    Scene *sc = (Scene *) s;
    Mesh m;
    m.P.resize(nV);
    for (uint32_t i = 0; i < nV; ++i) m.P[i] = V3(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
    if (N) { m.N.resize(nV); for (uint32_t i = 0; i < nV; ++i) m.N[i] = V3(N[3 * i], N[3 * i + 1], N[3 * i + 2]); }
    if (UV) m.UV.assign(UV, UV + 2 * nV);
    m.idx.assign(idx, idx + 3 * nT);
    m.bsdf = bsdf;
    if (radiance) {
        Emitter e; e.radiance = V3(radiance[0], radiance[1], radiance[2]); e.samplingWeight = samplingWeight;
        e.mesh = (int) sc->meshes.size();
        m.emitter = (int) sc->emitters.size();
        sc->emitters.push_back(e);
    }
    sc->meshes.push_back(std::move(m));
    return (int) sc->meshes.size() - 1;
}
/* Medium plugin instance (homogeneous / heterogeneous + phase function); the density grid is copied */
int orc_add_medium(void *s, const OrcMedium *m) {
    Scene *sc = (Scene *) s;
    OrcMedium mm = *m;
    sc->mediaData.emplace_back();
    if (m->type == 1 && m->density) {
        size_t n = (size_t) m->res[0] * m->res[1] * m->res[2];
        sc->mediaData.back().assign(m->density, m->density + n);
    }
    sc->media.push_back(mm);
    for (size_t i = 0; i < sc->media.size(); ++i) sc->media[i].density = sc->mediaData[i].empty() ? nullptr : sc->mediaData[i].data();
    return (int) sc->media.size() - 1;
}
/* <ref name="interior"/"exterior"> of a shape (shape.cpp:160-176) */
void orc_set_mesh_media(void *s, int mesh, int interior, int exterior) {
    Scene *sc = (Scene *) s; sc->meshes[mesh].interior = interior; sc->meshes[mesh].exterior = exterior;
}
/* component probes for the medium tests: n x (o, d, mint, maxt) rays */
void orc_medium_transmittance(void *s, int medium, uint64_t n, const float *rays, uint64_t seed, float *out) {
    Scene *sc = (Scene *) s;
    CounterSampler smp(1 << 16, 1, seed);
    for (uint64_t i = 0; i < n; ++i) {
        smp.generate((int) (i & 0xFFFF), (int) (i >> 16));
        const float *r = rays + 8 * i;
        Spectrum t = MediumEval(sc->media[medium]).evalTransmittance(Ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3], r[7]), &smp);
        out[3 * i] = t.x; out[3 * i + 1] = t.y; out[3 * i + 2] = t.z;
    }
}
void orc_medium_sample_distance(void *s, int medium, uint64_t n, const float *rays, uint64_t seed, float *out /* n x 12: ok, t, sigmaS rgb, transmittance rgb, pdfSuccess, pdfFailure, sigmaA.r, pad */) {
    Scene *sc = (Scene *) s;
    CounterSampler smp(1 << 16, 1, seed);
    for (uint64_t i = 0; i < n; ++i) {
        smp.generate((int) (i & 0xFFFF), (int) (i >> 16));
        const float *r = rays + 8 * i;
        MediumSamplingRecord mRec;
        bool ok = MediumEval(sc->media[medium]).sampleDistance(Ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3], r[7]), mRec, &smp);
        float *o = out + 12 * i;
        o[0] = ok ? 1.0f : 0.0f; o[1] = mRec.t; o[2] = mRec.sigmaS.x; o[3] = mRec.sigmaS.y; o[4] = mRec.sigmaS.z;
        o[5] = mRec.transmittance.x; o[6] = mRec.transmittance.y; o[7] = mRec.transmittance.z; o[8] = mRec.pdfSuccess; o[9] = mRec.pdfFailure;
        o[10] = mRec.sigmaA.x; o[11] = 0;
    }
}
void orc_medium_density(void *s, int medium, uint64_t n, const float *p, float *out) {
    Scene *sc = (Scene *) s;
    for (uint64_t i = 0; i < n; ++i) out[i] = MediumEval(sc->media[medium]).lookupDensity(V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}
void orc_phase(void *s, int medium, uint64_t n, const float *wi, const float *samples, float *out /* n x 5: wo xyz, pdf, eval(wi, wo) */) {
    Scene *sc = (Scene *) s;
    for (uint64_t i = 0; i < n; ++i) {
        struct Two : Sampler { float a, b; int k = 0; void generate(int, int) override {} void advance() override {}
            float next1D() override { return (k++ & 1) ? b : a; } void next2D(float &x, float &y) override { x = a; y = b; } } two;
        two.a = samples[2 * i]; two.b = samples[2 * i + 1];
        MediumEval me(sc->media[medium]);
        V3 w(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), wo; float pdf;
        me.phaseSample(w, wo, pdf, &two);
        float *o = out + 5 * i; o[0] = wo.x; o[1] = wo.y; o[2] = wo.z; o[3] = pdf; o[4] = me.phaseEval(w, wo);
    }
}
/* <emitter type="constant">: radiance, samplingWeight (constant.cpp:47-52); at most one environment emitter (scene.cpp:510-514) */
int orc_add_constant_emitter(void *s, const float *radiance, float samplingWeight) {
    Scene *sc = (Scene *) s;
    Emitter e; e.radiance = V3(radiance[0], radiance[1], radiance[2]); e.samplingWeight = samplingWeight; e.mesh = -1;
    sc->emitters.push_back(e);
    return (int) sc->emitters.size() - 1;
}
/* <emitter type="envmap">: linear float RGB pixels (row-major, top row first), scale, toWorld and its inverse, samplingWeight
   (envmap.cpp:106-181).  Returns the emitter index, or -1 when the map is black / not finite (envmap.cpp:311-315) */
int orc_add_envmap_emitter(void *s, int width, int height, const float *pixels, float scale, const float *toWorld, const float *toLocal, float samplingWeight) {
    Scene *sc = (Scene *) s;
    sc->envmaps.emplace_back();
    if (!sc->envmaps.back().build(width, height, pixels, scale, toWorld, toLocal)) { sc->envmaps.pop_back(); return -1; }
    Emitter e; e.radiance = V3(0.0f); e.samplingWeight = samplingWeight; e.mesh = -1; e.envmap = (int) sc->envmaps.size() - 1;
    sc->emitters.push_back(e);
    return (int) sc->emitters.size() - 1;
}
/* the pyramid and the tables of environment map `env`: info = levels, then (w, h) per level */
void orc_envmap_info(void *s, int env, int32_t *info, float *normalization) {
    const EnvMap &e = ((Scene *) s)->envmaps[env];
    info[0] = e.mip.levels;
    for (int l = 0; l < e.mip.levels; ++l) { info[1 + 2 * l] = e.mip.lw[l]; info[2 + 2 * l] = e.mip.lh[l]; }
    *normalization = e.normalization;
}
void orc_envmap_level(void *s, int env, int level, float *out) {
    const EnvMap &e = ((Scene *) s)->envmaps[env];
    memcpy(out, e.mip.pyramid[level].data(), e.mip.pyramid[level].size() * sizeof(float));
}
void orc_envmap_tables(void *s, int env, float *cdfRows, float *cdfCols, float *rowWeights) {
    const EnvMap &e = ((Scene *) s)->envmaps[env];
    memcpy(cdfRows, e.cdfRows.data(), e.cdfRows.size() * sizeof(float));
    memcpy(cdfCols, e.cdfCols.data(), e.cdfCols.size() * sizeof(float));
    memcpy(rowWeights, e.rowWeights.data(), e.rowWeights.size() * sizeof(float));
}
/* Scene::evalEnvironment for n rays after commit: rays 6n (o, d), or 18n (o, d, rxO, rxD, ryO, ryD) with differentials -> out 3n */
void orc_eval_environment(void *s, uint64_t n, int withDifferentials, const float *rays, float *out) {
    Scene *sc = (Scene *) s;
    const int stride = withDifferentials ? 18 : 6;
    for (uint64_t i = 0; i < n; ++i) {
        const float *r = rays + (size_t) stride * i;
        Ray ray(V3(r[0], r[1], r[2]), V3(r[3], r[4], r[5]));
        RayDiff rd;
        if (withDifferentials) { rd.has = true; rd.rxO = V3(r[6], r[7], r[8]); rd.rxD = V3(r[9], r[10], r[11]); rd.ryO = V3(r[12], r[13], r[14]); rd.ryD = V3(r[15], r[16], r[17]); }
        const Spectrum v = sc->envEmitter >= 0 ? sc->evalEnvironment(ray, &rd) : Spectrum(0.0f);
        out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
    }
}
/* Scene::pdfEmitterDirect for the environment emitter: ref 6n (ref, refN), d 3n -> out n */
void orc_pdf_environment_direct(void *s, uint64_t n, const float *ref, const float *d, float *out) {
    Scene *sc = (Scene *) s;
    for (uint64_t i = 0; i < n; ++i) {
        DRec dRec;
        dRec.ref = V3(ref[6 * i], ref[6 * i + 1], ref[6 * i + 2]); dRec.refN = V3(ref[6 * i + 3], ref[6 * i + 4], ref[6 * i + 5]);
        dRec.d = V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]); dRec.emitter = sc->envEmitter; dRec.solidAngle = true;
        out[i] = sc->envEmitter >= 0 ? sc->pdfEmitterDirect(dRec) : 0.0f;
    }
}
/* <shape type="shapegroup"> / <shape type="instance">: meshes added with orc_set_mesh_group live in the group's object space */
int orc_add_shapegroup(void *s) { Scene *sc = (Scene *) s; sc->groups.emplace_back(); return (int) sc->groups.size() - 1; }
void orc_set_mesh_group(void *s, int mesh, int group) { ((Scene *) s)->meshes[mesh].group = group; }
int orc_add_instance(void *s, int group, const float *M, const float *Minv) {
    Scene *sc = (Scene *) s;
    Scene::Instance in; in.group = group; memcpy(in.M, M, 64); memcpy(in.Minv, Minv, 64);
    sc->instances.push_back(in);
    return (int) sc->instances.size() - 1;
}
void orc_set_camera(void *s, const float *camToWorld, const float *sampleToCamera, float nearClip, float farClip, int W, int H) {
    Scene *sc = (Scene *) s;
    memcpy(sc->camToWorld, camToWorld, 64); memcpy(sc->sampleToCamera, sampleToCamera, 64);
    sc->nearClip = nearClip; sc->farClip = farClip; sc->W = W; sc->H = H;
}
/* <sensor type="thinlens">: apertureRadius, focusDistance (thinlens.cpp:132-142, sensor.cpp:162) */
void orc_set_thinlens(void *s, float apertureRadius, float focusDistance) { Scene *sc = (Scene *) s; sc->apertureRadius = apertureRadius; sc->focusDistance = focusDistance; }
void orc_commit(void *s, int useTree) { ((Scene *) s)->commit(useTree != 0); }
void orc_accel_info(void *s, uint64_t *out /* nTri, nNodes, nIndices, nLeaves, maxDepth */, float *aabb6) {
    Scene *sc = (Scene *) s;
    out[0] = sc->accel.tri.size(); out[1] = sc->accel.nodes.size(); out[2] = sc->accel.indices.size();
    out[3] = sc->accel.nLeaves; out[4] = (uint64_t) sc->accel.maxDepth;
    for (int i = 0; i < 3; ++i) { aabb6[i] = sc->accel.aabb.min[i]; aabb6[3 + i] = sc->accel.aabb.max[i]; }
}
void orc_get_triaccel(void *s, void *out48) { Scene *sc = (Scene *) s; memcpy(out48, sc->accel.tri.data(), sc->accel.tri.size() * sizeof(TriAccel)); }

/* rays: n x 8 floats (o.xyz, mint, d.xyz, maxt).  mode 0 closest (out t,u,v,prim), 1 occlusion (prim = 0/1).
 * accelMode: -1 scene default, 0 brute force, 1 tree */
void orc_trace(void *s, uint64_t n, const float *rays, int mode, int accelMode, float *t, float *u, float *v, uint32_t *prim) {
    Scene *sc = (Scene *) s;
    Accel &A = sc->accel;
    bool saved = A.useTree;
    if (accelMode >= 0) A.useTree = accelMode != 0;
    for (uint64_t i = 0; i < n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3], r[7]);
        if (mode == 0) {
            Hit h; h.t = kInf; h.u = h.v = 0; h.prim = 0xFFFFFFFFu;
            bool ok = A.rayIntersect(ray, h);
            t[i] = ok ? h.t : kInf; u[i] = ok ? h.u : 0; v[i] = ok ? h.v : 0; prim[i] = ok ? h.prim : 0xFFFFFFFFu;
        } else prim[i] = A.rayOccluded(ray) ? 1u : 0u;
    }
    A.useTree = saved;
}

/* primary rays for pixel-sample positions (n x 2) -> n x 8 */
void orc_camera_rays(void *s, uint64_t n, const float *pos, float *rays) {
    Scene *sc = (Scene *) s;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r = sc->sampleRay(pos[2 * i], pos[2 * i + 1]);
        float *o = rays + 8 * i;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.mint; o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = r.maxt;
    }
}

/* full intersection records for given rays: out 24 floats: p(3) geoN(3) shN(3) s(3) t(3) wi(3) t(1) mesh(1) prim(1) valid(1) pad(2) */
void orc_intersect_full(void *s, uint64_t n, const float *rays, float *out) {
    Scene *sc = (Scene *) s; OrcStats st{};
    for (uint64_t i = 0; i < n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3], r[7]);
        Intersection its; float *o = out + 24 * i; memset(o, 0, 96);
        if (sc->rayIntersect(ray, its, st)) {
            V3 vs[6] = {its.p, its.geoFrame.n, its.shFrame.n, its.shFrame.s, its.shFrame.t, its.wi};
            for (int k = 0; k < 6; ++k) { o[3 * k] = vs[k].x; o[3 * k + 1] = vs[k].y; o[3 * k + 2] = vs[k].z; }
            o[18] = its.t; o[19] = (float) its.mesh; o[20] = (float) its.prim; o[21] = 1.0f;
        }
    }
}

/* ---- bitmap textures (orc_texture.h) ---- */
int orc_add_texture(void *s, const OrcTextureDesc *d, const float *pixels) {
    Scene *sc = (Scene *) s;
    sc->textures.emplace_back();
    sc->textures.back().build(*d, pixels);
    return (int) sc->textures.size() - 1;
}
/* out: levels, then (w, h) per level */
void orc_texture_info(void *s, int tex, int32_t *out, float *maximum, float *bsdfScale) {
    const Texture &t = ((Scene *) s)->textures[tex];
    out[0] = t.levels;
    for (int l = 0; l < t.levels; ++l) { out[1 + 2 * l] = t.lw[l]; out[2 + 2 * l] = t.lh[l]; }
    *maximum = t.maximum; *bsdfScale = t.bsdfScale;
}
/* probe of roundToHalf (orc_texture.h): the storage rounding of the MIP pyramid */
void orc_half_round(uint64_t n, const float *in, float *out) { for (uint64_t i = 0; i < n; ++i) out[i] = roundToHalf(in[i]); }
void orc_texture_level(void *s, int tex, int level, float *out) {
    const Texture &t = ((Scene *) s)->textures[tex];
    memcpy(out, t.pyramid[level].data(), t.pyramid[level].size() * sizeof(float));
}
/* Texture2D::eval for n look-ups: uv (2n), partials (4n: dudx dudy dvdx dvdy) or NULL for the unfiltered path; out rgb (3n) */
void orc_texture_eval(void *s, int tex, uint64_t n, const float *uv, const float *partials, float *out) {
    const Texture &t = ((Scene *) s)->textures[tex];
    for (uint64_t i = 0; i < n; ++i) {
        const V3 r = partials ? t.eval(uv[2 * i], uv[2 * i + 1], true, partials[4 * i], partials[4 * i + 1], partials[4 * i + 2], partials[4 * i + 3])
                              : t.eval(uv[2 * i], uv[2 * i + 1], false, 0, 0, 0, 0);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
/* primary-ray intersections with uv partials (sampleRayDifferential + scaleDifferential + computePartials):
   pos (2n) -> out 8n: valid, u, v, dudx, dudy, dvdx, dvdy, mesh */
void orc_primary_partials(void *s, uint64_t n, const float *pos, int spp, float *out) {
    Scene *sc = (Scene *) s; OrcStats st{};
    const float scale = 1.0f / std::sqrt((float) spp);
    for (uint64_t i = 0; i < n; ++i) {
        RayDiff rd;
        Ray r = sc->sampleRay(pos[2 * i], pos[2 * i + 1], 0.5f, 0.5f, &rd, scale);
        Intersection its; float *o = out + 8 * i; memset(o, 0, 32);
        if (sc->rayIntersect(r, its, st)) {
            Scene::computePartials(its, r, rd);
            o[0] = 1; o[1] = its.tex.u; o[2] = its.tex.v; o[3] = its.tex.dudx; o[4] = its.tex.dudy; o[5] = its.tex.dvdx; o[6] = its.tex.dvdy; o[7] = (float) its.mesh;
        }
    }
}

/* ---- render ---- */
static void render_impl(Scene *sc, const OrcRenderParams *rp, float *film, OrcStats *stats,
                        float *perSample /* optional: W*H*(hi-lo)*4 floats Li.rgb, alpha */) {
    const int W = sc->W, H = sc->H;
    RFilter filter(rp->rfilter, rp->rfilterParam);
    const int bs = rp->blockSize > 0 ? rp->blockSize : 32;
    const int nbx = (W + bs - 1) / bs, nby = (H + bs - 1) / bs, nBlocks = nbx * nby;
    const int lo = rp->sampleLo, hi = rp->sampleHi > 0 ? rp->sampleHi : rp->spp;
    int nThreads = rp->threads > 0 ? rp->threads : (int) std::thread::hardware_concurrency();
    if (nThreads < 1) nThreads = 1;
    std::vector<std::unique_ptr<ImageBlock>> blocks(nBlocks);
    std::atomic<int> next(0);
    std::vector<OrcStats> tstats(nThreads);
    const bool useDiff = !sc->textures.empty() || !sc->envmaps.empty(); /* who reads them: bitmap textures (computePartials) and envmap::evalEnvironment */
    const float diffScaleFactor = 1.0f / std::sqrt((float) rp->spp); /* integrator.cpp:144-145 */
    auto worker = [&](int tid) {
        OrcStats st{};
        std::unique_ptr<Sampler> sampler;
        if (rp->sampler == 0) sampler.reset(new SobolSampler(&sc->sobol, rp->seed, W, H));
        else if (rp->sampler == 1) sampler.reset(new IndependentSampler(rp->seed + (uint64_t) tid));
        else sampler.reset(new CounterSampler(W, (uint32_t) rp->spp, rp->seed));
        for (;;) {
            int b = next.fetch_add(1);
            if (b >= nBlocks) break;
            int bx = b % nbx, by = b / nbx;
            int ox = bx * bs, oy = by * bs, sx = std::min(bs, W - ox), sy = std::min(bs, H - oy);
            std::unique_ptr<ImageBlock> blk(new ImageBlock(ox, oy, sx, sy, &filter));
            /* integrator.cpp:162-187; pixel order inside the block is scanline here (the reference
               uses a Hilbert curve, renderproc.cpp:79-81 -- affects float summation order only) */
            for (int y = oy; y < oy + sy; ++y)
                for (int x = ox; x < ox + sx; ++x) {
                    sampler->generate(x, y);
                    for (int j = 0; j < lo; ++j) sampler->advance(); /* skip to the shard start (sobol/counter: exact) */
                    for (int j = lo; j < hi; ++j) {
                        float ax, ay; sampler->next2D(ax, ay);
                        float spx = (float) x + ax, spy = (float) y + ay;
                        float apx = 0.5f, apy = 0.5f;
                        if (sc->apertureRadius > 0) sampler->next2D(apx, apy); /* needsApertureSample, integrator.cpp:173-174 */
                        RayDiff rd;
                        Ray ray = sc->sampleRay(spx, spy, apx, apy, useDiff ? &rd : nullptr, diffScaleFactor);
                        float alpha;
                        Spectrum spec = rp->integrator == 1 ? sc->LiVol(ray, sampler.get(), *rp, alpha, st, useDiff ? &rd : nullptr) : sc->Li(ray, sampler.get(), *rp, alpha, st, useDiff ? &rd : nullptr); /* sensor weight = 1 */
                        if (!blk->put(spx, spy, spec, alpha)) ++st.badSamples;
                        if (perSample) {
                            float *o = perSample + (((size_t) y * W + x) * (size_t) (hi - lo) + (size_t) (j - lo)) * 4;
                            o[0] = spec.x; o[1] = spec.y; o[2] = spec.z; o[3] = alpha;
                        }
                        ++st.samples;
                        sampler->advance();
                    }
                }
            if (rp->sampler == 0 && ((SobolSampler *) sampler.get())->dimOverflow) st.dimOverflow = 1;
            blocks[b] = std::move(blk);
        }
        tstats[tid] = st;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nThreads; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto &t : th) t.join();
    /* film->put(block): imageblock.h:103-107 -> Bitmap::accumulate, clipped to the film; merged in
       block order so the oracle is deterministic (the reference merges in completion order) */
    memset(film, 0, (size_t) W * H * 5 * sizeof(float));
    for (int b = 0; b < nBlocks; ++b) {
        ImageBlock &blk = *blocks[b];
        for (int y = 0; y < blk.bh(); ++y) {
            int fy = blk.oy - blk.border + y;
            if (fy < 0 || fy >= H) continue;
            for (int x = 0; x < blk.bw(); ++x) {
                int fx = blk.ox - blk.border + x;
                if (fx < 0 || fx >= W) continue;
                const float *src = blk.data.data() + ((size_t) y * blk.bw() + x) * 5;
                float *dst = film + ((size_t) fy * W + fx) * 5;
                for (int k = 0; k < 5; ++k) dst[k] += src[k];
            }
        }
    }
    if (stats) {
        OrcStats tot{};
        for (auto &s : tstats) {
            tot.samples += s.samples; tot.rays += s.rays; tot.shadowRays += s.shadowRays; tot.pathLengthSum += s.pathLengthSum;
            tot.nodeVisits += s.nodeVisits; tot.primTests += s.primTests; tot.badSamples += s.badSamples; tot.dimOverflow |= s.dimOverflow;
        }
        *stats = tot;
    }
}
void orc_render(void *s, const OrcRenderParams *rp, float *film, OrcStats *stats) { render_impl((Scene *) s, rp, film, stats, nullptr); }
void orc_render_samples(void *s, const OrcRenderParams *rp, float *film, OrcStats *stats, float *perSample) {
    render_impl((Scene *) s, rp, film, stats, perSample);
}
/* fmtconv.cpp:979-990: rgb = spec * (weight != 0 ? 1/weight : weight) */
void orc_develop(const float *film, int W, int H, float *rgb) {
    for (size_t i = 0; i < (size_t) W * H; ++i) {
        float weight = film[5 * i + 4], invWeight = (weight != 0) ? 1 / weight : weight;
        for (int k = 0; k < 3; ++k) rgb[3 * i + k] = film[5 * i + k] * invWeight;
    }
}
void orc_filter_table(int kind, float param, float *values32, float *radius, int *border) {
    RFilter f(kind, param); memcpy(values32, f.values, 128); *radius = f.radius; *border = f.borderSize;
}
/* splat n samples (pos 2, value 4: rgb+alpha) into a W x H film through 32x32 ImageBlocks */
/* one ImageBlock (layout of oracle/render_ref_shim.cpp::renderref_block_put): data (w + 2 border) x (h + 2 border) x 5 */
void orc_block_put(int ox, int oy, int w, int h, int kind, float param, int n, const float *pos, const float *val, float *data, int *ok) {
    RFilter filter(kind, param);
    ImageBlock blk(ox, oy, w, h, &filter);
    for (int i = 0; i < n; ++i) ok[i] = blk.put(pos[2 * i], pos[2 * i + 1], V3(val[4 * i], val[4 * i + 1], val[4 * i + 2]), val[4 * i + 3]) ? 1 : 0;
    memcpy(data, blk.data.data(), blk.data.size() * sizeof(float));
}
void orc_splat(int W, int H, int kind, float param, uint64_t n, const float *pos, const float *val, float *film) {
    RFilter filter(kind, param);
    const int bs = 32, nbx = (W + bs - 1) / bs, nby = (H + bs - 1) / bs;
    std::vector<std::unique_ptr<ImageBlock>> blocks((size_t) nbx * nby);
    for (int b = 0; b < nbx * nby; ++b) {
        int ox = (b % nbx) * bs, oy = (b / nbx) * bs;
        blocks[b].reset(new ImageBlock(ox, oy, std::min(bs, W - ox), std::min(bs, H - oy), &filter));
    }
    for (uint64_t i = 0; i < n; ++i) {
        int px = (int) std::floor(pos[2 * i]), py = (int) std::floor(pos[2 * i + 1]);
        if (px < 0 || py < 0 || px >= W || py >= H) continue;
        blocks[(size_t) (py / bs) * nbx + px / bs]->put(pos[2 * i], pos[2 * i + 1], V3(val[4 * i], val[4 * i + 1], val[4 * i + 2]), val[4 * i + 3]);
    }
    memset(film, 0, (size_t) W * H * 5 * sizeof(float));
    for (auto &bp : blocks) {
        ImageBlock &blk = *bp;
        for (int y = 0; y < blk.bh(); ++y) {
            int fy = blk.oy - blk.border + y; if (fy < 0 || fy >= H) continue;
            for (int x = 0; x < blk.bw(); ++x) {
                int fx = blk.ox - blk.border + x; if (fx < 0 || fx >= W) continue;
                for (int k = 0; k < 5; ++k) film[((size_t) fy * W + fx) * 5 + k] += blk.data[((size_t) y * blk.bw() + x) * 5 + k];
            }
        }
    }
}

/* ---- component entry points for fixtures / parity tests ---- */
void orc_sfmt_words(uint64_t seed, uint64_t n, uint64_t *out) { SFMT r; r.seed(seed); for (uint64_t i = 0; i < n; ++i) out[i] = r.nextULong(); }
void orc_sfmt_floats(uint64_t seed, uint64_t n, float *out) { SFMT r; r.seed(seed); for (uint64_t i = 0; i < n; ++i) out[i] = r.nextFloat(); }
uint64_t orc_tea(uint32_t v0, uint32_t v1, int rounds) { return sampleTEA(v0, v1, rounds); }
void orc_sobol_sample(const uint32_t *m32, uint64_t n, const uint64_t *index, const uint32_t *dim, uint32_t scramble, float *out) {
    SobolTables T; T.m32 = m32;
    for (uint64_t i = 0; i < n; ++i) out[i] = sobolSample(T, index[i], dim[i], scramble);
}
void orc_sobol_lookup(const uint64_t *vdc, const uint64_t *inv, uint32_t m, uint64_t n, const uint32_t *frame, const uint32_t *px,
                      const uint32_t *py, uint64_t scramble, uint64_t *out) {
    SobolTables T; T.vdc = vdc; T.inv = inv;
    for (uint64_t i = 0; i < n; ++i) out[i] = sobolLookUp(T, m, frame[i], px[i], py[i], scramble);
}
/* the first `ndim` sampler outputs of (pixel, sample) exactly as renderBlock + Li would draw them
 * (dims 0,1 through next2D's pixel rescale); sampler kinds as in OrcRenderParams */
void orc_sampler_stream(void *s, int kind, uint64_t seed, int W, int H, int spp, int px, int py, int sampleIdx, int ndim, float *out) {
    Scene *sc = (Scene *) s;
    std::unique_ptr<Sampler> sm;
    if (kind == 0) sm.reset(new SobolSampler(&sc->sobol, seed, W, H));
    else if (kind == 1) sm.reset(new IndependentSampler(seed));
    else sm.reset(new CounterSampler(W, (uint32_t) spp, seed));
    sm->generate(px, py);
    for (int j = 0; j < sampleIdx; ++j) sm->advance();
    int i = 0;
    if (ndim >= 2) { sm->next2D(out[0], out[1]); i = 2; }
    for (; i < ndim; ++i) out[i] = sm->next1D();
}

struct ReplaySampler : Sampler { /* test_chisquare.cpp:58-92 FakeSampler analogue */
    float v; explicit ReplaySampler(float x) : v(x) {}
    void generate(int, int) override {} void advance() override {}
    float next1D() override { return v; }
    void next2D(float &a, float &b) override { a = b = v; }
};
/* BSDF component calls on local-frame directions. bsdfs: array of nB descs, id = which to query.
 * wi/wo: n x 3; samples: n x 3 (2D sample + the extra 1D a rough dielectric draws).
 * out_sample: n x 10: wo(3) weight(3) pdf(1) sampledType(1) eta(1) pad */
void orc_bsdf_eval(const OrcBsdf *b, int nB, int id, uint64_t n, const float *wi, const float *wo, float *outRgb, float *outPdf) {
    BsdfSet bs{b, nB};
    for (uint64_t i = 0; i < n; ++i) {
        BRec r; r.wi = V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]); r.wo = V3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        Spectrum f = bs.eval(id, r); outRgb[3 * i] = f.x; outRgb[3 * i + 1] = f.y; outRgb[3 * i + 2] = f.z;
        outPdf[i] = bs.pdf(id, r);
    }
}
/* measure == EDiscrete: the delta components (test_chisquare.cpp:131-160 checks them the same way) */
void orc_bsdf_eval_discrete(const OrcBsdf *b, int nB, int id, uint64_t n, const float *wi, const float *wo, float *outRgb, float *outPdf) {
    BsdfSet bs{b, nB};
    for (uint64_t i = 0; i < n; ++i) {
        BRec r; r.wi = V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]); r.wo = V3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        Spectrum f = bs.eval(id, r, true); outRgb[3 * i] = f.x; outRgb[3 * i + 1] = f.y; outRgb[3 * i + 2] = f.z;
        outPdf[i] = bs.pdf(id, r, true);
    }
}
void orc_bsdf_sample(const OrcBsdf *b, int nB, int id, uint64_t n, const float *wi, const float *samples, float *out) {
    BsdfSet bs{b, nB};
    for (uint64_t i = 0; i < n; ++i) {
        ReplaySampler rs(samples[3 * i + 2]);
        BRec r; r.wi = V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]); r.sampler = &rs;
        float pdf = 0;
        Spectrum w = bs.sample(id, r, pdf, samples[3 * i], samples[3 * i + 1]);
        float *o = out + 10 * i;
        o[0] = r.wo.x; o[1] = r.wo.y; o[2] = r.wo.z; o[3] = w.x; o[4] = w.y; o[5] = w.z;
        o[6] = w.isZero() ? 0.0f : pdf; o[7] = (float) r.sampledType; o[8] = r.eta; o[9] = 0;
    }
}
uint32_t orc_bsdf_type(const OrcBsdf *b, int nB, int id) { BsdfSet bs{b, nB}; return bs.type(id); }
/* microfacet protocol of src/tests/test_microfacet.cpp:50-131.  out: n x 6: m(3), pdf from sample(), pdf(wi,m), eval(m) */
void orc_microfacet_sample(int type, float alphaU, float alphaV, int sampleVisible, uint64_t n, const float *wi, const float *samples, float *out) {
    Microfacet d(type, alphaU, alphaV, sampleVisible != 0);
    for (uint64_t i = 0; i < n; ++i) {
        V3 w(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        float pdf; V3 m = d.sample(w, samples[2 * i], samples[2 * i + 1], pdf);
        float *o = out + 6 * i; o[0] = m.x; o[1] = m.y; o[2] = m.z; o[3] = pdf; o[4] = d.pdf(w, m); o[5] = d.eval(m);
    }
}
void orc_microfacet_eval(int type, float alphaU, float alphaV, int sampleVisible, uint64_t n, const float *wi, const float *m, float *out /* n x 3: D, pdf, G1(wi,m) */) {
    Microfacet d(type, alphaU, alphaV, sampleVisible != 0);
    for (uint64_t i = 0; i < n; ++i) {
        V3 w(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), mm(m[3 * i], m[3 * i + 1], m[3 * i + 2]);
        out[3 * i] = d.eval(mm); out[3 * i + 1] = d.pdf(w, mm); out[3 * i + 2] = d.smithG1(w, mm);
    }
}
/* ---- component hooks with the argument layout of oracle/core_ref_shim.cpp (the same functions of the reference, compiled from
 * /root/reference): tests/test_oracle_reference_pins.py compares the two one to one ---- */
void orc_triaccel_load(int n, const float *tris, uint32_t *records, int *status) {
    for (int i = 0; i < n; ++i) {
        const float *t = tris + 9 * i;
        TriAccel a; memset(&a, 0, sizeof(a));
        status[i] = a.load(V3(t[0], t[1], t[2]), V3(t[3], t[4], t[5]), V3(t[6], t[7], t[8]));
        a.shapeIndex = 0; a.primIndex = 0;
        memcpy(records + 12 * i, &a, 48);
    }
}
void orc_triaccel_intersect(int n, const float *tris, const float *rays, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *t = tris + 9 * i, *r = rays + 8 * i;
        TriAccel a; float *o = out + 4 * i; o[0] = o[1] = o[2] = o[3] = 0;
        if (a.load(V3(t[0], t[1], t[2]), V3(t[3], t[4], t[5]), V3(t[6], t[7], t[8])) != 0) continue;
        Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3], r[7]);
        float u, v, tt;
        if (a.rayIntersect(ray, r[3], r[7], u, v, tt)) { o[0] = 1; o[1] = tt; o[2] = u; o[3] = v; }
    }
}
void orc_aabb_intersect(int n, const float *boxes, const float *rays, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *b = boxes + 6 * i, *r = rays + 8 * i;
        AABB box; box.min = V3(b[0], b[1], b[2]); box.max = V3(b[3], b[4], b[5]);
        Ray ray(V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3], r[7]);
        float nearT = 0, farT = 0;
        const bool hit = box.rayIntersect(ray, nearT, farT);
        out[3 * i] = hit ? 1.0f : 0.0f; out[3 * i + 1] = nearT; out[3 * i + 2] = farT;
    }
}
void orc_warp(int what, int n, const float *samples, float *out) {
    for (int i = 0; i < n; ++i) {
        const float sx = samples[2 * i], sy = samples[2 * i + 1]; float *o = out + 3 * i;
        if (what == 0) { const V3 v = squareToCosineHemisphere(sx, sy); o[0] = v.x; o[1] = v.y; o[2] = v.z; }
        else if (what == 1) { const V3 v = squareToUniformSphere(sx, sy); o[0] = v.x; o[1] = v.y; o[2] = v.z; }
        else if (what == 2) { squareToUniformDiskConcentric(sx, sy, o[0], o[1]); o[2] = 0; }
        else { squareToUniformTriangle(sx, sy, o[0], o[1]); o[2] = 0; }
    }
}
void orc_fresnel_dielectric_ext(int n, const float *cosThetaI, float eta, float *out) {
    for (int i = 0; i < n; ++i) { float ct; out[2 * i] = fresnelDielectricExt(cosThetaI[i], ct, eta); out[2 * i + 1] = ct; }
}
void orc_fresnel_conductor_exact_rgb(int n, const float *cosThetaI, const float *eta, const float *k, float *out) {
    const Spectrum e(eta[0], eta[1], eta[2]), kk(k[0], k[1], k[2]);
    for (int i = 0; i < n; ++i) { const Spectrum r = fresnelConductorExact(cosThetaI[i], e, kk); out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z; }
}
void orc_reflect(int n, const float *wi, const float *nrm, float *out) {
    for (int i = 0; i < n; ++i) { const V3 r = reflect(V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), V3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2])); out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z; }
}
void orc_refract(int n, const float *wi, const float *nrm, float eta, const float *cosThetaT, float *out) {
    for (int i = 0; i < n; ++i) { const V3 r = refract(V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), V3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), eta, cosThetaT[i]); out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z; }
}
void orc_coordinate_system(int n, const float *a, float *out) {
    for (int i = 0; i < n; ++i) { V3 b, c; coordinateSystem(V3(a[3 * i], a[3 * i + 1], a[3 * i + 2]), b, c); float *o = out + 6 * i; o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = c.x; o[4] = c.y; o[5] = c.z; }
}
void orc_shading_frame(int n, const float *nrm, const float *dpdu, float *out) {
    for (int i = 0; i < n; ++i) {
        Frame f;
        computeShadingFrame(V3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), V3(dpdu[3 * i], dpdu[3 * i + 1], dpdu[3 * i + 2]), f);
        float *o = out + 9 * i;
        o[0] = f.s.x; o[1] = f.s.y; o[2] = f.s.z; o[3] = f.t.x; o[4] = f.t.y; o[5] = f.t.z; o[6] = f.n.x; o[7] = f.n.y; o[8] = f.n.z;
    }
}
float orc_pmf(int m, const float *weights, int n, const float *samples, uint32_t *index, uint32_t *indexReuse, float *reused, float *cdf) {
    Discrete d;
    for (int i = 0; i < m; ++i) d.append(weights[i]);
    const float total = d.normalize();
    for (int i = 0; i < n; ++i) {
        index[i] = (uint32_t) d.sample(samples[i]);
        float s = samples[i];
        indexReuse[i] = (uint32_t) d.sampleReuse(s);
        reused[i] = s;
    }
    for (int i = 0; i < m; ++i) cdf[i] = d[i];
    return total;
}
/* rec 27n: p, geoFrame.n, dpdu, dpdv, ray.o, rxOrigin, ryOrigin, rxDirection, ryDirection -> out 4n: dudx dudy dvdx dvdy
   (layout of oracle/render_ref_shim.cpp::renderref_compute_partials) */
void orc_compute_partials(int n, const float *rec, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *r = rec + 27 * i;
        Intersection its;
        its.p = V3(r[0], r[1], r[2]); its.geoFrame.n = V3(r[3], r[4], r[5]);
        its.dpdu = V3(r[6], r[7], r[8]); its.dpdv = V3(r[9], r[10], r[11]);
        RayDiff rd; rd.has = true;
        rd.rxO = V3(r[15], r[16], r[17]); rd.ryO = V3(r[18], r[19], r[20]); rd.rxD = V3(r[21], r[22], r[23]); rd.ryD = V3(r[24], r[25], r[26]);
        Ray ray(V3(r[12], r[13], r[14]), V3(0, 0, 1));
        Scene::computePartials(its, ray, rd);
        out[4 * i] = its.tex.dudx; out[4 * i + 1] = its.tex.dudy; out[4 * i + 2] = its.tex.dvdx; out[4 * i + 3] = its.tex.dvdy;
    }
}
/* Triangle::sample (triangle.cpp:24-62) as TriMesh::samplePosition uses it: barycentric warp of the 2-D sample */
void orc_triangle_sample(int n, const float *tris, const float *samples, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *t = tris + 9 * i;
        const V3 p0(t[0], t[1], t[2]), p1(t[3], t[4], t[5]), p2(t[6], t[7], t[8]);
        float bx, by; squareToUniformTriangle(samples[2 * i], samples[2 * i + 1], bx, by);
        const V3 sideA = p1 - p0, sideB = p2 - p0;
        const V3 p = p0 + (sideA * bx) + (sideB * by);
        out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z;
    }
}

/* emitter direct sampling from reference points (n x 6: ref, refN) with 2D samples -> n x 12:
 * d(3) dist pdf value(3) visible p(3) */
void orc_sample_emitter_direct(void *s, uint64_t n, const float *ref, const float *samples, float *out) {
    Scene *sc = (Scene *) s; OrcStats st{};
    for (uint64_t i = 0; i < n; ++i) {
        DRec d; d.ref = V3(ref[6 * i], ref[6 * i + 1], ref[6 * i + 2]); d.refN = V3(ref[6 * i + 3], ref[6 * i + 4], ref[6 * i + 5]);
        Spectrum v = sc->sampleEmitterDirect(d, samples[2 * i], samples[2 * i + 1], st);
        float *o = out + 12 * i;
        o[0] = d.d.x; o[1] = d.d.y; o[2] = d.d.z; o[3] = d.dist; o[4] = v.isZero() ? 0.0f : d.pdf; o[5] = v.x; o[6] = v.y; o[7] = v.z;
        o[8] = v.isZero() ? 0.0f : 1.0f; o[9] = d.p.x; o[10] = d.p.y; o[11] = d.p.z;
    }
}
int orc_hardware_threads() { return (int) std::thread::hardware_concurrency(); }

} // extern "C"
