/* Thin C entry points over the REFERENCE's own core math, compiled where it lies under /root/reference (never copied) into
 * oracle/_ref/libcoreref.so by oracle/Makefile, behind the stand-in headers of oracle/shim_core/:
 *   src/bsdfs/microfacet.h                      MicrofacetDistribution: eval, sample (all / visible normals), pdf, smithG1, G
 *   include/mitsuba/render/triaccel.h           TriAccel::load, TriAccel::rayIntersect
 *   include/mitsuba/core/aabb.h                 AABB::rayIntersect(ray, nearT, farT)
 *   include/mitsuba/core/triangle.h + triangle.cpp   Triangle::rayIntersect (Moeller-Trumbore), Triangle::sample
 *   include/mitsuba/core/warp.h + warp.cpp      squareTo{CosineHemisphere, UniformSphere, UniformDiskConcentric, UniformTriangle}
 *   include/mitsuba/core/util.h + util.cpp      fresnelDielectricExt, fresnelConductorExact, refract, reflect, coordinateSystem,
 *                                               computeShadingFrame, fresnelDiffuseReflectance (with quad.cpp), solveLinearSystem2x2
 *   include/mitsuba/core/pmf.h                  DiscreteDistribution: normalize, sample, sampleReuse
 *   include/mitsuba/core/qmc.h + qmc.cpp        sampleTEA, radicalInverse2Single, sobol2Single
 * Used only to pin the oracle (tests/gen_golden.py -> tests/golden/core_ref.npz; tests/test_oracle_reference_pins.py). */
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/aabb.h>
#include <mitsuba/core/frame.h>
#include <mitsuba/core/pmf.h>
#include <mitsuba/core/qmc.h>
#include <mitsuba/core/random.h>
#include <mitsuba/core/triangle.h>
#include <mitsuba/core/warp.h>
#include <mitsuba/render/triaccel.h>
#include "microfacet.h" /* -I$(REF)/src/bsdfs */

namespace mitsuba {
/* util.cpp's stratified / latin-hypercube helpers draw from Random (src/libcore/random.cpp needs boost::filesystem streams);
   nothing here calls them */
Float Random::nextFloat() { return 0.5f; }
size_t Random::nextSize(size_t) { return 0; }
}

using namespace mitsuba;

#include <mitsuba/core/half.h>
extern "C" {
/* float -> half -> float through the reference's own half class (include/mitsuba/core/half.h:431-487, src/libcore/half.cpp:78-200): the
   storage format of BitmapTexture's MIP pyramid (bitmap.cpp:177-180, TMIPMap<Color3, Color3h>) */
void coreref_half_round(int n, const float *in, float *out) { for (int i = 0; i < n; ++i) out[i] = (float) half(in[i]); }


/* ---- microfacet.h: type 0 beckmann, 1 ggx, 2 phong; wi (3n), samples (2n) -> out 6n: m(3), pdf from sample(), pdf(wi, m), eval(m) */
void coreref_microfacet_sample(int type, float alphaU, float alphaV, int sampleVisible, int n, const float *wi, const float *samples, float *out) {
    MicrofacetDistribution d((MicrofacetDistribution::EType) type, alphaU, alphaV, sampleVisible != 0);
    for (int i = 0; i < n; ++i) {
        const Vector w(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        Float pdf;
        const Normal m = d.sample(w, Point2(samples[2 * i], samples[2 * i + 1]), pdf);
        float *o = out + 6 * i;
        o[0] = m.x; o[1] = m.y; o[2] = m.z; o[3] = pdf; o[4] = d.pdf(w, Vector(m)); o[5] = d.eval(Vector(m));
    }
}
/* wi (3n), m (3n) -> out 3n: eval(m), pdf(wi, m), smithG1(wi, m) */
void coreref_microfacet_eval(int type, float alphaU, float alphaV, int sampleVisible, int n, const float *wi, const float *m, float *out) {
    MicrofacetDistribution d((MicrofacetDistribution::EType) type, alphaU, alphaV, sampleVisible != 0);
    for (int i = 0; i < n; ++i) {
        const Vector w(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), mm(m[3 * i], m[3 * i + 1], m[3 * i + 2]);
        out[3 * i] = d.eval(mm); out[3 * i + 1] = d.pdf(w, mm); out[3 * i + 2] = d.smithG1(w, mm);
    }
}

/* ---- triaccel.h: tris 9n floats -> records 12n words (k, n_u, n_v, n_d, a_u, a_v, b_nu, b_nv, c_nu, c_nv, shapeIndex, primIndex), status n */
void coreref_triaccel_load(int n, const float *tris, uint32_t *records, int *status) {
    for (int i = 0; i < n; ++i) {
        const float *t = tris + 9 * i;
        TriAccel a;
        memset(&a, 0, sizeof(a));
        status[i] = a.load(Point(t[0], t[1], t[2]), Point(t[3], t[4], t[5]), Point(t[6], t[7], t[8]));
        a.shapeIndex = 0; a.primIndex = 0;
        memcpy(records + 12 * i, &a, 48);
    }
}
/* one triangle per ray: rays 8n (o, mint, d, maxt) -> out 4n: hit, t, u, v */
void coreref_triaccel_intersect(int n, const float *tris, const float *rays, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *t = tris + 9 * i, *r = rays + 8 * i;
        TriAccel a;
        float *o = out + 4 * i;
        o[0] = o[1] = o[2] = o[3] = 0;
        if (a.load(Point(t[0], t[1], t[2]), Point(t[3], t[4], t[5]), Point(t[6], t[7], t[8])) != 0) continue;
        Ray ray(Point(r[0], r[1], r[2]), Vector(r[4], r[5], r[6]), r[3], r[7], 0.0f);
        Float u, v, tt;
        if (a.rayIntersect(ray, r[3], r[7], u, v, tt)) { o[0] = 1; o[1] = tt; o[2] = u; o[3] = v; }
    }
}
/* Triangle::rayIntersect(p0, p1, p2, ray, u, v, t): out 4n: hit, t, u, v (no interval test inside) */
void coreref_triangle_intersect(int n, const float *tris, const float *rays, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *t = tris + 9 * i, *r = rays + 8 * i;
        Ray ray(Point(r[0], r[1], r[2]), Vector(r[4], r[5], r[6]), r[3], r[7], 0.0f);
        Float u, v, tt;
        float *o = out + 4 * i;
        o[0] = o[1] = o[2] = o[3] = 0;
        if (Triangle::rayIntersect(Point(t[0], t[1], t[2]), Point(t[3], t[4], t[5]), Point(t[6], t[7], t[8]), ray, u, v, tt)) { o[0] = 1; o[1] = tt; o[2] = u; o[3] = v; }
    }
}
/* Triangle::sample (triangle.cpp:24-62) without normals / texcoords: samples 2n -> out 3n positions */
void coreref_triangle_sample(int n, const float *tris, const float *samples, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *t = tris + 9 * i;
        Point P[3] = {Point(t[0], t[1], t[2]), Point(t[3], t[4], t[5]), Point(t[6], t[7], t[8])};
        Triangle tri; tri.idx[0] = 0; tri.idx[1] = 1; tri.idx[2] = 2;
        Normal nn; Point2 uv;
        const Point p = tri.sample(P, NULL, NULL, nn, uv, Point2(samples[2 * i], samples[2 * i + 1]));
        out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z;
    }
}
/* ---- aabb.h: boxes 6n (min, max), rays 8n -> out 3n: hit, nearT, farT */
void coreref_aabb_intersect(int n, const float *boxes, const float *rays, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *b = boxes + 6 * i, *r = rays + 8 * i;
        AABB box(Point(b[0], b[1], b[2]), Point(b[3], b[4], b[5]));
        Ray ray(Point(r[0], r[1], r[2]), Vector(r[4], r[5], r[6]), r[3], r[7], 0.0f);
        Float nearT = 0, farT = 0;
        const bool hit = box.rayIntersect(ray, nearT, farT);
        out[3 * i] = hit ? 1.0f : 0.0f; out[3 * i + 1] = nearT; out[3 * i + 2] = farT;
    }
}
/* ---- warp.cpp: what 0 cosine hemisphere, 1 uniform sphere, 2 concentric disk, 3 uniform triangle; samples 2n -> out 3n (z = 0 for 2-D results) */
void coreref_warp(int what, int n, const float *samples, float *out) {
    for (int i = 0; i < n; ++i) {
        const Point2 s(samples[2 * i], samples[2 * i + 1]);
        float *o = out + 3 * i;
        if (what == 0) { const Vector v = warp::squareToCosineHemisphere(s); o[0] = v.x; o[1] = v.y; o[2] = v.z; }
        else if (what == 1) { const Vector v = warp::squareToUniformSphere(s); o[0] = v.x; o[1] = v.y; o[2] = v.z; }
        else if (what == 2) { const Point2 p = warp::squareToUniformDiskConcentric(s); o[0] = p.x; o[1] = p.y; o[2] = 0; }
        else { const Point2 p = warp::squareToUniformTriangle(s); o[0] = p.x; o[1] = p.y; o[2] = 0; }
    }
}
/* ---- util.cpp */
void coreref_fresnel_dielectric_ext(int n, const float *cosThetaI, float eta, float *out /* 2n: F, cosThetaT */) {
    for (int i = 0; i < n; ++i) { Float ct; out[2 * i] = fresnelDielectricExt(cosThetaI[i], ct, eta); out[2 * i + 1] = ct; }
}
void coreref_fresnel_conductor_exact(int n, const float *cosThetaI, float eta, float k, float *out) {
    for (int i = 0; i < n; ++i) out[i] = fresnelConductorExact(cosThetaI[i], eta, k);
}
/* the Spectrum overload (util.cpp:739-761) that roughconductor / conductor call: eta, k RGB -> out 3n */
void coreref_fresnel_conductor_exact_rgb(int n, const float *cosThetaI, const float *eta, const float *k, float *out) {
    Spectrum e, kk;
    for (int c = 0; c < 3; ++c) { e[c] = eta[c]; kk[c] = k[c]; }
    for (int i = 0; i < n; ++i) { const Spectrum r = fresnelConductorExact(cosThetaI[i], e, kk); out[3 * i] = r[0]; out[3 * i + 1] = r[1]; out[3 * i + 2] = r[2]; }
}
/* wi (3n), normal (3n) -> out 3n */
void coreref_reflect(int n, const float *wi, const float *nrm, float *out) {
    for (int i = 0; i < n; ++i) { const Vector r = reflect(Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), Normal(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2])); out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z; }
}
/* refract(wi, n, eta, cosThetaT) (util.cpp: the variant the rough dielectric uses) */
void coreref_refract(int n, const float *wi, const float *nrm, float eta, const float *cosThetaT, float *out) {
    for (int i = 0; i < n; ++i) { const Vector r = refract(Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), Normal(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), eta, cosThetaT[i]); out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z; }
}
/* coordinateSystem(a, b, c): a (3n) -> out 6n (b, c) */
void coreref_coordinate_system(int n, const float *a, float *out) {
    for (int i = 0; i < n; ++i) { Vector b, c; coordinateSystem(Vector(a[3 * i], a[3 * i + 1], a[3 * i + 2]), b, c); float *o = out + 6 * i; o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = c.x; o[4] = c.y; o[5] = c.z; }
}
/* computeShadingFrame(n, dpdu, frame): -> out 9n (s, t, n) */
void coreref_shading_frame(int n, const float *nrm, const float *dpdu, float *out) {
    for (int i = 0; i < n; ++i) {
        Frame f;
        computeShadingFrame(Vector(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), Vector(dpdu[3 * i], dpdu[3 * i + 1], dpdu[3 * i + 2]), f);
        float *o = out + 9 * i;
        o[0] = f.s.x; o[1] = f.s.y; o[2] = f.s.z; o[3] = f.t.x; o[4] = f.t.y; o[5] = f.t.z; o[6] = f.n.x; o[7] = f.n.y; o[8] = f.n.z;
    }
}
float coreref_fresnel_diffuse_reflectance(float eta, int fast) { return fresnelDiffuseReflectance(eta, fast != 0); }
/* ---- pmf.h: weights (m) -> normalization; samples (n) -> index n, and sampleReuse: index + rescaled sample */
float coreref_pmf(int m, const float *weights, int n, const float *samples, uint32_t *index, uint32_t *indexReuse, float *reused, float *cdf) {
    DiscreteDistribution d;
    for (int i = 0; i < m; ++i) d.append(weights[i]);
    const Float total = d.normalize();
    for (int i = 0; i < n; ++i) {
        index[i] = (uint32_t) d.sample(samples[i]);
        Float s = samples[i];
        indexReuse[i] = (uint32_t) d.sampleReuse(s);
        reused[i] = s;
    }
    for (int i = 0; i < m; ++i) cdf[i] = d[i];
    return total;
}
/* ---- qmc.h */
uint64_t coreref_tea(uint32_t v0, uint32_t v1, int rounds) { return sampleTEA(v0, v1, rounds); }
float coreref_radical_inverse2(uint32_t n, uint32_t scramble) { return radicalInverse2Single(n, scramble); }
float coreref_sobol2(uint32_t n, uint32_t scramble) { return sobol2Single(n, scramble); }
}
