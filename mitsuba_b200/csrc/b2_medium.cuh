// Participating media on the `volpath` path:
//   HeterogeneousMedium, Woodcock tracking   src/medium/heterogeneous.cpp:546-660
//   HomogeneousMedium (balance/single/manual) src/medium/homogeneous.cpp:266-362
//   GridDataSource::lookupFloat              src/volume/gridvolume.cpp:336-371
//   isotropic / Henyey-Greenstein phase      src/phase/isotropic.cpp:62-79, src/phase/hg.cpp:76-115
#pragma once
#include "b2_math.cuh"
#include "b2_sampler.cuh"
#include "b2_trace.cuh"
#include "b2_types.h"

namespace b2 {

#define B2_INV_FOURPI 0.07957747154594766788f
// The reference's Woodcock loops are unbounded; a stuck random stream (Sobol' dimension overflow, a sample that is
// exactly 0 in empty space) would spin forever, so both this code and the oracle stop after this many steps.
#define B2_MAX_WOODCOCK_STEPS (1u << 20)

struct MediumRec { // include/mitsuba/render/medium.h:34-82
    float t;
    V3 p;
    Spectrum sigmaS, transmittance;
    float pdfFailure, pdfSuccess;
};

// warp.cpp:25-31
B2_DEV V3 squareToUniformSphere(float sx, float sy) {
    const float z = 1.0f - 2.0f * sy;
    const float r = safe_sqrt(1.0f - z * z);
    float sinPhi, cosPhi;
    sincosf(2.0f * B2_PI * sx, &sinPhi, &cosPhi);
    return V3(r * cosPhi, r * sinPhi, z);
}

B2_DEV float phaseEval(const DMedium &m, const V3 &wi, const V3 &wo) {
    if (m.phase == 0) return B2_INV_FOURPI;                       // isotropic.cpp:75-77
    const float temp = 1.0f + m.g * m.g + 2.0f * m.g * dot(wi, wo); // hg.cpp:105-108
    return B2_INV_FOURPI * (1 - m.g * m.g) / (temp * sqrtf(temp));
}
// sample(pRec, pdf, sampler): weight 1, draws one 2D sample
B2_DEV float phaseSample(const DMedium &m, const V3 &wi, V3 &wo, float &pdf, PathSampler &smp) {
    float sx, sy;
    smp.next2D(sx, sy);
    if (m.phase == 0) { // isotropic.cpp:69-73
        wo = squareToUniformSphere(sx, sy);
        pdf = B2_INV_FOURPI;
        return 1.0f;
    }
    float cosT; // hg.cpp:76-100
    if (fabsf(m.g) < B2_EPSILON) cosT = 1 - 2 * sx;
    else {
        const float sqrTerm = (1 - m.g * m.g) / (1 - m.g + 2 * m.g * sx);
        cosT = (1 + m.g * m.g - sqrTerm * sqrTerm) / (2 * m.g);
    }
    const float sinT = safe_sqrt(1.0f - cosT * cosT);
    float sinPhi, cosPhi;
    sincosf(2 * B2_PI * sy, &sinPhi, &cosPhi);
    Frame f;
    f.n = -wi;
    coordinateSystem(f.n, f.s, f.t);
    wo = f.toWorld(V3(sinT * cosPhi, sinT * sinPhi, cosT));
    pdf = phaseEval(m, wi, wo);
    return 1.0f;
}

// gridvolume.cpp:336-371
B2_DEV float lookupDensity(const DMedium &m, const V3 &wp) {
    const float *M = m.worldToGrid;
    const V3 p(M[0] * wp.x + M[1] * wp.y + M[2] * wp.z + M[3], M[4] * wp.x + M[5] * wp.y + M[6] * wp.z + M[7],
               M[8] * wp.x + M[9] * wp.y + M[10] * wp.z + M[11]);
    const int x1 = (int) floorf(p.x), y1 = (int) floorf(p.y), z1 = (int) floorf(p.z), x2 = x1 + 1, y2 = y1 + 1, z2 = z1 + 1;
    if (x1 < 0 || y1 < 0 || z1 < 0 || x2 >= m.res[0] || y2 >= m.res[1] || z2 >= m.res[2]) return 0.0f;
    const float fx = p.x - x1, fy = p.y - y1, fz = p.z - z1, _fx = 1.0f - fx, _fy = 1.0f - fy, _fz = 1.0f - fz;
    const float *D = m.density;
    const int rx = m.res[0], ry = m.res[1];
    const float d000 = __ldg(D + (z1 * ry + y1) * rx + x1), d001 = __ldg(D + (z1 * ry + y1) * rx + x2), d010 = __ldg(D + (z1 * ry + y2) * rx + x1),
                d011 = __ldg(D + (z1 * ry + y2) * rx + x2), d100 = __ldg(D + (z2 * ry + y1) * rx + x1), d101 = __ldg(D + (z2 * ry + y1) * rx + x2),
                d110 = __ldg(D + (z2 * ry + y2) * rx + x1), d111 = __ldg(D + (z2 * ry + y2) * rx + x2);
    return ((d000 * _fx + d001 * fx) * _fy + (d010 * _fx + d011 * fx) * fy) * _fz + ((d100 * _fx + d101 * fx) * _fy + (d110 * _fx + d111 * fx) * fy) * fz;
}

B2_DEV bool densityBox(const DMedium &m, const V3 &o, const V3 &d, float rayMint, float rayMaxt, float &mint, float &maxt) {
    const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    if (!aabbRayIntersect(m.aabbMin, m.aabbMax, o, d, dRcp, mint, maxt)) return false;
    mint = fmaxf(mint, rayMint);
    maxt = fminf(maxt, rayMaxt);
    return true;
}

// One shared copy of the Woodcock loop (heterogeneous.cpp:560-580 and :627-655).  Not inlined on purpose: lanes of a warp that walk
// through the medium for different reasons (distance sampling, shadow connection, emitter look-up) all execute this one loop body, so
// the hardware issues them together (measured +50 % on the smoke scene against per-call-site copies).
//   nWalks = 1: Medium::sampleDistance  -- returns 1 if the walk left [mint, maxt), else 0 with the collision in tHit / densityHit
//   nWalks = 2: Medium::evalTransmittance -- returns how many of the two walks got through
static __device__ __noinline__ int woodcockWalk(const DMedium &m, const V3 &o, const V3 &d, float mint, float maxt, int nWalks, PathSampler &smp,
                                               float &tHit, float &densityHit) {
    const float invMaxDensity = m.invMaxDensity;
    int escaped = 0;
    uint32_t steps = 0;
    for (int i = 0; i < nWalks; ++i) {
        float t = mint;
        while (true) {
            t -= logOneMinus(smp.next1D()) * invMaxDensity;
            if (t >= maxt) { ++escaped; break; }
            const float density = lookupDensity(m, o + d * t) * m.scale;
            if (density * invMaxDensity > smp.next1D()) { tHit = t; densityHit = density; break; }
            if (++steps > B2_MAX_WOODCOCK_STEPS) { tHit = t; densityHit = 0.0f; break; }
        }
    }
    return escaped;
}

// Medium::evalTransmittance(Ray(o, d, mint, maxt), sampler)
B2_DEV Spectrum mediumTransmittance(const DMedium &m, const V3 &o, const V3 &d, float rayMint, float rayMaxt, PathSampler &smp) {
    if (m.type == 0) { // homogeneous.cpp:266-273
        const float negLength = rayMint - rayMaxt;
        Spectrum tr;
        const float sx = m.sigmaA[0] + m.sigmaS[0], sy = m.sigmaA[1] + m.sigmaS[1], sz = m.sigmaA[2] + m.sigmaS[2];
        tr.x = sx != 0 ? fastexp(sx * negLength) : 1.0f;
        tr.y = sy != 0 ? fastexp(sy * negLength) : 1.0f;
        tr.z = sz != 0 ? fastexp(sz * negLength) : 1.0f;
        return tr;
    }
    // heterogeneous.cpp:546-585: two Woodcock walks, result = fraction that got through
    float mint, maxt;
    if (!densityBox(m, o, d, rayMint, rayMaxt, mint, maxt)) return Spectrum(1.0f);
    float tHit, densityHit;
    const int escaped = woodcockWalk(m, o, d, mint, maxt, 2, smp, tHit, densityHit);
    return Spectrum((float) escaped / 2);
}

// Medium::sampleDistance(Ray(o, d, mint, maxt), mRec, sampler)
B2_DEV bool mediumSampleDistance(const DMedium &m, const V3 &o, const V3 &d, float rayMint, float rayMaxt, MediumRec &mRec, PathSampler &smp) {
    if (m.type == 0) { // homogeneous.cpp:275-362
        float rnd = smp.next1D(), sampledDistance;
        float samplingDensity = m.samplingDensity;
        const float sigmaT[3] = {m.sigmaA[0] + m.sigmaS[0], m.sigmaA[1] + m.sigmaS[1], m.sigmaA[2] + m.sigmaS[2]};
        if (rnd < m.mediumSamplingWeight) {
            rnd /= m.mediumSamplingWeight;
            if (m.strategy == 0) {
                const int channel = min((int) (smp.next1D() * 3), 2);
                samplingDensity = sigmaT[channel];
            }
            sampledDistance = -logOneMinus(rnd) / samplingDensity;
        } else sampledDistance = B2_INF;
        const float distSurf = rayMaxt - rayMint;
        bool success = true;
        if (sampledDistance < distSurf) {
            mRec.t = sampledDistance + rayMint;
            mRec.p = o + d * mRec.t;
            mRec.sigmaS = V3(m.sigmaS[0], m.sigmaS[1], m.sigmaS[2]);
            if (mRec.p.x == o.x && mRec.p.y == o.y && mRec.p.z == o.z) success = false;
        } else { sampledDistance = distSurf; success = false; }
        if (m.strategy == 0) {
            mRec.pdfFailure = 0; mRec.pdfSuccess = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float tmp = fastexp(-sigmaT[i] * sampledDistance);
                mRec.pdfFailure += tmp;
                mRec.pdfSuccess += sigmaT[i] * tmp;
            }
            mRec.pdfFailure /= 3; mRec.pdfSuccess /= 3;
        } else {
            mRec.pdfFailure = fastexp(-samplingDensity * sampledDistance);
            mRec.pdfSuccess = samplingDensity * mRec.pdfFailure;
        }
        mRec.transmittance = V3(fastexp(sigmaT[0] * (-sampledDistance)), fastexp(sigmaT[1] * (-sampledDistance)), fastexp(sigmaT[2] * (-sampledDistance)));
        mRec.pdfSuccess = mRec.pdfSuccess * m.mediumSamplingWeight;
        mRec.pdfFailure = m.mediumSamplingWeight * mRec.pdfFailure + (1 - m.mediumSamplingWeight);
        if (maxComp(mRec.transmittance) < 1e-20f) mRec.transmittance = Spectrum(0.0f);
        return success;
    }
    // heterogeneous.cpp:613-658 (Woodcock): pdfs and transmittance are placeholders
    mRec.pdfFailure = 1.0f; mRec.pdfSuccess = 1.0f; mRec.transmittance = Spectrum(1.0f);
    float mint, maxt;
    if (!densityBox(m, o, d, rayMint, rayMaxt, mint, maxt)) return false;
    float tHit = 0.0f, densityAtT = 0.0f;
    if (woodcockWalk(m, o, d, mint, maxt, 1, smp, tHit, densityAtT)) return false;
    if (densityAtT == 0.0f) return false; // only after B2_MAX_WOODCOCK_STEPS (a real collision needs density > 0)
    mRec.t = tHit; mRec.p = o + d * tHit;
    mRec.sigmaS = V3(m.albedo[0], m.albedo[1], m.albedo[2]) * densityAtT;
    mRec.transmittance = Spectrum(1.0f / densityAtT);
    if (!isfinite(mRec.transmittance.x)) mRec.transmittance = Spectrum(0.0f);
    return true;
}

} // namespace b2
