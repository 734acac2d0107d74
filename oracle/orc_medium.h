/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle), never linked into the product.
 *
 * Participating media on the `volpath` path, float32, operation order of the reference:
 *   HeterogeneousMedium, Woodcock tracking   src/medium/heterogeneous.cpp:546-660 (evalTransmittance, sampleDistance)
 *   HomogeneousMedium                        src/medium/homogeneous.cpp:266-362 (strategies `balance`, `single`, `manual`)
 *   GridDataSource::lookupFloat              src/volume/gridvolume.cpp:186-199 (worldToGrid), :336-371 (trilinear)
 *   Isotropic / Henyey-Greenstein phase      src/phase/isotropic.cpp:62-79, src/phase/hg.cpp:76-115
 *
 * Parity status: UNPINNED by the reference (it ships no test of media or of volpath); pinned here against closed
 * forms (tests/test_oracle_volpath.py: Beer-Lambert slabs, trilinear lookups, phase normalisation, furnace).
 */
#pragma once
#include "orc_math.h"
#include "orc_sampler.h"
#include <cmath>
#include <cstdint>

extern "C" {
typedef struct OrcMedium {
    int32_t type;           /* 0 homogeneous, 1 heterogeneous (method woodcock; `simpson` is not on the path) */
    int32_t phase;          /* 0 isotropic, 1 hg */
    float g;                /* hg.cpp:49 */
    float sigmaA[3], sigmaS[3]; /* homogeneous (medium.cpp:30-36); unused by heterogeneous */
    int32_t strategy;       /* homogeneous.cpp:186-222: 0 balance (default), 1 single, 2 manual */
    float samplingDensity;  /* single: sigmaT[channel]; manual: property */
    float mediumSamplingWeight; /* homogeneous.cpp:160-183 (after the max(.,0.5) clamp) */
    float scale;            /* heterogeneous.cpp:185 */
    float albedo[3];        /* heterogeneous: constvolume `albedo` */
    int32_t res[3];         /* gridvolume resolution (x, y, z) */
    float worldToGrid[12];  /* rows of the affine map m_worldToGrid (gridvolume.cpp:186-193) */
    float aabbMin[3], aabbMax[3]; /* m_aabb: world-space box of the transformed data box (gridvolume.cpp:197-199) */
    const float *density;   /* res.x * res.y * res.z float32 values in [0, 1], x fastest */
} OrcMedium;
}

namespace orc {

/* warp.cpp:25-31 */
inline V3 squareToUniformSphere(float sx, float sy) {
    float z = 1.0f - 2.0f * sy;
    float r = safe_sqrt(1.0f - z * z);
    float sinPhi, cosPhi;
    sincosf(2.0f * kPi * sx, &sinPhi, &cosPhi);
    return V3(r * cosPhi, r * sinPhi, z);
}
static const float kInvFourPi = 0.07957747154594766788f;

struct MediumSamplingRecord { /* include/mitsuba/render/medium.h:34-82 */
    float t = 0; V3 p; Spectrum sigmaA, sigmaS; float pdfFailure = 1, pdfSuccess = 1; Spectrum transmittance;
};

struct MediumEval {
    const OrcMedium &m;
    explicit MediumEval(const OrcMedium &mm) : m(mm) {}

    /* ---- phase functions: wi = direction the light came from reversed (-ray.d), wo = new direction ---- */
    float phaseEval(const V3 &wi, const V3 &wo) const {
        if (m.phase == 0) return kInvFourPi;                      /* isotropic.cpp:75-77 */
        float temp = 1.0f + m.g * m.g + 2.0f * m.g * dot(wi, wo); /* hg.cpp:105-108 */
        return kInvFourPi * (1 - m.g * m.g) / (temp * std::sqrt(temp));
    }
    /* sample(pRec, pdf, sampler): returns the weight (1) and the pdf; draws one 2D sample */
    float phaseSample(const V3 &wi, V3 &wo, float &pdf, Sampler *sampler) const {
        float sx, sy; sampler->next2D(sx, sy);
        if (m.phase == 0) { /* isotropic.cpp:69-73 */
            wo = squareToUniformSphere(sx, sy);
            pdf = kInvFourPi;
            return 1.0f;
        }
        float cosTheta; /* hg.cpp:76-100 */
        if (std::abs(m.g) < kEpsilon) cosTheta = 1 - 2 * sx;
        else {
            float sqrTerm = (1 - m.g * m.g) / (1 - m.g + 2 * m.g * sx);
            cosTheta = (1 + m.g * m.g - sqrTerm * sqrTerm) / (2 * m.g);
        }
        float sinTheta = safe_sqrt(1.0f - cosTheta * cosTheta), sinPhi, cosPhi;
        sincosf(2 * kPi * sy, &sinPhi, &cosPhi);
        wo = Frame(-wi).toWorld(V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta));
        pdf = phaseEval(wi, wo);
        return 1.0f;
    }

    /* ---- gridvolume.cpp:336-371 ---- */
    float lookupDensity(const V3 &wp) const {
        const float *M = m.worldToGrid;
        const V3 p(M[0] * wp.x + M[1] * wp.y + M[2] * wp.z + M[3], M[4] * wp.x + M[5] * wp.y + M[6] * wp.z + M[7],
                   M[8] * wp.x + M[9] * wp.y + M[10] * wp.z + M[11]);
        const int x1 = (int) std::floor(p.x), y1 = (int) std::floor(p.y), z1 = (int) std::floor(p.z), x2 = x1 + 1, y2 = y1 + 1, z2 = z1 + 1;
        if (x1 < 0 || y1 < 0 || z1 < 0 || x2 >= m.res[0] || y2 >= m.res[1] || z2 >= m.res[2]) return 0;
        const float fx = p.x - x1, fy = p.y - y1, fz = p.z - z1, _fx = 1.0f - fx, _fy = 1.0f - fy, _fz = 1.0f - fz;
        const float *D = m.density;
        const int rx = m.res[0], ry = m.res[1];
        const float d000 = D[(z1 * ry + y1) * rx + x1], d001 = D[(z1 * ry + y1) * rx + x2], d010 = D[(z1 * ry + y2) * rx + x1],
                    d011 = D[(z1 * ry + y2) * rx + x2], d100 = D[(z2 * ry + y1) * rx + x1], d101 = D[(z2 * ry + y1) * rx + x2],
                    d110 = D[(z2 * ry + y2) * rx + x1], d111 = D[(z2 * ry + y2) * rx + x2];
        return ((d000 * _fx + d001 * fx) * _fy + (d010 * _fx + d011 * fx) * fy) * _fz +
               ((d100 * _fx + d101 * fx) * _fy + (d110 * _fx + d111 * fx) * fy) * fz;
    }

    bool densityBox(const Ray &ray, float &mint, float &maxt) const {
        AABB b; b.min = V3(m.aabbMin[0], m.aabbMin[1], m.aabbMin[2]); b.max = V3(m.aabbMax[0], m.aabbMax[1], m.aabbMax[2]);
        if (!b.rayIntersect(ray, mint, maxt)) return false;
        mint = std::max(mint, ray.mint);
        maxt = std::min(maxt, ray.maxt);
        return true;
    }

    /* Medium::evalTransmittance(ray, sampler) */
    Spectrum evalTransmittance(const Ray &ray, Sampler *sampler) const {
        if (m.type == 0) { /* homogeneous.cpp:266-273 */
            float negLength = ray.mint - ray.maxt;
            Spectrum tr;
            for (int i = 0; i < 3; ++i) {
                float sT = m.sigmaA[i] + m.sigmaS[i];
                tr[i] = sT != 0 ? fastexp(sT * negLength) : 1.0f;
            }
            return tr;
        }
        /* heterogeneous.cpp:546-585: two Woodcock walks, result = fraction that got through */
        float mint, maxt;
        if (!densityBox(ray, mint, maxt)) return Spectrum(1.0f);
        const float invMaxDensity = 1.0f / (m.scale * 1.0f); /* gridvolume.cpp:583-585: maximum value 1 */
        const int nSamples = 2;
        float result = 0;
        uint32_t steps = 0; /* not in the reference: both restatement and CUDA code stop a stuck random stream after 2^20 steps */
        for (int i = 0; i < nSamples; ++i) {
            float t = mint;
            while (true) {
                t -= fastlog(1 - sampler->next1D()) * invMaxDensity;
                if (t >= maxt) { result += 1; break; }
                V3 p = ray(t);
                float density = lookupDensity(p) * m.scale;
                if (density * invMaxDensity > sampler->next1D()) break;
                if (++steps > (1u << 20)) break;
            }
        }
        return Spectrum(result / nSamples);
    }

    /* Medium::sampleDistance(ray, mRec, sampler) */
    bool sampleDistance(const Ray &ray, MediumSamplingRecord &mRec, Sampler *sampler) const {
        if (m.type == 0) { /* homogeneous.cpp:275-362 */
            float rand = sampler->next1D(), sampledDistance;
            float samplingDensity = m.samplingDensity;
            const float sigmaT[3] = {m.sigmaA[0] + m.sigmaS[0], m.sigmaA[1] + m.sigmaS[1], m.sigmaA[2] + m.sigmaS[2]};
            if (rand < m.mediumSamplingWeight) {
                rand /= m.mediumSamplingWeight;
                if (m.strategy == 0) {
                    int channel = std::min((int) (sampler->next1D() * 3), 2);
                    samplingDensity = sigmaT[channel];
                }
                sampledDistance = -fastlog(1 - rand) / samplingDensity;
            } else sampledDistance = kInf;
            float distSurf = ray.maxt - ray.mint;
            bool success = true;
            if (sampledDistance < distSurf) {
                mRec.t = sampledDistance + ray.mint;
                mRec.p = ray(mRec.t);
                mRec.sigmaA = V3(m.sigmaA[0], m.sigmaA[1], m.sigmaA[2]);
                mRec.sigmaS = V3(m.sigmaS[0], m.sigmaS[1], m.sigmaS[2]);
                if (mRec.p.x == ray.o.x && mRec.p.y == ray.o.y && mRec.p.z == ray.o.z) success = false;
            } else { sampledDistance = distSurf; success = false; }
            if (m.strategy == 0) {
                mRec.pdfFailure = 0; mRec.pdfSuccess = 0;
                for (int i = 0; i < 3; ++i) {
                    float tmp = fastexp(-sigmaT[i] * sampledDistance);
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
            if (mRec.transmittance.max() < 1e-20f) mRec.transmittance = Spectrum(0.0f);
            return success;
        }
        /* heterogeneous.cpp:613-658 (Woodcock): pdfs and transmittance are placeholders */
        mRec.pdfFailure = 1.0f; mRec.pdfSuccess = 1.0f; mRec.transmittance = Spectrum(1.0f);
        float mint, maxt;
        if (!densityBox(ray, mint, maxt)) return false;
        const float invMaxDensity = 1.0f / (m.scale * 1.0f);
        float t = mint, densityAtT = 0;
        bool success = false;
        uint32_t steps = 0;
        while (true) {
            t -= fastlog(1 - sampler->next1D()) * invMaxDensity;
            if (t >= maxt) break;
            V3 p = ray(t);
            densityAtT = lookupDensity(p) * m.scale;
            if (densityAtT * invMaxDensity > sampler->next1D()) {
                mRec.t = t; mRec.p = p;
                Spectrum albedo(m.albedo[0], m.albedo[1], m.albedo[2]);
                mRec.sigmaS = albedo * densityAtT;
                mRec.sigmaA = Spectrum(densityAtT) - mRec.sigmaS;
                mRec.transmittance = Spectrum(densityAtT != 0.0f ? 1.0f / densityAtT : 0);
                if (!std::isfinite(mRec.transmittance[0])) mRec.transmittance = Spectrum(0.0f);
                success = true;
                break;
            }
            if (++steps > (1u << 20)) break;
        }
        return success && mRec.pdfSuccess > 0;
    }
};

} // namespace orc
