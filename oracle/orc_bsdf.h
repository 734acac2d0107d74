/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
 * BSDFs named by BASELINE.json's north_star: diffuse (src/bsdfs/diffuse.cpp:110-150),
 * roughconductor (src/bsdfs/roughconductor.cpp:257-420), roughdielectric
 * (src/bsdfs/roughdielectric.cpp:270-614), coating (src/bsdfs/coating.cpp:208-376) and the
 * MicrofacetDistribution helper (src/bsdfs/microfacet.h:45-721).  All parameters are constants
 * (ConstantSpectrumTexture / ConstantFloatTexture), as in the configs of BASELINE.json. */
#pragma once
#include "orc_math.h"
#include "orc_sampler.h"
#include "orc_texture.h"

extern "C" {
/* plain-C description of one BSDF node; `nested` indexes the same array (coating only) */
typedef struct OrcBsdf {
    int32_t type;          /* 0 diffuse, 1 roughconductor, 2 roughdielectric, 3 coating, 4 null (src/bsdfs/null.cpp), 5 twosided,
                              6 dielectric, 7 conductor, 8 plastic (src/bsdfs/{twosided,dielectric,conductor,plastic}.cpp) */
    int32_t distr;         /* 0 beckmann, 1 ggx, 2 phong/as   (microfacet.h:48-57) */
    int32_t sampleVisible; /* microfacet.h:138 default true; forced false for phong :145-148 */
    int32_t nested;        /* coating: index of the nested BSDF; else -1 */
    float alphaU, alphaV;  /* before the 1e-4 clamp of microfacet.h:70-71 */
    float eta;             /* dielectric/coating: intIOR/extIOR; conductor: unused */
    float thickness;       /* coating.cpp:126 */
    float reflectance[3];  /* diffuse: reflectance; others: specularReflectance */
    float transmittance[3];/* roughdielectric specularTransmittance */
    float etaC[3], kC[3];  /* roughconductor eta, k (already divided by extEta, roughconductor.cpp:189-190) */
    float sigmaA[3];       /* coating.cpp:129-130 */
    int32_t nested2;       /* twosided: BSDF of the back side (= nested when only one child was given, twosided.cpp:89-90) */
    float diffuseReflectance[3]; /* plastic.cpp:158-159 */
    float fdrInt, fdrExt;  /* plastic.cpp:194-195 fresnelDiffuseReflectance(1/eta), (eta), integrated on the host */
    float specSamplingWeight;    /* plastic.cpp:199-202 */
    int32_t nonlinear;     /* plastic.cpp:161 */
    int32_t texture;       /* index of a `bitmap` texture (src/textures/bitmap.cpp) bound to `reflectance` of diffuse, `specularReflectance` of roughconductor / conductor or `diffuseReflectance` of plastic, -1 = constant */
} OrcBsdf;
}

namespace orc {

/* include/mitsuba/render/bsdf.h:224-285 */
enum BsdfFlags {
    ENull = 0x00001, EDiffuseReflection = 0x00002, EDiffuseTransmission = 0x00004, EGlossyReflection = 0x00008,
    EGlossyTransmission = 0x00010, EDeltaReflection = 0x00020, EDeltaTransmission = 0x00040,
    EDelta1DReflection = 0x00080, EDelta1DTransmission = 0x00100,
    EAnisotropic = 0x01000, ESpatiallyVarying = 0x02000, ENonSymmetric = 0x04000, EFrontSide = 0x08000,
    EBackSide = 0x10000, EUsesSampler = 0x20000,
    EReflection = EDiffuseReflection | EDeltaReflection | EDelta1DReflection | EGlossyReflection,
    ETransmission = EDiffuseTransmission | EDeltaTransmission | EDelta1DTransmission | EGlossyTransmission | ENull,
    EDiffuse = EDiffuseReflection | EDiffuseTransmission, EGlossy = EGlossyReflection | EGlossyTransmission,
    ESmooth = EDiffuse | EGlossy, EDelta = ENull | EDeltaReflection | EDeltaTransmission,
    EDelta1D = EDelta1DReflection | EDelta1DTransmission, EAll = EDiffuse | EGlossy | EDelta | EDelta1D
};

/* src/bsdfs/microfacet.h */
struct Microfacet {
    int type; float alphaU, alphaV; bool visible; float exponentU = 0, exponentV = 0;
    Microfacet(int t, float aU, float aV, bool sv) : type(t), alphaU(aU), alphaV(aV), visible(sv) {
        alphaU = std::max(alphaU, 1e-4f); alphaV = std::max(alphaV, 1e-4f); /* :70-71 */
        if (type == 2) computePhongExponent();
    }
    void computePhongExponent() { /* :693-696 */
        exponentU = std::max(2.0f / (alphaU * alphaU) - 2.0f, 0.0f);
        exponentV = std::max(2.0f / (alphaV * alphaV) - 2.0f, 0.0f);
    }
    bool isIsotropic() const { return alphaU == alphaV; }
    void scaleAlpha(float v) { alphaU *= v; alphaV *= v; if (type == 2) computePhongExponent(); } /* :183-188 */
    float interpolatePhongExponent(const V3 &v) const { /* :554-565 */
        const float sinTheta2 = Frame::sinTheta2(v);
        if (isIsotropic() || sinTheta2 <= kRcpOverflow) return exponentU;
        float invSinTheta2 = 1 / sinTheta2;
        float cosPhi2 = v.x * v.x * invSinTheta2, sinPhi2 = v.y * v.y * invSinTheta2;
        return exponentU * cosPhi2 + exponentV * sinPhi2;
    }
    float eval(const V3 &m) const { /* :191-234 */
        if (Frame::cosTheta(m) <= 0) return 0.0f;
        float cosTheta2 = Frame::cosTheta2(m);
        float beckmannExponent = ((m.x * m.x) / (alphaU * alphaU) + (m.y * m.y) / (alphaV * alphaV)) / cosTheta2;
        float result;
        switch (type) {
            case 0: result = fastexp(-beckmannExponent) / (kPi * alphaU * alphaV * cosTheta2 * cosTheta2); break;
            case 1: { float root = (1.0f + beckmannExponent) * cosTheta2;
                      result = 1.0f / (kPi * alphaU * alphaV * root * root); } break;
            default: { float exponent = interpolatePhongExponent(m);
                      result = std::sqrt((exponentU + 2) * (exponentV + 2)) * kInvTwoPi * std::pow(Frame::cosTheta(m), exponent); } break;
        }
        if (result * Frame::cosTheta(m) < 1e-20f) result = 0;
        return result;
    }
    float projectRoughness(const V3 &v) const { /* :541-551 */
        float invSinTheta2 = 1 / Frame::sinTheta2(v);
        if (isIsotropic() || invSinTheta2 <= 0) return alphaU;
        float cosPhi2 = v.x * v.x * invSinTheta2, sinPhi2 = v.y * v.y * invSinTheta2;
        return std::sqrt(cosPhi2 * alphaU * alphaU + sinPhi2 * alphaV * alphaV);
    }
    float smithG1(const V3 &v, const V3 &m) const { /* :477-514 */
        if (dot(v, m) * Frame::cosTheta(v) <= 0) return 0.0f;
        float tanTheta = std::abs(Frame::tanTheta(v));
        if (tanTheta == 0.0f) return 1.0f;
        float alpha = projectRoughness(v);
        if (type == 1) {
            float root = alpha * tanTheta;
            return 2.0f / (1.0f + hypot2(1.0f, root));
        }
        float a = 1.0f / (alpha * tanTheta);
        if (a >= 1.6f) return 1.0f;
        float aSqr = a * a;
        return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
    }
    float G(const V3 &wi, const V3 &wo, const V3 &m) const { return smithG1(wi, m) * smithG1(wo, m); }
    void sampleFirstQuadrant(float u1, float &phi, float &exponent) const { /* :699-708 */
        float cosPhi, sinPhi;
        phi = std::atan(std::sqrt((exponentU + 2.0f) / (exponentV + 2.0f)) * std::tan(kPi * u1 * 0.5f));
        sincos(phi, &sinPhi, &cosPhi);
        exponent = exponentU * cosPhi * cosPhi + exponentV * sinPhi * sinPhi;
    }
    V3 sampleAll(float sx, float sy, float &pdf) const { /* :287-395 */
        float cosThetaM = 0.0f, sinPhiM, cosPhiM, alphaSqr;
        if (type == 0 || type == 1) {
            if (isIsotropic()) {
                sincos((2.0f * kPi) * sy, &sinPhiM, &cosPhiM);
                alphaSqr = alphaU * alphaU;
            } else {
                float phiM = std::atan(alphaV / alphaU * std::tan(kPi + 2 * kPi * sy)) + kPi * std::floor(2 * sy + 0.5f);
                sincos(phiM, &sinPhiM, &cosPhiM);
                float cosSc = cosPhiM / alphaU, sinSc = sinPhiM / alphaV;
                alphaSqr = 1.0f / (cosSc * cosSc + sinSc * sinSc);
            }
            if (type == 0) {
                float tanThetaMSqr = alphaSqr * -fastlog(1.0f - sx);
                cosThetaM = 1.0f / std::sqrt(1.0f + tanThetaMSqr);
                pdf = (1.0f - sx) / (kPi * alphaU * alphaV * cosThetaM * cosThetaM * cosThetaM);
            } else {
                float tanThetaMSqr = alphaSqr * sx / (1.0f - sx);
                cosThetaM = 1.0f / std::sqrt(1.0f + tanThetaMSqr);
                float temp = 1 + tanThetaMSqr / alphaSqr;
                pdf = kInvPi / (alphaU * alphaV * cosThetaM * cosThetaM * cosThetaM * temp * temp);
            }
        } else {
            float phiM, exponent;
            if (isIsotropic()) {
                phiM = (2.0f * kPi) * sy; exponent = exponentU;
            } else {
                if (sy < 0.25f) { sampleFirstQuadrant(4 * sy, phiM, exponent); }
                else if (sy < 0.5f) { sampleFirstQuadrant(4 * (0.5f - sy), phiM, exponent); phiM = kPi - phiM; }
                else if (sy < 0.75f) { sampleFirstQuadrant(4 * (sy - 0.5f), phiM, exponent); phiM += kPi; }
                else { sampleFirstQuadrant(4 * (1 - sy), phiM, exponent); phiM = 2 * kPi - phiM; }
            }
            sincos(phiM, &sinPhiM, &cosPhiM);
            cosThetaM = std::pow(sx, 1.0f / (exponent + 2.0f));
            pdf = std::sqrt((exponentU + 2.0f) * (exponentV + 2.0f)) * kInvTwoPi * std::pow(cosThetaM, exponent + 1.0f);
        }
        if (pdf < 1e-20f) pdf = 0;
        float sinThetaM = std::sqrt(std::max(0.0f, 1 - cosThetaM * cosThetaM));
        return V3(sinThetaM * cosPhiM, sinThetaM * sinPhiM, cosThetaM);
    }
    float pdfAll(const V3 &m) const { return eval(m) * Frame::cosTheta(m); } /* :404-407 */
    void sampleVisible11(float thetaI, float sx, float sy, float &slopeX, float &slopeY) const { /* :573-690 */
        const float SQRT_PI_INV = 1 / std::sqrt(kPi);
        if (type == 0) {
            if (thetaI < 1e-4f) {
                float sinPhi, cosPhi;
                float r = std::sqrt(-fastlog(1.0f - sx));
                sincos(2 * kPi * sy, &sinPhi, &cosPhi);
                slopeX = r * cosPhi; slopeY = r * sinPhi; return;
            }
            float tanThetaI = std::tan(thetaI);
            float cotThetaI = 1 / tanThetaI;
            float a = -1, c = erf_as(cotThetaI);
            float sample_x = std::max(sx, 1e-6f);
            float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
            float b = c - (1 + c) * std::pow(1 - sample_x, fit);
            float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * std::exp(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5f * (a + c);
                float invErf = erfinv(b);
                float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * std::exp(-invErf * invErf)) - sample_x;
                float derivative = normalization * (1 - invErf * tanThetaI);
                if (std::abs(value) < 1e-5f) break;
                if (value > 0) c = b; else a = b;
                b -= value / derivative;
            }
            slopeX = erfinv(b);
            slopeY = erfinv(2.0f * std::max(sy, 1e-6f) - 1.0f);
        } else {
            if (thetaI < 1e-4f) {
                float sinPhi, cosPhi;
                float r = safe_sqrt(sx / (1 - sx));
                sincos(2 * kPi * sy, &sinPhi, &cosPhi);
                slopeX = r * cosPhi; slopeY = r * sinPhi; return;
            }
            float tanThetaI = std::tan(thetaI);
            float a = 1 / tanThetaI;
            float G1 = 2.0f / (1.0f + safe_sqrt(1.0f + 1.0f / (a * a)));
            float A = 2.0f * sx / G1 - 1.0f;
            if (std::abs(A) == 1) A -= signum(A) * kEpsilon;
            float tmp = 1.0f / (A * A - 1.0f);
            float B = tanThetaI;
            float D = safe_sqrt(B * B * tmp * tmp - (A * A - B * B) * tmp);
            float slope_x_1 = B * tmp - D;
            float slope_x_2 = B * tmp + D;
            slopeX = (A < 0.0f || slope_x_2 > 1.0f / tanThetaI) ? slope_x_1 : slope_x_2;
            float S;
            if (sy > 0.5f) { S = 1.0f; sy = 2.0f * (sy - 0.5f); }
            else { S = -1.0f; sy = 2.0f * (0.5f - sy); }
            float z = (sy * (sy * (sy * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) + 0.000152998850436920f) /
                      (sy * (sy * (sy * (sy * 0.169507819808272f - 0.397203533833404f) - 0.232500544458471f) + 1.0f) - 0.539825872510702f);
            slopeY = S * z * std::sqrt(1.0f + slopeX * slopeX);
        }
    }
    V3 sampleVisible(const V3 &_wi, float sx, float sy) const { /* :421-459 */
        V3 wi = normalize(V3(alphaU * _wi.x, alphaV * _wi.y, _wi.z));
        float theta = 0, phi = 0;
        if (wi.z < 0.99999f) { theta = std::acos(wi.z); phi = std::atan2(wi.y, wi.x); }
        float sinPhi, cosPhi;
        sincos(phi, &sinPhi, &cosPhi);
        float slx, sly;
        sampleVisible11(theta, sx, sy, slx, sly);
        float rx = cosPhi * slx - sinPhi * sly, ry = sinPhi * slx + cosPhi * sly;
        rx *= alphaU; ry *= alphaV;
        float normalization = 1.0f / std::sqrt(rx * rx + ry * ry + 1.0f);
        return V3(-rx * normalization, -ry * normalization, normalization);
    }
    float pdfVisible(const V3 &wi, const V3 &m) const { /* :462-466 */
        if (Frame::cosTheta(wi) == 0) return 0.0f;
        return smithG1(wi, m) * absDot(wi, m) * eval(m) / std::abs(Frame::cosTheta(wi));
    }
    V3 sample(const V3 &wi, float sx, float sy, float &pdf) const { /* :236-246 */
        V3 m;
        if (visible) { m = sampleVisible(wi, sx, sy); pdf = pdfVisible(wi, m); }
        else m = sampleAll(sx, sy, pdf);
        return m;
    }
    float pdf(const V3 &wi, const V3 &m) const { return visible ? pdfVisible(wi, m) : pdfAll(m); } /* :266-271 */
};

/* BSDFSamplingRecord (include/mitsuba/render/bsdf.h:123-192) restricted to what `path` sets:
 * typeMask = EAll, component = -1, mode = ERadiance. */
struct TexCtx { /* what `m_reflectance->eval(bRec.its)` reads of the intersection: uv and, once computePartials() ran, the uv partials */
    const Texture *textures = nullptr;
    float u = 0, v = 0; bool hasUVPartials = false; float dudx = 0, dudy = 0, dvdx = 0, dvdy = 0;
};
struct BRec {
    V3 wi, wo; float eta = 1; uint32_t sampledType = 0; Sampler *sampler = nullptr; const TexCtx *its = nullptr;
};

struct BsdfSet {
    const OrcBsdf *b; int n;
    static Microfacet distr(const OrcBsdf &d) { return Microfacet(d.distr, d.alphaU, d.alphaV, d.sampleVisible != 0); }

    /* combined type flags as BSDF::configure() ORs the components (bsdf.cpp) */
    uint32_t type(int id) const {
        const OrcBsdf &d = b[id];
        switch (d.type) {
            case 0: if (d.texture >= 0) return EDiffuseReflection | EFrontSide | ESpatiallyVarying; /* a texture that is not all black */
                    return (std::max(std::max(d.reflectance[0], d.reflectance[1]), d.reflectance[2]) > 0) ? (EDiffuseReflection | EFrontSide) : 0; /* diffuse.cpp:98-103 */
            case 1: return EGlossyReflection | EFrontSide | (d.alphaU != d.alphaV ? EAnisotropic : 0);                  /* roughconductor.cpp:229-238 */
            case 2: return EGlossyReflection | EGlossyTransmission | EFrontSide | EBackSide | EUsesSampler | ENonSymmetric | (d.alphaU != d.alphaV ? EAnisotropic : 0); /* roughdielectric.cpp:240-255 */
            case 4: return ENull | EFrontSide | EBackSide;                                                              /* null.cpp:38-43 */
            case 5: return ((type(d.nested) & ~EBackSide) | EFrontSide) | ((type(d.nested2) & ~EFrontSide) | EBackSide);  /* twosided.cpp:96-102 */
            case 6: return EDeltaReflection | EDeltaTransmission | EFrontSide | EBackSide | ENonSymmetric;               /* dielectric.cpp:190-194 */
            case 7: return EDeltaReflection | EFrontSide;                                                               /* conductor.cpp:181-183 */
            case 8: return EDeltaReflection | EDiffuseReflection | EFrontSide;                                          /* plastic.cpp:211-215 */
            default: return type(d.nested) | EDeltaReflection | EFrontSide | EBackSide;                                  /* coating.cpp:160-171 */
        }
    }

    /* m_reflectance->eval(bRec.its) (diffuse.cpp:115,148) / m_specularReflectance->eval(bRec.its) (roughconductor.cpp:285,369,415; conductor.cpp:221-256):
       constant, or the bitmap texture at the intersection */
    static V3 reflectance(const OrcBsdf &d, const BRec &r) {
        if ((d.type == 0 || d.type == 1 || d.type == 7) && d.texture >= 0 && r.its) { /* diffuse `reflectance`; roughconductor / conductor `specularReflectance` */
            const TexCtx &t = *r.its;
            return t.textures[d.texture].eval(t.u, t.v, t.hasUVPartials, t.dudx, t.dudy, t.dvdx, t.dvdy);
        }
        return V3(d.reflectance[0], d.reflectance[1], d.reflectance[2]);
    }
    /* m_diffuseReflectance->eval(bRec.its) of plastic (plastic.cpp:271,304,415): constant or the bitmap texture */
    static V3 diffuseReflectance(const OrcBsdf &d, const BRec &r) {
        if (d.type == 8 && d.texture >= 0 && r.its) {
            const TexCtx &t = *r.its;
            return t.textures[d.texture].eval(t.u, t.v, t.hasUVPartials, t.dudx, t.dudy, t.dvdx, t.dvdy);
        }
        return V3(d.diffuseReflectance[0], d.diffuseReflectance[1], d.diffuseReflectance[2]);
    }
    /* BSDF::usesRayDifferentials(): diffuse.cpp:101, plastic.cpp:196-197, twosided.cpp:91-92; the other plugins take constants here */
    bool usesRayDifferentials(int id) const {
        const OrcBsdf &d = b[id];
        if (d.type == 0 || d.type == 8 || d.type == 1 || d.type == 7) return d.texture >= 0; /* roughconductor.cpp:223-227, conductor.cpp:178-179 */
        if (d.type == 5) return usesRayDifferentials(d.nested) || usesRayDifferentials(d.nested2);
        if (d.type == 3) return usesRayDifferentials(d.nested); /* coating.cpp:172-174 */
        return false;
    }

    Spectrum eval(int id, const BRec &r, bool discrete = false) const {
        const OrcBsdf &d = b[id];
        const V3 R = reflectance(d, r);
        switch (d.type) {
        case 0: { /* diffuse.cpp:110-118 */
            if (discrete || type(id) == 0 || Frame::cosTheta(r.wi) <= 0 || Frame::cosTheta(r.wo) <= 0) return Spectrum(0.0f);
            return R * (kInvPi * Frame::cosTheta(r.wo));
        }
        case 1: { /* roughconductor.cpp:257-297 */
            if (discrete || Frame::cosTheta(r.wi) <= 0 || Frame::cosTheta(r.wo) <= 0) return Spectrum(0.0f);
            V3 H = normalize(r.wo + r.wi);
            Microfacet ds = distr(d);
            const float D = ds.eval(H);
            if (D == 0) return Spectrum(0.0f);
            const Spectrum F = fresnelConductorExact(dot(r.wi, H), V3(d.etaC[0], d.etaC[1], d.etaC[2]), V3(d.kC[0], d.kC[1], d.kC[2])) * R;
            const float G = ds.G(r.wi, r.wo, H);
            float model = D * G / (4.0f * Frame::cosTheta(r.wi));
            return F * model;
        }
        case 2: { /* roughdielectric.cpp:270-349 */
            if (discrete || Frame::cosTheta(r.wi) == 0) return Spectrum(0.0f);
            const float m_eta = d.eta, m_invEta = 1 / d.eta;
            bool reflect = Frame::cosTheta(r.wi) * Frame::cosTheta(r.wo) > 0;
            V3 H;
            if (reflect) H = normalize(r.wo + r.wi);
            else { float eta = Frame::cosTheta(r.wi) > 0 ? m_eta : m_invEta; H = normalize(r.wi + r.wo * eta); }
            H *= signum(Frame::cosTheta(H));
            Microfacet ds = distr(d);
            const float D = ds.eval(H);
            if (D == 0) return Spectrum(0.0f);
            const float F = fresnelDielectricExt(dot(r.wi, H), m_eta);
            const float G = ds.G(r.wi, r.wo, H);
            if (reflect) {
                float value = F * D * G / (4.0f * std::abs(Frame::cosTheta(r.wi)));
                return R * value;
            } else {
                float eta = Frame::cosTheta(r.wi) > 0.0f ? m_eta : m_invEta;
                float sqrtDenom = dot(r.wi, H) + eta * dot(r.wo, H);
                float value = ((1 - F) * D * G * eta * eta * dot(r.wi, H) * dot(r.wo, H)) / (Frame::cosTheta(r.wi) * sqrtDenom * sqrtDenom);
                float factor = Frame::cosTheta(r.wi) > 0 ? m_invEta : m_eta; /* mode == ERadiance */
                return V3(d.transmittance[0], d.transmittance[1], d.transmittance[2]) * std::abs(value * factor * factor);
            }
        }
        case 4: return Spectrum(discrete ? 1.0f : 0.0f); /* null.cpp:45-47 (typeMask contains ENull) */
        case 5: { /* twosided.cpp:109-121 */
            if (Frame::cosTheta(r.wi) > 0) return eval(d.nested, r, discrete);
            BRec b = r; b.wi.z *= -1; b.wo.z *= -1;
            return eval(d.nested2, b, discrete);
        }
        case 6: { /* dielectric.cpp:229-255 */
            float cosThetaT;
            float F = fresnelDielectricExt(Frame::cosTheta(r.wi), cosThetaT, d.eta);
            const float invEta = 1 / d.eta;
            if (Frame::cosTheta(r.wi) * Frame::cosTheta(r.wo) >= 0) {
                if (!discrete || std::abs(dot(V3(-r.wi.x, -r.wi.y, r.wi.z), r.wo) - 1) > kDeltaEpsilon) return Spectrum(0.0f);
                return R * F;
            } else {
                float scale = -(cosThetaT < 0 ? invEta : d.eta);
                if (!discrete || std::abs(dot(V3(scale * r.wi.x, scale * r.wi.y, cosThetaT), r.wo) - 1) > kDeltaEpsilon) return Spectrum(0.0f);
                float factor = cosThetaT < 0 ? invEta : d.eta; /* ERadiance */
                return V3(d.transmittance[0], d.transmittance[1], d.transmittance[2]) * factor * factor * (1 - F);
            }
        }
        case 7: { /* conductor.cpp:221-235 */
            if (!discrete || Frame::cosTheta(r.wi) <= 0 || Frame::cosTheta(r.wo) <= 0 ||
                std::abs(dot(V3(-r.wi.x, -r.wi.y, r.wi.z), r.wo) - 1) > kDeltaEpsilon) return Spectrum(0.0f);
            return R * fresnelConductorExact(Frame::cosTheta(r.wi), V3(d.etaC[0], d.etaC[1], d.etaC[2]), V3(d.kC[0], d.kC[1], d.kC[2]));
        }
        case 8: { /* plastic.cpp:243-279 */
            if (Frame::cosTheta(r.wo) <= 0 || Frame::cosTheta(r.wi) <= 0) return Spectrum(0.0f);
            float Fi = fresnelDielectricExt(Frame::cosTheta(r.wi), d.eta);
            if (discrete) {
                if (std::abs(dot(V3(-r.wi.x, -r.wi.y, r.wi.z), r.wo) - 1) < kDeltaEpsilon) return R * Fi;
                return Spectrum(0.0f);
            }
            float Fo = fresnelDielectricExt(Frame::cosTheta(r.wo), d.eta);
            Spectrum diff = diffuseReflectance(d, r);
            if (d.nonlinear) diff = diff / (Spectrum(1.0f) - diff * d.fdrInt);
            else diff /= 1 - d.fdrInt;
            const float invEta2 = 1 / (d.eta * d.eta);
            return diff * (squareToCosineHemispherePdf(r.wo) * invEta2 * (1 - Fi) * (1 - Fo));
        }
        default: { /* coating.cpp:208-248 */
            const float m_eta = d.eta, m_invEta = 1 / d.eta;
            bool sampleNested = (type(d.nested) & EAll) != 0;
            if (discrete && std::abs(dot(V3(-r.wi.x, -r.wi.y, r.wi.z), r.wo) - 1) < kDeltaEpsilon)
                return R * fresnelDielectricExt(std::abs(Frame::cosTheta(r.wi)), m_eta);
            else if (sampleNested) {
                float R12, R21;
                BRec ri = r;
                ri.wi = refractIn(d, r.wi, R12);
                ri.wo = refractIn(d, r.wo, R21);
                if (R12 == 1 || R21 == 1) return Spectrum(0.0f);
                Spectrum result = eval(d.nested, ri, discrete) * (1 - R12) * (1 - R21);
                Spectrum sigmaA = V3(d.sigmaA[0], d.sigmaA[1], d.sigmaA[2]) * d.thickness;
                if (!sigmaA.isZero())
                    result *= expSpec(-sigmaA * (1 / std::abs(Frame::cosTheta(ri.wi)) + 1 / std::abs(Frame::cosTheta(ri.wo))));
                if (!discrete)
                    result *= m_invEta * m_invEta * Frame::cosTheta(r.wo) / Frame::cosTheta(ri.wo);
                return result;
            }
            return Spectrum(0.0f);
        }
        }
    }

    /* coating.cpp:193-205 */
    static V3 refractIn(const OrcBsdf &d, const V3 &wi, float &R) {
        float cosThetaT, invEta = 1 / d.eta;
        R = fresnelDielectricExt(std::abs(Frame::cosTheta(wi)), cosThetaT, d.eta);
        return V3(invEta * wi.x, invEta * wi.y, -signum(Frame::cosTheta(wi)) * cosThetaT);
    }
    static V3 refractOut(const OrcBsdf &d, const V3 &wi, float &R) {
        float cosThetaT, invEta = 1 / d.eta;
        R = fresnelDielectricExt(std::abs(Frame::cosTheta(wi)), cosThetaT, invEta);
        return V3(d.eta * wi.x, d.eta * wi.y, -signum(Frame::cosTheta(wi)) * cosThetaT);
    }
    float specularSamplingWeight(const OrcBsdf &d) const { /* coating.cpp:177-181 */
        V3 avg = V3(d.sigmaA[0], d.sigmaA[1], d.sigmaA[2]) * (-2 * d.thickness);
        float avgAbsorption = expSpec(avg).average();
        return 1.0f / (avgAbsorption + 1.0f);
    }

    float pdf(int id, const BRec &r, bool discrete = false) const {
        const OrcBsdf &d = b[id];
        switch (d.type) {
        case 0: /* diffuse.cpp:120-128 */
            if (discrete || type(id) == 0 || Frame::cosTheta(r.wi) <= 0 || Frame::cosTheta(r.wo) <= 0) return 0.0f;
            return squareToCosineHemispherePdf(r.wo);
        case 1: { /* roughconductor.cpp:299-326 */
            if (discrete || Frame::cosTheta(r.wi) <= 0 || Frame::cosTheta(r.wo) <= 0) return 0.0f;
            V3 H = normalize(r.wo + r.wi);
            Microfacet ds = distr(d);
            if (ds.visible) return ds.eval(H) * ds.smithG1(r.wi, H) / (4.0f * Frame::cosTheta(r.wi));
            else return ds.pdf(r.wi, H) / (4 * absDot(r.wo, H));
        }
        case 2: { /* roughdielectric.cpp:351-417 */
            if (discrete) return 0.0f;
            const float m_eta = d.eta, m_invEta = 1 / d.eta;
            bool reflect = Frame::cosTheta(r.wi) * Frame::cosTheta(r.wo) > 0;
            V3 H; float dwh_dwo;
            if (reflect) {
                H = normalize(r.wo + r.wi);
                dwh_dwo = 1.0f / (4.0f * dot(r.wo, H));
            } else {
                float eta = Frame::cosTheta(r.wi) > 0 ? m_eta : m_invEta;
                H = normalize(r.wi + r.wo * eta);
                float sqrtDenom = dot(r.wi, H) + eta * dot(r.wo, H);
                dwh_dwo = (eta * eta * dot(r.wo, H)) / (sqrtDenom * sqrtDenom);
            }
            H *= signum(Frame::cosTheta(H));
            Microfacet sd = distr(d);
            if (!sd.visible) sd.scaleAlpha(1.2f - 0.2f * std::sqrt(std::abs(Frame::cosTheta(r.wi))));
            float prob = sd.pdf(signum(Frame::cosTheta(r.wi)) * r.wi, H);
            float F = fresnelDielectricExt(dot(r.wi, H), m_eta);
            prob *= reflect ? F : (1 - F);
            return std::abs(prob * dwh_dwo);
        }
        case 4: return discrete ? 1.0f : 0.0f; /* null.cpp:49-51 */
        case 5: { /* twosided.cpp:123-135 */
            if (r.wi.z > 0) return pdf(d.nested, r, discrete);
            BRec b = r; b.wi.z *= -1; b.wo.z *= -1;
            return pdf(d.nested2, b, discrete);
        }
        case 6: { /* dielectric.cpp:257-279 */
            float cosThetaT;
            float F = fresnelDielectricExt(Frame::cosTheta(r.wi), cosThetaT, d.eta);
            const float invEta = 1 / d.eta;
            if (Frame::cosTheta(r.wi) * Frame::cosTheta(r.wo) >= 0) {
                if (!discrete || std::abs(dot(V3(-r.wi.x, -r.wi.y, r.wi.z), r.wo) - 1) > kDeltaEpsilon) return 0.0f;
                return F;
            } else {
                float scale = -(cosThetaT < 0 ? invEta : d.eta);
                if (!discrete || std::abs(dot(V3(scale * r.wi.x, scale * r.wi.y, cosThetaT), r.wo) - 1) > kDeltaEpsilon) return 0.0f;
                return 1 - F;
            }
        }
        case 7: /* conductor.cpp:237-250 */
            if (!discrete || Frame::cosTheta(r.wi) <= 0 || Frame::cosTheta(r.wo) <= 0 ||
                std::abs(dot(V3(-r.wi.x, -r.wi.y, r.wi.z), r.wo) - 1) > kDeltaEpsilon) return 0.0f;
            return 1.0f;
        case 8: { /* plastic.cpp:281-309 */
            if (Frame::cosTheta(r.wo) <= 0 || Frame::cosTheta(r.wi) <= 0) return 0.0f;
            float Fi = fresnelDielectricExt(Frame::cosTheta(r.wi), d.eta);
            float w = d.specSamplingWeight;
            float probSpecular = (Fi * w) / (Fi * w + (1 - Fi) * (1 - w));
            if (discrete) {
                if (std::abs(dot(V3(-r.wi.x, -r.wi.y, r.wi.z), r.wo) - 1) < kDeltaEpsilon) return probSpecular;
                return 0.0f;
            }
            return squareToCosineHemispherePdf(r.wo) * (1 - probSpecular);
        }
        default: { /* coating.cpp:250-286 */
            const float m_invEta = 1 / d.eta;
            bool sampleNested = (type(d.nested) & EAll) != 0;
            float R12;
            V3 wiPrime = refractIn(d, r.wi, R12);
            float w = specularSamplingWeight(d);
            float probSpecular = (R12 * w) / (R12 * w + (1 - R12) * (1 - w));
            if (discrete && std::abs(dot(V3(-r.wi.x, -r.wi.y, r.wi.z), r.wo) - 1) < kDeltaEpsilon)
                return sampleNested ? probSpecular : 1.0f;
            else if (sampleNested) {
                float R21;
                BRec ri = r;
                ri.wi = wiPrime;
                ri.wo = refractIn(d, r.wo, R21);
                if (R12 == 1 || R21 == 1) return 0.0f;
                float p = pdf(d.nested, ri, discrete);
                if (!discrete) p *= m_invEta * m_invEta * Frame::cosTheta(r.wo) / Frame::cosTheta(ri.wo);
                return p * (1 - probSpecular);
            }
            return 0.0f;
        }
        }
    }

    /* sample(bRec, pdf, sample): returns weight = f*cos/pdf; r.wo, r.eta, r.sampledType set */
    Spectrum sample(int id, BRec &r, float &pdfOut, float sx, float sy) const {
        const OrcBsdf &d = b[id];
        const V3 R = reflectance(d, r);
        switch (d.type) {
        case 0: { /* diffuse.cpp:141-150 */
            if (type(id) == 0 || Frame::cosTheta(r.wi) <= 0) return Spectrum(0.0f);
            r.wo = squareToCosineHemisphere(sx, sy);
            r.eta = 1.0f; r.sampledType = EDiffuseReflection;
            pdfOut = squareToCosineHemispherePdf(r.wo);
            return R;
        }
        case 1: { /* roughconductor.cpp:372-420 */
            if (Frame::cosTheta(r.wi) < 0) return Spectrum(0.0f);
            Microfacet ds = distr(d);
            V3 m = ds.sample(r.wi, sx, sy, pdfOut);
            if (pdfOut == 0) return Spectrum(0.0f);
            r.wo = reflect(r.wi, m);
            r.eta = 1.0f; r.sampledType = EGlossyReflection;
            if (Frame::cosTheta(r.wo) <= 0) return Spectrum(0.0f);
            Spectrum F = fresnelConductorExact(dot(r.wi, m), V3(d.etaC[0], d.etaC[1], d.etaC[2]), V3(d.kC[0], d.kC[1], d.kC[2])) * R;
            float weight;
            if (ds.visible) weight = ds.smithG1(r.wo, m);
            else weight = ds.eval(m) * ds.G(r.wi, r.wo, m) * dot(r.wi, m) / (pdfOut * Frame::cosTheta(r.wi));
            pdfOut /= 4.0f * dot(r.wo, m);
            return F * weight;
        }
        case 2: { /* roughdielectric.cpp:515-614 */
            const float m_eta = d.eta, m_invEta = 1 / d.eta;
            bool sampleReflection = true;
            Microfacet ds = distr(d);
            Microfacet sd = ds;
            if (!ds.visible) sd.scaleAlpha(1.2f - 0.2f * std::sqrt(std::abs(Frame::cosTheta(r.wi))));
            float microfacetPDF;
            const V3 m = sd.sample(signum(Frame::cosTheta(r.wi)) * r.wi, sx, sy, microfacetPDF);
            if (microfacetPDF == 0) return Spectrum(0.0f);
            pdfOut = microfacetPDF;
            float cosThetaT;
            float F = fresnelDielectricExt(dot(r.wi, m), cosThetaT, m_eta);
            Spectrum weight(1.0f);
            if (r.sampler->next1D() > F) { sampleReflection = false; pdfOut *= 1 - F; }
            else pdfOut *= F;
            float dwh_dwo;
            if (sampleReflection) {
                r.wo = reflect(r.wi, m);
                r.eta = 1.0f; r.sampledType = EGlossyReflection;
                if (Frame::cosTheta(r.wi) * Frame::cosTheta(r.wo) <= 0) return Spectrum(0.0f);
                weight *= R;
                dwh_dwo = 1.0f / (4.0f * dot(r.wo, m));
            } else {
                if (cosThetaT == 0) return Spectrum(0.0f);
                r.wo = refract(r.wi, m, m_eta, cosThetaT);
                r.eta = cosThetaT < 0 ? m_eta : m_invEta;
                r.sampledType = EGlossyTransmission;
                if (Frame::cosTheta(r.wi) * Frame::cosTheta(r.wo) >= 0) return Spectrum(0.0f);
                float factor = cosThetaT < 0 ? m_invEta : m_eta;
                weight *= V3(d.transmittance[0], d.transmittance[1], d.transmittance[2]) * (factor * factor);
                float sqrtDenom = dot(r.wi, m) + r.eta * dot(r.wo, m);
                dwh_dwo = (r.eta * r.eta * dot(r.wo, m)) / (sqrtDenom * sqrtDenom);
            }
            if (ds.visible) weight *= ds.smithG1(r.wo, m);
            else weight *= std::abs(ds.eval(m) * ds.G(r.wi, r.wo, m) * dot(r.wi, m) / (microfacetPDF * Frame::cosTheta(r.wi)));
            pdfOut *= std::abs(dwh_dwo);
            return weight;
        }
        case 4: { /* null.cpp:65-76 */
            r.wo = -r.wi; r.sampledType = ENull; r.eta = 1.0f; pdfOut = 1;
            return Spectrum(1.0f);
        }
        case 5: { /* twosided.cpp:162-184 */
            bool flipped = false;
            if (Frame::cosTheta(r.wi) < 0) { r.wi.z *= -1; flipped = true; }
            Spectrum result = sample(flipped ? d.nested2 : d.nested, r, pdfOut, sx, sy);
            if (flipped) {
                r.wi.z *= -1;
                if (!result.isZero() && pdfOut != 0) r.wo.z *= -1;
            }
            return result;
        }
        case 6: { /* dielectric.cpp:281-310 (both components enabled) */
            float cosThetaT;
            float F = fresnelDielectricExt(Frame::cosTheta(r.wi), cosThetaT, d.eta);
            const float invEta = 1 / d.eta;
            if (sx <= F) {
                r.sampledType = EDeltaReflection;
                r.wo = V3(-r.wi.x, -r.wi.y, r.wi.z);
                r.eta = 1.0f;
                pdfOut = F;
                return R;
            } else {
                r.sampledType = EDeltaTransmission;
                float scale = -(cosThetaT < 0 ? invEta : d.eta);
                r.wo = V3(scale * r.wi.x, scale * r.wi.y, cosThetaT);
                r.eta = cosThetaT < 0 ? d.eta : invEta;
                pdfOut = 1 - F;
                float factor = cosThetaT < 0 ? invEta : d.eta;
                return V3(d.transmittance[0], d.transmittance[1], d.transmittance[2]) * (factor * factor);
            }
        }
        case 7: { /* conductor.cpp:268-283 */
            if (Frame::cosTheta(r.wi) <= 0) return Spectrum(0.0f);
            r.sampledType = EDeltaReflection;
            r.wo = V3(-r.wi.x, -r.wi.y, r.wi.z);
            r.eta = 1.0f;
            pdfOut = 1;
            return R * fresnelConductorExact(Frame::cosTheta(r.wi), V3(d.etaC[0], d.etaC[1], d.etaC[2]), V3(d.kC[0], d.kC[1], d.kC[2]));
        }
        case 8: { /* plastic.cpp:377-424 (both components enabled) */
            if (Frame::cosTheta(r.wi) <= 0) return Spectrum(0.0f);
            float Fi = fresnelDielectricExt(Frame::cosTheta(r.wi), d.eta);
            r.eta = 1.0f;
            float w = d.specSamplingWeight;
            float probSpecular = (Fi * w) / (Fi * w + (1 - Fi) * (1 - w));
            if (sx < probSpecular) {
                r.sampledType = EDeltaReflection;
                r.wo = V3(-r.wi.x, -r.wi.y, r.wi.z);
                pdfOut = probSpecular;
                return R * Fi / probSpecular;
            } else {
                r.sampledType = EDiffuseReflection;
                r.wo = squareToCosineHemisphere((sx - probSpecular) / (1 - probSpecular), sy);
                float Fo = fresnelDielectricExt(Frame::cosTheta(r.wo), d.eta);
                Spectrum diff = diffuseReflectance(d, r);
                if (d.nonlinear) diff = diff / (Spectrum(1.0f) - diff * d.fdrInt);
                else diff /= 1 - d.fdrInt;
                pdfOut = (1 - probSpecular) * squareToCosineHemispherePdf(r.wo);
                const float invEta2 = 1 / (d.eta * d.eta);
                return diff * (invEta2 * (1 - Fi) * (1 - Fo) / (1 - probSpecular));
            }
        }
        default: { /* coating.cpp:288-371 */
            const float m_invEta = 1 / d.eta;
            bool sampleNested = (type(d.nested) & EAll) != 0;
            float R12;
            V3 wiPrime = refractIn(d, r.wi, R12);
            float w = specularSamplingWeight(d);
            float probSpecular = (R12 * w) / (R12 * w + (1 - R12) * (1 - w));
            bool choseSpecular = true;
            if (sampleNested) {
                if (sx < probSpecular) sx /= probSpecular;
                else { sx = (sx - probSpecular) / (1 - probSpecular); choseSpecular = false; }
            }
            if (choseSpecular) {
                r.sampledType = EDeltaReflection;
                r.wo = V3(-r.wi.x, -r.wi.y, r.wi.z);
                r.eta = 1.0f;
                pdfOut = sampleNested ? probSpecular : 1.0f;
                return R * (R12 / pdfOut);
            } else {
                if (R12 == 1.0f) return Spectrum(0.0f);
                V3 wiBackup = r.wi;
                r.wi = wiPrime;
                Spectrum result = sample(d.nested, r, pdfOut, sx, sy);
                r.wi = wiBackup;
                if (result.isZero()) return Spectrum(0.0f);
                V3 woPrime = r.wo;
                Spectrum sigmaA = V3(d.sigmaA[0], d.sigmaA[1], d.sigmaA[2]) * d.thickness;
                if (!sigmaA.isZero())
                    result *= expSpec(-sigmaA * (1 / std::abs(Frame::cosTheta(wiPrime)) + 1 / std::abs(Frame::cosTheta(woPrime))));
                float R21;
                r.wo = refractOut(d, woPrime, R21);
                if (R21 == 1.0f) return Spectrum(0.0f);
                pdfOut *= 1.0f - probSpecular;
                result /= 1.0f - probSpecular;
                result *= (1 - R12) * (1 - R21);
                if (!(r.sampledType & EDelta)) /* BSDF::getMeasure(sampledType) == ESolidAngle */
                    pdfOut *= m_invEta * m_invEta * Frame::cosTheta(r.wo) / Frame::cosTheta(woPrime);
                return result;
            }
        }
        }
    }
};

} // namespace orc
