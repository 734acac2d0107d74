// Device math for the wavefront path tracer: vectors, frames, warps, Fresnel terms.
// Each routine states the Mitsuba 0.6 source it has to agree with (file:line under the reference
// tree); operation ORDER is kept (scalar division = multiply by reciprocal, vector.h operator/),
// because the parity build (-fmad=false) is compared against an IEEE-strict CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define B2_DEV __device__ __forceinline__
#define B2_HD __host__ __device__ __forceinline__

namespace b2 {

// include/mitsuba/core/constants.h:28-31,51-87
#define B2_EPSILON 1e-4f
#define B2_SHADOW_EPSILON 1e-3f
#define B2_DELTA_EPSILON 1e-3f
#define B2_PI 3.14159265358979323846f
#define B2_INV_PI 0.31830988618379067154f
#define B2_INV_TWOPI 0.15915494309189533577f
#define B2_ONE_MINUS_EPS 0x1.fffffep-1f
#define B2_RCPOVERFLOW 0x1p-128f
#define B2_INF __int_as_float(0x7f800000)

struct V3 {
    float x, y, z;
    B2_HD V3() {}
    B2_HD V3(float a) : x(a), y(a), z(a) {}
    B2_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
};
B2_HD V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
B2_HD V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
B2_HD V3 operator-(const V3 &a) { return V3(-a.x, -a.y, -a.z); }
B2_HD V3 operator*(const V3 &a, float f) { return V3(a.x * f, a.y * f, a.z * f); }
B2_HD V3 operator*(float f, const V3 &a) { return V3(a.x * f, a.y * f, a.z * f); }
B2_HD V3 operator*(const V3 &a, const V3 &b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
B2_HD V3 operator/(const V3 &a, const V3 &b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
B2_HD V3 operator/(const V3 &a, float f) { float r = 1.0f / f; return V3(a.x * r, a.y * r, a.z * r); }
B2_HD float dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
B2_HD float absDot(const V3 &a, const V3 &b) { return fabsf(dot(a, b)); }
B2_HD V3 cross(const V3 &a, const V3 &b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
B2_HD float lengthSquared(const V3 &a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
B2_HD float length(const V3 &a) { return sqrtf(lengthSquared(a)); }
B2_HD V3 normalize(const V3 &a) { return a / length(a); }
B2_HD bool isZero(const V3 &a) { return a.x == 0 && a.y == 0 && a.z == 0; }
B2_HD float maxComp(const V3 &a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
B2_HD float average(const V3 &a) { return (a.x + a.y + a.z) * (1.0f / 3.0f); }
B2_HD float comp(const V3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
typedef V3 Spectrum;

// include/mitsuba/core/math.h:185-237: on Linux/x86_64 the reference evaluates exp/log in double and
// rounds; expf/logf differ from that by <= 1 ulp, inside the stated tolerance.
// math.h:174-200: on Linux/x86-64 the reference evaluates exp/log in double and rounds once; the parity build does the same,
// the throughput build uses the single-precision intrinsics
#ifdef B2_FAST_TRI
B2_DEV float fastexp(float v) { return expf(v); }
B2_DEV float fastlog(float v) { return logf(v); }
// log(1 - u) of the free-path sampling: under --use_fast_math logf is __logf, whose absolute error (2^-21) is a relative error
// of 1e-4..1e-3 for the small steps u << 1; log1pf keeps full relative accuracy
B2_DEV float logOneMinus(float u) { return log1pf(-u); }
#else
B2_DEV float fastexp(float v) { return (float) exp((double) v); }
B2_DEV float fastlog(float v) { return (float) log((double) v); }
B2_DEV float logOneMinus(float u) { return fastlog(1 - u); }
#endif
B2_DEV float safe_sqrt(float v) { return sqrtf(fmaxf(0.0f, v)); }
B2_DEV float signum(float v) { return copysignf(1.0f, v); }
B2_DEV V3 expSpec(const V3 &v) { return V3(fastexp(v.x), fastexp(v.y), fastexp(v.z)); }
B2_DEV V3 safeSqrtSpec(const V3 &v) { return V3(safe_sqrt(v.x), safe_sqrt(v.y), safe_sqrt(v.z)); }

// src/libcore/math.cpp:25-53
B2_DEV float erfinv_giles(float x) {
    float w = -fastlog((1.0f - x) * (1.0f + x));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = 3.43273939e-07f + p * w;
        p = -3.5233877e-06f + p * w;
        p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w;
        p = -0.00125372503f + p * w;
        p = -0.00417768164f + p * w;
        p = 0.246640727f + p * w;
        p = 1.50140941f + p * w;
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = 0.000100950558f + p * w;
        p = 0.00134934322f + p * w;
        p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w;
        p = -0.0076224613f + p * w;
        p = 0.00943887047f + p * w;
        p = 1.00167406f + p * w;
        p = 2.83297682f + p * w;
    }
    return p * x;
}
// src/libcore/math.cpp:55-72
B2_DEV float erf_as(float x) {
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f,
                p = 0.3275911f;
    float sign = signum(x);
    x = fabsf(x);
    float t = 1.0f / (1.0f + p * x);
    float y = 1.0f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * fastexp(-x * x);
    return sign * y;
}
// src/libcore/math.cpp:74-86
B2_DEV float hypot2(float a, float b) {
    float r;
    if (fabsf(a) > fabsf(b)) {
        r = b / a;
        r = fabsf(a) * sqrtf(1.0f + r * r);
    } else if (b != 0.0f) {
        r = a / b;
        r = fabsf(b) * sqrtf(1.0f + r * r);
    } else {
        r = 0.0f;
    }
    return r;
}

// src/libcore/util.cpp:592-601
B2_DEV void coordinateSystem(const V3 &a, V3 &b, V3 &c) {
    if (fabsf(a.x) > fabsf(a.y)) {
        float invLen = 1.0f / sqrtf(a.x * a.x + a.z * a.z);
        c = V3(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(a.y * a.y + a.z * a.z);
        c = V3(0.0f, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}

// include/mitsuba/core/frame.h:37-130
struct Frame {
    V3 s, t, n;
    B2_DEV V3 toLocal(const V3 &v) const { return V3(dot(v, s), dot(v, t), dot(v, n)); }
    B2_DEV V3 toWorld(const V3 &v) const { return s * v.x + t * v.y + n * v.z; }
};
B2_DEV float cosTheta(const V3 &v) { return v.z; }
B2_DEV float cosTheta2(const V3 &v) { return v.z * v.z; }
B2_DEV float sinTheta2(const V3 &v) { return 1.0f - v.z * v.z; }
B2_DEV float tanTheta(const V3 &v) {
    float temp = 1 - v.z * v.z;
    if (temp <= 0.0f) return 0.0f;
    return sqrtf(temp) / v.z;
}
// src/libcore/util.cpp:603-608
B2_DEV void computeShadingFrame(const V3 &n, const V3 &dpdu, Frame &frame) {
    frame.n = n;
    frame.s = normalize(dpdu - frame.n * dot(frame.n, dpdu));
    frame.t = cross(frame.n, frame.s);
}

// src/libcore/warp.cpp:81-103
B2_DEV void squareToUniformDiskConcentric(float sx, float sy, float &px, float &py) {
    float r1 = 2.0f * sx - 1.0f;
    float r2 = 2.0f * sy - 1.0f;
    float phi, r;
    if (r1 == 0 && r2 == 0) {
        r = phi = 0;
    } else if (r1 * r1 > r2 * r2) {
        r = r1;
        phi = (B2_PI / 4.0f) * (r2 / r1);
    } else {
        r = r2;
        phi = (B2_PI / 2.0f) - (r1 / r2) * (B2_PI / 4.0f);
    }
    float cosPhi, sinPhi;
    sincosf(phi, &sinPhi, &cosPhi);
    px = r * cosPhi;
    py = r * sinPhi;
}
// src/libcore/warp.cpp:43-52
B2_DEV V3 squareToCosineHemisphere(float sx, float sy) {
    float px, py;
    squareToUniformDiskConcentric(sx, sy, px, py);
    float z = safe_sqrt(1.0f - px * px - py * py);
    if (z == 0) z = 1e-10f;
    return V3(px, py, z);
}
B2_DEV float squareToCosineHemispherePdf(const V3 &d) { return B2_INV_PI * cosTheta(d); }
// src/libcore/warp.cpp:76-79
B2_DEV void squareToUniformTriangle(float sx, float sy, float &bx, float &by) {
    float a = safe_sqrt(1.0f - sx);
    bx = 1 - a;
    by = a * sy;
}

// src/libcore/util.cpp:651-681
B2_DEV float fresnelDielectricExt(float cosThetaI_, float &cosThetaT_, float eta) {
    if (eta == 1) {
        cosThetaT_ = -cosThetaI_;
        return 0.0f;
    }
    float scale = (cosThetaI_ > 0) ? 1 / eta : eta, cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) {
        cosThetaT_ = 0.0f;
        return 1.0f;
    }
    float cosThetaI = fabsf(cosThetaI_);
    float cosThetaT = sqrtf(cosThetaTSqr);
    float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
B2_DEV float fresnelDielectricExt(float cosThetaI, float eta) {
    float c;
    return fresnelDielectricExt(cosThetaI, c, eta);
}
// src/libcore/util.cpp:739-761
B2_DEV Spectrum fresnelConductorExact(float cosThetaI, const Spectrum &eta, const Spectrum &k) {
    float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    Spectrum temp1 = eta * eta - k * k - Spectrum(sinThetaI2), a2pb2 = safeSqrtSpec(temp1 * temp1 + k * k * eta * eta * 4.0f),
             a = safeSqrtSpec((a2pb2 + temp1) * 0.5f);
    Spectrum term1 = a2pb2 + Spectrum(cosThetaI2), term2 = a * (2 * cosThetaI);
    Spectrum Rs2 = (term1 - term2) / (term1 + term2);
    Spectrum term3 = a2pb2 * cosThetaI2 + Spectrum(sinThetaI4), term4 = term2 * sinThetaI2;
    Spectrum Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5f * (Rp2 + Rs2);
}
// src/libcore/util.cpp:763-772
B2_DEV V3 reflect(const V3 &wi, const V3 &n) { return (2 * dot(wi, n)) * n - wi; }
B2_DEV V3 refract(const V3 &wi, const V3 &n, float eta, float cosThetaT) {
    if (cosThetaT < 0) eta = 1 / eta;
    return n * (dot(wi, n) * eta + cosThetaT) - wi * eta;
}

} // namespace b2
