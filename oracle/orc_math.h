/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under mitsuba_b200/
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it (as the checker / CPU baseline).
 *
 * Scalar float32 restatement of the Mitsuba 0.6 math the `path` integrator touches.
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Compiled with -ffp-contract=off and without -ffast-math: it is the IEEE-strict reading of the
 * reference source (the reference build itself uses -funsafe-math-optimizations, so it is not
 * bit-reproducible against itself).
 *
 * PARITY STATUS: "parity unpinned" at image level (the reference holds no golden image and cannot
 * be built here).  Pinned pieces: SFMT19937 (197-word KAT), Sobol' (against the
 * reference's own sobolseq.cpp compiled into oracle/_ref), BSDF/microfacet chi^2 + 3-way protocol. */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>

namespace orc {

/* include/mitsuba/core/constants.h:28-31,51-87 (SINGLE_PRECISION build) */
static const float kEpsilon = 1e-4f;
static const float kShadowEpsilon = 1e-3f;
static const float kDeltaEpsilon = 1e-3f;
static const float kPi = 3.14159265358979323846f;
static const float kInvPi = 0.31830988618379067154f;
static const float kInvTwoPi = 0.15915494309189533577f;
static const float kOneMinusEps = 0x1.fffffep-1f;
static const float kRcpOverflow = 0x1p-128f;
static const float kInf = std::numeric_limits<float>::infinity();

/* include/mitsuba/core/vector.h / point.h / normal.h: scalar division multiplies by the
 * reciprocal (vector.h operator/(Scalar)); normalize(v) = v / v.length() (vector.h:625-627). */
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a) : x(a), y(a), z(a) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return (&x)[i]; }
    float &operator[](int i) { return (&x)[i]; }
    V3 operator+(const V3 &v) const { return V3(x + v.x, y + v.y, z + v.z); }
    V3 operator-(const V3 &v) const { return V3(x - v.x, y - v.y, z - v.z); }
    V3 operator-() const { return V3(-x, -y, -z); }
    V3 operator*(float f) const { return V3(x * f, y * f, z * f); }
    V3 operator*(const V3 &v) const { return V3(x * v.x, y * v.y, z * v.z); }
    V3 operator/(const V3 &v) const { return V3(x / v.x, y / v.y, z / v.z); }
    V3 operator/(float f) const { float r = 1.0f / f; return V3(x * r, y * r, z * r); }
    V3 &operator+=(const V3 &v) { x += v.x; y += v.y; z += v.z; return *this; }
    V3 &operator*=(float f) { x *= f; y *= f; z *= f; return *this; }
    V3 &operator*=(const V3 &v) { x *= v.x; y *= v.y; z *= v.z; return *this; }
    V3 &operator/=(float f) { float r = 1.0f / f; x *= r; y *= r; z *= r; return *this; }
    float lengthSquared() const { return x * x + y * y + z * z; }
    float length() const { return std::sqrt(lengthSquared()); }
    bool isZero() const { return x == 0 && y == 0 && z == 0; }
    float max() const { return std::max(std::max(x, y), z); }      /* Spectrum::max() */
    float average() const { return (x + y + z) * (1.0f / 3.0f); } /* spectrum.h average(): sum * (1/N) */
};
inline V3 operator*(float f, const V3 &v) { return V3(v.x * f, v.y * f, v.z * f); }
inline float dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float absDot(const V3 &a, const V3 &b) { return std::abs(dot(a, b)); }
inline V3 cross(const V3 &a, const V3 &b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline V3 normalize(const V3 &v) { return v / v.length(); }
typedef V3 Spectrum; /* SPECTRUM_SAMPLES=3 RGB build (build/config-linux-gcc.py:7) */
inline V3 expSpec(const V3 &v) { /* TSpectrum::exp(): math::fastexp per channel */
    return V3((float) ::exp((double) v.x), (float) ::exp((double) v.y), (float) ::exp((double) v.z));
}
inline V3 safeSqrtSpec(const V3 &v) {
    return V3(std::sqrt(std::max(0.0f, v.x)), std::sqrt(std::max(0.0f, v.y)), std::sqrt(std::max(0.0f, v.z)));
}

/* include/mitsuba/core/math.h:185-237 (Linux/x86_64 branch: exp/log go through double) */
inline float fastexp(float v) { return (float) ::exp((double) v); }
inline float fastlog(float v) { return (float) ::log((double) v); }
inline float safe_sqrt(float v) { return std::sqrt(std::max(0.0f, v)); }
inline float safe_acos(float v) { return std::acos(std::min(1.0f, std::max(-1.0f, v))); } /* math.h:250-252 */
inline float signum(float v) { return copysignf(1.0f, v); }
inline void sincos(float t, float *s, float *c) { ::sincosf(t, s, c); }

/* src/libcore/math.cpp:25-53 (Giles erfinv) */
inline float erfinv(float x) {
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
        w = std::sqrt(w) - 3.0f;
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
/* src/libcore/math.cpp:55-72 (A&S 7.1.26 erf) */
inline float erf_as(float x) {
    float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f,
          a5 = 1.061405429f, p = 0.3275911f;
    float sign = signum(x);
    x = std::abs(x);
    float t = 1.0f / (1.0f + p * x);
    float y = 1.0f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * fastexp(-x * x);
    return sign * y;
}
/* src/libcore/math.cpp:74-86 */
inline float hypot2(float a, float b) {
    float r;
    if (std::abs(a) > std::abs(b)) {
        r = b / a;
        r = std::abs(a) * std::sqrt(1.0f + r * r);
    } else if (b != 0.0f) {
        r = a / b;
        r = std::abs(b) * std::sqrt(1.0f + r * r);
    } else {
        r = 0.0f;
    }
    return r;
}

/* src/libcore/util.cpp:592-601 */
inline void coordinateSystem(const V3 &a, V3 &b, V3 &c) {
    if (std::abs(a.x) > std::abs(a.y)) {
        float invLen = 1.0f / std::sqrt(a.x * a.x + a.z * a.z);
        c = V3(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / std::sqrt(a.y * a.y + a.z * a.z);
        c = V3(0.0f, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}

/* include/mitsuba/core/frame.h:37-130 */
struct Frame {
    V3 s, t, n;
    Frame() {}
    explicit Frame(const V3 &nn) : n(nn) { coordinateSystem(n, s, t); }
    V3 toLocal(const V3 &v) const { return V3(dot(v, s), dot(v, t), dot(v, n)); }
    V3 toWorld(const V3 &v) const { return s * v.x + t * v.y + n * v.z; }
    static float cosTheta(const V3 &v) { return v.z; }
    static float cosTheta2(const V3 &v) { return v.z * v.z; }
    static float sinTheta2(const V3 &v) { return 1.0f - v.z * v.z; }
    static float tanTheta(const V3 &v) {
        float temp = 1 - v.z * v.z;
        if (temp <= 0.0f) return 0.0f;
        return std::sqrt(temp) / v.z;
    }
};
/* src/libcore/util.cpp:603-608 */
inline void computeShadingFrame(const V3 &n, const V3 &dpdu, Frame &frame) {
    frame.n = n;
    frame.s = normalize(dpdu - frame.n * dot(frame.n, dpdu));
    frame.t = cross(frame.n, frame.s);
}

/* src/libcore/warp.cpp:81-103 */
inline void squareToUniformDiskConcentric(float sx, float sy, float &px, float &py) {
    float r1 = 2.0f * sx - 1.0f;
    float r2 = 2.0f * sy - 1.0f;
    float phi, r;
    if (r1 == 0 && r2 == 0) {
        r = phi = 0;
    } else if (r1 * r1 > r2 * r2) {
        r = r1;
        phi = (kPi / 4.0f) * (r2 / r1);
    } else {
        r = r2;
        phi = (kPi / 2.0f) - (r1 / r2) * (kPi / 4.0f);
    }
    float cosPhi, sinPhi;
    sincos(phi, &sinPhi, &cosPhi);
    px = r * cosPhi;
    py = r * sinPhi;
}
/* src/libcore/warp.cpp:43-52 */
inline V3 squareToCosineHemisphere(float sx, float sy) {
    float px, py;
    squareToUniformDiskConcentric(sx, sy, px, py);
    float z = safe_sqrt(1.0f - px * px - py * py);
    if (z == 0) z = 1e-10f;
    return V3(px, py, z);
}
/* include/mitsuba/core/warp.h squareToCosineHemispherePdf: INV_PI * cosTheta */
inline float squareToCosineHemispherePdf(const V3 &d) { return kInvPi * Frame::cosTheta(d); }
/* src/libcore/warp.cpp:76-79 */
inline void squareToUniformTriangle(float sx, float sy, float &bx, float &by) {
    float a = safe_sqrt(1.0f - sx);
    bx = 1 - a;
    by = a * sy;
}

/* src/libcore/util.cpp:651-681 */
inline float fresnelDielectricExt(float cosThetaI_, float &cosThetaT_, float eta) {
    if (eta == 1) {
        cosThetaT_ = -cosThetaI_;
        return 0.0f;
    }
    float scale = (cosThetaI_ > 0) ? 1 / eta : eta,
          cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) {
        cosThetaT_ = 0.0f;
        return 1.0f;
    }
    float cosThetaI = std::abs(cosThetaI_);
    float cosThetaT = std::sqrt(cosThetaTSqr);
    float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
inline float fresnelDielectricExt(float cosThetaI, float eta) {
    float c;
    return fresnelDielectricExt(cosThetaI, c, eta);
}
/* src/libcore/util.cpp:739-761 (spectral variant) */
inline Spectrum fresnelConductorExact(float cosThetaI, const Spectrum &eta, const Spectrum &k) {
    float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    Spectrum temp1 = eta * eta - k * k - Spectrum(sinThetaI2),
             a2pb2 = safeSqrtSpec(temp1 * temp1 + k * k * eta * eta * 4),
             a = safeSqrtSpec((a2pb2 + temp1) * 0.5f);
    Spectrum term1 = a2pb2 + Spectrum(cosThetaI2), term2 = a * (2 * cosThetaI);
    Spectrum Rs2 = (term1 - term2) / (term1 + term2);
    Spectrum term3 = a2pb2 * cosThetaI2 + Spectrum(sinThetaI4), term4 = term2 * sinThetaI2;
    Spectrum Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5f * (Rp2 + Rs2);
}
/* src/libcore/util.cpp:763-772 */
inline V3 reflect(const V3 &wi, const V3 &n) { return 2 * dot(wi, n) * n - wi; }
inline V3 refract(const V3 &wi, const V3 &n, float eta, float cosThetaT) {
    if (cosThetaT < 0) eta = 1 / eta;
    return n * (dot(wi, n) * eta + cosThetaT) - wi * eta;
}

} // namespace orc
