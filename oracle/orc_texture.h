/* TEST INFRASTRUCTURE ONLY -- CPU restatement of Mitsuba 0.6's `bitmap` texture.  Nothing under
 * mitsuba_b200/ may include or link this file.
 *
 * Follows, in single precision (Float = float):
 *   MIP pyramid construction          include/mitsuba/render/mipmap.h:155-303 (TMIPMap constructor)
 *   separable resampling              include/mitsuba/core/rfilter.h:107-190 (Resampler ctor), :209-330 (resampleAndClamp, lookup)
 *                                     src/libcore/bitmap.cpp:2230-2329 (x pass, then y pass, each clamped to [0, maxValue])
 *   2-lobe Lanczos filter             src/rfilters/lanczos.cpp:43-56 (chosen at src/textures/bitmap.cpp:281-287)
 *   texel / box / bilinear / EWA      include/mitsuba/render/mipmap.h:499-571, :586-608, :638-721, :767-838
 *   uv transform + filter choice      src/librender/texture.cpp:124-133 (Texture2D::eval), src/textures/bitmap.cpp:400-421,:452-465
 *   energy conservation of the BSDF   src/librender/bsdf.cpp:88-111 (scale texture 0.99 / max)
 * Pinning: both halves are checked bit for bit against the reference's own code, compiled from /root/reference into oracle/_ref
 * behind stand-in headers for the rest of libcore (oracle/Makefile):
 *   librfilterref.so  Resampler<float> (rfilter.h) + LanczosSincFilter (src/rfilters/lanczos.cpp)  -> tests/golden/resample_ref.npz
 *   libmipmapref.so   TMIPMap<Color3, Color3>::eval / evalBilinear / evalBox / evalEWA (mipmap.h with the real barray.h, spectrum.h,
 *                     math.h, src/libcore/math.cpp)                                                 -> tests/golden/mipmap_ref.npz
 * (tests/gen_golden.py writes the fixtures, tests/test_oracle_texture.py compares).  Texture2D's uv transform, the partials of
 * Intersection::computePartials and the sensors' ray differentials sit in files that need all of librender; they are pinned by
 * closed forms (finite differences of neighbouring pixels, uv scale / offset identities).
 * Image file decoding (PNG/JPEG/EXR) is not part of the path: pixels arrive as linear float, 1 or 3 channels, row-major, top row
 * first (the layout Bitmap::convert(..., EFloat, gamma 1) hands to the MIP map).
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "orc_math.h"

extern "C" {
typedef struct OrcTextureDesc {
    int32_t width, height, channels; /* channels: 1 (luminance) or 3 (RGB) */
    int32_t filterType;              /* 0 nearest, 1 bilinear, 2 trilinear, 3 ewa        (bitmap.cpp:213-230) */
    int32_t wrapU, wrapV;            /* 0 repeat, 1 clamp, 2 mirror, 3 zero, 4 one        (bitmap.cpp:324-338) */
    float maxAnisotropy;             /* bitmap.cpp:232-235 (forced to 1 unless ewa) */
    float uoffset, voffset, uscale, vscale; /* texture.cpp:82-95 */
} OrcTextureDesc;
}

namespace orc {

enum { TexNearest = 0, TexBilinear = 1, TexTrilinear = 2, TexEWA = 3 };
enum { WrapRepeat = 0, WrapClamp = 1, WrapMirror = 2, WrapZero = 3, WrapOne = 4 };
static const int kMipLutSize = 64; /* mipmap.h:37 */

inline int floorToInt(float v) { return (int) std::floor(v); }
inline int ceilToInt(float v) { return (int) std::ceil(v); }
inline int modulo(int a, int b) { int r = a % b; return r < 0 ? r + b : r; } /* math.h modulo */

/* src/rfilters/lanczos.cpp:43-56 with lobes = 2 */
inline float lanczos2(float x) {
    x = std::abs(x);
    if (x < 1e-4f) return 1.0f; /* Epsilon, constants.h single precision */
    else if (x > 2.0f) return 0.0f;
    float x1 = (float) (M_PI * x);
    float x2 = x1 / 2.0f;
    return (std::sin(x1) * std::sin(x2)) / (x1 * x2);
}

/* rfilter.h:107-190 (resampling mode only; the pyramid never filters at equal size) + :209-330 */
struct Resampler {
    int bc, sourceRes, targetRes, taps;
    std::vector<int> start;
    std::vector<float> weights;
    Resampler(int bc_, int src, int trg) : bc(bc_), sourceRes(src), targetRes(trg) {
        float filterRadius = 2.0f, scale = 1.0f, invScale = 1.0f;
        if (trg < src) { scale = (float) src / (float) trg; invScale = 1 / scale; filterRadius *= scale; }
        taps = ceilToInt(filterRadius * 2);
        start.resize(trg); weights.resize((size_t) taps * trg);
        for (int i = 0; i < trg; i++) {
            float center = (i + 0.5f) / trg * src;
            start[i] = floorToInt(center - filterRadius + 0.5f);
            float sum = 0;
            for (int j = 0; j < taps; j++) {
                float pos = start[i] + j + 0.5f - center;
                float weight = lanczos2(pos * invScale);
                weights[(size_t) i * taps + j] = weight;
                sum += weight;
            }
            float normalization = 1.0f / sum;
            for (int j = 0; j < taps; j++) weights[(size_t) i * taps + j] *= normalization;
        }
    }
    float lookup(const float *source, int pos, size_t stride, int offset) const { /* rfilter.h lookup() */
        if (pos < 0 || pos >= sourceRes) {
            switch (bc) {
                case WrapClamp: pos = std::min(std::max(pos, 0), sourceRes - 1); break;
                case WrapRepeat: pos = modulo(pos, sourceRes); break;
                case WrapMirror: pos = modulo(pos, 2 * sourceRes); if (pos >= sourceRes) pos = 2 * sourceRes - pos - 1; break;
                case WrapZero: return 0.0f;
                case WrapOne: return 1.0f;
            }
        }
        return source[stride * pos + offset];
    }
    void resampleAndClamp(const float *source, size_t sourceStride, float *target, size_t targetStride, int channels, float mn, float mx) const {
        for (int i = 0; i < targetRes; ++i)
            for (int ch = 0; ch < channels; ++ch) {
                float result = 0;
                for (int j = 0; j < taps; ++j) result += lookup(source, start[i] + j, sourceStride * channels, ch) * weights[(size_t) i * taps + j];
                target[(size_t) i * targetStride * channels + ch] = std::min(mx, std::max(mn, result));
            }
    }
};

/* float -> half -> float, round to nearest even, overflow to infinity (half.h:431-487 fast path, half.cpp:78-200 denormals / overflow):
 * BitmapTexture stores its pyramid as TMIPMap<Color3, Color3h> (bitmap.cpp:177-180); every level is resampled in float from the previous
 * *float* level and only rounded when it is stored (mipmap.h:226-230, :262-264, barray.h:88-92).  Pinned against the reference's half
 * class in oracle/_ref/libcoreref.so (tests/golden/half_ref.npz). */
static inline float roundToHalf(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    const uint32_t sign = x & 0x80000000u;
    uint32_t a = x & 0x7fffffffu;
    if (a >= 0x7f800000u) return f;                              /* infinities, NaNs */
    if (a < 0x38800000u) {                                       /* below 2^-14: half denormals, spacing 2^-24 = the float spacing in [0.5, 1) */
        float m; memcpy(&m, &a, 4);
        volatile float t = m + 0.5f;                             /* rounds to nearest even at 2^-24 */
        m = t - 0.5f;
        memcpy(&a, &m, 4);
    } else {
        a += 0x00000fffu + ((a >> 13) & 1u);                     /* round the significand to 10 bits, carry into the exponent */
        a &= 0xffffe000u;
        if (a >= 0x47800000u) a = 0x7f800000u;                   /* 65520 and above: infinity */
    }
    a |= sign;
    float r; memcpy(&r, &a, 4);
    return r;
}

struct Texture {
    OrcTextureDesc d;
    int levels = 0;
    std::vector<std::vector<float>> pyramid; /* [level][(y * w + x) * channels + c] */
    std::vector<int> lw, lh;
    std::vector<float> ratioX, ratioY;
    float weightLut[kMipLutSize];
    float maximum = 0;   /* component-wise maximum of level 0 (mipmap.h:229-241) */
    float bsdfScale = 1; /* bsdf.cpp:88-111 */

    /* maxValue: the upper clamp of the resampling passes (mipmap.h:156-157, :262: 1 for `bitmap` textures, infinity for `envmap`) */
    void build(const OrcTextureDesc &desc, const float *pixels, float maxValue = 1.0f) {
        d = desc;
        if (d.filterType != TexEWA) d.maxAnisotropy = 1.0f;
        const int ch = d.channels;
        std::vector<float> cur(pixels, pixels + (size_t) d.width * d.height * ch);
        for (float &v : cur) v = std::max(v, 0.0f); /* clampNegative, mipmap.h:231-239 */
        maximum = 0;
        for (float v : cur) maximum = std::max(maximum, v);
        bsdfScale = maximum > 1.0f ? 0.99f * (1.0f / maximum) : 1.0f;
        int w = d.width, h = d.height;
        pyramid.clear(); lw.clear(); lh.clear(); ratioX.clear(); ratioY.clear();
        pyramid.push_back(cur); lw.push_back(w); lh.push_back(h); ratioX.push_back(1); ratioY.push_back(1);
        if (d.filterType != TexNearest && d.filterType != TexBilinear) {
            while (w > 1 || h > 1) {
                const int nw = std::max(1, (w + 1) / 2), nh = std::max(1, (h + 1) / 2);
                std::vector<float> next((size_t) nw * nh * ch);
                const std::vector<float> *src = &pyramid.back();
                std::vector<float> temp;
                if (w != nw) { /* x pass, bitmap.cpp:2258-2293 */
                    Resampler r(d.wrapU, w, nw);
                    std::vector<float> &dst = (h == nh) ? next : temp;
                    if (h != nh) temp.resize((size_t) nw * h * ch);
                    for (int y = 0; y < h; ++y) r.resampleAndClamp(src->data() + (size_t) y * w * ch, 1, dst.data() + (size_t) y * nw * ch, 1, ch, 0.0f, maxValue);
                    src = &dst;
                }
                if (h != nh) { /* y pass, :2296-2327 */
                    Resampler r(d.wrapV, h, nh);
                    for (int x = 0; x < nw; ++x) r.resampleAndClamp(src->data() + (size_t) x * ch, nw, next.data() + (size_t) x * ch, nw, ch, 0.0f, maxValue);
                }
                w = nw; h = nh;
                pyramid.push_back(next); lw.push_back(w); lh.push_back(h);
                ratioX.push_back((float) w / (float) d.width); ratioY.push_back((float) h / (float) d.height);
            }
        }
        levels = (int) pyramid.size();
        for (auto &lvl : pyramid) for (float &v : lvl) v = roundToHalf(v); /* stored as half (see roundToHalf); the chain above ran on floats */
        for (int i = 0; i < kMipLutSize; ++i) { /* mipmap.h:296-302 */
            float r2 = (float) i / (float) (kMipLutSize - 1);
            weightLut[i] = fastexp(-2.0f * r2) - fastexp(-2.0f);
        }
    }

    V3 texel(int level, int x, int y) const { /* evalTexel, mipmap.h:499-558 */
        const int sx = lw[level], sy = lh[level];
        if (x < 0 || x >= sx) {
            switch (d.wrapU) {
                case WrapRepeat: x = modulo(x, sx); break;
                case WrapClamp: x = std::min(std::max(x, 0), sx - 1); break;
                case WrapMirror: x = modulo(x, 2 * sx); if (x >= sx) x = 2 * sx - x - 1; break;
                case WrapZero: return V3(0.0f);
                case WrapOne: return V3(1.0f);
            }
        }
        if (y < 0 || y >= sy) {
            switch (d.wrapV) {
                case WrapRepeat: y = modulo(y, sy); break;
                case WrapClamp: y = std::min(std::max(y, 0), sy - 1); break;
                case WrapMirror: y = modulo(y, 2 * sy); if (y >= sy) y = 2 * sy - y - 1; break;
                case WrapZero: return V3(0.0f);
                case WrapOne: return V3(1.0f);
            }
        }
        const float *p = pyramid[level].data() + ((size_t) y * sx + x) * d.channels;
        return d.channels == 3 ? V3(p[0], p[1], p[2]) : V3(p[0]);
    }
    V3 evalBox(int level, float u, float v) const { /* :561-564 */
        return texel(level, floorToInt(u * lw[level]), floorToInt(v * lh[level]));
    }
    V3 evalBilinear(int level, float uu, float vv) const { /* :570-591 */
        if (!std::isfinite(uu) || !std::isfinite(vv)) return V3(0.0f);
        if (level >= levels) return evalBox(levels - 1, uu, vv);
        const float u = uu * lw[level] - 0.5f, v = vv * lh[level] - 0.5f;
        const int xPos = floorToInt(u), yPos = floorToInt(v);
        const float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
        return texel(level, xPos, yPos) * dx2 * dy2 + texel(level, xPos, yPos + 1) * dx2 * dy1 + texel(level, xPos + 1, yPos) * dx1 * dy2 +
               texel(level, xPos + 1, yPos + 1) * dx1 * dy1;
    }
    V3 evalEWA(int level, float uu, float vv, float A, float B, float C) const { /* :767-838 */
        if (!std::isfinite(A + B + C + uu + vv)) return V3(0.0f);
        if (level >= levels) return evalBox(levels - 1, uu, vv);
        const float u = uu * lw[level] - 0.5f, v = vv * lh[level] - 0.5f;
        A /= ratioX[level] * ratioX[level];
        B /= ratioX[level] * ratioY[level];
        C /= ratioY[level] * ratioY[level];
        const float invDet = 1.0f / (-B * B + 4.0f * A * C), deltaU = 2.0f * std::sqrt(C * invDet), deltaV = 2.0f * std::sqrt(A * invDet);
        const int u0 = ceilToInt(u - deltaU), u1 = floorToInt(u + deltaU), v0 = ceilToInt(v - deltaV), v1 = floorToInt(v + deltaV);
        const float As = A * kMipLutSize, Bs = B * kMipLutSize, Cs = C * kMipLutSize;
        V3 result(0.0f);
        float denominator = 0.0f;
        const float ddq = 2 * As, uu0 = (float) u0 - u;
        for (int vt = v0; vt <= v1; ++vt) {
            const float vvv = (float) vt - v;
            float q = As * uu0 * uu0 + (Bs * uu0 + Cs * vvv) * vvv;
            float dq = As * (2 * uu0 + 1) + Bs * vvv;
            for (int ut = u0; ut <= u1; ++ut) {
                if (q < (float) kMipLutSize) {
                    const uint32_t qi = (uint32_t) q;
                    if (qi < (uint32_t) kMipLutSize) {
                        const float weight = weightLut[(int) q];
                        result += texel(level, ut, vt) * weight;
                        denominator += weight;
                    }
                }
                q += dq;
                dq += ddq;
            }
        }
        if (denominator == 0) return evalBilinear(level, uu, vv);
        return result / denominator;
    }
    /* TMIPMap::eval(uv, d0, d1), mipmap.h:638-721 */
    V3 evalFiltered(float u, float v, float d0x, float d0y, float d1x, float d1y) const {
        if (d.filterType == TexNearest) return evalBox(0, u, v);
        if (d.filterType == TexBilinear) return evalBilinear(0, u, v);
        const float du0 = d0x * lw[0], dv0 = d0y * lh[0], du1 = d1x * lw[0], dv1 = d1y * lh[0];
        float A = dv0 * dv0 + dv1 * dv1, B = -2.0f * (du0 * dv0 + du1 * dv1), C = du0 * du0 + du1 * du1, F = A * C - B * B * 0.25f;
        const float root = hypot2(A - C, B), Aprime = 0.5f * (A + C - root), Cprime = 0.5f * (A + C + root);
        float majorRadius = Aprime != 0 ? std::sqrt(F / Aprime) : 0, minorRadius = Cprime != 0 ? std::sqrt(F / Cprime) : 0;
        if (d.filterType == TexTrilinear || !(minorRadius > 0) || !(majorRadius > 0) || F < 0) {
            const float level = log2f_mts(std::max(majorRadius, 1e-4f));
            const int ilevel = floorToInt(level);
            if (ilevel < 0) return evalBilinear(0, u, v);
            const float a = level - ilevel;
            return evalBilinear(ilevel, u, v) * (1.0f - a) + evalBilinear(ilevel + 1, u, v) * a;
        }
        if (minorRadius * d.maxAnisotropy < majorRadius) {
            minorRadius = majorRadius / d.maxAnisotropy;
            const float theta = 0.5f * std::atan(B / (A - C));
            const float sinTheta = std::sin(theta), cosTheta = std::cos(theta);
            const float a2 = majorRadius * majorRadius, b2 = minorRadius * minorRadius, sinTheta2 = sinTheta * sinTheta, cosTheta2 = cosTheta * cosTheta,
                        sin2Theta = 2 * sinTheta * cosTheta;
            A = a2 * cosTheta2 + b2 * sinTheta2;
            B = (a2 - b2) * sin2Theta;
            C = a2 * sinTheta2 + b2 * cosTheta2;
            F = a2 * b2;
        }
        const float scale = 1.0f / F;
        A *= scale; B *= scale; C *= scale;
        const float level = std::max(0.0f, log2f_mts(minorRadius));
        const int ilevel = (int) level;
        const float a = level - ilevel;
        if (majorRadius < 1 || !(A > 0 && C > 0)) return evalBilinear(ilevel, u, v);
        return evalEWA(ilevel, u, v, A, B, C) * (1.0f - a) + evalEWA(ilevel + 1, u, v, A, B, C) * a;
    }
    /* BitmapTexture::eval(uv), bitmap.cpp:400-421 */
    V3 evalUnfiltered(float u, float v) const { return d.filterType != TexNearest ? evalBilinear(0, u, v) : evalBox(0, u, v); }

    /* Texture2D::eval(its, filter = true), texture.cpp:124-133, times the BSDF's energy-conservation scale */
    V3 eval(float itsU, float itsV, bool hasUVPartials, float dudx, float dudy, float dvdx, float dvdy) const {
        const float u = itsU * d.uscale + d.uoffset, v = itsV * d.vscale + d.voffset;
        V3 r = hasUVPartials ? evalFiltered(u, v, dudx * d.uscale, dvdx * d.vscale, dudy * d.uscale, dvdy * d.vscale) : evalUnfiltered(u, v);
        return r * bsdfScale;
    }

    static float hypot2(float a, float b) { /* src/libcore/math.cpp:74-86 */
        float r;
        if (std::abs(a) > std::abs(b)) { r = b / a; r = std::abs(a) * std::sqrt(1.0f + r * r); }
        else if (b != 0.0f) { r = a / b; r = std::abs(b) * std::sqrt(1.0f + r * r); }
        else r = 0.0f;
        return r;
    }
    static float log2f_mts(float value) { /* math.cpp:103-106 */
        const float invLn2 = 1.0f / std::log(2.0f);
        return fastlog(value) * invLn2;
    }
};

} // namespace orc
