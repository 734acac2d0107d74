#include "mipmap.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "b2mts.h"

namespace b2host {

namespace {

const double kPi = 3.14159265358979323846;

float lanczos2(float x) { // LanczosSincFilter::eval, lobes = 2 (radius 2)
    x = std::fabs(x);
    if (x < 1e-4f) return 1.0f;
    if (x > 2.0f) return 0.0f;
    const float x1 = (float) (kPi * x), x2 = x1 / 2.0f;
    return (std::sin(x1) * std::sin(x2)) / (x1 * x2);
}

int wrapIndex(int mode, int pos, int res, float &constant, bool &isConstant) {
    isConstant = false;
    if (pos >= 0 && pos < res) return pos;
    auto mod = [](int a, int b) { const int r = a % b; return r < 0 ? r + b : r; };
    switch (mode) {
        case 0: return mod(pos, res);
        case 1: return std::min(std::max(pos, 0), res - 1);
        case 2: { int p = mod(pos, 2 * res); return p >= res ? 2 * res - p - 1 : p; }
        case 3: isConstant = true; constant = 0.0f; return 0;
        default: isConstant = true; constant = 1.0f; return 0;
    }
}

// one axis of Resampler<float> in resampling mode: per target sample the first source tap and `taps` normalised weights
struct AxisKernel {
    int taps = 0;
    std::vector<int> first;
    std::vector<float> weight;
    AxisKernel(int src, int trg) {
        float radius = 2.0f, invScale = 1.0f;
        if (trg < src) {
            const float scale = (float) src / (float) trg;
            invScale = 1 / scale;
            radius *= scale;
        }
        taps = (int) std::ceil(radius * 2);
        first.resize(trg);
        weight.resize((size_t) taps * trg);
        for (int i = 0; i < trg; ++i) {
            const float center = (i + 0.5f) / trg * src;
            first[i] = (int) std::floor(center - radius + 0.5f);
            float *wt = &weight[(size_t) i * taps];
            float sum = 0;
            for (int j = 0; j < taps; ++j) {
                const float pos = first[i] + j + 0.5f - center;
                wt[j] = lanczos2(pos * invScale);
                sum += wt[j];
            }
            const float normalization = 1.0f / sum;
            for (int j = 0; j < taps; ++j) wt[j] *= normalization;
        }
    }
};

// resample `count` lines along one axis: element (line, i, ch) sits at src[line * lineStride + i * elemStride + ch]
void resampleAxis(const AxisKernel &k, int mode, int srcRes, int trgRes, const float *src, float *dst, int count, size_t srcLineStride, size_t srcElemStride,
                  size_t dstLineStride, size_t dstElemStride, int channels, float maxValue) {
    for (int line = 0; line < count; ++line) {
        const float *s = src + line * srcLineStride;
        float *d = dst + line * dstLineStride;
        for (int i = 0; i < trgRes; ++i) {
            const float *wt = &k.weight[(size_t) i * k.taps];
            for (int ch = 0; ch < channels; ++ch) {
                float result = 0;
                for (int j = 0; j < k.taps; ++j) {
                    float c; bool isC;
                    const int pos = wrapIndex(mode, k.first[i] + j, srcRes, c, isC);
                    result += (isC ? c : s[pos * srcElemStride + ch]) * wt[j];
                }
                d[i * dstElemStride + ch] = std::min(maxValue, std::max(0.0f, result));
            }
        }
    }
}

} // namespace

// The reference's bitmap texture keeps its pyramid in half precision (bitmap.cpp:177-180: TMIPMap<Color3, Color3h>): each level is
// resampled in float from the previous float level and rounded to half -- to nearest even, overflow to infinity (half.h:431-487,
// half.cpp:78-200) -- when it is stored (mipmap.h:226-230, :262-264).  The device arrays stay float (the values are half-representable).
static inline float roundToHalf(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    const uint32_t sign = x & 0x80000000u;
    uint32_t a = x & 0x7fffffffu;
    if (a >= 0x7f800000u) return f;
    if (a < 0x38800000u) { // half denormals: spacing 2^-24, the float spacing in [0.5, 1)
        float m; memcpy(&m, &a, 4);
        volatile float t = m + 0.5f;
        m = t - 0.5f;
        memcpy(&a, &m, 4);
    } else {
        a += 0x00000fffu + ((a >> 13) & 1u);
        a &= 0xffffe000u;
        if (a >= 0x47800000u) a = 0x7f800000u;
    }
    a |= sign;
    float r; memcpy(&r, &a, 4);
    return r;
}

static void buildMipPyramidFloat(const float *pixels, int width, int height, int channels, int wrapU, int wrapV, bool pyramid, MipPyramid &out, float maxValue);
void buildMipPyramid(const float *pixels, int width, int height, int channels, int wrapU, int wrapV, bool pyramid, MipPyramid &out, float maxValue) {
    buildMipPyramidFloat(pixels, width, height, channels, wrapU, wrapV, pyramid, out, maxValue);
    for (auto &lvl : out.level) for (float &v : lvl) v = roundToHalf(v);
}

static void buildMipPyramidFloat(const float *pixels, int width, int height, int channels, int wrapU, int wrapV, bool pyramid, MipPyramid &out, float maxValue) {
    out = MipPyramid();
    out.channels = channels;
    std::vector<float> base(pixels, pixels + (size_t) width * height * channels);
    float mx = 0.0f;
    for (float &v : base) { v = std::max(v, 0.0f); mx = std::max(mx, v); }
    out.maximum = mx;
    out.w.push_back(width); out.h.push_back(height);
    out.level.push_back(std::move(base));
    if (!pyramid) return;
    int w = width, h = height;
    while (w > 1 || h > 1) {
        const int nw = std::max(1, (w + 1) / 2), nh = std::max(1, (h + 1) / 2);
        const std::vector<float> &prev = out.level.back();
        std::vector<float> next((size_t) nw * nh * channels), tmp;
        const float *src = prev.data();
        if (w != nw) { // rows first
            float *dst = next.data();
            if (h != nh) { tmp.resize((size_t) nw * h * channels); dst = tmp.data(); }
            resampleAxis(AxisKernel(w, nw), wrapU, w, nw, src, dst, h, (size_t) w * channels, channels, (size_t) nw * channels, channels, channels, maxValue);
            src = dst;
        }
        if (h != nh) // then columns
            resampleAxis(AxisKernel(h, nh), wrapV, h, nh, src, next.data(), nw, channels, (size_t) nw * channels, channels, (size_t) nw * channels, channels, maxValue);
        w = nw; h = nh;
        out.w.push_back(w); out.h.push_back(h);
        out.level.push_back(std::move(next));
    }
}

void ewaWeightTable(float *lut64) {
    for (int i = 0; i < 64; ++i) {
        const float r2 = (float) i / 63.0f;
        lut64[i] = (float) std::exp((double) (-2.0f * r2)) - (float) std::exp((double) -2.0f);
    }
}

} // namespace b2host

// Host-only entry point (no device needed): one level of the pyramid b2_scene_commit would build for this texture
extern "C" int b2_mipmap_level(const b2_texture_desc *t, int level, int *levels, int *width, int *height, float *out) {
    if (!t || !t->pixels || t->width <= 0 || t->height <= 0 || (t->channels != 1 && t->channels != 3)) return B2_ERR_INVALID;
    b2host::MipPyramid mp;
    b2host::buildMipPyramid(t->pixels, t->width, t->height, t->channels, t->wrap_u, t->wrap_v, t->filter_type >= B2_TEX_TRILINEAR, mp);
    if (level < 0 || level >= (int) mp.level.size()) return B2_ERR_INVALID;
    if (levels) *levels = (int) mp.level.size();
    if (width) *width = mp.w[level];
    if (height) *height = mp.h[level];
    if (out) memcpy(out, mp.level[level].data(), mp.level[level].size() * sizeof(float));
    return B2_OK;
}
