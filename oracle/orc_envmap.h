/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's `envmap` emitter (src/emitters/envmap.cpp), used by tests/, smoke() and
 * bench.py's CPU legs as the checker.  Nothing under mitsuba_b200/ may include, link or call it.
 *
 * What is restated, with the reference lines each piece follows:
 *   - storage: a TMIPMap<Spectrum, SpectrumHalf> with ERepeat / EClamp boundaries, EWA filtering, maxAnisotropy 10, resampled with the
 *     2-lobe Lanczos filter and no upper clamp (envmap.cpp:139-175)               -> orc::Texture (orc_texture.h) with maxValue = inf
 *   - the marginal / conditional CDF tables over luminance x sin(theta) (envmap.cpp:260-329)
 *   - evalEnvironment with and without ray differentials (envmap.cpp:380-410)
 *   - internalSampleDirection / internalPdfDirection (envmap.cpp:567-630) and the sampleReuse helper (:651-656)
 *   - sampleDirect / pdfDirect / fillDirectSamplingRecord (envmap.cpp:516-560, :359-374) live with the other emitters in mts_oracle.cpp
 * Image file decoding is not part of the path: pixels arrive as linear float RGB, row-major, top row first.
 * Pinned against the reference's own class compiled into oracle/_ref/libpathref.so (tests/test_oracle_reference_pins.py). */
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "orc_math.h"
#include "orc_texture.h"

namespace orc {

struct EnvMap {
    Texture mip;
    int w = 0, h = 0;
    std::vector<float> cdfRows, cdfCols, rowWeights;
    float normalization = 0, scale = 1, pixelSizeX = 0, pixelSizeY = 0;
    float toWorld[16], toLocal[16]; /* m_worldTransform and its inverse (row-major 4x4) */

    static float luminance(const V3 &c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; } /* spectrum.h:725-727 */
    static V3 xfVector(const float *M, const V3 &v) {
        return V3(M[0] * v.x + M[1] * v.y + M[2] * v.z, M[4] * v.x + M[5] * v.y + M[6] * v.z, M[8] * v.x + M[9] * v.y + M[10] * v.z);
    }

    /* false: the map is black or holds a non-finite value (envmap.cpp:311-315 raise an error) */
    bool build(int width, int height, const float *pixels, float scale_, const float *toWorld_, const float *toLocal_) {
        OrcTextureDesc d;
        d.width = width; d.height = height; d.channels = 3;
        d.filterType = TexEWA; d.wrapU = WrapRepeat; d.wrapV = WrapClamp; d.maxAnisotropy = 10.0f;
        d.uoffset = d.voffset = 0; d.uscale = d.vscale = 1;
        mip.build(d, pixels, std::numeric_limits<float>::infinity());
        scale = scale_;
        memcpy(toWorld, toWorld_, sizeof(toWorld)); memcpy(toLocal, toLocal_, sizeof(toLocal));
        w = width; h = height;
        /* envmap.cpp:263-321 */
        cdfCols.assign((size_t) (w + 1) * h, 0.0f);
        cdfRows.assign((size_t) h + 1, 0.0f);
        rowWeights.assign((size_t) h, 0.0f);
        size_t colPos = 0, rowPos = 0;
        float rowSum = 0.0f;
        cdfRows[rowPos++] = 0;
        for (int y = 0; y < h; ++y) {
            float colSum = 0;
            cdfCols[colPos++] = 0;
            for (int x = 0; x < w; ++x) {
                colSum += luminance(mip.texel(0, x, y));
                cdfCols[colPos++] = colSum;
            }
            const float norm = 1.0f / colSum;
            for (int x = 1; x < w; ++x) cdfCols[colPos - x - 1] *= norm;
            cdfCols[colPos - 1] = 1.0f;
            const float weight = std::sin((y + 0.5f) * kPi / h);
            rowWeights[y] = weight;
            rowSum += colSum * weight;
            cdfRows[rowPos++] = rowSum;
        }
        const float norm = 1.0f / rowSum;
        for (int y = 1; y < h; ++y) cdfRows[rowPos - y - 1] *= norm;
        cdfRows[rowPos - 1] = 1.0f;
        if (rowSum == 0 || !std::isfinite(rowSum)) return false;
        normalization = 1.0f / (rowSum * (2 * kPi / w) * (kPi / h));
        pixelSizeX = 2 * kPi / w; pixelSizeY = kPi / h;
        return true;
    }

    /* envmap.cpp:651-656 */
    static uint32_t sampleReuse(const float *cdf, uint32_t size, float &sample) {
        const float *entry = std::lower_bound(cdf, cdf + size + 1, sample);
        const uint32_t index = std::min((uint32_t) std::max((ptrdiff_t) 0, entry - cdf - 1), size - 1);
        sample = (sample - cdf[index]) / (cdf[index + 1] - cdf[index]);
        return index;
    }
    static float intervalToTent(float sample) { /* src/libcore/warp.cpp:143-155 */
        float sign;
        if (sample < 0.5f) { sign = 1; sample *= 2; } else { sign = -1; sample = 2 * (sample - 0.5f); }
        return sign * (1 - std::sqrt(sample));
    }

    /* evalEnvironment (envmap.cpp:380-410): d = ray direction in world space; rxD / ryD = the differential directions or NULL */
    V3 evalEnvironment(const V3 &dWorld, const V3 *rxD, const V3 *ryD) const {
        const V3 v = xfVector(toLocal, dWorld);
        const float uvx = std::atan2(v.x, -v.z) * kInvTwoPi, uvy = safe_acos(v.y) * kInvPi;
        V3 value;
        if (!rxD) value = mip.evalBilinear(0, uvx, uvy);
        else {
            const V3 dvdx = xfVector(toLocal, *rxD) - v, dvdy = xfVector(toLocal, *ryD) - v;
            const float t1 = kInvTwoPi / (v.x * v.x + v.z * v.z), t2 = -kInvPi / std::max(safe_sqrt(1.0f - v.y * v.y), kEpsilon);
            value = mip.evalFiltered(uvx, uvy, t1 * (dvdx.z * v.x - dvdx.x * v.z), t2 * dvdx.y, t1 * (dvdy.z * v.x - dvdy.x * v.z), t2 * dvdy.y);
        }
        return value * scale;
    }

    /* the four-texel interpolation both helpers share (envmap.cpp:579-591 / :615-629): value1 + value2 and the row-weighted luminance */
    void interpolate(int xPos, int yPos, float dx1, float dy1, V3 &value, float &weightedLum) const {
        const float dx2 = 1.0f - dx1, dy2 = 1.0f - dy1;
        const V3 value1 = mip.texel(0, xPos, yPos) * dx2 * dy2 + mip.texel(0, xPos + 1, yPos) * dx1 * dy2;
        const V3 value2 = mip.texel(0, xPos, yPos + 1) * dx2 * dy1 + mip.texel(0, xPos + 1, yPos + 1) * dx1 * dy1;
        value = value1 + value2;
        weightedLum = luminance(value1) * rowWeights[std::min(std::max(yPos, 0), h - 1)] + luminance(value2) * rowWeights[std::min(std::max(yPos + 1, 0), h - 1)];
    }

    /* internalSampleDirection (envmap.cpp:567-598): d in the emitter's frame */
    void sampleDirection(float sx, float sy, V3 &d, V3 &value, float &pdf) const {
        const uint32_t row = sampleReuse(cdfRows.data(), (uint32_t) h, sy), col = sampleReuse(cdfCols.data() + (size_t) row * (w + 1), (uint32_t) w, sx);
        const float posX = (float) col + intervalToTent(sx), posY = (float) row + intervalToTent(sy);
        const int xPos = floorToInt(posX), yPos = floorToInt(posY);
        float lum;
        interpolate(xPos, yPos, posX - xPos, posY - yPos, value, lum);
        value = value * scale;
        pdf = lum * normalization;
        const float phi = pixelSizeX * (posX + 0.5f), theta = pixelSizeY * (posY + 0.5f);
        float sinPhi, cosPhi, sinTheta, cosTheta;
        sincos(phi, &sinPhi, &cosPhi); sincos(theta, &sinTheta, &cosTheta);
        d = V3(sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta);
        pdf /= std::max(std::abs(sinTheta), kEpsilon);
    }

    /* internalPdfDirection (envmap.cpp:601-630): d in the emitter's frame */
    float pdfDirection(const V3 &d) const {
        const float uvx = std::atan2(d.x, -d.z) * kInvTwoPi, uvy = safe_acos(d.y) * kInvPi;
        if (!std::isfinite(uvx) || !std::isfinite(uvy)) return 0.0f;
        const float u = uvx * w - 0.5f, v = uvy * h - 0.5f;
        const int xPos = floorToInt(u), yPos = floorToInt(v);
        V3 value; float lum;
        interpolate(xPos, yPos, u - xPos, v - yPos, value, lum);
        const float sinTheta = safe_sqrt(1 - d.y * d.y);
        return lum * normalization / std::max(std::abs(sinTheta), kEpsilon);
    }
};

} // namespace orc
