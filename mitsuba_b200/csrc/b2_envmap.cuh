// `envmap` emitter on the device: a latitude-longitude radiance map around the scene
//   EnvironmentMap::evalEnvironment                         src/emitters/envmap.cpp:380-410
//   EnvironmentMap::internalSampleDirection / sampleReuse   src/emitters/envmap.cpp:567-598, :651-656
//   EnvironmentMap::internalPdfDirection                    src/emitters/envmap.cpp:601-630
//   EnvironmentMap::sampleDirect / pdfDirect                src/emitters/envmap.cpp:516-560 (joined to the scene in b2_kernels.inl)
// The MIP pyramid (Lanczos-2 resampling without an upper clamp, half-precision storage, envmap.cpp:139-175) and the marginal / conditional
// CDF tables (envmap.cpp:260-329) are scene preparation and are built by the host at commit.
// Every entry point is one out-of-line copy per kernel: only rays that leave the scene or pick the map as their light run them.
#pragma once
#include "b2_texture.cuh"

namespace b2 {

B2_DEV V3 envToLocal(const DEnvMap &e, const V3 &v) {
    return V3(e.toLocal[0] * v.x + e.toLocal[1] * v.y + e.toLocal[2] * v.z, e.toLocal[3] * v.x + e.toLocal[4] * v.y + e.toLocal[5] * v.z,
              e.toLocal[6] * v.x + e.toLocal[7] * v.y + e.toLocal[8] * v.z);
}
B2_DEV V3 envToWorld(const DEnvMap &e, const V3 &v) {
    return V3(e.toWorld[0] * v.x + e.toWorld[1] * v.y + e.toWorld[2] * v.z, e.toWorld[3] * v.x + e.toWorld[4] * v.y + e.toWorld[5] * v.z,
              e.toWorld[6] * v.x + e.toWorld[7] * v.y + e.toWorld[8] * v.z);
}
B2_DEV float envLuminance(const V3 &c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; } // spectrum.h:725-727

// evalEnvironment: d = world-space ray direction; hasDiff: the ray is a sensor ray and rxD / ryD are its differential directions
static __device__ __noinline__ Spectrum envEval(const DEnvMap &e, const float *lut, V3 d, bool hasDiff, V3 rxD, V3 ryD) {
    const V3 v = envToLocal(e, d);
    const float uvx = atan2f(v.x, -v.z) * B2_INV_TWOPI, uvy = acosf(fminf(1.0f, fmaxf(-1.0f, v.y))) * B2_INV_PI;
    V3 value;
    if (!hasDiff) value = texBilinear(e.tex, 0, uvx, uvy);
    else {
        const V3 dvdx = envToLocal(e, rxD) - v, dvdy = envToLocal(e, ryD) - v;
        const float t1 = B2_INV_TWOPI / (v.x * v.x + v.z * v.z), t2 = -B2_INV_PI / fmaxf(safe_sqrt(1.0f - v.y * v.y), B2_EPSILON);
        value = texFiltered(e.tex, lut, uvx, uvy, t1 * (dvdx.z * v.x - dvdx.x * v.z), t2 * dvdx.y, t1 * (dvdy.z * v.x - dvdy.x * v.z), t2 * dvdy.y);
    }
    return value * e.scale;
}

// the four texels both helpers interpolate (envmap.cpp:579-591 / :615-629): value1 + value2 and the row-weighted luminance
B2_DEV void envInterpolate(const DEnvMap &e, int xPos, int yPos, float dx1, float dy1, V3 &value, float &weightedLum) {
    const float dx2 = 1.0f - dx1, dy2 = 1.0f - dy1;
    const V3 value1 = texTexel(e.tex, 0, xPos, yPos) * dx2 * dy2 + texTexel(e.tex, 0, xPos + 1, yPos) * dx1 * dy2;
    const V3 value2 = texTexel(e.tex, 0, xPos, yPos + 1) * dx2 * dy1 + texTexel(e.tex, 0, xPos + 1, yPos + 1) * dx1 * dy1;
    value = value1 + value2;
    weightedLum = envLuminance(value1) * __ldg(e.rowWeights + min(max(yPos, 0), e.h - 1)) + envLuminance(value2) * __ldg(e.rowWeights + min(max(yPos + 1, 0), e.h - 1));
}

// envmap.cpp:651-656: std::lower_bound over cdf[0..size], index clamp, sample reuse
B2_DEV uint32_t envSampleReuse(const float *__restrict__ cdf, uint32_t size, float &sample) {
    uint32_t lo = 0, cnt = size + 1;
    while (cnt > 0) {
        const uint32_t step = cnt >> 1, it = lo + step;
        if (__ldg(cdf + it) < sample) { lo = it + 1; cnt -= step + 1; }
        else cnt = step;
    }
    const uint32_t index = min((uint32_t) max(0, (int) lo - 1), size - 1);
    const float c0 = __ldg(cdf + index), c1 = __ldg(cdf + index + 1);
    sample = (sample - c0) / (c1 - c0);
    return index;
}
B2_DEV float envIntervalToTent(float sample) { // src/libcore/warp.cpp:143-155
    float sign;
    if (sample < 0.5f) { sign = 1; sample *= 2; } else { sign = -1; sample = 2 * (sample - 0.5f); }
    return sign * (1 - sqrtf(sample));
}

// internalSampleDirection: d in the map's own frame, value = radiance (scaled), pdf = solid-angle density
static __device__ __noinline__ void envSampleDirection(const DEnvMap &e, float sx, float sy, V3 &d, Spectrum &value, float &pdf) {
    const uint32_t row = envSampleReuse(e.cdfRows, (uint32_t) e.h, sy);
    const uint32_t col = envSampleReuse(e.cdfCols + (size_t) row * (e.w + 1), (uint32_t) e.w, sx);
    const float posX = (float) col + envIntervalToTent(sx), posY = (float) row + envIntervalToTent(sy);
    const int xPos = (int) floorf(posX), yPos = (int) floorf(posY);
    float lum;
    envInterpolate(e, xPos, yPos, posX - xPos, posY - yPos, value, lum);
    value = value * e.scale;
    pdf = lum * e.normalization;
    float sinPhi, cosPhi, sinTheta, cosTheta;
    sincosf(e.pixelSizeX * (posX + 0.5f), &sinPhi, &cosPhi);
    sincosf(e.pixelSizeY * (posY + 0.5f), &sinTheta, &cosTheta);
    d = V3(sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta);
    pdf /= fmaxf(fabsf(sinTheta), B2_EPSILON);
}

// internalPdfDirection: d in the map's own frame
static __device__ __noinline__ float envPdfDirection(const DEnvMap &e, V3 d) {
    const float uvx = atan2f(d.x, -d.z) * B2_INV_TWOPI, uvy = acosf(fminf(1.0f, fmaxf(-1.0f, d.y))) * B2_INV_PI;
    if (!isfinite(uvx) || !isfinite(uvy)) return 0.0f;
    const float u = uvx * e.w - 0.5f, v = uvy * e.h - 0.5f;
    const int xPos = (int) floorf(u), yPos = (int) floorf(v);
    V3 value;
    float lum;
    envInterpolate(e, xPos, yPos, u - xPos, v - yPos, value, lum);
    const float sinTheta = safe_sqrt(1 - d.y * d.y);
    return lum * e.normalization / fmaxf(fabsf(sinTheta), B2_EPSILON);
}

} // namespace b2
