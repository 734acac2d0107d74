// `bitmap` texture look-ups on the device:
//   TMIPMap::evalTexel / evalBox / evalBilinear / evalEWA / eval   include/mitsuba/render/mipmap.h:499-571,586-608,638-721,767-838
//   BitmapTexture::eval(uv) / eval(uv, d0, d1)                      src/textures/bitmap.cpp:400-421,452-465
//   Texture2D::eval(its, filter)                                    src/librender/texture.cpp:124-133
//   Intersection::computePartials                                   src/librender/intersection.cpp:23-85
// The pyramid itself (Lanczos-2 resampling, mipmap.h:155-303) is scene preparation and is built by the host at commit.
#pragma once
#include "b2_math.cuh"
#include "b2_types.h"

namespace b2 {

#define B2_MIPMAP_LUT_SIZE 64 // mipmap.h:37

B2_DEV int texModulo(int a, int b) { // math::modulo
    const int r = a % b;
    return r < 0 ? r + b : r;
}

// boundary handling of one coordinate; returns false when the texel is the constant 0 / 1 of the zero / one modes
B2_DEV bool texWrap(int mode, int size, int &x, float &constant) {
    if (x >= 0 && x < size) return true;
    switch (mode) {
        case 0: x = texModulo(x, size); return true;
        case 1: x = min(max(x, 0), size - 1); return true;
        case 2: x = texModulo(x, 2 * size); if (x >= size) x = 2 * size - x - 1; return true;
        case 3: constant = 0.0f; return false;
        default: constant = 1.0f; return false;
    }
}

B2_DEV V3 texTexel(const DTexture &t, int level, int x, int y) {
    const int sx = t.lw[level], sy = t.lh[level];
    float c = 0.0f;
    if (!texWrap(t.wrapU, sx, x, c)) return V3(c);
    if (!texWrap(t.wrapV, sy, y, c)) return V3(c);
    const size_t idx = (size_t) t.off[level] + (size_t) y * sx + x;
    if (t.channels == 3) {
        const float4 v = __ldg((const float4 *) t.data + idx);
        return V3(v.x, v.y, v.z);
    }
    return V3(__ldg((const float *) t.data + idx));
}

B2_DEV V3 texBox(const DTexture &t, int level, float u, float v) {
    return texTexel(t, level, (int) floorf(u * t.lw[level]), (int) floorf(v * t.lh[level]));
}

B2_DEV V3 texBilinear(const DTexture &t, int level, float uu, float vv) {
    if (!isfinite(uu) || !isfinite(vv)) return V3(0.0f);
    if (level >= t.levels) return texBox(t, t.levels - 1, uu, vv);
    const float u = uu * t.lw[level] - 0.5f, v = vv * t.lh[level] - 0.5f;
    const int xPos = (int) floorf(u), yPos = (int) floorf(v);
    const float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
    return texTexel(t, level, xPos, yPos) * dx2 * dy2 + texTexel(t, level, xPos, yPos + 1) * dx2 * dy1 + texTexel(t, level, xPos + 1, yPos) * dx1 * dy2 +
           texTexel(t, level, xPos + 1, yPos + 1) * dx1 * dy1;
}

B2_DEV V3 texEWA(const DTexture &t, const float *lut, int level, float uu, float vv, float A, float B, float C) {
    if (!isfinite(A + B + C + uu + vv)) return V3(0.0f);
    if (level >= t.levels) return texBox(t, t.levels - 1, uu, vv);
    const float u = uu * t.lw[level] - 0.5f, v = vv * t.lh[level] - 0.5f;
    const float ratioX = (float) t.lw[level] / (float) t.lw[0], ratioY = (float) t.lh[level] / (float) t.lh[0]; // m_sizeRatio
    A /= ratioX * ratioX;
    B /= ratioX * ratioY;
    C /= ratioY * ratioY;
    const float invDet = 1.0f / (-B * B + 4.0f * A * C), deltaU = 2.0f * sqrtf(C * invDet), deltaV = 2.0f * sqrtf(A * invDet);
    const int u0 = (int) ceilf(u - deltaU), u1 = (int) floorf(u + deltaU), v0 = (int) ceilf(v - deltaV), v1 = (int) floorf(v + deltaV);
    const float As = A * B2_MIPMAP_LUT_SIZE, Bs = B * B2_MIPMAP_LUT_SIZE, Cs = C * B2_MIPMAP_LUT_SIZE;
    V3 result(0.0f);
    float denominator = 0.0f;
    const float ddq = 2 * As, uu0 = (float) u0 - u;
    for (int vt = v0; vt <= v1; ++vt) {
        const float vd = (float) vt - v;
        float q = As * uu0 * uu0 + (Bs * uu0 + Cs * vd) * vd;
        float dq = As * (2 * uu0 + 1) + Bs * vd;
        for (int ut = u0; ut <= u1; ++ut) {
            if (q < (float) B2_MIPMAP_LUT_SIZE) {
                const uint32_t qi = (uint32_t) q; // round-off can leave q slightly below 0: that converts to entry 0 here and in the reference
                if (qi < B2_MIPMAP_LUT_SIZE) {
                    const float weight = __ldg(lut + qi);
                    result = result + texTexel(t, level, ut, vt) * weight;
                    denominator += weight;
                }
            }
            q += dq;
            dq += ddq;
        }
    }
    if (denominator == 0) return texBilinear(t, level, uu, vv);
    return result / denominator;
}

B2_DEV float texLog2(float v) { // math::log2 (math.cpp:103-106)
    const float invLn2 = 1.0f / 0.693147182464599609375f; // 1 / logf(2)
    return fastlog(v) * invLn2;
}

// TMIPMap::eval(uv, d0, d1)
B2_DEV V3 texFiltered(const DTexture &t, const float *lut, float u, float v, float d0x, float d0y, float d1x, float d1y) {
    if (t.filter == 0) return texBox(t, 0, u, v);
    if (t.filter == 1) return texBilinear(t, 0, u, v);
    const float du0 = d0x * t.lw[0], dv0 = d0y * t.lh[0], du1 = d1x * t.lw[0], dv1 = d1y * t.lh[0];
    float A = dv0 * dv0 + dv1 * dv1, B = -2.0f * (du0 * dv0 + du1 * dv1), C = du0 * du0 + du1 * du1, F = A * C - B * B * 0.25f;
    const float root = hypot2(A - C, B), Aprime = 0.5f * (A + C - root), Cprime = 0.5f * (A + C + root);
    float majorRadius = Aprime != 0 ? sqrtf(F / Aprime) : 0.0f, minorRadius = Cprime != 0 ? sqrtf(F / Cprime) : 0.0f;
    if (t.filter == 2 || !(minorRadius > 0) || !(majorRadius > 0) || F < 0) {
        const float level = texLog2(fmaxf(majorRadius, B2_EPSILON));
        const int ilevel = (int) floorf(level);
        if (ilevel < 0) return texBilinear(t, 0, u, v);
        const float a = level - ilevel;
        return texBilinear(t, ilevel, u, v) * (1.0f - a) + texBilinear(t, ilevel + 1, u, v) * a;
    }
    if (minorRadius * t.maxAnisotropy < majorRadius) { // enlarge skinny ellipses (mipmap.h:673-697)
        minorRadius = majorRadius / t.maxAnisotropy;
        const float theta = 0.5f * atanf(B / (A - C));
        float sinTheta, cosTheta;
        sincosf(theta, &sinTheta, &cosTheta);
        const float a2 = majorRadius * majorRadius, b2 = minorRadius * minorRadius, sinTheta2 = sinTheta * sinTheta, cosTheta2 = cosTheta * cosTheta,
                    sin2Theta = 2 * sinTheta * cosTheta;
        A = a2 * cosTheta2 + b2 * sinTheta2;
        B = (a2 - b2) * sin2Theta;
        C = a2 * sinTheta2 + b2 * cosTheta2;
        F = a2 * b2;
    }
    const float scale = 1.0f / F;
    A *= scale; B *= scale; C *= scale;
    const float level = fmaxf(0.0f, texLog2(minorRadius));
    const int ilevel = (int) level;
    const float a = level - ilevel;
    if (majorRadius < 1 || !(A > 0 && C > 0)) return texBilinear(t, ilevel, u, v);
    return texEWA(t, lut, ilevel, u, v, A, B, C) * (1.0f - a) + texEWA(t, lut, ilevel + 1, u, v, A, B, C) * a;
}

// uv and uv partials of one intersection as the texture look-up reads them (shape.h:147-165)
struct TexCoord {
    float u, v;
    bool hasUVPartials;
    float dudx, dudy, dvdx, dvdy;
};

// Texture2D::eval(its, filter = true) times the energy-conservation scale of the BSDF that owns the texture.  One copy per kernel
// (not inlined): a look-up is a long, divergent piece of code that only textured hits run.
static __device__ __noinline__ Spectrum texEval(const DTexture &t, const float *lut, const TexCoord &c) {
    const float u = c.u * t.uscale + t.uoffset, v = c.v * t.vscale + t.voffset;
    V3 r;
    if (c.hasUVPartials) r = texFiltered(t, lut, u, v, c.dudx * t.uscale, c.dvdx * t.vscale, c.dudy * t.uscale, c.dvdy * t.vscale);
    else r = t.filter != 0 ? texBilinear(t, 0, u, v) : texBox(t, 0, u, v);
    return r * t.bsdfScale;
}

B2_DEV bool solveLinearSystem2x2(float a00, float a01, float a10, float a11, float b0, float b1, float &x0, float &x1) { // util.cpp:527-539
    const float det = a00 * a11 - a01 * a10;
    if (fabsf(det) <= 2.93873587705571876e-39f) return false; // RCPOVERFLOW_FLT
    const float inverse = 1.0f / det;
    x0 = (a11 * b0 - a01 * b1) * inverse;
    x1 = (a00 * b1 - a10 * b0) * inverse;
    return true;
}

// Intersection::computePartials for a ray with differentials (rxOrigin = ryOrigin = o)
B2_DEV void computeUVPartials(const V3 &p, const V3 &n, const V3 &dpdu, const V3 &dpdv, const V3 &o, const V3 &rxD, const V3 &ryD, TexCoord &c) {
    c.hasUVPartials = true;
    c.dudx = c.dvdx = c.dudy = c.dvdy = 0.0f;
    if (isZero(dpdu) && isZero(dpdv)) return;
    const float pp = dot(n, p), pox = dot(n, o), prx = dot(n, rxD), pry = dot(n, ryD);
    if (prx == 0 || pry == 0) return;
    const float tx = (pp - pox) / prx, ty = (pp - pox) / pry;
    const float absX = fabsf(n.x), absY = fabsf(n.y), absZ = fabsf(n.z);
    int a0, a1;
    if (absX > absY && absX > absZ) { a0 = 1; a1 = 2; }
    else if (absY > absZ) { a0 = 0; a1 = 2; }
    else { a0 = 0; a1 = 1; }
    const V3 px = o + rxD * tx, py = o + ryD * ty;
    const float A00 = comp(dpdu, a0), A01 = comp(dpdv, a0), A10 = comp(dpdu, a1), A11 = comp(dpdv, a1);
    float x0, x1;
    if (solveLinearSystem2x2(A00, A01, A10, A11, comp(px, a0) - comp(p, a0), comp(px, a1) - comp(p, a1), x0, x1)) { c.dudx = x0; c.dvdx = x1; }
    else { c.dudx = 1; c.dvdx = 0; }
    if (solveLinearSystem2x2(A00, A01, A10, A11, comp(py, a0) - comp(p, a0), comp(py, a1) - comp(p, a1), x0, x1)) { c.dudy = x0; c.dvdy = x1; }
    else c.dudy = 1; // intersection.cpp:82 assigns dudy twice; dvdy keeps its previous value (0: the record was just created)
}

} // namespace b2
