// MIP pyramid of a `bitmap` texture, built on the host at scene commit the way the reference builds it when the plugin is
// constructed (include/mitsuba/render/mipmap.h:155-303): level 0 = the image with negative values clamped, every further level a
// separable 2-lobe Lanczos resampling of the previous one to ceil(size / 2), clamped to [0, 1] (mipmap.h:269; bitmap.cpp:2230-2329;
// rfilter.h:107-330; src/rfilters/lanczos.cpp:43-56).
#pragma once
#include <cstdint>
#include <vector>

namespace b2host {

struct MipPyramid {
    int channels = 0;
    std::vector<int> w, h;                  // per level
    std::vector<std::vector<float>> level;  // level[l][(y * w[l] + x) * channels + c]
    float maximum = 0.0f;                   // largest component of level 0 (mipmap.h:229, after clampNegative)
};

// wrap modes: 0 repeat, 1 clamp, 2 mirror, 3 zero, 4 one.  `pyramid` false: level 0 only (filter types nearest / bilinear, mipmap.h:185,246)
// maxValue: upper clamp of the resampled levels (mipmap.h:156-157, :262): 1 for `bitmap` textures, infinity for `envmap` (envmap.cpp:172-175)
void buildMipPyramid(const float *pixels, int width, int height, int channels, int wrapU, int wrapV, bool pyramid, MipPyramid &out, float maxValue = 1.0f);

// EWA filter weights (mipmap.h:296-302)
void ewaWeightTable(float *lut64);

} // namespace b2host
