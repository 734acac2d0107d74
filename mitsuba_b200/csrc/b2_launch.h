// Host-callable launchers of the wavefront kernels.  The kernel translation unit is compiled twice
// (b2::parity with -fmad=false, b2::fast with FMA contraction); both export the same functions.
#pragma once
#include "b2_types.h"

#define B2_TRACE_BLOCK 256
#define B2_SHADE_BLOCK 128
#define B2_BIN_BITS 17
#define B2_NBINS (1u << B2_BIN_BITS)

namespace b2 {

struct LaunchCfg {
    int numSMs = 0;
    size_t traceSmem = 0;
    int gridExtend = 0, gridExtendSort = 0, gridOccluded = 0, gridTrace = 0, gridGenerate = 0, gridVolstep = 0, gridVolLockstep = 0;
    int volLockstep = 1; // B2_VOL_LOCKSTEP=0: the ticketed k_volstep instead of k_volstep_lockstep
    int gridShade[5] = {0, 0, 0, 0, 0};
    int gridShadeTex = 0; // k_shade<-1, TEX = true> (textured scenes)
    int gridShadeTexCls[4] = {0, 0, 0, 0}; // k_shade<c, TEX = true>: class-sorted dispatch of textured / environment-mapped scenes
    // flat-leaf variants (shared-memory resident scenes): k_extend_flat / k_occluded_flat
    size_t flatSmem = 0;
    int gridExtendFlat = 0, gridExtendFlatSort = 0, gridOccludedFlat = 0;
};

#define B2_DECLARE_LAUNCHERS(NS)                                                                                               \
    namespace NS {                                                                                                             \
    void KernelSet_init(LaunchCfg &cfg, const DScene &sc, int numSMs);                                                         \
    void launch_generate(const LaunchCfg &, const DScene &, const DPool &, const DRender &, const DFilter &, cudaStream_t);    \
    void launch_extend(const LaunchCfg &, const DScene &, const DPool &, const DRender &, bool sort, cudaStream_t);            \
    void launch_shade(const LaunchCfg &, const DScene &, const DPool &, const DRender &, int cls, bool queued, cudaStream_t);  \
    void launch_occluded(const LaunchCfg &, const DScene &, const DPool &, const DRender &, cudaStream_t);                     \
    void launch_volstep(const LaunchCfg &, const DScene &, const DPool &, const DRender &, cudaStream_t);                      \
    void launch_medium_probe(const LaunchCfg &, const DScene &, int medium, int what, uint64_t n, const float *in,             \
                             uint64_t seed, float *out, cudaStream_t);                                                         \
    void launch_film_pack(const LaunchCfg &, const float4 *rgba, const float *w, float *out, size_t n, cudaStream_t);          \
    void launch_trace(const LaunchCfg &, const DScene &, const float4 *rays, float4 *out, uint64_t n, bool shadow, bool count, \
                      unsigned long long *counters, const uint32_t *order, cudaStream_t);                                      \
    void launch_bin(const LaunchCfg &, const DScene &, const DPool &, const float4 *rays, uint32_t n, uint32_t *keys,          \
                    uint32_t *hist, uint32_t *order, cudaStream_t);                                                            \
    void launch_bsdf_eval(const LaunchCfg &, const DScene &, int mat, uint64_t n, const float *wi, const float *wo,            \
                          float *rgb, float *pdf, cudaStream_t);                                                               \
    void launch_bsdf_sample(const LaunchCfg &, const DScene &, int mat, uint64_t n, const float *wi, const float *samples,     \
                            float *out, cudaStream_t);                                                                         \
    void launch_emitter_direct(const LaunchCfg &, const DScene &, uint64_t n, const float *ref, const float *samples,          \
                               float *out, cudaStream_t);                                                                      \
    void launch_camera_rays(const LaunchCfg &, const DScene &, uint64_t n, const float *pos, float *rays, cudaStream_t);       \
    void launch_texture_probe(const LaunchCfg &, const DScene &, int what, int tex, int hasPartials, float diffScale, uint64_t n, const float *in,   \
                              float *out, cudaStream_t);                                                                    \
    void launch_envmap_probe(const LaunchCfg &, const DScene &, int what, uint64_t n, const float *in, float *out, cudaStream_t); \
    void launch_sampler_stream(const DScene &, const DRender &, int px, int py, int sampleIdx, int ndim, float *out,           \
                               cudaStream_t);                                                                                  \
    void launch_splat(const LaunchCfg &, const DFilter &, int W, int H, uint64_t n, const float *pos, const float *val,        \
                      float4 *rgba, float *wgt, cudaStream_t);                                                                 \
    }

B2_DECLARE_LAUNCHERS(parity)
B2_DECLARE_LAUNCHERS(fast)

} // namespace b2
