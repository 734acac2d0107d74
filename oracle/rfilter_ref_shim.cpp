/* Thin C entry points over the REFERENCE's own Resampler<float> (include/mitsuba/core/rfilter.h:107-449) and LanczosSincFilter
 * (src/rfilters/lanczos.cpp), compiled where they lie under /root/reference (never copied) into oracle/_ref/librfilterref.so by
 * oracle/Makefile.  Used only to pin the oracle's MIP-pyramid resampler (orc_texture.h) and to generate tests/golden/resample_ref.npz. */
#include <mitsuba/core/cobject.h> /* oracle/shim_rfilter stand-in */
#include "lanczos.cpp"            /* -I$(REF)/src/rfilters */

namespace mitsuba {
/* out-of-line members of ReconstructionFilter that live in src/libcore/rfilter.cpp (which needs Stream / Properties for real);
   the Resampler only calls getRadius() and the virtual eval() */
ReconstructionFilter::ReconstructionFilter(const Properties &props) : ConfigurableObject(props), m_radius(0), m_scaleFactor(0), m_borderSize(0) {}
ReconstructionFilter::ReconstructionFilter(Stream *stream, InstanceManager *manager) : ConfigurableObject(stream, manager) {}
ReconstructionFilter::~ReconstructionFilter() {}
void ReconstructionFilter::configure() {}
void ReconstructionFilter::serialize(Stream *, InstanceManager *) const {}
}

using namespace mitsuba;

extern "C" {
/* boundary condition in the reference's enum order: 0 clamp, 1 repeat, 2 mirror, 3 zero, 4 one */
void rfref_resample(int bc, int lobes, int srcRes, int trgRes, const float *src, int srcStride, float *dst, int dstStride, int channels, int clampResult) {
    Properties props;
    props.lobes = lobes;
    LanczosSincFilter filter(props);
    Resampler<float> r(&filter, (ReconstructionFilter::EBoundaryCondition) bc, srcRes, trgRes);
    if (clampResult) r.resampleAndClamp(src, (size_t) srcStride, dst, (size_t) dstStride, channels, 0.0f, 1.0f);
    else r.resample(src, (size_t) srcStride, dst, (size_t) dstStride, channels);
}
float rfref_lanczos(int lobes, float x) {
    Properties props;
    props.lobes = lobes;
    LanczosSincFilter filter(props);
    return filter.eval(x);
}
}
