/* Thin C entry points over the REFERENCE's own TMIPMap look-ups (include/mitsuba/render/mipmap.h:499-838: evalTexel, evalBox,
 * evalBilinear, evalEWA, eval), compiled where they lie under /root/reference (never copied) into oracle/_ref/libmipmapref.so by
 * oracle/Makefile.  The pyramid levels are handed in by the caller and enter the real class through its cache-file constructor
 * (mipmap.h:320-380) over an in-memory "file" laid out with the reference's BlockedArray.  Used only to pin the oracle's look-ups
 * (orc_texture.h) and to generate tests/golden/mipmap_ref.npz. */
#include <mitsuba/render/mipmap.h>

namespace mitsuba {
namespace stats {
StatsCounter avgEWASamples, clampedAnisotropy, mipStorage, filteredLookups;
}
/* math::hypot2 / math::log2 come from the reference's own src/libcore/math.cpp, compiled alongside (see Makefile) */
std::ostream &operator<<(std::ostream &os, const ReconstructionFilter::EBoundaryCondition &value) { return os << (int) value; } /* only toString() uses it */
}

using namespace mitsuba;
typedef TMIPMap<Color3, Color3> MIPMap3;

/* the private header struct of the cache file, reachable through a derived class */
struct Builder : public MIPMap3 {
    typedef MIPMap3::MIPMapHeader Header;
    typedef MIPMap3::Array2DType Array;
};

extern "C" {
/* levels: nLevels RGB images (float, 3 per texel, row-major), sizes (w, h) per level; bc in the reference's enum order
 * (0 clamp, 1 repeat, 2 mirror, 3 zero, 4 one); filterType 0 nearest, 1 bilinear, 2 trilinear, 3 ewa */
void *mipref_create(int nLevels, const int *sizes, const float *const *levels, int bcu, int bcv, int filterType, float maxAnisotropy) {
    size_t padding = sizeof(Builder::Header) % MTS_MIPMAP_CACHE_ALIGNMENT;
    if (padding) padding = MTS_MIPMAP_CACHE_ALIGNMENT - padding;
    size_t total = sizeof(Builder::Header) + padding;
    for (int l = 0; l < nLevels; ++l) total += Builder::Array::bufferSize(Vector2i(sizes[2 * l], sizes[2 * l + 1]));
    uint8_t *buf = (uint8_t *) allocAligned(total);
    memset(buf, 0, total);
    Builder::Header header;
    memset(&header, 0, sizeof(header));
    memcpy(header.identifier, "MIP", 3);
    header.version = MTS_MIPMAP_CACHE_VERSION;
    header.pixelFormat = (uint8_t) Bitmap::ERGB;
    header.levels = (uint8_t) nLevels;
    header.bcu = (uint8_t) bcu; header.bcv = (uint8_t) bcv;
    header.filterType = (uint8_t) filterType;
    header.gamma = 1.0f;
    header.width = sizes[0]; header.height = sizes[1];
    memcpy(buf, &header, sizeof(header));
    uint8_t *ptr = buf + sizeof(Builder::Header) + padding;
    for (int l = 0; l < nLevels; ++l) {
        Builder::Array a;
        a.map(ptr, Vector2i(sizes[2 * l], sizes[2 * l + 1]));
        a.init((const Color3 *) levels[l]); /* linear -> blocked layout with the reference's own indexing */
        ptr += a.getBufferSize();
    }
    MemoryMappedFile::registeredData() = buf;
    MemoryMappedFile::registeredSize() = total;
    return new MIPMap3(fs::path("memory"), maxAnisotropy);
}
void mipref_eval(void *m, int n, const float *uv, const float *d0d1 /* n x 4: d0.x d0.y d1.x d1.y */, float *out) {
    const MIPMap3 *mm = (const MIPMap3 *) m;
    for (int i = 0; i < n; ++i) {
        Color3 v = mm->eval(Point2(uv[2 * i], uv[2 * i + 1]), Vector2(d0d1[4 * i], d0d1[4 * i + 1]), Vector2(d0d1[4 * i + 2], d0d1[4 * i + 3]));
        out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2];
    }
}
void mipref_eval_bilinear(void *m, int level, int n, const float *uv, float *out) {
    const MIPMap3 *mm = (const MIPMap3 *) m;
    for (int i = 0; i < n; ++i) {
        Color3 v = mm->evalBilinear(level, Point2(uv[2 * i], uv[2 * i + 1]));
        out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2];
    }
}
void mipref_eval_box(void *m, int level, int n, const float *uv, float *out) {
    const MIPMap3 *mm = (const MIPMap3 *) m;
    for (int i = 0; i < n; ++i) {
        Color3 v = mm->evalBox(level, Point2(uv[2 * i], uv[2 * i + 1]));
        out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2];
    }
}
}
