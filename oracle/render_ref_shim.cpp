/* Thin C entry points over more of the REFERENCE's own code, compiled where it lies under /root/reference (never copied) into
 * oracle/_ref/librenderref.so by oracle/Makefile, behind the stand-in headers of oracle/shim_core/:
 *   src/librender/intersection.cpp                        Intersection::computePartials
 *   src/libcore/rfilter.cpp + src/rfilters/{gaussian,box}.cpp   ReconstructionFilter::configure (the 32-entry table), radius, border
 *   include/mitsuba/render/imageblock.h + imageblock.cpp  ImageBlock::put(pos, spec, alpha): the filtered splat
 *   src/samplers/sobol.cpp (+ sobolseq.cpp, qmc.cpp, librender's sampler.cpp)   SobolSampler: generate / advance / next1D / next2D
 * Scaffolding (not the reference): Bitmap's constructor and clear() (src/libcore/bitmap.cpp needs OpenEXR / libpng / libjpeg) get a
 * plain zeroed float buffer; base-class members of ConfigurableObject / SerializableObject.
 * Used only to pin the oracle (tests/gen_golden.py -> tests/golden/render_ref.npz; tests/test_oracle_reference_pins.py). */
#include <mitsuba/render/shape.h>
#include <mitsuba/render/imageblock.h>
#include <mitsuba/render/sampler.h>
#include <mitsuba/core/rfilter.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/random.h>

namespace mitsuba {
/* ---- scaffolding ---- */
ConfigurableObject::ConfigurableObject(Stream *, InstanceManager *) {}
void ConfigurableObject::setParent(ConfigurableObject *) {}
void ConfigurableObject::addChild(const std::string &, ConfigurableObject *) {}
void ConfigurableObject::configure() {}
void ConfigurableObject::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(ConfigurableObject, true, SerializableObject)
SerializableObject::SerializableObject(Stream *, InstanceManager *) {}
MTS_IMPLEMENT_CLASS(SerializableObject, true, Object)
void InstanceManager::serialize(Stream *, const SerializableObject *) {}
Float Stream::readFloat() { return 0; }
void Stream::writeFloat(Float) {}
void Stream::readFloatArray(Float *, size_t) {}
void Stream::writeFloatArray(const Float *, size_t) {}
size_t Stream::readSize() { return 0; }
void Stream::writeSize(size_t) {}
void Stream::read(void *, size_t) {}
void Stream::write(const void *, size_t) {}
size_t Stream::getPos() const { return 0; }
size_t Stream::getSize() const { return 0; }
void Stream::seek(size_t) {}
void Stream::flush() {}
void Stream::truncate(size_t) {}
bool Stream::canRead() const { return false; }
bool Stream::canWrite() const { return false; }
int Stream::readInt() { return 0; }
void Stream::writeInt(int) {}
unsigned int Stream::readUInt() { return 0; }
void Stream::writeUInt(unsigned int) {}
bool Stream::readBool() { return false; }
void Stream::writeBool(bool) {}
unsigned long long Stream::readULong() { return 0; }
void Stream::writeULong(unsigned long long) {}
void Stream::readULongArray(uint64_t *, size_t) {}
void Stream::writeULongArray(const uint64_t *, size_t) {}
template <> int Stream::readElement<int>() { return 0; }
template <> void Stream::writeElement<int>(int) {}
MTS_IMPLEMENT_CLASS(WorkResult, true, Object)
/* Bitmap storage for ImageBlock: ESpectrumAlphaWeight in float32 = SPECTRUM_SAMPLES + 2 = 5 channels (bitmap.cpp's updateChannelCount) */
Bitmap::Bitmap(EPixelFormat pFmt, EComponentFormat cFmt, const Vector2i &size, uint8_t channelCount, uint8_t *)
    : m_pixelFormat(pFmt), m_componentFormat(cFmt), m_size(size), m_data(NULL), m_gamma(1.0f), m_channelCount(channelCount), m_ownsData(true) {
    if (pFmt == ESpectrumAlphaWeight) m_channelCount = SPECTRUM_SAMPLES + 2;
    m_data = (uint8_t *) calloc((size_t) size.x * size.y * m_channelCount, sizeof(float));
}
Bitmap::~Bitmap() { if (m_data && m_ownsData) free(m_data); }
void Bitmap::clear() { memset(m_data, 0, (size_t) m_size.x * m_size.y * m_channelCount * sizeof(float)); }
std::string Bitmap::toString() const { return "Bitmap"; }
MTS_IMPLEMENT_CLASS(Bitmap, false, Object)
}

using namespace mitsuba;
extern "C" void *CreateInstance_gaussian(const Properties &props);
extern "C" void *CreateInstance_box(const Properties &props);
extern "C" void *CreateInstance_sobol(const Properties &props);

static ReconstructionFilter *makeFilter(int kind) {
    Properties props;
    ReconstructionFilter *f = (ReconstructionFilter *) (kind == 0 ? CreateInstance_box(props) : CreateInstance_gaussian(props));
    f->configure();
    return f;
}

extern "C" {
/* rec 27n: p, geoFrame.n, dpdu, dpdv, ray.o, rxOrigin, ryOrigin, rxDirection, ryDirection -> out 4n: dudx dudy dvdx dvdy */
void renderref_compute_partials(int n, const float *rec, float *out) {
    for (int i = 0; i < n; ++i) {
        const float *r = rec + 27 * i;
        Intersection its;
        its.p = Point(r[0], r[1], r[2]);
        its.geoFrame = Frame(Normal(r[3], r[4], r[5]));
        its.geoFrame.n = Normal(r[3], r[4], r[5]);
        its.dpdu = Vector(r[6], r[7], r[8]); its.dpdv = Vector(r[9], r[10], r[11]);
        its.hasUVPartials = false;
        its.dudx = its.dudy = its.dvdx = its.dvdy = 0;
        RayDifferential ray;
        ray.o = Point(r[12], r[13], r[14]); ray.d = Vector(0, 0, 1);
        ray.rxOrigin = Point(r[15], r[16], r[17]); ray.ryOrigin = Point(r[18], r[19], r[20]);
        ray.rxDirection = Vector(r[21], r[22], r[23]); ray.ryDirection = Vector(r[24], r[25], r[26]);
        ray.hasDifferentials = true;
        its.computePartials(ray);
        out[4 * i] = its.dudx; out[4 * i + 1] = its.dudy; out[4 * i + 2] = its.dvdx; out[4 * i + 3] = its.dvdy;
    }
}
/* kind 0 box, 1 gaussian (plugin defaults): the discretised table through evalDiscretized, radius, border */
void renderref_filter_table(int kind, float *values32, float *radius, int *border) {
    ReconstructionFilter *f = makeFilter(kind);
    *radius = f->getRadius(); *border = f->getBorderSize();
    /* evalDiscretized(x) = m_values[min((int) |x * m_scaleFactor|, 31)] with m_scaleFactor = 31 / radius (rfilter.cpp:37-57) */
    const Float scale = MTS_FILTER_RESOLUTION / f->getRadius();
    for (int i = 0; i <= MTS_FILTER_RESOLUTION; ++i) values32[i] = f->evalDiscretized((i + 0.5f) / scale);
}
/* one ImageBlock of w x h pixels at offset (ox, oy): n samples pos (2n, film coordinates) val (4n: rgb, alpha)
 * -> data (w + 2 border) x (h + 2 border) x 5, ok n */
void renderref_block_put(int ox, int oy, int w, int h, int kind, int n, const float *pos, const float *val, float *data, int *ok) {
    ReconstructionFilter *f = makeFilter(kind);
    ref<ImageBlock> blk = new ImageBlock(Bitmap::ESpectrumAlphaWeight, Vector2i(w, h), f);
    blk->setOffset(Point2i(ox, oy));
    blk->clear();
    for (int i = 0; i < n; ++i) {
        Spectrum s; s[0] = val[4 * i]; s[1] = val[4 * i + 1]; s[2] = val[4 * i + 2];
        ok[i] = blk->put(Point2(pos[2 * i], pos[2 * i + 1]), s, val[4 * i + 3]) ? 1 : 0;
    }
    const Bitmap *bmp = blk->getBitmap();
    memcpy(data, bmp->getFloat32Data(), sizeof(float) * 5 * (size_t) (w + 2 * f->getBorderSize()) * (h + 2 * f->getBorderSize()));
}
/* SobolSampler as SamplingIntegrator drives it: setFilmResolution(blocked), generate(pixel), advance() x sampleIdx, next2D, next1D... */
void renderref_sobol_stream(uint64_t scramble, int W, int H, int spp, int px, int py, int sampleIdx, int ndim, float *out) {
    Properties props;
    props.setInteger("sampleCount", spp);
    props.setInteger("scramble", (int) scramble);
    Sampler *s = (Sampler *) CreateInstance_sobol(props);
    s->configure();
    s->setFilmResolution(Vector2i(W, H), true);
    s->generate(Point2i(px, py));
    for (int j = 0; j < sampleIdx; ++j) s->advance();
    int i = 0;
    if (ndim >= 2) { const Point2 p = s->next2D(); out[0] = p.x; out[1] = p.y; i = 2; }
    for (; i < ndim; ++i) out[i] = s->next1D();
}
}
