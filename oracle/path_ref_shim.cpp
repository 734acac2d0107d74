/* A minimal Mitsuba 0.6 `path` renderer assembled from the REFERENCE's own sources, compiled where they lie under /root/reference
 * (never copied) into oracle/_ref/libpathref.so by oracle/Makefile, behind the stand-in headers of oracle/shim_core/:
 *   src/integrators/path/path.cpp (MIPathTracer::Li) + src/librender/integrator.cpp (SamplingIntegrator::renderBlock)
 *   src/librender/{scene,skdtree,trimesh,shape,emitter,sensor,film,bsdf,texture,medium,phase,subsurface,sampler,intersection,
 *                  imageblock,shader}.cpp with gkdtree.h / sahkdtree3.h / triaccel.h / records.inl
 *   src/sensors/perspective.cpp, src/emitters/area.cpp, src/samplers/{sobol,independent}.cpp + sobolseq.cpp,
 *   src/rfilters/{gaussian,box}.cpp, the nine BSDF plugins of src/bsdfs/, src/libcore/{util,warp,math,quad,qmc,triangle,transform,
 *   aabb,rfilter,random,spectrum,timer}.cpp
 *   src/integrators/path/volpath.cpp, src/medium/{heterogeneous,homogeneous}.cpp, src/volume/{gridvolume,constvolume}.cpp,
 *   src/phase/{isotropic,hg}.cpp, src/sensors/thinlens.cpp, src/emitters/constant.cpp, src/shapes/{shapegroup,instance}.cpp
 * The scene graph is built the way the XML loader would build it (plugins through their own CreateInstance + Properties, addChild,
 * configure, Scene::initialize builds the SAH kd-tree) and rendered block by block with SamplingIntegrator::renderBlock.
 *
 * Scaffolding (not the reference): everything listed under "scaffolding" below -- the threading / scheduling / serialization /
 * plugin-manager / OpenGL layers are reduced to no-ops, Bitmap is a plain float buffer, the film is a stand-in that only carries
 * the resolution and the reconstruction filter; blocks are accumulated into the output in block order by this file.
 * Used only to pin the oracle at image level (tests/gen_golden.py -> tests/golden/path_ref.npz). */
#include <mitsuba/render/scene.h>
#include <mitsuba/render/renderproc.h>
#include <mitsuba/render/renderqueue.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/sched.h>
#include <mitsuba/core/statistics.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/fstream.h>
#include <mitsuba/render/mipmap.h>
#include <mitsuba/hw/gputexture.h>
#include <mitsuba/hw/basicshader.h>
#include <mitsuba/hw/renderer.h>
#include <mitsuba/core/zstream.h>
#include <mitsuba/core/lock.h>
#include <mitsuba/core/tls.h>
#include <mitsuba/render/medium.h>
#include <mitsuba/render/phase.h>
#include <mitsuba/render/volume.h>
#include "orc_sampler.h" /* orc::CounterSampler: this repository's counter-based `independent` stream, served to the reference integrator */

namespace mitsuba {
/* ---------------------------------------------- scaffolding ---------------------------------------------- */
ConfigurableObject::ConfigurableObject(Stream *, InstanceManager *) {}
void ConfigurableObject::setParent(ConfigurableObject *) {}
void ConfigurableObject::addChild(const std::string &, ConfigurableObject *) {}
void ConfigurableObject::configure() {}
void ConfigurableObject::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(ConfigurableObject, true, SerializableObject)
SerializableObject::SerializableObject(Stream *, InstanceManager *) {}
MTS_IMPLEMENT_CLASS(SerializableObject, true, Object)
void NetworkedObject::bindUsedResources(ParallelProcess *) const {}
void NetworkedObject::wakeup(ConfigurableObject *, std::map<std::string, SerializableObject *> &) {}
void NetworkedObject::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(NetworkedObject, true, ConfigurableObject)
MTS_IMPLEMENT_CLASS(WorkResult, true, Object)
SerializableObject *InstanceManager::getInstance(Stream *) { return NULL; }
void InstanceManager::serialize(Stream *, const SerializableObject *) {}
ref<PluginManager> PluginManager::m_instance;
/* the only plugin-manager requests on this path: ThinLens::createShape (thinlens.cpp:522-534) asks for a `disk` and
   ConstantBackgroundEmitter::createShape (constant.cpp:67-92) for a `sphere`, both called by Scene::initializeBidirectional to represent
   the aperture / the environment to the bidirectional integrators (Scene::rayIntersectAll); the path tracer never intersects these
   "special shapes", so an empty stand-in shape is handed back (the bounding-sphere set-up in constant.cpp's createShape is the
   reference's own code and runs) */
class StandinApertureShape : public Shape {
public:
    StandinApertureShape(const Properties &p) : Shape(p) {}
    void configure() {}
    AABB getAABB() const { return AABB(); }
    size_t getPrimitiveCount() const { return 0; }
    size_t getEffectivePrimitiveCount() const { return 0; }
    const Class *getClass() const { return Shape::m_theClass; }
};
extern "C" void *CreateInstance_scale(const Properties &props);
ConfigurableObject *PluginManager::createObject(const Class *, const Properties &props) {
    /* `scale`: the ScalingTexture BSDF::ensureEnergyConservation wraps around a texture whose maximum exceeds 1 (bsdf.cpp:88-111): the
       reference's own src/textures/scale.cpp */
    if (props.getPluginName() == "scale") return (ConfigurableObject *) (Texture *) CreateInstance_scale(props);
    return props.getPluginName() == "disk" || props.getPluginName() == "sphere" ? new StandinApertureShape(props) : NULL;
}
ref<Scheduler> Scheduler::m_scheduler;
SerializableObject *Scheduler::getResource(int, int) { return NULL; }
int Scheduler::registerResource(SerializableObject *) { return 0; }
bool Scheduler::unregisterResource(int) { return true; }
bool Scheduler::wait(const ParallelProcess *) { return true; }
bool Scheduler::cancel(ParallelProcess *, bool) { return true; }
bool Scheduler::schedule(ParallelProcess *) { return true; }
size_t Scheduler::getCoreCount() const { return 1; }
Float RenderQueue::getRenderTime(const RenderJob *) const { return 0; }
void RenderQueue::signalRefresh(const RenderJob *) {} /* GUI notification after Film::setBitmap: no listeners here */
/* single-threaded: locks and condition variables do nothing, "thread-local" storage is one object */
struct Mutex::MutexPrivate {};
Mutex::Mutex() {}
Mutex::~Mutex() {}
MTS_IMPLEMENT_CLASS(Mutex, false, Object)
void Mutex::lock() {}
void Mutex::unlock() {}
struct ConditionVariable::ConditionVariablePrivate {};
ConditionVariable::ConditionVariable(Mutex *) {}
ConditionVariable::~ConditionVariable() {}
MTS_IMPLEMENT_CLASS(ConditionVariable, false, Object)
void ConditionVariable::wait() {}
void ConditionVariable::signal() {}
void ConditionVariable::broadcast() {}
namespace detail {
struct ThreadLocalBase::ThreadLocalPrivate { ConstructFunctor construct; void *value; };
ThreadLocalBase::ThreadLocalBase(const ConstructFunctor &c, const DestructFunctor &) : d(new ThreadLocalPrivate()) { d->construct = c; d->value = NULL; }
ThreadLocalBase::~ThreadLocalBase() {}
void *ThreadLocalBase::get(bool &existed) { existed = d->value != NULL; if (!existed) d->value = d->construct(); return d->value; }
}
StatsCounter::StatsCounter(const std::string &, const std::string &, EStatsType, uint64_t, uint64_t) {}
StatsCounter::~StatsCounter() {}
Shader *Renderer::registerShaderForResource(const HWResource *) { return NULL; }
void Renderer::unregisterShaderForResource(const HWResource *) {}
ConstantSpectrumTexture::ConstantSpectrumTexture(Stream *stream, InstanceManager *manager) : Texture(stream, manager) {}
Shader *ConstantSpectrumTexture::createShader(Renderer *) const { return NULL; }
ref<Bitmap> ConstantSpectrumTexture::getBitmap(const Vector2i &) const { return NULL; }
void ConstantSpectrumTexture::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(ConstantSpectrumTexture, false, Texture)
ConstantFloatTexture::ConstantFloatTexture(Stream *stream, InstanceManager *manager) : Texture(stream, manager) {}
Shader *ConstantFloatTexture::createShader(Renderer *) const { return NULL; }
ref<Bitmap> ConstantFloatTexture::getBitmap(const Vector2i &) const { return NULL; }
void ConstantFloatTexture::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(ConstantFloatTexture, false, Texture)
/* a static transform only (src/libcore/track.cpp needs Eigen) */
AnimatedTransform::AnimatedTransform(Stream *) {}
AnimatedTransform::AnimatedTransform(const AnimatedTransform *trafo) : m_transform(trafo->m_transform) {} /* track.cpp:8-16 without tracks */
void AnimatedTransform::TransformFunctor::operator()(const Float &, Transform &) const {}
AABB AnimatedTransform::getTranslationBounds() const { AABB b; b.expandBy(m_transform(Point(0.0f))); return b; }
void AnimatedTransform::serialize(Stream *) const {}
std::string AnimatedTransform::toString() const { return "AnimatedTransform"; }
AnimatedTransform::~AnimatedTransform() {}
/* static transforms only (no animation tracks): the no-track branches of track.cpp:123-129 and :254-264 */
AABB AnimatedTransform::getSpatialBounds(const AABB &aabb) const { AABB r; for (int j = 0; j < 8; ++j) r.expandBy(m_transform(aabb.getCorner(j))); return r; }
void AnimatedTransform::prependScale(const Vector &scale) { m_transform = m_transform * Transform::scale(scale); } /* track.cpp:222-223; only ThinLens::createShape (bidirectional set-up, never called here) uses it */
void AnimatedTransform::collectKeyframes(std::set<Float> &result) const { result.insert((Float) 0); }
MTS_IMPLEMENT_CLASS(AnimatedTransform, false, Object)
ref<const AnimatedTransform> Properties::getAnimatedTransform(const std::string &k, const Transform &d) const { return new AnimatedTransform(getTransform(k, d)); }
ref<const AnimatedTransform> Properties::getAnimatedTransform(const std::string &k, const AnimatedTransform *d) const { return hasProperty(k) || !d ? new AnimatedTransform(getTransform(k, Transform())) : d; }
ref<const AnimatedTransform> Properties::getAnimatedTransform(const std::string &k) const { return new AnimatedTransform(getTransform(k)); }
std::ostream &operator<<(std::ostream &os, const ETransportMode &m) { return os << (int) m; }
/* envmap.cpp's image-file path (constructor else-branch, serialize) and its OpenGL shader: never reached through the cache-file route
   (FileStream itself is the reference's own src/libcore/fstream.cpp) */
void GPUTexture::initAndRelease() {}
std::ostream &operator<<(std::ostream &os, const Bitmap::EPixelFormat &v) { return os << (int) v; } /* TMIPMap::toString only */
std::ostream &operator<<(std::ostream &os, const Bitmap::EComponentFormat &v) { return os << (int) v; }
Bitmap::Bitmap(EFileFormat, Stream *, const std::string &) { throw std::runtime_error("Bitmap: image files are not readable here"); }
void Bitmap::scale(Float) { throw std::runtime_error("Bitmap::scale: not available"); } /* ScalingTexture::getBitmap only */
ref<Bitmap> Bitmap::extractChannel(int) const { throw std::runtime_error("Bitmap::extractChannel: not available"); }
std::string Bitmap::getChannelName(int) const { return ""; }
ref<Bitmap> Bitmap::expand() { throw std::runtime_error("Bitmap::expand: not available"); }
ref<Bitmap> Bitmap::convert(EPixelFormat, EComponentFormat, Float, Float, Spectrum::EConversionIntent) { throw std::runtime_error("Bitmap::convert: not available"); }
ref<Bitmap> Bitmap::resample(const ReconstructionFilter *, ReconstructionFilter::EBoundaryCondition, ReconstructionFilter::EBoundaryCondition, const Vector2i &, Float, Float) const { throw std::runtime_error("Bitmap::resample: not available"); }
void Bitmap::write(EFileFormat, Stream *, int) const { throw std::runtime_error("Bitmap::write: not available"); }
/* Bitmap: a zeroed float buffer (src/libcore/bitmap.cpp needs OpenEXR / libpng / libjpeg) */
Bitmap::Bitmap(EPixelFormat pFmt, EComponentFormat cFmt, const Vector2i &size, uint8_t channelCount, uint8_t *)
    : m_pixelFormat(pFmt), m_componentFormat(cFmt), m_size(size), m_data(NULL), m_gamma(1.0f), m_channelCount(channelCount), m_ownsData(true) {
    /* Bitmap::updateChannelCount (bitmap.cpp): the formats that occur here */
    switch (pFmt) {
        case ELuminance: m_channelCount = 1; break;
        case ELuminanceAlpha: m_channelCount = 2; break;
        case ERGB: case EXYZ: m_channelCount = 3; break;
        case ERGBA: case EXYZA: m_channelCount = 4; break;
        case ESpectrum: m_channelCount = SPECTRUM_SAMPLES; break;
        case ESpectrumAlpha: m_channelCount = SPECTRUM_SAMPLES + 1; break;
        case ESpectrumAlphaWeight: m_channelCount = SPECTRUM_SAMPLES + 2; break;
        default: break; /* EMultiChannel: as passed */
    }
    m_data = (uint8_t *) calloc((size_t) size.x * size.y * m_channelCount, sizeof(float));
}
Bitmap::~Bitmap() { if (m_data && m_ownsData) free(m_data); }
void Bitmap::clear() { memset(m_data, 0, (size_t) m_size.x * m_size.y * m_channelCount * sizeof(float)); }
std::string Bitmap::toString() const { return "Bitmap"; }
MTS_IMPLEMENT_CLASS(Bitmap, false, Object)
/* Stream: nothing is (de)serialised */
float Stream::readSingle() { float v = 0; read(&v, sizeof(v)); return v; } /* little-endian host = the .vol byte order */ double Stream::readDouble() { return 0; } void Stream::writeSingle(float) {} void Stream::writeDouble(double) {}
int Stream::readInt() { int v = 0; read(&v, sizeof(v)); return v; } void Stream::writeInt(int) {} float Stream::readFloat() { float v = 0; read(&v, sizeof(v)); return v; } void Stream::writeFloat(float) {}
void Stream::readFloatArray(float *, size_t) {} void Stream::writeFloatArray(const float *, size_t) {}
std::string Stream::readString() { return ""; } void Stream::writeString(const std::string &) {} void Stream::read(void *, size_t) {} void Stream::write(const void *, size_t) {}
void Stream::readSingleArray(float *, size_t) {} void Stream::writeSingleArray(const float *, size_t) {}
void Stream::readUIntArray(unsigned int *, size_t) {} void Stream::writeUIntArray(const unsigned int *, size_t) {}
void Stream::readDoubleArray(double *, size_t) {} void Stream::writeDoubleArray(const double *, size_t) {}
void Stream::seek(size_t) {} void Stream::setByteOrder(int) {} void Stream::copyTo(Stream *, long long) {}
void Stream::readULongArray(uint64_t *, size_t) {} void Stream::writeULongArray(const uint64_t *, size_t) {}
unsigned int Stream::readUInt() { return 0; } void Stream::writeUInt(unsigned int) {} size_t Stream::readSize() { return 0; } void Stream::writeSize(size_t) {}
bool Stream::readBool() { return false; } void Stream::writeBool(bool) {} short Stream::readShort() { return 0; } void Stream::writeShort(short) {}
long long Stream::readLong() { return 0; } void Stream::writeLong(long long) {} unsigned long long Stream::readULong() { return 0; } void Stream::writeULong(unsigned long long) {}
void Stream::skip(size_t n) { seek(getPos() + n); } void Stream::flush() {} void Stream::truncate(size_t) {} bool Stream::canRead() const { return true; } bool Stream::canWrite() const { return false; }
size_t Stream::getPos() const { return 0; }
size_t Stream::getSize() const { return 0; }
template <typename T> void Stream::readArray(T *, size_t) {}
template <typename T> void Stream::writeArray(const T *, size_t) {}
template <typename T> T Stream::readElement() { return T(); }
template <typename T> void Stream::writeElement(T) {}
template void Stream::readArray<float>(float *, size_t); template void Stream::writeArray<float>(const float *, size_t);
template void Stream::readArray<int>(int *, size_t); template void Stream::writeArray<int>(const int *, size_t);
template void Stream::readArray<unsigned int>(unsigned int *, size_t); template void Stream::writeArray<unsigned int>(const unsigned int *, size_t);
template void Stream::readArray<unsigned long>(unsigned long *, size_t); template void Stream::writeArray<unsigned long>(const unsigned long *, size_t);
template void Stream::readArray<double>(double *, size_t); template void Stream::writeArray<double>(const double *, size_t);
template int Stream::readElement<int>(); template void Stream::writeElement<int>(int);
template float Stream::readElement<float>(); template void Stream::writeElement<float>(float);
ZStream::ZStream(Stream *child, EStreamType, int) : m_childStream(child) {}
ZStream::~ZStream() {}
std::string ZStream::toString() const { return "ZStream"; }
void ZStream::read(void *, size_t) {} void ZStream::write(const void *, size_t) {} void ZStream::seek(size_t) {} size_t ZStream::getPos() const { return 0; }
size_t ZStream::getSize() const { return 0; } void ZStream::truncate(size_t) {} void ZStream::flush() {} bool ZStream::canWrite() const { return false; } bool ZStream::canRead() const { return false; }
MTS_IMPLEMENT_CLASS(ZStream, false, Stream)
}

/* SamplingIntegrator::render() (integrator.cpp:96-128, never called here: the blocks are rendered directly) constructs a
   BlockedRenderProcess; its constructor symbol is satisfied without dragging in the scheduler classes */
extern "C" void _ZN7mitsuba20BlockedRenderProcessC1EPKNS_9RenderJobEPNS_11RenderQueueEi() { abort(); }

using namespace mitsuba;

#define DECL(name) extern "C" void *CreateInstance_##name(const Properties &props);
DECL(diffuse) DECL(roughconductor) DECL(roughdielectric) DECL(coating) DECL(dielectric) DECL(conductor) DECL(plastic) DECL(twosided) DECL(null)
DECL(gaussian) DECL(box) DECL(sobol) DECL(independent) DECL(path) DECL(perspective) DECL(area)
DECL(thinlens) DECL(constant) DECL(shapegroup) DECL(instance) DECL(envmap) DECL(bitmap)
DECL(volpath) DECL(heterogeneous) DECL(homogeneous) DECL(gridvolume) DECL(constvolume) DECL(isotropic) DECL(hg)

/* The `independent` sampler of this repository is a counter-based stream, not the reference's SFMT: to compare volpath
 * images sample for sample the reference integrator is handed that stream through the real Sampler interface. */
class CounterSamplerPlugin : public Sampler {
public:
    CounterSamplerPlugin(int W, size_t spp, uint64_t seed) : Sampler(Properties()), m_impl(W, (uint32_t) spp, seed) { m_sampleCount = spp; }
    void generate(const Point2i &pos) { m_impl.generate(pos.x, pos.y); m_sampleIndex = 0; }
    void advance() { m_impl.advance(); ++m_sampleIndex; }
    Float next1D() { return m_impl.next1D(); }
    Point2 next2D() { float a, b; m_impl.next2D(a, b); return Point2(a, b); }
    ref<Sampler> clone() { return this; }
    void setSampleIndex(size_t) {}
    std::string toString() const { return "CounterSampler"; }
    const Class *getClass() const { return Sampler::m_theClass; }
private:
    orc::CounterSampler m_impl;
};

/* carries resolution + reconstruction filter to the sensor / integrator (hdrfilm.cpp needs the Bitmap file writers) */
class StandinFilm : public Film {
public:
    StandinFilm(const Properties &props) : Film(props) {}
    void clear() {}
    void put(const ImageBlock *) {}
    void setBitmap(const Bitmap *bitmap, Float) { m_last = const_cast<Bitmap *>(bitmap); } /* an integrator that renders the whole film (b200path) hands it over here */
    ref<Bitmap> m_last;
    void addBitmap(const Bitmap *, Float) {}
    void setDestinationFile(const fs::path &, uint32_t) {}
    void develop(const Scene *, Float) {}
    bool develop(const Point2i &, const Vector2i &, const Point2i &, Bitmap *) const { return false; }
    bool destinationExists(const fs::path &) const { return false; }
    bool hasAlpha() const { return true; } /* the `rgba` film: RadianceQueryRecord::EOpacity stays set (integrator.cpp:160-161) */
    std::string toString() const { return "StandinFilm"; }
    const Class *getClass() const { return Film::m_theClass; }
};

struct PathRef {
    ref<Scene> scene;
    ref<Sensor> sensor;
    ref<Sampler> sampler;
    ref<Integrator> integrator;
    ref<Film> film;
    int W, H;
    std::vector<Object *> keep;
};

namespace {
template <typename MIP> struct MipCacheWriter : public MIP { typedef typename MIP::MIPMapHeader Header; typedef typename MIP::Array2DType Array; };
template <typename MIP, typename Value> bool writeMipCache(const std::string &mip, uint64_t timestamp, uint8_t pixelFormat, int nLevels, const int *sizes, const float *const *levels,
                                                          const float *original, int bcu, int bcv, int filterType) {
    typedef MipCacheWriter<MIP> W;
    size_t padding = sizeof(typename W::Header) % MTS_MIPMAP_CACHE_ALIGNMENT;
    if (padding) padding = MTS_MIPMAP_CACHE_ALIGNMENT - padding;
    size_t total = sizeof(typename W::Header) + padding;
    for (int l = 0; l < nLevels; ++l) total += W::Array::bufferSize(Vector2i(sizes[2 * l], sizes[2 * l + 1]));
    uint8_t *buf = (uint8_t *) allocAligned(total);
    memset(buf, 0, total);
    typename W::Header header;
    memset(&header, 0, sizeof(header));
    memcpy(header.identifier, "MIP", 3);
    header.version = MTS_MIPMAP_CACHE_VERSION;
    header.pixelFormat = pixelFormat;
    header.levels = (uint8_t) nLevels;
    header.bcu = (uint8_t) bcu; header.bcv = (uint8_t) bcv;
    header.filterType = (uint8_t) filterType;
    header.gamma = 1.0f;
    header.width = sizes[0]; header.height = sizes[1];
    header.timestamp = timestamp;
    uint8_t *ptr = buf + sizeof(typename W::Header) + padding;
    for (int l = 0; l < nLevels; ++l) {
        typename W::Array a;
        a.map(ptr, Vector2i(sizes[2 * l], sizes[2 * l + 1]));
        if (l == 0 && original) { Value mn, mx, avg; a.init((const Value *) original, mn, mx, avg); header.minimum = mn; header.maximum = mx; header.average = avg; }
        a.init((const Value *) levels[l]);
        ptr += a.getBufferSize();
    }
    memcpy(buf, &header, sizeof(header));
    FILE *f = fopen(mip.c_str(), "wb");
    if (!f) { freeAligned(buf); return false; }
    fwrite(buf, 1, total, f); fclose(f);
    freeAligned(buf);
    return true;
}
}
namespace mitsuba { extern const Float CIE_wavelengths[471]; extern const Float CIE_X_entries[471]; extern const Float CIE_Y_entries[471]; extern const Float CIE_Z_entries[471]; }
extern "C" {
void *pathref_new() {
    PathRef *p = new PathRef();
    p->scene = new Scene(Properties("scene"));
    return p;
}
/* same layout as bsdfref_create (oracle/bsdf_ref_shim.cpp) */
void *pathref_bsdf2(int plugin, int nf, const char **fk, const float *fv, int ns, const char **sk, const char **sv, int nb, const char **bk, const int *bv,
                    int nsp, const char **spk, const float *spv, void *child, void *child2, void *texture, const char *textureName);
void *pathref_bsdf(int plugin, int nf, const char **fk, const float *fv, int ns, const char **sk, const char **sv, int nb, const char **bk, const int *bv,
                   int nsp, const char **spk, const float *spv, void *child, void *child2) {
    return pathref_bsdf2(plugin, nf, fk, fv, ns, sk, sv, nb, bk, bv, nsp, spk, spv, child, child2, NULL, NULL);
}
/* the same with a Texture child named like the parameter it binds to (<texture name="reflectance" type="bitmap">: BSDF::addChild) */
void *pathref_bsdf2(int plugin, int nf, const char **fk, const float *fv, int ns, const char **sk, const char **sv, int nb, const char **bk, const int *bv,
                    int nsp, const char **spk, const float *spv, void *child, void *child2, void *texture, const char *textureName) {
    static const char *const pluginNames[] = {"diffuse", "roughconductor", "roughdielectric", "coating", "null", "twosided", "dielectric", "conductor", "plastic"};
    Properties props(plugin >= 0 && plugin < 9 ? pluginNames[plugin] : ""); /* the XML loader names the plugin in the Properties (scenehandler.cpp) */
    for (int i = 0; i < nf; ++i) props.setFloat(fk[i], fv[i]);
    for (int i = 0; i < ns; ++i) props.setString(sk[i], sv[i]);
    for (int i = 0; i < nb; ++i) props.setBoolean(bk[i], bv[i] != 0);
    for (int i = 0; i < nsp; ++i) { Spectrum s; s[0] = spv[3 * i]; s[1] = spv[3 * i + 1]; s[2] = spv[3 * i + 2]; props.setSpectrum(spk[i], s); }
    BSDF *b = NULL;
    switch (plugin) {
        case 0: b = (BSDF *) CreateInstance_diffuse(props); break;
        case 1: b = (BSDF *) CreateInstance_roughconductor(props); break;
        case 2: b = (BSDF *) CreateInstance_roughdielectric(props); break;
        case 3: b = (BSDF *) CreateInstance_coating(props); break;
        case 4: b = (BSDF *) CreateInstance_null(props); break;
        case 5: b = (BSDF *) CreateInstance_twosided(props); break;
        case 6: b = (BSDF *) CreateInstance_dielectric(props); break;
        case 7: b = (BSDF *) CreateInstance_conductor(props); break;
        case 8: b = (BSDF *) CreateInstance_plastic(props); break;
    }
    if (!b) return NULL;
    if (child) b->addChild("", (ConfigurableObject *) (BSDF *) child);
    if (child2) b->addChild("", (ConfigurableObject *) (BSDF *) child2);
    if (texture) b->addChild(textureName, (ConfigurableObject *) (Texture *) texture);
    b->configure();
    return b;
}
/* phase 0 isotropic / 1 hg(g) */
static PhaseFunction *makePhase(int phase, float g) {
    Properties pp(phase == 1 ? "hg" : "isotropic");
    if (phase == 1) pp.setFloat("g", g);
    PhaseFunction *ph = (PhaseFunction *) (phase == 1 ? CreateInstance_hg(pp) : CreateInstance_isotropic(pp));
    ph->configure();
    return ph;
}
/* `homogeneous` medium (src/medium/homogeneous.cpp): sigmaA / sigmaS RGB; strategy "balance" | "single" | "manual" */
void *pathref_medium_homogeneous(const float *sigmaA, const float *sigmaS, const char *strategy, float samplingDensity, float mediumSamplingWeight, int phase, float g) {
    Properties mp("homogeneous");
    Spectrum a, sc;
    for (int i = 0; i < 3; ++i) { a[i] = sigmaA[i]; sc[i] = sigmaS[i]; }
    mp.setSpectrum("sigmaA", a); mp.setSpectrum("sigmaS", sc);
    mp.setString("strategy", strategy);
    if (samplingDensity > 0) mp.setFloat("samplingDensity", samplingDensity);
    if (mediumSamplingWeight >= 0) mp.setFloat("mediumSamplingWeight", mediumSamplingWeight);
    Medium *m = (Medium *) CreateInstance_homogeneous(mp);
    m->addChild("", makePhase(phase, g));
    m->configure();
    return m;
}
/* `heterogeneous` medium, method woodcock (src/medium/heterogeneous.cpp): density = `gridvolume` read from a .vol file with an
 * optional toWorld (row-major 4 x 4, NULL = identity), albedo = `constvolume`; scale */
void *pathref_medium_heterogeneous(const char *volFile, const float *toWorld, const float *albedo, float scale, int phase, float g) {
    Properties gp("gridvolume");
    gp.setString("filename", volFile);
    if (toWorld) { Matrix4x4 M; for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M.m[r][c] = toWorld[4 * r + c]; gp.setTransform("toWorld", Transform(M)); }
    VolumeDataSource *density = (VolumeDataSource *) CreateInstance_gridvolume(gp);
    density->configure();
    Properties cp("constvolume");
    Spectrum al; for (int i = 0; i < 3; ++i) al[i] = albedo[i];
    cp.setSpectrum("value", al);
    VolumeDataSource *alb = (VolumeDataSource *) CreateInstance_constvolume(cp);
    alb->configure();
    Properties mp("heterogeneous");
    mp.setString("method", "woodcock");
    mp.setFloat("scale", scale);
    Medium *m = (Medium *) CreateInstance_heterogeneous(mp);
    m->addChild("density", density);
    m->addChild("albedo", alb);
    m->addChild("", makePhase(phase, g));
    m->configure();
    return m;
}
/* a TriMesh as the loaders leave it (positions, optional normals / texcoords, triangles), its BSDF, optionally an area emitter */
void pathref_add_mesh_media(void *h, const float *P, const float *N, const float *UV, int nV, const uint32_t *idx, int nT, void *bsdf, const float *radiance, float samplingWeight,
                            void *interior, void *exterior);
void pathref_add_mesh(void *h, const float *P, const float *N, const float *UV, int nV, const uint32_t *idx, int nT, void *bsdf, const float *radiance, float samplingWeight) {
    pathref_add_mesh_media(h, P, N, UV, nV, idx, nT, bsdf, radiance, samplingWeight, NULL, NULL);
}
/* bsdf may be NULL for an index-matched medium boundary: Shape::configure then creates the `null` BSDF itself (shape.cpp:64-68) through
   the plugin manager, which this build does not have -- so the `null` plugin is attached here */
static TriMesh *makeMesh(const float *P, const float *N, const float *UV, int nV, const uint32_t *idx, int nT, void *bsdf, const float *radiance, float samplingWeight,
                         void *interior, void *exterior) {
    TriMesh *mesh = new TriMesh("mesh", (size_t) nT, (size_t) nV, N != NULL, UV != NULL, false, false, N == NULL /* face normals, skdtree.h:383-399 */);
    memcpy(mesh->getVertexPositions(), P, sizeof(float) * 3 * nV);
    if (N) memcpy(mesh->getVertexNormals(), N, sizeof(float) * 3 * nV);
    if (UV) memcpy(mesh->getVertexTexcoords(), UV, sizeof(float) * 2 * nV);
    memcpy(mesh->getTriangles(), idx, sizeof(uint32_t) * 3 * nT);
    if (!bsdf) { Properties np("null"); BSDF *nb = (BSDF *) CreateInstance_null(np); nb->configure(); bsdf = nb; }
    mesh->addChild("", (ConfigurableObject *) (BSDF *) bsdf);
    if (interior) mesh->addChild("interior", (ConfigurableObject *) (Medium *) interior);
    if (exterior) mesh->addChild("exterior", (ConfigurableObject *) (Medium *) exterior);
    if (radiance) {
        Properties ep("area");
        Spectrum s; s[0] = radiance[0]; s[1] = radiance[1]; s[2] = radiance[2];
        ep.setSpectrum("radiance", s);
        ep.setFloat("samplingWeight", samplingWeight);
        Emitter *em = (Emitter *) CreateInstance_area(ep);
        mesh->addChild("", em);
        em->setParent(mesh);      /* the scene loader calls setParent() after addChild() (scenehandler.cpp); AreaLight keeps the shape */
        em->configure();
    }
    mesh->configure();
    return mesh;
}
void pathref_add_mesh_media(void *h, const float *P, const float *N, const float *UV, int nV, const uint32_t *idx, int nT, void *bsdf, const float *radiance, float samplingWeight,
                            void *interior, void *exterior) {
    PathRef *p = (PathRef *) h;
    ref<TriMesh> mesh = makeMesh(P, N, UV, nV, idx, nT, bsdf, radiance, samplingWeight, interior, exterior);
    p->scene->addChild("", mesh);
    p->keep.push_back(mesh);
}
/* <shape type="shapegroup"> (src/shapes/shapegroup.cpp): the member meshes (object space) are added with ShapeGroup::addChild and
 * ShapeGroup::configure builds the group's own SAH kd-tree; the group itself is a compound of nothing (Scene::addChild would add no
 * element), so it is only kept alive here */
void *pathref_shapegroup_new(void *h) {
    PathRef *p = (PathRef *) h;
    Shape *g = (Shape *) CreateInstance_shapegroup(Properties("shapegroup"));
    g->incRef();
    p->keep.push_back(g);
    return g;
}
void pathref_shapegroup_add_mesh(void *group, const float *P, const float *N, const float *UV, int nV, const uint32_t *idx, int nT, void *bsdf) {
    Shape *g = (Shape *) group;
    TriMesh *mesh = makeMesh(P, N, UV, nV, idx, nT, bsdf, NULL, 1.0f, NULL, NULL);
    mesh->incRef();
    g->addChild("", mesh);
}
void pathref_shapegroup_configure(void *group) { ((Shape *) group)->configure(); }
/* <shape type="instance"> (src/shapes/instance.cpp) with a `toWorld` (row-major 4 x 4) referencing a configured shapegroup */
void pathref_add_instance(void *h, void *group, const float *toWorld) {
    PathRef *p = (PathRef *) h;
    Properties ip("instance");
    Matrix4x4 M;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M.m[r][c] = toWorld[4 * r + c];
    ip.setTransform("toWorld", Transform(M));
    ref<Shape> inst = (Shape *) CreateInstance_instance(ip);
    inst->addChild("", (ConfigurableObject *) (Shape *) group);
    inst->configure();
    p->scene->addChild("", inst);
    p->keep.push_back(inst);
}
/* the inverse the reference derives for a `toWorld` (Transform::Transform(const Matrix4x4 &) -> Matrix4x4::invert, transform.h / matrix.h):
 * matrix set-up is host work, the oracle is handed the same numbers (row-major 4 x 4) */
void pathref_transform_inverse(const float *toWorld, float *out16) {
    Matrix4x4 M;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M.m[r][c] = toWorld[4 * r + c];
    const Matrix4x4 &I = Transform(M).getInverseMatrix();
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out16[4 * r + c] = I.m[r][c];
}
/* <emitter type="constant"> (src/emitters/constant.cpp) */
void pathref_add_constant_emitter(void *h, const float *radiance, float samplingWeight) {
    PathRef *p = (PathRef *) h;
    Properties ep("constant");
    Spectrum s; s[0] = radiance[0]; s[1] = radiance[1]; s[2] = radiance[2];
    ep.setSpectrum("radiance", s);
    ep.setFloat("samplingWeight", samplingWeight);
    ref<Emitter> em = (Emitter *) CreateInstance_constant(ep);
    em->configure();
    p->scene->addChild("", em);
    p->keep.push_back(em);
}
/* <texture type="bitmap"> (src/textures/bitmap.cpp) the way a second Mitsuba run constructs it: from the MIP map cache file next to the image
 * (bitmap.cpp:236-244 -> mipmap.h:320-380).  `levels`: the pyramid (float, `channels` per texel: 1 = luminance, 3 = RGB; one level for the
 * nearest / bilinear filters); `original`: the decoded image before the half-precision rounding -- the cache header carries its minimum,
 * maximum and average (BlockedArray::init, barray.h:102-124), which getMaximum() / getAverage() later hand to the BSDFs.  Layout, float ->
 * half conversion and the statistics are the reference's own code; every look-up afterwards is BitmapTexture's. */
void *pathref_bitmap_texture(const char *stem, int channels, int nLevels, const int *sizes, const float *const *levels, const float *original,
                             const char *filterType, const char *wrapU, const char *wrapV, float maxAnisotropy, float uoffset, float voffset, float uscale, float vscale) {
    const std::string img = std::string(stem) + ".img", mip = std::string(stem) + ".mip";
    { FILE *f = fopen(img.c_str(), "wb"); if (!f) return NULL; fputs("placeholder for the decoded image handed in as a MIP map cache\n", f); fclose(f); }
    boost::system::error_code ec;
    const uint64_t timestamp = (uint64_t) fs::last_write_time(fs::path(img), ec);
    auto wrapOf = [](const std::string &m) { return m == "repeat" ? ReconstructionFilter::ERepeat : m == "clamp" ? ReconstructionFilter::EClamp : m == "mirror" ? ReconstructionFilter::EMirror
                                                   : (m == "zero" || m == "black") ? ReconstructionFilter::EZero : ReconstructionFilter::EOne; };
    const std::string ft(filterType);
    const int filt = ft == "ewa" ? (int) EEWA : ft == "trilinear" ? (int) ETrilinear : ft == "bilinear" ? (int) EBilinear : (int) ENearest;
    typedef TSpectrum<Float, 1> C1; typedef TSpectrum<Float, 3> C3;
    const bool ok = channels == 3
        ? writeMipCache<TMIPMap<C3, TSpectrum<half, 3> >, C3>(mip, timestamp, (uint8_t) Bitmap::ERGB, nLevels, sizes, levels, original, wrapOf(wrapU), wrapOf(wrapV), filt)
        : writeMipCache<TMIPMap<C1, TSpectrum<half, 1> >, C1>(mip, timestamp, (uint8_t) Bitmap::ELuminance, nLevels, sizes, levels, original, wrapOf(wrapU), wrapOf(wrapV), filt);
    if (!ok) return NULL;
    Properties tp("bitmap");
    tp.setString("filename", img);
    tp.setString("filterType", filterType);
    tp.setString("wrapModeU", wrapU); tp.setString("wrapModeV", wrapV);
    tp.setFloat("maxAnisotropy", maxAnisotropy);
    tp.setFloat("uoffset", uoffset); tp.setFloat("voffset", voffset); tp.setFloat("uscale", uscale); tp.setFloat("vscale", vscale);
    Texture *t = (Texture *) CreateInstance_bitmap(tp);
    t->configure();
    return t;
}
/* <emitter type="envmap"> (src/emitters/envmap.cpp).  The image file readers (OpenEXR, RGBE, ...) are not here, so the pyramid enters the
 * real class the way a second Mitsuba run reads it: through its MIP map cache file (envmap.cpp:143-147 -> mipmap.h:320-380).  The caller
 * hands in the pyramid levels (float RGB, row-major, level l of size sizes[2l] x sizes[2l+1]); they are written next to a placeholder image
 * `<stem>.img` as `<stem>.mip` with the header validateCacheFile (mipmap.h:405-446) expects and the reference's own BlockedArray layout and
 * float -> half conversion (TSpectrum<half>, src/libcore/half.cpp).  CDF tables, look-ups and sampling are then all the reference's. */
namespace {
typedef TMIPMap<Spectrum, TSpectrum<half, SPECTRUM_SAMPLES> > EnvMIP;
struct EnvCacheWriter : public EnvMIP { typedef EnvMIP::MIPMapHeader Header; typedef EnvMIP::Array2DType Array; };
}
int pathref_add_envmap(void *h, const char *stem, int nLevels, const int *sizes, const float *const *levels, const float *toWorld, float scale, float samplingWeight) {
    PathRef *p = (PathRef *) h;
    const std::string img = std::string(stem) + ".img", mip = std::string(stem) + ".mip";
    { FILE *f = fopen(img.c_str(), "wb"); if (!f) return -1; fputs("placeholder for the decoded image handed in as a MIP map cache\n", f); fclose(f); }
    boost::system::error_code ec;
    const uint64_t timestamp = (uint64_t) fs::last_write_time(fs::path(img), ec);
    size_t padding = sizeof(EnvCacheWriter::Header) % MTS_MIPMAP_CACHE_ALIGNMENT;
    if (padding) padding = MTS_MIPMAP_CACHE_ALIGNMENT - padding;
    size_t total = sizeof(EnvCacheWriter::Header) + padding;
    for (int l = 0; l < nLevels; ++l) total += EnvCacheWriter::Array::bufferSize(Vector2i(sizes[2 * l], sizes[2 * l + 1]));
    uint8_t *buf = (uint8_t *) allocAligned(total);
    memset(buf, 0, total);
    EnvCacheWriter::Header header;
    memset(&header, 0, sizeof(header));
    memcpy(header.identifier, "MIP", 3);
    header.version = MTS_MIPMAP_CACHE_VERSION;
    header.pixelFormat = (uint8_t) Bitmap::ERGB;
    header.levels = (uint8_t) nLevels;
    header.bcu = (uint8_t) ReconstructionFilter::ERepeat; header.bcv = (uint8_t) ReconstructionFilter::EClamp;
    header.filterType = (uint8_t) EEWA;
    header.gamma = 1.0f;
    header.width = sizes[0]; header.height = sizes[1];
    header.timestamp = timestamp;
    memcpy(buf, &header, sizeof(header));
    uint8_t *ptr = buf + sizeof(EnvCacheWriter::Header) + padding;
    for (int l = 0; l < nLevels; ++l) {
        EnvCacheWriter::Array a;
        a.map(ptr, Vector2i(sizes[2 * l], sizes[2 * l + 1]));
        a.init((const Spectrum *) levels[l]);
        ptr += a.getBufferSize();
    }
    { FILE *f = fopen(mip.c_str(), "wb"); if (!f) { freeAligned(buf); return -1; } fwrite(buf, 1, total, f); fclose(f); }
    freeAligned(buf);
    Properties ep("envmap");
    ep.setString("filename", img);
    ep.setFloat("scale", scale);
    ep.setFloat("samplingWeight", samplingWeight);
    if (toWorld) {
        Matrix4x4 M;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M.m[r][c] = toWorld[4 * r + c];
        ep.setTransform("toWorld", Transform(M));
    }
    ref<Emitter> em = (Emitter *) CreateInstance_envmap(ep);
    em->configure();
    p->scene->addChild("", em);
    p->keep.push_back(em);
    return 0;
}
/* Emitter::getBitmap of the environment emitter (EnvironmentMap: level 0 of its half-precision pyramid, envmap.cpp:632-634): size in wh,
 * RGB floats in out (may be NULL to query the size).  What the b200path plugin marshals. */
int pathref_environment_bitmap(void *h, int *wh, float *out) {
    PathRef *p = (PathRef *) h;
    const Emitter *e = p->scene->getEnvironmentEmitter();
    if (!e) return -1;
    ref<Bitmap> bm = e->getBitmap(Vector2i(0));
    if (bm == NULL || bm->getPixelFormat() != Bitmap::ERGB || bm->getComponentFormat() != Bitmap::EFloat16) return -2;
    wh[0] = bm->getSize().x; wh[1] = bm->getSize().y;
    if (out) { const half *src = (const half *) bm->getData(); for (size_t k = 0; k < (size_t) wh[0] * wh[1] * 3; ++k) out[k] = (float) src[k]; }
    return 0;
}
/* Scene::evalEnvironment for n rays: rays 6n (o, d) without differentials, or 18n (o, d, rxO, rxD, ryO, ryD) with them -> out 3n */
void pathref_eval_environment(void *h, int n, int withDifferentials, const float *rays, float *out) {
    PathRef *p = (PathRef *) h;
    const int stride = withDifferentials ? 18 : 6;
    for (int i = 0; i < n; ++i) {
        const float *r = rays + (size_t) stride * i;
        RayDifferential ray(Point(r[0], r[1], r[2]), Vector(r[3], r[4], r[5]), 0.0f);
        if (withDifferentials) {
            ray.rxOrigin = Point(r[6], r[7], r[8]); ray.rxDirection = Vector(r[9], r[10], r[11]);
            ray.ryOrigin = Point(r[12], r[13], r[14]); ray.ryDirection = Vector(r[15], r[16], r[17]);
            ray.hasDifferentials = true;
        }
        const Spectrum v = p->scene->evalEnvironment(ray);
        out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2];
    }
}
/* Scene::pdfEmitterDirect for the environment emitter: ref 6n (ref, refN), d 3n -> out n (solid-angle density incl. the emitter choice) */
void pathref_pdf_environment_direct(void *h, int n, const float *ref, const float *d, float *out) {
    PathRef *p = (PathRef *) h;
    for (int i = 0; i < n; ++i) {
        DirectSamplingRecord dRec(Point(ref[6 * i], ref[6 * i + 1], ref[6 * i + 2]), 0.0f);
        dRec.refN = Normal(ref[6 * i + 3], ref[6 * i + 4], ref[6 * i + 5]);
        dRec.d = Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        dRec.measure = ESolidAngle;
        dRec.object = p->scene->getEnvironmentEmitter();
        out[i] = p->scene->getEnvironmentEmitter() ? p->scene->pdfEmitterDirect(dRec) : 0.0f;
    }
}
/* perspective sensor + film + sampler + path integrator; rfilter 0 box / 1 gaussian; sampler 0 sobol / 1 independent */
void pathref_setup2(void *h, const float *toWorld, float fov, float nearClip, float farClip, int W, int H, int rfilter, int samplerKind, int spp, uint64_t scramble,
                    int maxDepth, int rrDepth, int strictNormals, int hideEmitters, int integratorKind);
void pathref_setup(void *h, const float *toWorld, float fov, float nearClip, float farClip, int W, int H, int rfilter, int samplerKind, int spp, uint64_t scramble,
                   int maxDepth, int rrDepth, int strictNormals, int hideEmitters) {
    pathref_setup2(h, toWorld, fov, nearClip, farClip, W, H, rfilter, samplerKind, spp, scramble, maxDepth, rrDepth, strictNormals, hideEmitters, 0);
}
/* samplerKind 0 sobol, 1 independent (SFMT), 2 this repository's counter stream; integratorKind 0 path, 1 volpath */
void pathref_setup3(void *h, const float *toWorld, float fov, float nearClip, float farClip, int W, int H, int rfilter, int samplerKind, int spp, uint64_t scramble,
                    int maxDepth, int rrDepth, int strictNormals, int hideEmitters, int integratorKind, float apertureRadius, float focusDistance);
void pathref_setup2(void *h, const float *toWorld, float fov, float nearClip, float farClip, int W, int H, int rfilter, int samplerKind, int spp, uint64_t scramble,
                    int maxDepth, int rrDepth, int strictNormals, int hideEmitters, int integratorKind) {
    pathref_setup3(h, toWorld, fov, nearClip, farClip, W, H, rfilter, samplerKind, spp, scramble, maxDepth, rrDepth, strictNormals, hideEmitters, integratorKind, 0.0f, 0.0f);
}
void pathref_setup4(void *h, const float *toWorld, float fov, float nearClip, float farClip, int W, int H, int rfilter, int samplerKind, int spp, uint64_t scramble,
                    int maxDepth, int rrDepth, int strictNormals, int hideEmitters, int integratorKind, float apertureRadius, float focusDistance,
                    int cropX, int cropY, int cropW, int cropH);
/* apertureRadius > 0: the `thinlens` sensor (src/sensors/thinlens.cpp) with that aperture and focusDistance (<= 0: the plugin's default) */
void pathref_setup3(void *h, const float *toWorld, float fov, float nearClip, float farClip, int W, int H, int rfilter, int samplerKind, int spp, uint64_t scramble,
                    int maxDepth, int rrDepth, int strictNormals, int hideEmitters, int integratorKind, float apertureRadius, float focusDistance) {
    pathref_setup4(h, toWorld, fov, nearClip, farClip, W, H, rfilter, samplerKind, spp, scramble, maxDepth, rrDepth, strictNormals, hideEmitters, integratorKind,
                   apertureRadius, focusDistance, 0, 0, W, H);
}
/* film crop window (film.cpp:36-47): W x H is the full film; the render (blocks, sample positions, output buffer) covers cropW x cropH */
void pathref_setup4(void *h, const float *toWorld, float fov, float nearClip, float farClip, int W, int H, int rfilter, int samplerKind, int spp, uint64_t scramble,
                    int maxDepth, int rrDepth, int strictNormals, int hideEmitters, int integratorKind, float apertureRadius, float focusDistance,
                    int cropX, int cropY, int cropW, int cropH) {
    PathRef *p = (PathRef *) h;
    p->W = cropW; p->H = cropH;
    Properties fp("hdrfilm");
    fp.setInteger("width", W); fp.setInteger("height", H);
    if (cropX != 0 || cropY != 0 || cropW != W || cropH != H) {
        fp.setInteger("cropOffsetX", cropX); fp.setInteger("cropOffsetY", cropY);
        fp.setInteger("cropWidth", cropW); fp.setInteger("cropHeight", cropH);
    }
    p->film = new StandinFilm(fp);
    Properties rp(rfilter == 0 ? "box" : "gaussian");
    ReconstructionFilter *rf = (ReconstructionFilter *) (rfilter == 0 ? CreateInstance_box(rp) : CreateInstance_gaussian(rp));
    rf->configure();
    p->film->addChild("", rf);
    p->film->configure();
    Properties sp(samplerKind == 0 ? "sobol" : "independent");
    sp.setInteger("sampleCount", spp);
    sp.setInteger("scramble", (int) scramble);
    if (samplerKind == 2) p->sampler = new CounterSamplerPlugin(cropW, (size_t) spp, scramble);
    else p->sampler = (Sampler *) (samplerKind == 0 ? CreateInstance_sobol(sp) : CreateInstance_independent(sp));
    p->sampler->configure();
    Properties cp(apertureRadius > 0 ? "thinlens" : "perspective");
    Matrix4x4 M;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) M.m[r][c] = toWorld[4 * r + c];
    cp.setTransform("toWorld", Transform(M));
    cp.setFloat("fov", fov); cp.setFloat("nearClip", nearClip); cp.setFloat("farClip", farClip);
    if (apertureRadius > 0) {
        cp.setFloat("apertureRadius", apertureRadius);
        if (focusDistance > 0) cp.setFloat("focusDistance", focusDistance);
        p->sensor = (Sensor *) CreateInstance_thinlens(cp);
    } else
        p->sensor = (Sensor *) CreateInstance_perspective(cp);
    p->sensor->addChild("", p->film);
    p->sensor->addChild("", p->sampler);
    p->sensor->configure();
    Properties ip("path");
    ip.setInteger("maxDepth", maxDepth); ip.setInteger("rrDepth", rrDepth);
    ip.setBoolean("strictNormals", strictNormals != 0); ip.setBoolean("hideEmitters", hideEmitters != 0);
    p->integrator = (Integrator *) (integratorKind == 1 ? CreateInstance_volpath(ip) : CreateInstance_path(ip));
    p->integrator->configure();
    p->scene->addChild("", p->sensor);
    p->scene->addChild("", p->integrator);
    p->scene->configure();
    p->scene->getKDTree()->setParallelBuild(false);
    p->scene->initialize();
    p->integrator->configureSampler(p->scene, p->sampler);
}
/* m_sampleToCamera as PerspectiveCameraImpl::configure derives it (perspective.cpp:146-153; the member itself is private to the plugin):
 * the same chain of the reference's own Transform operations, incl. the crop window.  Row-major 4 x 4. */
void pathref_sample_to_camera(void *h, float *out16) {
    PathRef *p = (PathRef *) h;
    const PerspectiveCamera *cam = static_cast<const PerspectiveCamera *>(p->sensor.get());
    const Float aspect = cam->getAspect();
    const Vector2i &filmSize = p->film->getSize(), &cropSize = p->film->getCropSize();
    const Point2i &cropOffset = p->film->getCropOffset();
    const Vector2 relSize((Float) cropSize.x / (Float) filmSize.x, (Float) cropSize.y / (Float) filmSize.y);
    const Point2 relOffset((Float) cropOffset.x / (Float) filmSize.x, (Float) cropOffset.y / (Float) filmSize.y);
    const Transform cameraToSample =
          Transform::scale(Vector(1.0f / relSize.x, 1.0f / relSize.y, 1.0f))
        * Transform::translate(Vector(-relOffset.x, -relOffset.y, 0.0f))
        * Transform::scale(Vector(-0.5f, -0.5f * aspect, 1.0f))
        * Transform::translate(Vector(-1.0f, -1.0f / aspect, 0.0f))
        * Transform::perspective(cam->getXFov(), cam->getNearClip(), cam->getFarClip());
    const Transform sampleToCamera = cameraToSample.inverse();
    const Matrix4x4 &M = sampleToCamera.getMatrix();
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out16[4 * r + c] = M.m[r][c];
}
/* Sensor::sampleRay for film positions: pos 2n -> rays 8n (o, mint, d, maxt) */
void pathref_camera_rays(void *h, int n, const float *pos, float *rays) {
    PathRef *p = (PathRef *) h;
    for (int i = 0; i < n; ++i) {
        Ray ray;
        p->sensor->sampleRay(ray, Point2(pos[2 * i], pos[2 * i + 1]), Point2(0.5f), 0.5f);
        float *o = rays + 8 * i;
        o[0] = ray.o.x; o[1] = ray.o.y; o[2] = ray.o.z; o[3] = ray.mint; o[4] = ray.d.x; o[5] = ray.d.y; o[6] = ray.d.z; o[7] = ray.maxt;
    }
}
/* Scene::rayIntersect(ray, its) (kd-tree traversal + ShapeKDTree::fillIntersectionRecord): rays 8n (o, mint, d, maxt) -> out 24n:
 * p geoN shN s t wi (3 each), t, -, primIndex, valid, pad 2 (layout of the oracle's orc_intersect_full; the mesh index is not comparable) */
void pathref_intersect(void *h, int n, const float *rays, float *out) {
    PathRef *p = (PathRef *) h;
    for (int i = 0; i < n; ++i) {
        const float *r = rays + 8 * i;
        Ray ray(Point(r[0], r[1], r[2]), Vector(r[4], r[5], r[6]), r[3], r[7], 0.0f);
        Intersection its;
        float *o = out + 24 * i;
        memset(o, 0, 96);
        if (p->scene->rayIntersect(ray, its)) {
            const Vector vs[6] = {Vector(its.p), Vector(its.geoFrame.n), Vector(its.shFrame.n), its.shFrame.s, its.shFrame.t, its.wi};
            for (int k = 0; k < 6; ++k) { o[3 * k] = vs[k].x; o[3 * k + 1] = vs[k].y; o[3 * k + 2] = vs[k].z; }
            o[18] = its.t; o[20] = (float) its.primIndex; o[21] = 1.0f;
        }
    }
}
/* Scene::sampleEmitterDirect(dRec, sample, testVisibility = true): ref 6n (ref, refN), samples 2n -> out 12n: d(3) dist pdf value(3) ok p(3)
 * (layout of the oracle's orc_sample_emitter_direct) */
void pathref_sample_emitter_direct(void *h, int n, const float *ref, const float *samples, float *out) {
    PathRef *p = (PathRef *) h;
    for (int i = 0; i < n; ++i) {
        DirectSamplingRecord dRec(Point(ref[6 * i], ref[6 * i + 1], ref[6 * i + 2]), 0.0f);
        dRec.refN = Normal(ref[6 * i + 3], ref[6 * i + 4], ref[6 * i + 5]);
        const Spectrum value = p->scene->sampleEmitterDirect(dRec, Point2(samples[2 * i], samples[2 * i + 1]), true);
        float *o = out + 12 * i;
        o[0] = dRec.d.x; o[1] = dRec.d.y; o[2] = dRec.d.z; o[3] = dRec.dist; o[4] = dRec.pdf;
        o[5] = value[0]; o[6] = value[1]; o[7] = value[2]; o[8] = value.isZero() ? 0.0f : 1.0f; o[9] = dRec.p.x; o[10] = dRec.p.y; o[11] = dRec.p.z;
    }
}
/* The CIE 1931 2-degree observer table the reference integrates against (src/libcore/spectrum.cpp:29-34,743-1141: 471 entries, 360..830 nm,
 * 1 nm): out 471 x 4 = wavelength, xbar, ybar, zbar.  Used by tools/extract_cie_table.py to write mitsuba_b200/data/cie1931_xyz_1nm.txt
 * (public standard data, stored like the Sobol' tables) and by the test that pins that file. */
int pathref_cie_tables(float *out) {
    for (int i = 0; i < 471; ++i) { out[4 * i] = CIE_wavelengths[i]; out[4 * i + 1] = CIE_X_entries[i]; out[4 * i + 2] = CIE_Y_entries[i]; out[4 * i + 3] = CIE_Z_entries[i]; }
    return 471;
}
/* <spectrum value="l0:v0, l1:v1, ..."> / <spectrum filename="x.spd"> as the scene loader turns them into a property (scenehandler.cpp:557-611):
 * InterpolatedSpectrum, zeroExtend(), Spectrum::fromContinuousSpectrum (spectrum.cpp:172-186: adaptive Gauss-Lobatto quadrature of the product
 * with the matching functions), clampNegative. */
int pathref_spectrum_to_rgb(int n, const float *wavelengths, const float *values, int zeroExtend, float *rgb) {
    try {
        InterpolatedSpectrum interp((size_t) n);
        for (int i = 0; i < n; ++i) interp.append(wavelengths[i], values[i]);
        if (zeroExtend) interp.zeroExtend();
        Spectrum discrete;
        discrete.fromContinuousSpectrum(interp);
        discrete.clampNegative();
        for (int i = 0; i < 3; ++i) rgb[i] = discrete[i];
        return 0;
    } catch (...) { return 1; }
}
/* The RGB (eta, k) a conductor plugin derives from material="<name>" (roughconductor.cpp:174-190, conductor.cpp:160-176): the reference's own
 * InterpolatedSpectrum file reader + Spectrum::fromContinuousSpectrum (src/libcore/spectrum.cpp) on <dataDir>/ior/<name>.{eta,k}.spd.
 * Used by tools/extract_conductor_presets.py to generate mitsuba_b200/data/conductor_presets.txt and by the test that pins that table. */
int pathref_conductor_preset(const char *dataDir, const char *material, float *eta3, float *k3) {
    try {
        const std::string base = std::string(dataDir) + "/ior/" + material;
        Spectrum eta, k;
        eta.fromContinuousSpectrum(InterpolatedSpectrum(fs::path(base + ".eta.spd")));
        k.fromContinuousSpectrum(InterpolatedSpectrum(fs::path(base + ".k.spd")));
        for (int i = 0; i < 3; ++i) { eta3[i] = eta[i]; k3[i] = k[i]; }
        return 0;
    } catch (...) { return 1; }
}
#ifdef WITH_B200_SHIM
/* The same Scene object rendered through the Mitsuba-side plugin of this repository (mitsuba_b200/host/b200_integrator.cpp, class
 * B200PathTracer: Integrator::render -> C-ABI of libb2mts.so -> film through Film::setBitmap) instead of MIPathTracer + renderBlock.
 * Returns 0 and fills out (H x W x 5), or 1 with the exception text in err. */
extern "C" void *CreateInstance_b200path(const Properties &props);
int pathref_render_b200(void *h, int device, int parity, float *out, char *err, int errLen) {
    PathRef *p = (PathRef *) h;
    struct ThrowScope { ThrowScope() { standinLogThrows() = true; } ~ThrowScope() { standinLogThrows() = false; } } scope; /* Log(EError) throws in here */
    try {
        const Properties &op = p->integrator->getProperties();
        Properties ip("b200path");
        ip.setInteger("maxDepth", op.getInteger("maxDepth", -1)); ip.setInteger("rrDepth", op.getInteger("rrDepth", 5));
        ip.setBoolean("strictNormals", op.getBoolean("strictNormals", false)); ip.setBoolean("hideEmitters", op.getBoolean("hideEmitters", false));
        ip.setInteger("device", device); ip.setBoolean("parity", parity != 0);
        ref<Integrator> integ = (Integrator *) CreateInstance_b200path(ip);
        integ->configure();
        if (!integ->render(p->scene, NULL, NULL, 0, 0, 0)) throw std::runtime_error("render() was cancelled");
        StandinFilm *film = static_cast<StandinFilm *>(p->film.get());
        if (!film->m_last) throw std::runtime_error("the integrator did not hand a bitmap to the film");
        memcpy(out, film->m_last->getFloat32Data(), sizeof(float) * 5 * (size_t) p->W * p->H);
        return 0;
    } catch (const std::exception &e) {
        if (err && errLen > 0) { strncpy(err, e.what(), (size_t) errLen - 1); err[errLen - 1] = 0; }
        return 1;
    }
}
#endif
/* film out: H x W x 5 (rgb, alpha, weight), blocks of 32 x 32 rendered by SamplingIntegrator::renderBlock with the pixels of a block in
 * scanline order and accumulated here in block order (the reference's scheduler hands out Hilbert-ordered pixels and merges blocks in
 * completion order: float summation order only) */
void pathref_render_blocks(void *h, int first, int step, float *out);
void pathref_render(void *h, float *out) { pathref_render_blocks(h, 0, 1, out); }
/* the blocks with index == first (mod step): lets several processes share one image (bench.py --impl reference) */
void pathref_render_blocks(void *h, int first, int step, float *out) {
    PathRef *p = (PathRef *) h;
    int blockIndex = 0;
    SamplingIntegrator *integrator = static_cast<SamplingIntegrator *>(p->integrator.get());
    const ReconstructionFilter *rf = p->film->getReconstructionFilter();
    const int bs = 32, border = rf->getBorderSize(), W = p->W, H = p->H;
    memset(out, 0, sizeof(float) * 5 * (size_t) W * H);
    bool stop = false;
    for (int oy = 0; oy < H; oy += bs)
        for (int ox = 0; ox < W; ox += bs) {
            if ((blockIndex++ % step) != first) continue;
            const int sx = std::min(bs, W - ox), sy = std::min(bs, H - oy);
            ref<ImageBlock> block = new ImageBlock(Bitmap::ESpectrumAlphaWeight, Vector2i(sx, sy), rf);
            block->setOffset(Point2i(ox, oy));
            std::vector<TPoint2<uint8_t> > points;
            for (int y = 0; y < sy; ++y) for (int x = 0; x < sx; ++x) points.push_back(TPoint2<uint8_t>((uint8_t) x, (uint8_t) y));
            integrator->renderBlock(p->scene, p->sensor, p->sampler, block, stop, points);
            const float *data = block->getBitmap()->getFloat32Data();
            const int bw = sx + 2 * border, bh = sy + 2 * border;
            for (int y = 0; y < bh; ++y) {
                const int fy = oy - border + y;
                if (fy < 0 || fy >= H) continue;
                for (int x = 0; x < bw; ++x) {
                    const int fx = ox - border + x;
                    if (fx < 0 || fx >= W) continue;
                    for (int k = 0; k < 5; ++k) out[((size_t) fy * W + fx) * 5 + k] += data[((size_t) y * bw + x) * 5 + k];
                }
            }
        }
}
}
