/* Thin C entry points over the REFERENCE's own BSDF plugins -- src/bsdfs/{diffuse,roughconductor,roughdielectric,coating,dielectric,
 * conductor,plastic,twosided,null}.cpp with microfacet.h / ior.h and src/libcore/{util,warp,math,quad}.cpp -- compiled where they
 * lie under /root/reference (never copied) into oracle/_ref/libbsdfref.so by oracle/Makefile, behind the stand-in headers of
 * oracle/shim_core/.  Each plugin keeps its own entry point (CreateInstance, renamed per plugin on the command line), is
 * instantiated from a Properties object exactly as the plugin manager would do it, and its eval() / pdf() / sample() are called
 * through the real BSDF interface (include/mitsuba/render/bsdf.h) with a real BSDFSamplingRecord.
 *
 * What is NOT the reference here: the out-of-line members of the base classes BSDF, Texture, Shader and of the two constant
 * textures (src/librender/{bsdf,texture,shader}.cpp and src/libhw/basicshader.cpp pull in the whole scene graph / OpenGL layer);
 * they are given the few lines the plugins rely on (flag OR-ing in BSDF::configure, getEta() = 1).  None of them touches eval / pdf /
 * sample.  Used only to pin the oracle (tests/gen_golden.py -> tests/golden/bsdf_ref.npz; tests/test_oracle_reference_pins.py). */
#include <mitsuba/render/bsdf.h>
#include <mitsuba/render/sampler.h>
#include <mitsuba/render/texture.h>
#include <mitsuba/hw/basicshader.h>
#include <mitsuba/core/random.h>

namespace mitsuba {
/* ---- scaffolding: base-class members that live in librender / libhw ---- */
BSDF::BSDF(const Properties &props) : ConfigurableObject(props) {
    m_ensureEnergyConservation = props.getBoolean("ensureEnergyConservation", true);
    m_usesRayDifferentials = false;
}
BSDF::BSDF(Stream *stream, InstanceManager *manager) : ConfigurableObject(stream, manager) { m_ensureEnergyConservation = true; m_usesRayDifferentials = false; }
BSDF::~BSDF() {}
void BSDF::setParent(ConfigurableObject *) {}
void BSDF::addChild(const std::string &name, ConfigurableObject *obj) { ConfigurableObject::addChild(name, obj); }
void BSDF::configure() { m_combinedType = 0; for (size_t i = 0; i < m_components.size(); ++i) m_combinedType |= m_components[i]; }
Float BSDF::getEta() const { return 1.0f; }
Frame BSDF::getFrame(const Intersection &its) const { return its.shFrame; }
void BSDF::getFrameDerivative(const Intersection &, Frame &, Frame &) const {}
Spectrum BSDF::getDiffuseReflectance(const Intersection &) const { return Spectrum(0.0f); }
Texture *BSDF::ensureEnergyConservation(Texture *texture, const std::string &, Float) const { return texture; } /* inputs here never exceed 1 */
std::pair<Texture *, Texture *> BSDF::ensureEnergyConservation(Texture *a, Texture *b, const std::string &, const std::string &, Float) const { return std::make_pair(a, b); }
void BSDF::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(BSDF, true, ConfigurableObject)

Texture::Texture(const Properties &props) : ConfigurableObject(props) {}
Texture::Texture(Stream *stream, InstanceManager *manager) : ConfigurableObject(stream, manager) {}
Texture::~Texture() {}
Vector3i Texture::getResolution() const { return Vector3i(0); }
Spectrum Texture::eval(const Intersection &, bool) const { return Spectrum(0.0f); }
void Texture::evalGradient(const Intersection &, Spectrum *) const {}
Spectrum Texture::getAverage() const { return Spectrum(0.0f); }
Spectrum Texture::getMinimum() const { return Spectrum(0.0f); }
Spectrum Texture::getMaximum() const { return Spectrum(0.0f); }
bool Texture::isConstant() const { return false; }
bool Texture::isMonochromatic() const { return false; }
bool Texture::usesRayDifferentials() const { return false; }
ref<Bitmap> Texture::getBitmap(const Vector2i &) const { return NULL; }
ref<Texture> Texture::expand() { return this; }
void Texture::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(Texture, true, ConfigurableObject)

ConstantSpectrumTexture::ConstantSpectrumTexture(Stream *stream, InstanceManager *manager) : Texture(stream, manager) {}
Shader *ConstantSpectrumTexture::createShader(Renderer *) const { return NULL; }
ref<Bitmap> ConstantSpectrumTexture::getBitmap(const Vector2i &) const { return NULL; }
void ConstantSpectrumTexture::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(ConstantSpectrumTexture, true, Texture)
ConstantFloatTexture::ConstantFloatTexture(Stream *stream, InstanceManager *manager) : Texture(stream, manager) {}
Shader *ConstantFloatTexture::createShader(Renderer *) const { return NULL; }
ref<Bitmap> ConstantFloatTexture::getBitmap(const Vector2i &) const { return NULL; }
void ConstantFloatTexture::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(ConstantFloatTexture, true, Texture)

Shader *Renderer::registerShaderForResource(const HWResource *) { return NULL; }
void Renderer::unregisterShaderForResource(const HWResource *) {}

/* BSDFSamplingRecord's constructors are inline functions of render/records.inl, a file that only compiles together with scene.h;
   they are pure member-initialiser lists (records.inl:24-38) */
BSDFSamplingRecord::BSDFSamplingRecord(const Intersection &its_, Sampler *sampler_, ETransportMode mode_)
    : its(its_), sampler(sampler_), wi(its_.wi), mode(mode_), typeMask(BSDF::EAll), component(-1), sampledType(0), sampledComponent(-1) {}
BSDFSamplingRecord::BSDFSamplingRecord(const Intersection &its_, const Vector &wo_, ETransportMode mode_)
    : its(its_), sampler(NULL), wi(its_.wi), wo(wo_), mode(mode_), typeMask(BSDF::EAll), component(-1), sampledType(0), sampledComponent(-1) {}
BSDFSamplingRecord::BSDFSamplingRecord(const Intersection &its_, const Vector &wi_, const Vector &wo_, ETransportMode mode_)
    : its(its_), sampler(NULL), wi(wi_), wo(wo_), mode(mode_), typeMask(BSDF::EAll), component(-1), sampledType(0), sampledComponent(-1) {}
/* ConfigurableObject's members live in src/libcore/properties.cpp (boost::variant) */
ConfigurableObject::ConfigurableObject(Stream *, InstanceManager *) {}
void ConfigurableObject::setParent(ConfigurableObject *) {}
void ConfigurableObject::addChild(const std::string &, ConfigurableObject *) {}
void ConfigurableObject::configure() {}
void ConfigurableObject::serialize(Stream *, InstanceManager *) const {}
MTS_IMPLEMENT_CLASS(ConfigurableObject, true, SerializableObject)
SerializableObject::SerializableObject(Stream *, InstanceManager *) {}
MTS_IMPLEMENT_CLASS(SerializableObject, true, Object)
Float BSDF::getRoughness(const Intersection &, int) const { return 0.0f; }
Float ContinuousSpectrum::average(Float, Float) const { return 0.0f; }
std::string ContinuousSpectrum::toString() const { return ""; }
void InstanceManager::serialize(Stream *, const SerializableObject *) {}
InterpolatedSpectrum::InterpolatedSpectrum(const fs::path &) {}
Float InterpolatedSpectrum::eval(Float) const { return 0; }
Float InterpolatedSpectrum::average(Float, Float) const { return 0; }
std::string InterpolatedSpectrum::toString() const { return ""; }
void Spectrum::fromContinuousSpectrum(const ContinuousSpectrum &) {}
std::string Spectrum::toString() const { return ""; }
Float Random::nextFloat() { return 0.5f; }
size_t Random::nextSize(size_t) { return 0; }
void Stream::writeFloat(float) {}
void Stream::writeBool(bool) {}
void Stream::writeUInt(unsigned int) {}
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
template <> void Stream::writeArray<float>(const float *, size_t) {}
}

using namespace mitsuba;

/* the plugins' own entry points (MTS_EXPORT_PLUGIN, cobject.h:99-107), renamed per plugin with -DCreateInstance=... */
#define DECL(name) extern "C" void *CreateInstance_##name(const Properties &props);
DECL(diffuse) DECL(roughconductor) DECL(roughdielectric) DECL(coating) DECL(dielectric) DECL(conductor) DECL(plastic) DECL(twosided) DECL(null)

/* a Sampler that replays a given stream: the rough dielectric and the plastic / coating plugins draw extra numbers from bRec.sampler */
class ReplaySampler : public Sampler {
public:
    ReplaySampler() : Sampler(Properties()), m_values(NULL), m_count(0), m_pos(0) {}
    void set(const float *v, size_t n) { m_values = v; m_count = n; m_pos = 0; }
    Float next1D() { return m_pos < m_count ? m_values[m_pos++] : 0.5f; }
    Point2 next2D() { Float a = next1D(), b = next1D(); return Point2(a, b); }
    ref<Sampler> clone() { return this; }
    std::string toString() const { return "ReplaySampler"; }
    const Class *getClass() const { return NULL; }
private:
    const float *m_values; size_t m_count, m_pos;
};
extern "C" {

/* plugin: 0 diffuse, 1 roughconductor, 2 roughdielectric, 3 coating, 4 null, 5 twosided, 6 dielectric, 7 conductor, 8 plastic.
 * Properties are handed over as parallel key/value arrays (floats, strings, booleans, RGB spectra); child: nested BSDF or NULL. */
void *bsdfref_create(int plugin, int nf, const char **fk, const float *fv, int ns, const char **sk, const char **sv, int nb, const char **bk, const int *bv,
                     int nsp, const char **spk, const float *spv, void *child, void *child2) {
    Properties props;
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
    b->configure();
    return b;
}
unsigned int bsdfref_type(void *bsdf) { return ((BSDF *) bsdf)->getType(); }

static void makeIts(Intersection &its) {
    its.p = Point(0.0f); its.t = 1.0f;
    its.geoFrame = Frame(Vector(1, 0, 0), Vector(0, 1, 0), Normal(0, 0, 1));
    its.shFrame = its.geoFrame;
    its.uv = Point2(0.5f, 0.5f);
    its.dpdu = Vector(1, 0, 0); its.dpdv = Vector(0, 1, 0);
    its.hasUVPartials = false;
    its.time = 0;
}
/* eval + pdf for n (wi, wo) pairs in the local frame; measure 0 = ESolidAngle, 1 = EDiscrete; out rgb (3n), pdf (n) */
void bsdfref_eval(void *bsdf, int n, const float *wi, const float *wo, int discrete, float *outRgb, float *outPdf) {
    const BSDF *b = (const BSDF *) bsdf;
    Intersection its; makeIts(its);
    for (int i = 0; i < n; ++i) {
        its.wi = Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        BSDFSamplingRecord bRec(its, Vector(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), ERadiance);
        const EMeasure m = discrete ? EDiscrete : ESolidAngle;
        const Spectrum f = b->eval(bRec, m);
        outRgb[3 * i] = f[0]; outRgb[3 * i + 1] = f[1]; outRgb[3 * i + 2] = f[2];
        outPdf[i] = b->pdf(bRec, m);
    }
}
/* sample(bRec, pdf, sample): samples 3n (2-D sample + one number for bRec.sampler->next1D/2D) -> out 10n: wo(3) weight(3) pdf sampledType eta - */
void bsdfref_sample(void *bsdf, int n, const float *wi, const float *samples, float *out) {
    const BSDF *b = (const BSDF *) bsdf;
    Intersection its; makeIts(its);
    static ReplaySampler sampler;
    for (int i = 0; i < n; ++i) {
        its.wi = Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        float extra[4] = {samples[3 * i + 2], samples[3 * i + 2], samples[3 * i + 2], samples[3 * i + 2]};
        sampler.set(extra, 4);
        BSDFSamplingRecord bRec(its, &sampler, ERadiance);
        Float pdf = 0;
        const Spectrum w = b->sample(bRec, pdf, Point2(samples[3 * i], samples[3 * i + 1]));
        float *o = out + 10 * i;
        o[0] = bRec.wo.x; o[1] = bRec.wo.y; o[2] = bRec.wo.z; o[3] = w[0]; o[4] = w[1]; o[5] = w[2]; o[6] = pdf;
        o[7] = (float) bRec.sampledType; o[8] = bRec.eta; o[9] = 0;
        if (w.isZero()) { o[0] = o[1] = o[2] = 0; o[7] = 0; o[8] = 0; } /* wo / sampledType are unspecified after a failed sample */
    }
}
}
