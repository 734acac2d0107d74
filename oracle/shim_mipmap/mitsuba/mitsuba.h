/* Minimal stand-in for <mitsuba/mitsuba.h>, used ONLY to compile the reference's texture look-up code where it lies under
 * /root/reference into oracle/_ref/libmipmapref.so (see oracle/Makefile, oracle/mipmap_ref_shim.cpp):
 *   include/mitsuba/render/mipmap.h   TMIPMap: evalTexel / evalBox / evalBilinear / evalEWA / eval  (the real header)
 *   include/mitsuba/core/barray.h     BlockedArray                                                  (the real header)
 *   include/mitsuba/core/rfilter.h    ReconstructionFilter::EBoundaryCondition                      (the real header)
 *   include/mitsuba/core/spectrum.h   TSpectrum / Color3 arithmetic                                 (the real header)
 *   include/mitsuba/core/math.h       floorToInt, ceilToInt, modulo, clamp                          (the real header)
 *   src/libcore/math.cpp semantics of hypot2 / log2 are provided by the reference's own math.cpp     (compiled alongside)
 * Everything in this directory is scaffolding that those files expect from the rest of libcore (macros, Float, Object, ref<>,
 * Vector2 / Point2 / Vector2i, a Bitmap that is only named, a MemoryMappedFile over a memory buffer, empty statistics).
 * Test infrastructure only. */
#pragma once
#include <assert.h>
#include <stdint.h>
#include <string.h>
#include <sys/types.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define MTS_EXPORT_CORE
#define MTS_EXPORT_RENDER
#define MTS_DECLARE_CLASS() \
    static Class *m_theClass; \
    virtual const Class *getClass() const;
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super)
#define MTS_EXPORT_PLUGIN(name, descr)
#define EXPECT_NOT_TAKEN(a) (a)
#define EXPECT_TAKEN(a) (a)
#define FINLINE inline
#define SAssert(cond) assert(cond)
#define Assert(cond) assert(cond)
#define AssertEx(cond, msg) assert(cond)
#define SAssertEx(cond, msg) assert(cond)
#ifndef SINGLE_PRECISION
#define SINGLE_PRECISION 1
#endif
#if defined(__linux) && !defined(__LINUX__)
#define __LINUX__ /* include/mitsuba/core/platform.h:68-69 -- selects the double-precision fastexp / fastlog of math.h:175-199 */
#endif
#define SPECTRUM_SAMPLES 3
#define Epsilon 1e-4f          /* include/mitsuba/core/constants.h:28 (single precision) */
#define RCPOVERFLOW 2.93873587705571876e-39f
#define MTS_NAMESPACE_IS_STANDIN 1

namespace mitsuba {
typedef float Float;
enum ELogLevel { ETrace = 0, EDebug = 100, EInfo = 200, EWarn = 300, EError = 400 };
inline void standinLog(ELogLevel, const char *, ...) {}
}
#define Log(level, ...) ::mitsuba::standinLog(level, __VA_ARGS__)
#define SLog(level, ...) ::mitsuba::standinLog(level, __VA_ARGS__)

#include <mitsuba/core/math.h> /* the reference's header */
#include <boost/filesystem/fstream.hpp> /* stand-in */

namespace mitsuba {
namespace fs = boost::filesystem;
using std::endl;
class Stream { /* only named by inline (un)serialisation code that the shim never instantiates */
public:
    Float readFloat();
    void writeFloat(Float);
    void readFloatArray(Float *, size_t);
    void writeFloatArray(const Float *, size_t);
    float readSingle();
    void writeSingle(float);
    void readSingleArray(float *, size_t);
    void writeSingleArray(const float *, size_t);
    template <typename T> void readArray(T *, size_t);
    template <typename T> void writeArray(const T *, size_t);
};
class InstanceManager;
class Properties;
class Class {
public:
    Class(const char *, bool, const char *) {}
};
class Object {
public:
    virtual ~Object() {}
    void incRef() const {}
    void decRef() const {}
    virtual std::string toString() const { return ""; }
};
class ConfigurableObject : public Object {
public:
    ConfigurableObject() {}
    ConfigurableObject(const Properties &) {}
    ConfigurableObject(Stream *, InstanceManager *) {}
};
class SerializableObject : public Object {};
template <typename T> class ref {
public:
    ref() : m_ptr(NULL) {}
    ref(T *p) : m_ptr(p) {}
    T *operator->() const { return m_ptr; }
    T *get() const { return m_ptr; }
    operator T *() const { return m_ptr; }
private:
    T *m_ptr;
};
/* vector.h / point.h: only .x/.y, construction and operator[] are used by mipmap.h / barray.h */
template <typename T> struct TVec2 {
    T x, y;
    TVec2() : x(0), y(0) {}
    TVec2(T v) : x(v), y(v) {}
    TVec2(T x_, T y_) : x(x_), y(y_) {}
    T &operator[](int i) { return (&x)[i]; }
    const T &operator[](int i) const { return (&x)[i]; }
};
typedef TVec2<Float> Vector2;
typedef TVec2<Float> Point2;
typedef TVec2<int> Vector2i;
inline void *allocAligned(size_t size) { void *p = NULL; if (posix_memalign(&p, 64, size ? size : 64)) return NULL; return p; }
inline void freeAligned(void *p) { free(p); }
inline std::string memString(size_t) { return ""; }
inline std::string formatString(const char *fmt, ...) { return fmt; }
inline std::string indent(const std::string &s) { return s; }
}
