/* Minimal stand-in for <mitsuba/core/cobject.h>, used ONLY to compile two pieces of the reference where they lie under
 * /root/reference into oracle/_ref/librfilterref.so (see oracle/Makefile):
 *   include/mitsuba/core/rfilter.h   -- class ReconstructionFilter and the Resampler<Scalar> template (the real header)
 *   src/rfilters/lanczos.cpp         -- LanczosSincFilter (the real plugin source)
 * Everything here is scaffolding those two files expect from the rest of libcore (macros, Float, an empty ConfigurableObject,
 * a Properties that answers getInteger); include/mitsuba/core/math.h is the reference's own header.  Test infrastructure only. */
#pragma once
#include <assert.h>
#include <stdint.h>
#include <sys/types.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <limits>
#include <string>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define MTS_EXPORT_CORE
#define MTS_DECLARE_CLASS()
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super)
#define MTS_EXPORT_PLUGIN(name, descr)
#define EXPECT_NOT_TAKEN(a) (a)
#define EXPECT_TAKEN(a) (a)
#define FINLINE inline
#define SAssert(cond) assert(cond)
#ifndef SINGLE_PRECISION
#define SINGLE_PRECISION 1
#endif
#if defined(__linux) && !defined(__LINUX__)
#define __LINUX__ /* include/mitsuba/core/platform.h:68-69 -- selects the double-precision fastexp / fastlog of math.h:175-199 */
#endif
#define Epsilon 1e-4f /* include/mitsuba/core/constants.h:28 (single precision) */

namespace mitsuba { typedef float Float; }
#include <mitsuba/core/math.h> /* the reference's header: floorToInt, ceilToInt, modulo, clamp */

namespace mitsuba {
class Stream;
class InstanceManager;
class Properties {
public:
    int lobes = 3;
    int getInteger(const std::string &, int) const { return lobes; }
};
class ConfigurableObject {
public:
    ConfigurableObject(const Properties &) {}
    ConfigurableObject(Stream *, InstanceManager *) {}
    virtual ~ConfigurableObject() {}
};
inline std::string formatString(const char *fmt, ...) { return fmt; }
}
