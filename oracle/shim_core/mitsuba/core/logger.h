/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
namespace mitsuba {
enum ELogLevel { ETrace = 0, EDebug = 100, EInfo = 200, EWarn = 300, EError = 400 };
/* Logger::log (src/libcore/logger.cpp:100-147) throws std::runtime_error for EError.  Here messages are dropped, and EError only throws
   while standinLogThrows() is set (the harness of the b200path plugin sets it around Integrator::render): the scene-building entry points
   are extern "C" functions called from Python, which an exception must not cross */
inline bool &standinLogThrows() { static bool flag = false; return flag; }
inline void standinLog(ELogLevel level, const char *fmt, ...) {
    if (level < EError || !standinLogThrows()) return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw std::runtime_error(buf);
}
}
#define Log(level, ...) ::mitsuba::standinLog(level, __VA_ARGS__)
#define SLog(level, ...) ::mitsuba::standinLog(level, __VA_ARGS__)
#define Assert(cond) assert(cond)
#define SAssert(cond) assert(cond)
#define AssertEx(cond, msg) assert(cond)
#define SAssertEx(cond, msg) assert(cond)
#define NotImplementedError(name) assert(false)
