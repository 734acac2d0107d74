/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
namespace mitsuba {
enum ELogLevel { ETrace = 0, EDebug = 100, EInfo = 200, EWarn = 300, EError = 400 };
/* like Logger::log (src/libcore/logger.cpp:100-147): messages below EError are dropped here, EError throws std::runtime_error */
inline void standinLog(ELogLevel level, const char *fmt, ...) {
    if (level < EError) return;
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
