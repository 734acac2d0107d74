/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
namespace mitsuba { enum ELogLevel { ETrace = 0, EDebug = 100, EInfo = 200, EWarn = 300, EError = 400 }; inline void standinLog(ELogLevel, const char *, ...) {} }
#define Log(level, ...) ::mitsuba::standinLog(level, __VA_ARGS__)
#define SLog(level, ...) ::mitsuba::standinLog(level, __VA_ARGS__)
#define Assert(cond) assert(cond)
#define SAssert(cond) assert(cond)
#define AssertEx(cond, msg) assert(cond)
#define SAssertEx(cond, msg) assert(cond)
#define NotImplementedError(name) assert(false)
