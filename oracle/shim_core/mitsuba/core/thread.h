/* Stand-in header (test infrastructure only, see oracle/shim_core/README). */
#pragma once
#include <mitsuba/core/object.h>
#include <boost/filesystem.hpp>
namespace fs = boost::filesystem;
namespace mitsuba {
class FileResolver : public Object { public: fs::path resolve(const fs::path &p) const { return p; } };
class Thread : public Object { public: static Thread *getThread() { static Thread t; return &t; } FileResolver *getFileResolver() { static FileResolver r; return &r; } enum EThreadPriority { EIdlePriority = 0, ELowestPriority, ELowPriority, ENormalPriority, EHighPriority, EHighestPriority, ERealtimePriority }; }; }
