/* Stand-in header (test infrastructure only, see oracle/shim_core/README). */
#pragma once
#include <mitsuba/core/object.h>
#include <boost/filesystem.hpp>
namespace fs = boost::filesystem;
namespace mitsuba {
class FileResolver : public Object { public: fs::path resolve(const fs::path &p) const { return p; } };
class Logger : public Object { public: void setLogLevel(int) {} void logProgress(Float, const std::string &, const std::string &, const std::string &, const void *) {} void log(int, const Class *, const char *, int, const char *, ...) {} int getLogLevel() const { return 400; } };
class Thread : public Object { public: Thread() {} Thread(const std::string &) {} virtual void run() {} void start() { run(); } void join() {} void setCritical(bool) {} void setPriority(int) {} Logger *getLogger() { static Logger l; return &l; } static int getID() { return 0; } const std::string &getName() const { static std::string n; return n; } static Thread *getThread() { static Thread t; return &t; } FileResolver *getFileResolver() { static FileResolver r; return &r; } enum EThreadPriority { EIdlePriority = 0, ELowestPriority, ELowPriority, ENormalPriority, EHighPriority, EHighestPriority, ERealtimePriority }; }; }
