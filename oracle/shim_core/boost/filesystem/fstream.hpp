/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
#include <fstream>
#include <string>
#include <cstdio>
#include <unistd.h>
#include <sys/stat.h>
#include <ctime>
namespace boost { namespace filesystem {
class path { public: path() {} path(const char *s) : m_s(s) {} path(const std::string &s) : m_s(s) {} bool empty() const { return m_s.empty(); } const std::string &string() const { return m_s; } const char *c_str() const { return m_s.c_str(); }
  path filename() const { return *this; } path extension() const { return path(); } path parent_path() const { return path(); } path operator/(const path &o) const { return path(m_s + "/" + o.m_s); } bool is_absolute() const { return false; }
  path &replace_extension(const path &e) { size_t dot = m_s.find_last_of('.'), sl = m_s.find_last_of('/'); if (dot != std::string::npos && (sl == std::string::npos || dot > sl)) m_s.erase(dot); m_s += e.m_s; return *this; }
  private: std::string m_s; };
} namespace system { class error_code { public: error_code() : m_v(0) {} int value() const { return m_v; } void assign(int v) { m_v = v; } private: int m_v; }; } namespace filesystem {
inline std::time_t last_write_time(const path &p, system::error_code &ec) { struct stat st; if (::stat(p.string().c_str(), &st)) { ec.assign(1); return 0; } return st.st_mtime; }
inline bool exists(const path &p) { return ::access(p.string().c_str(), F_OK) == 0; }
inline size_t file_size(const path &p) { std::ifstream f(p.string().c_str(), std::ios::binary | std::ios::ate); return f ? (size_t) f.tellg() : 0; }
inline bool remove(const path &p) { return ::remove(p.string().c_str()) == 0; }
inline void resize_file(const path &p, size_t n) { if (::truncate(p.string().c_str(), (off_t) n)) {} }
class ifstream : public std::ifstream { public: ifstream() {} ifstream(const path &p) : std::ifstream(p.string().c_str()) {} };
class ofstream : public std::ofstream { public: ofstream() {} ofstream(const path &p, std::ios_base::openmode m = std::ios_base::out) : std::ofstream(p.string().c_str(), m) {} };
} }
