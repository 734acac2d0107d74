/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
#include <algorithm>
#include <string>
namespace boost { inline std::string to_lower_copy(std::string s) { std::transform(s.begin(), s.end(), s.begin(), ::tolower); return s; } inline void to_lower(std::string &s) { std::transform(s.begin(), s.end(), s.begin(), ::tolower); } }
namespace boost { inline bool ends_with(const std::string &s, const std::string &e) { return s.size() >= e.size() && s.compare(s.size() - e.size(), e.size(), e) == 0; } inline bool starts_with(const std::string &s, const std::string &e) { return s.compare(0, e.size(), e) == 0; } }
