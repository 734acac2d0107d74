/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
#include <mitsuba/mitsuba.h>
namespace mitsuba {
class Properties {
public:
    std::map<std::string, std::string> s; std::map<std::string, float> f; std::map<std::string, bool> b;
    std::string id; const std::string &getID() const { return id; } void setID(const std::string &v) { id = v; }
    bool hasProperty(const std::string &k) const { return s.count(k) || f.count(k) || b.count(k); }
    std::string getString(const std::string &k) const { return s.at(k); }
    std::string getString(const std::string &k, const std::string &d) const { return s.count(k) ? s.at(k) : d; }
    Float getFloat(const std::string &k) const { return f.at(k); }
    Float getFloat(const std::string &k, Float d) const { return f.count(k) ? f.at(k) : d; }
    bool getBoolean(const std::string &k) const { return b.at(k); }
    bool getBoolean(const std::string &k, bool d) const { return b.count(k) ? b.at(k) : d; }
};
}
