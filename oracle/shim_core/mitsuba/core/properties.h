/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here.
 * A Properties that holds floats, booleans, strings and spectra -- what the BSDF plugin constructors query. */
#pragma once
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/transform.h>
namespace mitsuba {
class AnimatedTransform;
class Properties {
public:
    enum EPropertyType { EBoolean = 0, EInteger, EFloat, EPoint, EVector, ETransform, EAnimatedTransform, ESpectrum, EString, EData };
    struct Data { uint8_t *ptr; size_t size; };                                                   /* properties.h:73-84 */
    Data getData(const std::string &) const { Data d; d.ptr = NULL; d.size = 0; return d; }      /* never set here: hasProperty() says no */
    Properties() {}
    Properties(const std::string &pluginName) : m_pluginName(pluginName) {}
    const std::string &getPluginName() const { return m_pluginName; }
    void setPluginName(const std::string &n) { m_pluginName = n; }
    const std::string &getID() const { return m_id; }
    void setID(const std::string &v) { m_id = v; }
    bool hasProperty(const std::string &k) const { return s.count(k) || f.count(k) || b.count(k) || sp.count(k) || i.count(k); }
    EPropertyType getType(const std::string &k) const { return f.count(k) ? EFloat : (s.count(k) ? EString : (b.count(k) ? EBoolean : (i.count(k) ? EInteger : ESpectrum))); }
    std::string getString(const std::string &k) const { return s.at(k); }
    std::string getString(const std::string &k, const std::string &d) const { return s.count(k) ? s.at(k) : d; }
    std::string getAsString(const std::string &k) const { return getString(k, ""); }
    std::string getAsString(const std::string &k, const std::string &d) const { return getString(k, d); }
    Float getFloat(const std::string &k) const { return f.at(k); }
    Float getFloat(const std::string &k, Float d) const { return f.count(k) ? f.at(k) : d; }
    size_t getSize(const std::string &k) const { return (size_t) i.at(k); }
    size_t getSize(const std::string &k, size_t d) const { return i.count(k) ? (size_t) i.at(k) : d; }
    int getInteger(const std::string &k) const { return i.at(k); }
    int getInteger(const std::string &k, int d) const { return i.count(k) ? i.at(k) : d; }
    bool getBoolean(const std::string &k) const { return b.at(k); }
    bool getBoolean(const std::string &k, bool d) const { return b.count(k) ? b.at(k) : d; }
    Spectrum getSpectrum(const std::string &k) const { return sp.at(k); }
    Spectrum getSpectrum(const std::string &k, const Spectrum &d) const { return sp.count(k) ? sp.at(k) : d; }
    void setString(const std::string &k, const std::string &v, bool = true) { s[k] = v; }
    void setFloat(const std::string &k, Float v, bool = true) { f[k] = v; }
    void setInteger(const std::string &k, int v, bool = true) { i[k] = v; }
    void setBoolean(const std::string &k, bool v, bool = true) { b[k] = v; }
    void setSpectrum(const std::string &k, const Spectrum &v, bool = true) { sp[k] = v; }
    Transform getTransform(const std::string &k) const { return tr.at(k); }
    Transform getTransform(const std::string &k, const Transform &d) const { return tr.count(k) ? tr.at(k) : d; }
    void setTransform(const std::string &k, const Transform &v, bool = true) { tr[k] = v; }
    ref<const AnimatedTransform> getAnimatedTransform(const std::string &k) const;                       /* defined by the shim that needs it */
    ref<const AnimatedTransform> getAnimatedTransform(const std::string &k, const AnimatedTransform *d) const;
    ref<const AnimatedTransform> getAnimatedTransform(const std::string &k, const Transform &d) const;
    void setAnimatedTransform(const std::string &k, const AnimatedTransform *, bool = true) {}
    Point getPoint(const std::string &k) const { return pt.at(k); }
    Point getPoint(const std::string &k, const Point &d) const { return pt.count(k) ? pt.at(k) : d; }
    Vector getVector(const std::string &k) const { return Vector(pt.at(k)); }
    Vector getVector(const std::string &k, const Vector &d) const { return pt.count(k) ? Vector(pt.at(k)) : d; }
    void setPoint(const std::string &k, const Point &v, bool = true) { pt[k] = v; }
    bool removeProperty(const std::string &k) { return f.erase(k) + s.erase(k) + b.erase(k) + i.erase(k) + sp.erase(k) + tr.erase(k) > 0; }
    void putPropertyNames(std::vector<std::string> &) const {}
    std::vector<std::string> getUnqueried() const { return std::vector<std::string>(); }
    void markQueried(const std::string &) const {}
    std::string toString() const { return m_pluginName; }
private:
    std::string m_pluginName, m_id;
    std::map<std::string, std::string> s;
    std::map<std::string, Float> f;
    std::map<std::string, int> i;
    std::map<std::string, bool> b;
    std::map<std::string, Spectrum> sp;
    std::map<std::string, Transform> tr;
    std::map<std::string, Point> pt;
};
}
