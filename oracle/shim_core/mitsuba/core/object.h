/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here.
 * Object + a small run-time class registry (the reference's class.h / object.h) so that `derivesFrom(MTS_CLASS(BSDF))` works. */
#pragma once
#include <map>
#include <string>
namespace mitsuba {
class Class {
public:
    Class(const std::string &name, bool, const std::string &super, void * = NULL, void * = NULL) : m_name(name), m_super(super) { registry()[name] = this; }
    bool derivesFrom(const Class *c) const {
        for (const Class *k = this; k;) {
            if (k == c) return true;
            std::map<std::string, Class *>::const_iterator it = registry().find(k->m_super);
            k = it == registry().end() ? NULL : it->second;
        }
        return false;
    }
    const std::string &getName() const { return m_name; }
private:
    static std::map<std::string, Class *> &registry() { static std::map<std::string, Class *> r; return r; }
    std::string m_name, m_super;
};
class Object {
public:
    virtual ~Object() {}
    void incRef() const {}
    void decRef(bool = true) const {}
    virtual std::string toString() const { return ""; }
    virtual const Class *getClass() const { return NULL; }
};
}
#define MTS_DECLARE_CLASS() static Class *m_theClass; virtual const Class *getClass() const;
#define MTS_CLASS(x) x::m_theClass
#define MTS_IMPLEMENT_CLASS(name, abstract, super) Class *name::m_theClass = new Class(#name, abstract, #super); const Class *name::getClass() const { return m_theClass; }
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super) MTS_IMPLEMENT_CLASS(name, abstract, super)
#define MTS_IMPLEMENT_CLASS_I(name, abstract, super) MTS_IMPLEMENT_CLASS(name, abstract, super)
#define MTS_IMPLEMENT_CLASS_IS(name, abstract, super) MTS_IMPLEMENT_CLASS(name, abstract, super)
