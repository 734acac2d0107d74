/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
#include <string>
namespace mitsuba {
class Class {
public:
    Class(const char *, bool, const char *, void * = NULL, void * = NULL) {}
    bool derivesFrom(const Class *) const { return false; }
    std::string getName() const { return ""; }
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
#define MTS_CLASS(x) ((const ::mitsuba::Class *) NULL)
#define MTS_IMPLEMENT_CLASS(name, abstract, super) Class *name::m_theClass = NULL; const Class *name::getClass() const { return NULL; }
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super) Class *name::m_theClass = NULL; const Class *name::getClass() const { return NULL; }
#define MTS_IMPLEMENT_CLASS_I(name, abstract, super) Class *name::m_theClass = NULL; const Class *name::getClass() const { return NULL; }
#define MTS_IMPLEMENT_CLASS_IS(name, abstract, super) Class *name::m_theClass = NULL; const Class *name::getClass() const { return NULL; }
#define MTS_EXPORT_PLUGIN(name, descr)
