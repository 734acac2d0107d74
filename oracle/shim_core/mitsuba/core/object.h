/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
namespace mitsuba {
class Class { public: Class(const char *, bool, const char *) {} };
class Object { public: virtual ~Object() {} void incRef() const {} void decRef() const {} virtual std::string toString() const { return ""; } };
}
#define MTS_DECLARE_CLASS() static Class *m_theClass; virtual const Class *getClass() const;
#define MTS_IMPLEMENT_CLASS(name, abstract, super)
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super)
#define MTS_EXPORT_PLUGIN(name, descr)
