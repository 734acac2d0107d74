/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
namespace mitsuba { template <typename T> class ref { public: ref() : p(NULL) {} ref(T *q) : p(q) {} T *operator->() const { return p; } T *get() const { return p; } operator T *() const { return p; } std::string toString() const { return p ? p->toString() : std::string("ref[null]"); } private: T *p; }; }
#include <vector>
namespace mitsuba { template <typename T> class ref_vector : public std::vector<ref<T> > { public: ref_vector() {} ref_vector(size_t n) : std::vector<ref<T> >(n) {} void ensureUnique() {} bool contains(const T *o) const { for (size_t i = 0; i < this->size(); ++i) if ((*this)[i].get() == o) return true; return false; } }; }
