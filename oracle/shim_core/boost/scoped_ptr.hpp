/* Stand-in header (test infrastructure only, see oracle/shim_core/README). */
#pragma once
#include <memory>
namespace boost { template <typename T> class scoped_ptr : public std::unique_ptr<T> { public: using std::unique_ptr<T>::unique_ptr; }; }
