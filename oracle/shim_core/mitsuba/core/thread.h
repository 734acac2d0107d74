/* Stand-in header (test infrastructure only, see oracle/shim_core/README). */
#pragma once
#include <mitsuba/mitsuba.h>
namespace mitsuba { class Thread : public Object { public: enum EThreadPriority { EIdlePriority = 0, ELowestPriority, ELowPriority, ENormalPriority, EHighPriority, EHighestPriority, ERealtimePriority }; }; }
