/* Stand-in header (test infrastructure only, see oracle/shim_core/README). */
#pragma once
#include <mitsuba/core/thread.h>
