/* Stand-in header (test infrastructure only, see oracle/shim_core/README): scaffolding that lets the reference's own sources and
 * headers compile where they lie under /root/reference, without boost or the rest of libcore.  No algorithm lives here. */
#pragma once
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/ray.h>
#include <mitsuba/core/aabb.h>
#include <mitsuba/core/triangle.h>
