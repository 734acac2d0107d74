/* stand-in: everything lives in the stand-in <mitsuba/mitsuba.h> */
#pragma once
#include <mitsuba/mitsuba.h>
