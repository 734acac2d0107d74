#pragma once
#include <mitsuba/mitsuba.h>
namespace mitsuba { class Timer : public Object { public: unsigned int getMilliseconds() const { return 0; } }; }
