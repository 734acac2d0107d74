/* stand-in for <mitsuba/core/statistics.h>: counters that count nothing */
#pragma once
#include <mitsuba/mitsuba.h>
namespace mitsuba {
enum EStatsType { ENumberValue = 0, EByteCount, EPercentage, EAverage };
class StatsCounter {
public:
    StatsCounter() {}
    StatsCounter(const char *, const char *, EStatsType = ENumberValue) {}
    StatsCounter &operator+=(size_t) { return *this; }
    StatsCounter &operator++() { return *this; }
    void incrementBase(size_t = 1) {}
};
}
