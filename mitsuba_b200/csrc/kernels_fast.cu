// Throughput build of the wavefront kernels: FMA contraction on (nvcc default).  See b2_kernels.inl.
#define B2_KNS fast
#ifndef B2_NO_FAST_TRI   // (A/B switch: the throughput build with the TriAccel test)
#define B2_FAST_TRI 1   // plane-form triangle test (b2_trace.cuh: triPlaneIntersect)
#endif
#include "b2_kernels.inl"
