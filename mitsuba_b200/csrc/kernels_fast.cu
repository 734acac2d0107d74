// Throughput build of the wavefront kernels: FMA contraction on (nvcc default).  See b2_kernels.inl.
#define B2_KNS fast
#define B2_FAST_TRI 1   // plane-form triangle test (b2_trace.cuh: triPlaneIntersect)
#include "b2_kernels.inl"
