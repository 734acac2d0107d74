// Throughput build of the wavefront kernels: FMA contraction on (nvcc default).  See b2_kernels.inl.
#define B2_KNS fast
#include "b2_kernels.inl"
