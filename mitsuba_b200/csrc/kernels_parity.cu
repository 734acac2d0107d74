// Parity build of the wavefront kernels: compiled with -fmad=false so that every a*b+c rounds twice,
// as in the IEEE-strict reading of the reference source.  See b2_kernels.inl.
#define B2_KNS parity
#include "b2_kernels.inl"
