/* Minimal stand-in for <mitsuba/mitsuba.h>, used ONLY to compile the reference's
 * src/samplers/sobolseq.{h,cpp} (third-party, MIT) into oracle/_ref/ -- see oracle/Makefile.
 * sobolseq.h needs exactly three things from the real header (sobolseq.h:24,57,89):
 * mitsuba::Float, ONE_MINUS_EPS_FLT/DBL and SINGLE_PRECISION.                                 */
#pragma once
#include <stdint.h>
#include <assert.h>
#include <algorithm>
#ifndef SINGLE_PRECISION
#define SINGLE_PRECISION 1
#endif
#define ONE_MINUS_EPS_FLT 0x1.fffffep-1f
#define ONE_MINUS_EPS_DBL 0x1.fffffffffffff7p-1
namespace mitsuba { typedef float Float; }
