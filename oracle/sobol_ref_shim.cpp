/* Thin C entry points over the REFERENCE's own sobolseq.{h,cpp}, compiled where they lie under
 * /root/reference (never copied) into oracle/_ref/libsobolref.so by oracle/Makefile.  Used only to
 * validate the oracle's Sobol' restatement and to generate tests/golden/sobol_ref.npz. */
#include <mitsuba/mitsuba.h>
#include "sobolseq.h"
extern "C" {
float sobolref_sample(uint64_t index, uint32_t dim, uint32_t scramble) { return sobol::sampleSingle(index, dim, scramble); }
uint64_t sobolref_look_up(uint32_t m, uint32_t frame, uint32_t px, uint32_t py, uint64_t scramble) {
    return sobol::look_up(m, frame, px, py, scramble);
}
}
