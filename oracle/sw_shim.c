/*
 * sw_shim.c — C door into the UNMODIFIED reference Smith-Waterman (common/base/lib/sw_avx/sw_avx.c, sw_avx2.c, compiled in
 * place by oracle/Makefile).  TEST INFRASTRUCTURE ONLY (see sw_oracle.h).  Glue only: fills lib_sw_avx the way
 * AlignmentUtils::create_read_aligned_to_ref does (genotype/utils/alignment_utils.cpp:574-579) and copies the result out.
 */
#include <string.h>

#include "smithwaterman_common.h"

/* isa: 0 = sw_avx2_run, 1 = sw_avx512_run, -1 = what sw_avx_init binds on this CPU.  is_assemble picks the parameter set
 * (sw_avx.c:13-14, 432-436).  Returns the CIGAR length, -1 when the reference gave up (sequence >= 2048), -2 on bad use. */
int ref_sw_align(int isa, int is_assemble, const uint8_t* seq1, int len1, const uint8_t* seq2, int len2, uint32_t* cigar, int cigar_cap,
                 int* alignment_offset, int* score, int* bound_isa)
{
    if (len1 < 0 || len2 < 0 || len1 > 2048 || len2 > 2048) return -2;
    p_lib_sw_avx sw = sw_avx_init((uint8_t)is_assemble);
    if (!sw) return -2;
    if (bound_isa) *bound_isa = sw->avx_function == sw_avx512_run ? 1 : 0;
    sw->len1 = (int16_t)len1;
    memcpy(sw->seq1, seq1, (size_t)len1);
    sw->len2 = (int16_t)len2;
    memcpy(sw->seq2, seq2, (size_t)len2);
    sw->alignment_offset = -7;
    sw->cigar_count = 0;
    if (isa == 1) sw_avx512_run(sw); else if (isa == 0) sw_avx2_run(sw); else sw->avx_function(sw);
    int n = (int)sw->cigar_count;
    *alignment_offset = sw->alignment_offset;
    *score = sw->score;
    if ((len1 > len2 ? len1 : len2) >= 2048) n = -1;
    else {
        if (n > cigar_cap) n = cigar_cap;
        memcpy(cigar, sw->cigar, (size_t)n * sizeof(uint32_t));
    }
    sw_avx_finit(sw);
    return n;
}
