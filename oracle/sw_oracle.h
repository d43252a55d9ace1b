/*
 * sw_oracle.h — CPU restatement of the Smith-Waterman aligner Rovaca runs right after the PairHMM path
 * (SURVEY.md §8f item 3): AlignmentUtils::create_read_aligned_to_ref (genotype/utils/alignment_utils.cpp:563-656) calls
 * sw->avx_function(sw) (common/base/include/smithwaterman_common.h:95) = sw_avx512_run / sw_avx2_run
 * (common/base/lib/sw_avx/sw_avx.c:492-504, sw_avx2.c) to realign every read to its best haplotype.
 *
 * TEST INFRASTRUCTURE ONLY (see phmm_oracle.h).  Parity status: PINNED against the reference's own sw_avx.c / sw_avx2.c,
 * compiled in place into oracle/_ref (oracle/Makefile) — tests/test_sw_oracle.py demands identical score, alignment
 * offset and CIGAR; vectors from the unmodified reference are committed as tests/golden/sw.npz.
 *
 * Citations are relative to /root/reference/src/haplotypecaller/common/base/.
 */
#ifndef SW_ORACLE_H_
#define SW_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sw_parameters (include/smithwaterman_common.h:52-58); the two sets the reference uses (lib/sw_avx/sw_avx.c:13-14):
 * assembly {200,-150,-260,-11}, read-to-haplotype realignment {10,-15,-30,-5}. */
typedef struct {
    int32_t match, mismatch, gap_open, gap_extend;
} orc_sw_params;

/* One alignment of seq2 (the read, len2 columns) against seq1 (the haplotype, len1 rows): the anti-diagonal Gotoh fill
 * with a 4-bit back-track code per cell (sw_avx.c:40-76,78-224), best cell on the last row / last column with the
 * reference's tie rules (:196-217), trace-back and CIGAR compaction (:252-361).  cigar[k] = length << 4 | op with op
 * 0 = M, 1 = I, 2 = D, 4 = S (include/hc_marco.h:22-26).  Returns the number of CIGAR elements written (<= cigar_cap),
 * or -1 when max(len1, len2) >= 2048, where the reference gives up (sw_avx.c:496-499: alignment_offset = 1, no CIGAR). */
int32_t orc_sw_align(const uint8_t* seq1, int32_t len1, const uint8_t* seq2, int32_t len2, const orc_sw_params* params,
                     uint32_t* cigar, int32_t cigar_cap, int32_t* alignment_offset, int32_t* score);

#ifdef __cplusplus
}
#endif
#endif
