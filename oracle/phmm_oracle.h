/*
 * phmm_oracle.h — CPU restatement of Rovaca's PairHMM likelihood path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker or the reported CPU baseline.  The product path (rovaca_b200/) never
 * links, imports or calls it and has no CPU fallback.
 *
 * Parity status: PINNED.  The reference ships no PairHMM golden vectors, so
 * this restatement is pinned against the reference itself: oracle/Makefile compiles the
 * reference's own kernel TUs, host wrapper (pairhmm_internal.cpp) and ReadRecord from
 * /root/reference *in place* into oracle/_ref/, and tests/test_oracle_vs_ref.py demands bitwise
 * equality; the vectors so produced are committed under tests/golden/ with their generator.
 *
 * All citations are relative to /root/reference/src/haplotypecaller/pairhmm/.
 */
#ifndef PHMM_ORACLE_H_
#define PHMM_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* PcrIndelModel, ../common/enum.h:8 */
enum { ORC_PCR_NONE = 0, ORC_PCR_HOSTILE = 1, ORC_PCR_AGGRESSIVE = 2, ORC_PCR_CONSERVATIVE = 3 };

/* Arithmetic variant of the FP32 recurrence:
 *   ORC_FMA_AVX512 — the contraction GCC applies to the reference's AVX-512 TUs (default path)
 *   ORC_FMA_NONE   — the un-contracted AVX2 TU (core/avx2_s.cpp) */
enum { ORC_FMA_AVX512 = 1, ORC_FMA_NONE = 0 };

/* Build the lookup tables once (core/context.h:13-190).  Idempotent, thread-safe after return. */
void orc_init(void);

/* Table accessors (for uploading the very same numbers to the GPU in tests, and for checks). */
const float*  orc_ph2pr_f32(void);          /* [128]   core/context.h:144-146 */
const double* orc_ph2pr_f64(void);          /* [128]   core/context.h:102-104 */
const float*  orc_mm_tri_f32(void);         /* [8256]  matchToMatchProb, max qual 127; core/context.h:49-59 */
const double* orc_mm_tri_f64(void);         /* [8256] */

/* One pair, raw forward sum.  Inputs are the *prepared* per-base arrays (already quality-capped
 * and PCR-adjusted); every quality is used as x & 127 (pairhmm_internal.cpp:466-469).
 * Base matching follows the one-hot rule of the default path (pairhmm_internal.cpp:27-41):
 * A,C,G,T match themselves, N matches anything, any other byte matches nothing. */
float  orc_pair_f32(const uint8_t* hap, uint32_t hap_len, const uint8_t* rs, const uint8_t* q,
                    const uint8_t* ins, const uint8_t* del, const uint8_t* gcp, uint32_t read_len,
                    int fma_mode);
double orc_pair_f64(const uint8_t* hap, uint32_t hap_len, const uint8_t* rs, const uint8_t* q,
                    const uint8_t* ins, const uint8_t* del, const uint8_t* gcp, uint32_t read_len);

/* Underflow policy + log10 conversion (pairhmm_internal.cpp:578-584, 674-689). */
double orc_log10_likelihood(float raw_f32, double raw_f64);

/* Read preparation (pairhmm_internal.cpp:434-443, 545-557). */
void orc_cap_quals(uint8_t* qual, uint32_t n, uint8_t mapq, int32_t min_quality_threshold);
int32_t orc_find_tandem_repeat_units(const uint8_t* bases, uint32_t n, uint32_t offset);   /* :65-119 */
void orc_apply_pcr_error_model(const uint8_t* bases, uint8_t* gap_qual, uint32_t n, int pcr_model); /* :121-159 */

/* A read / haplotype as the path sees it (what ReadRecord / Haplotype expose to it). */
typedef struct {
    const uint8_t* bases;   /* ASCII, read_record.cpp:176-182 */
    const uint8_t* qual;    /* raw phred, read_record.cpp:184-189 */
    const uint8_t* ins;     /* NULL -> 45 ('-'), read_record.cpp:191-196 */
    const uint8_t* del;     /* NULL -> aliases the PCR-adjusted ins array, pairhmm_internal.cpp:425,559 */
    const uint8_t* gcp;     /* NULL -> 10 ('+'-33), read_record.cpp:205-210 */
    uint32_t len;
    uint8_t mapq;
} orc_read;

typedef struct {
    const uint8_t* bases;
    uint32_t len;
} orc_hap;

/* Whole-region driver = call_pairhmm (pairhmm_internal.cpp:505-594 / 596-753; both give the same
 * numbers).  out is [n_haps][n_reads] row-major doubles; on return the first
 * *n_kept columns of every row are valid (in input read order, filtered reads removed), the
 * tail columns are NaN (pairhmm_internal.cpp:516,589-591).  kept_idx[0..*n_kept) receives the
 * original indices of the surviving reads.  raw_f32 / raw_f64 (optional, [n_reads][n_haps])
 * receive the per-pair raw sums; raw_f64 is NaN where the FP64 branch was not taken.
 * Returns the number of pairs that took the FP64 branch. */
int64_t orc_region(const orc_read* reads, uint32_t n_reads, const orc_hap* haps, uint32_t n_haps,
                   int32_t min_quality_threshold, int pcr_model, int fma_mode, double* out,
                   uint32_t* kept_idx, uint32_t* n_kept, float* raw_f32, double* raw_f64);

/* AlleleLikelihoods::search_best_allele (genotype/allele_likelihoods.hpp:589-657) as best_alleles_breaking_ties (:364-383)
 * calls it for every evidence of a sample: can_be_reference = true, one priority per allele (may be NULL = no
 * priorities), log10 likelihoods (informative threshold 0.2, :27).  values is the [n_alleles][stride] matrix call_pairhmm
 * returned; evidence runs over its first n_evidence columns.  Per evidence: best / second-best allele index, their
 * likelihoods (second = -inf when both indices are equal, :651-652) and BestAllele's confidence (:59-62). */
void orc_best_alleles(const double* values, uint32_t n_alleles, uint32_t stride, uint32_t n_evidence, const double* priorities,
                      uint32_t* best, uint32_t* second, double* best_lik, double* second_lik, double* confidence);

/* AlleleLikelihoods::marginal_likelihoods_direct (genotype/allele_likelihoods.hpp:702-732) for one sample:
 * new_values[n][r] starts at -inf and takes every old_values[a][r] that is greater, a over the old alleles whose
 * old2new_idx_map entry is n (entries < 0 = s_missing_index are skipped). */
void orc_marginalize(const double* old_values, uint32_t old_allele_count, uint32_t stride, uint32_t evidence_count,
                     const int32_t* old2new_idx_map, uint32_t new_allele_count, double* new_values);

#ifdef __cplusplus
}
#endif
#endif
