// phmm_log10.h — raw FP32 sums -> the reference's log10 doubles, a region at a time (phmm_log10.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

// v[i] = (double)(log10f(f[i]) - LOG10_INITIAL_CONSTANT_F) (pairhmm_internal.cpp:583) for i < n; returns how many f[i] are
// below MIN_ACCEPTED (1e-28f) — those take the FP64 branch (:578-580) and the caller overwrites their v[i].
// mode: 0 = libm's log10f, 1 = the fdlibm formula around libm's logf, 2 = the AVX-512 restatement (phmm_log10_pick_mode),
// 3 = the restatement one value at a time (the device epilogue's arithmetic; for checks).
size_t phmm_log10_block(const float* f, double* v, size_t n, int mode);
// The fastest mode whose bits equal this host's libm (tested here and now); PHMM_B200_LIBM_LOG10F in the environment pins 0.
int phmm_log10_pick_mode();
// For tests: inputs in [0.5, 2) (every `stride`-th float) on which the restated logf differs from libm's; UINT64_MAX without AVX-512.
uint64_t phmm_log10_logf_mismatches(uint32_t stride);
// normalize_likelihoods (:53-63), filter_poorly_modelled_evidence (:161-182) and transpose_likelihood_matrix (:184-193) of one
// region: v[nr][nh] log10 values -> out[nh][nr] with the surviving reads in front and a NaN tail, kept_idx[] = their original
// indices; returns how many survive.  mode 2 = AVX-512 (same bytes; tests/test_host_logic.py), else scalar.
uint32_t phmm_postprocess_region(const double* v, uint32_t nr, uint32_t nh, const uint32_t* read_len, double* out, uint32_t* kept_idx,
                                 int mode);
// Does the restatement of log10f the DEVICE epilogue uses (same operations, phmm_kernels.cu) give this host's libm bits?
bool phmm_log10_restated_matches_libm();
// AlleleLikelihoods::search_best_allele (genotype/allele_likelihoods.hpp:589-657; can_be_reference = true, priorities as in
// best_alleles_breaking_ties :364-383, may be null) for the first `kept` columns of lik[nh][nr]: best[k], second[k].
uint32_t phmm_best_haplotypes(const double* lik, uint32_t nr, uint32_t nh, uint32_t kept, const double* priority, uint32_t* best,
                              uint32_t* second);
