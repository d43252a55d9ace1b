/*
 * phmm_b200.h — C ABI of the B200-native PairHMM likelihood engine (drop-in for Rovaca's call_pairhmm).
 *
 * Every entry point cites the reference interface it replaces; paths are relative to
 * /root/reference/src/haplotypecaller/.  The library is libphmm_b200.so (rovaca_b200/); it needs an
 * NVIDIA GPU of compute capability 10.0 (sm_100a) and has NO CPU fallback: without a device every
 * call fails with PHMM_ERR_CUDA and phmm_last_error() says why.
 *
 * Results are bit-identical to the reference's CPU path on the same inputs: the raw forward sums
 * match rovaca/avx_512_float.cc / core/avx512_s.cpp (or core/avx2_s.cpp with PHMM_FMA_NONE) and the
 * post-processed matrix matches call_old_pairhmm / call_new_pairhmm002 byte for byte, including
 * the NaN tail and the set of filtered reads.
 */
#ifndef PHMM_B200_H_
#define PHMM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHMM_OK 0
#define PHMM_ERR_ARG (-1)      /* bad argument */
#define PHMM_ERR_CUDA (-2)     /* CUDA runtime / no usable device */
#define PHMM_ERR_MEM (-3)      /* host or device allocation failed */
#define PHMM_ERR_STATE (-4)    /* unknown ticket, engine shut down */

/* PcrIndelModel — common/enum.h:8 (same numeric values). */
#define PHMM_PCR_NONE 0
#define PHMM_PCR_HOSTILE 1
#define PHMM_PCR_AGGRESSIVE 2
#define PHMM_PCR_CONSERVATIVE 3

/* Which compiled flavour of the reference's FP32 kernel to reproduce bit for bit:
 * the FMA-contracted AVX-512 translation units (what init_pairhmm_ptr binds on an AVX-512 host,
 * pairhmm/pairhmm_engine.cpp:37-52) or the un-contracted AVX2 one (core/avx2_s.cpp). */
#define PHMM_FMA_DEFAULT 0   /* = PHMM_FMA_AVX512 */
#define PHMM_FMA_AVX512 1
#define PHMM_FMA_NONE 2

typedef struct phmm_engine phmm_engine;

/* Replaces: PairhmmArgument (haplotypecaller_args.h:35-39) + the one-time binding done by
 * rovaca::init_pairhmm_ptr (pairhmm/pairhmm_engine.cpp:21-53).  Zero-initialise for defaults. */
typedef struct phmm_config
{
    int32_t device;               /* CUDA ordinal; -1 = current device */
    int32_t fma_mode;             /* PHMM_FMA_* (0 = the AVX-512 flavour) */
    uint32_t n_workers;           /* host worker threads = batches in flight, per device (0 = cores / visible GPUs, clamped to 2..12) */
    uint64_t batch_cells;         /* seal a batch at this many read x haplotype cells (0 = 32e9) */
    uint32_t linger_us;           /* phmm_wait lets an open batch collect regions this long (0 = 200) */
    uint32_t cta_warps;           /* warps per forward-kernel CTA, 1..4; 0 = chosen per launch; tuning knob */
    uint32_t pipeline_depth;      /* batches phmm_submit/phmm_wait keep in flight before they start batching up (0 = 6) */
    uint32_t lanes_per_read;      /* 8, 16 or 32 lanes per read in the forward kernel; 0 = by batch size (small batches on an idle
                                     device get the wide, low-latency layouts); tuning / test knob */
    uint32_t fp64_mode;           /* PHMM_FP64_*: when pairs whose FP32 sum underflows (< 1e-28f) are recomputed in FP64 */
    uint32_t epilogue;            /* PHMM_EPILOGUE_*: where log10 / normalize / filter / transpose / best-haplotype search run */
    uint32_t rsv[4];
} phmm_config;

/* The reference's host side after its kernels (pairhmm_internal.cpp:578-584, :53-63, :161-182, :184-193) and the per-read
 * best-haplotype search the next step opens with (genotype/allele_likelihoods.hpp:589-657) run either on the host or in
 * one more kernel on the device.  log10f is whatever the HOST's libm says, so the device (and the host's vector path) use a
 * restatement of glibc's algorithm that is tested bit for bit against this host's libm when the engine is created; AUTO
 * takes the device only if that test passes and only for batches of 32 768 pairs or more (a smaller batch — the few regions
 * synchronous callers have in flight — is done sooner by the host's vector epilogue than one more launch takes); DEVICE
 * takes it for every batch and fails phmm_create if the test does not pass. */
#define PHMM_EPILOGUE_AUTO 0u
#define PHMM_EPILOGUE_HOST 1u
#define PHMM_EPILOGUE_DEVICE 2u

/* pairhmm_internal.cpp:578-584 recomputes an underflowing pair in FP64.  The value it gets is below -64 (log10) and never
 * reaches the returned matrix: filter_poorly_modelled_evidence drops the read (best < -8, or -4 under 51 bp) or
 * normalize_likelihoods caps the entry at best - 4.5 (SURVEY.md §8a "Consequence of a11-a14"); the reference's own FP64
 * kernels return 0 for every such pair (a12) and its matrix is the same.  So by default the recompute runs only for a
 * caller that can see the value (phmm_result.raw_f64 / phmm_flat_result.raw_f64). */
#define PHMM_FP64_WHEN_OBSERVABLE 0u
#define PHMM_FP64_ALWAYS 1u

/* What the path reads of a ReadRecord (genotype/read_record.h:67-103, read_record.cpp:176-210). */
typedef struct phmm_read
{
    const uint8_t* bases;         /* ASCII, decode_base() */
    const uint8_t* qual;          /* raw phred, qual() / decode_qual() */
    const uint8_t* ins;           /* insertion GOP; NULL = 45 ('-'), ins_gops() */
    const uint8_t* del;           /* deletion GOP; NULL = the PCR-adjusted insertion array, as the
                                     reference aliases it (pairhmm/pairhmm_internal.cpp:425,559) */
    const uint8_t* gcp;           /* gap continuation; NULL = 10 ('+'-33), gap_conts() */
    uint32_t len;                 /* seq_length() */
    uint8_t mapq;                 /* mapping_quality() */
} phmm_read;

/* What the path reads of a Haplotype (genotype/allele.h:104,107): get_display_string(), length(). */
typedef struct phmm_hap
{
    const uint8_t* bases;
    uint32_t len;
    double priority;              /* only read when best_hap is asked for: breaks ties between haplotypes whose likelihoods
                                     are within 0.2 of the best (allele_likelihoods.hpp:624-645); the reference passes
                                     (is_reference ? 1 : 0) + (1 - cigar elements), assembly_based_caller_utils.cpp:150-155 */
} phmm_hap;

/* One call of rovaca::call_pairhmm(hs, rs, min_quality_threshold, pcr_option, pool)
 * (pairhmm/pairhmm_engine.h:14-15). */
typedef struct phmm_region
{
    const phmm_read* reads;
    uint32_t n_reads;
    const phmm_hap* haps;
    uint32_t n_haps;
    int32_t min_quality_threshold;    /* -Q, default 18 (rovaca_tool_args.h:194) */
    int32_t pcr_model;                /* PHMM_PCR_*, default CONSERVATIVE (rovaca_tool_args.h:199) */
} phmm_region;

/* The DoubleVector2D call_pairhmm returns + its side effect on `rs`, in caller-owned memory
 * (the adapter builds the pmr vectors from it). */
typedef struct phmm_result
{
    double* likelihoods;   /* [n_haps][n_reads] row-major.  First n_kept columns of each row valid,
                              tail columns NaN (pairhmm_internal.cpp:516,589-591) */
    uint32_t* kept_idx;    /* [n_reads]: original indices of the reads filter_poorly_modelled_evidence
                              (pairhmm_internal.cpp:161-182) left in `rs`, in order */
    uint32_t n_kept;       /* out */
    float* raw_f32;        /* optional [n_reads][n_haps]: raw forward sums (dump-level parity) */
    double* raw_f64;       /* optional [n_reads][n_haps]: FP64 recompute, NaN where not taken */
    uint32_t n_fp64;       /* out: pairs that took the FP64 branch (pairhmm_internal.cpp:578) */
    uint32_t* best_hap;    /* optional [n_reads], first n_kept valid: for the k-th surviving read the haplotype
                              AlleleLikelihoods::search_best_allele (genotype/allele_likelihoods.hpp:589-657, as called by
                              best_alleles_breaking_ties :364-383 with phmm_hap.priority) names best ... */
    uint32_t* second_hap;  /* ... and second best; required when best_hap is given.  BestAllele's likelihood, second-best
                              likelihood and confidence are likelihoods[best][k], likelihoods[second][k] (-inf if the two
                              indices are equal) and their difference (:651-656) */
} phmm_result;

/* Many regions at once, flat arrays (what a batching caller / the bench hands over).  Region g owns
 * reads [reg_read_off[g], reg_read_off[g+1]) and haplotypes [reg_hap_off[g], reg_hap_off[g+1]);
 * read k has its bytes at read_bases + read_off[k] (same offset in every read blob). */
typedef struct phmm_flat_batch
{
    uint32_t n_regions;
    const uint64_t* reg_read_off;     /* [n_regions + 1] */
    const uint64_t* reg_hap_off;      /* [n_regions + 1] */
    const uint32_t* read_len;
    const uint8_t* read_mapq;
    const uint64_t* read_off;
    const uint8_t* read_bases;
    const uint8_t* read_qual;
    const uint8_t* read_ins;          /* NULL = default */
    const uint8_t* read_del;          /* NULL = alias */
    const uint8_t* read_gcp;          /* NULL = default */
    const uint32_t* hap_len;
    const uint64_t* hap_off;
    const uint8_t* hap_bases;
    int32_t min_quality_threshold;
    int32_t pcr_model;
    uint32_t flags;                   /* PHMM_FLAT_* */
    const double* hap_priority;       /* optional, per haplotype: phmm_hap.priority */
} phmm_flat_batch;

/* The byte blobs (read_bases, read_qual, read_ins/del/gcp, hap_bases) live in page-locked host memory (cudaHostAlloc,
 * cudaHostRegister, a pinned torch tensor): the engine copies them to the device straight from there instead of gathering
 * them into its own staging arena first. */
#define PHMM_FLAT_PINNED 1u

/* Region g writes its [n_haps][n_reads] matrix at likelihoods + lik_off[g] (lik_off[g] = sum over
 * earlier regions of n_haps*n_reads), its kept indices at kept_idx + reg_read_off[g]. */
typedef struct phmm_flat_result
{
    double* likelihoods;
    const uint64_t* lik_off;          /* [n_regions + 1] */
    uint32_t* kept_idx;               /* [total reads] */
    uint32_t* n_kept;                 /* [n_regions] */
    float* raw_f32;                   /* optional, [read][hap] per region at lik_off[g] */
    double* raw_f64;                  /* optional */
    uint64_t n_fp64;                  /* out */
    uint32_t* best_hap;               /* optional, [total reads] at reg_read_off[g]: phmm_result.best_hap */
    uint32_t* second_hap;             /* optional, with best_hap */
} phmm_flat_result;

typedef struct phmm_stats
{
    uint64_t regions, pairs, cells;   /* processed so far */
    uint64_t fp64_pairs;              /* pairs recomputed in FP64 */
    uint64_t slow_pairs;              /* pairs outside the wavefront kernel's envelope (per-pair kernel) */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t batches;
    double gpu_ms;                    /* device time with kernels of the path running: the UNION of the batches' kernel spans
                                         (CUDA events; batches of different workers overlap on the device), summed over devices */
    uint64_t kernel_launches;
    double host_plan_ms;              /* sum over batches of host planning + packing time (worker threads, wall) */
    double host_epilogue_ms;          /* sum over batches of the host epilogue (log10, cap, filter, transpose), or of the
                                         hand-over of the device epilogue's output */
    uint64_t device_epilogue_batches; /* batches whose epilogue ran on the device (PHMM_EPILOGUE_*) */
} phmm_stats;

/* Replaces rovaca::init_pairhmm_ptr (pairhmm/pairhmm_engine.cpp:21): builds the libm tables of
 * core/context.h:13-190 on the host, uploads them, starts the workers. */
int phmm_create(const phmm_config* cfg, phmm_engine** out);
void phmm_destroy(phmm_engine* e);

/* One engine in front of several GPUs of the box.  The reference is ONE multi-threaded process that binds the path once
 * (haplotypecaller.cpp:49) and calls it from its pool workers (:97, haplotypecaller_engine.cpp:250); with this form such a
 * process uses every GPU: each sealed batch of regions (phmm_submit / phmm_wait) and each sub-batch of a flat stream
 * (phmm_run_flat) goes to the device with the least outstanding work, the results land in the caller's host buffers as with
 * one device.  Regions are independent, so there is no collective and no peer access, and the result
 * of a region does not depend on the device that computed it.  devices[k] are CUDA ordinals (-1 = the current device; an
 * ordinal may be listed twice: two independent queues on one GPU); n_devices = 0 means every visible device.  cfg->device is
 * ignored, cfg->n_workers counts per device (0 = cores / max(devices used, devices visible), clamped to 2..12). */
int phmm_create_multi(const phmm_config* cfg, const int32_t* devices, uint32_t n_devices, phmm_engine** out);
uint32_t phmm_device_count(const phmm_engine* e);
int phmm_device_ordinal(const phmm_engine* e, uint32_t index);   /* CUDA ordinal of the engine's index-th device, -1 if out of range */

/* Replaces one synchronous rovaca::call_pairhmm call (haplotypecaller_engine.cpp:250).
 * = phmm_submit + phmm_wait.  Thread-safe; efficient only when many callers are in flight.
 * While the engine has room, a caller's own region is planned, launched and finished on the CALLING thread (no hand-over to
 * a worker and back); that thread's current CUDA device is switched to the engine's for the call and restored before it
 * returns. */
int phmm_call_region(phmm_engine* e, const phmm_region* region, phmm_result* result);

/* Asynchronous form of the same call, so that the caller's region loop
 * (haplotypecaller_engine.cpp:145-284) can keep the GPU full: inputs and `result` must stay valid
 * until phmm_wait(ticket) returns. */
int phmm_submit(phmm_engine* e, const phmm_region* region, phmm_result* result, uint64_t* ticket);
int phmm_wait(phmm_engine* e, uint64_t ticket);
int phmm_flush(phmm_engine* e);

/* A whole stream of regions in one call, pipelined over the workers (bench / batch callers). */
int phmm_run_flat(phmm_engine* e, const phmm_flat_batch* batch, phmm_flat_result* result);

/* AlleleLikelihoods::marginalize -> marginal_likelihoods_direct (genotype/allele_likelihoods.hpp:189-216, 702-732) on a matrix
 * phmm_call_region returned: out[n][e] = the largest values[a][e] over the old alleles a with old_to_new[a] == n, -inf when
 * there is none (old_to_new[a] < 0: allele a maps to nothing, :717).  values is [n_old][stride], out [n_new][out_stride],
 * e runs over the first n_evidence columns.  This one runs on the host, on a matrix the caller already holds;
 * phmm_resident_marginalize below does the same for many sites on matrices still on the device. */
int phmm_marginalize(const double* values, uint32_t n_old, uint32_t stride, uint32_t n_evidence, const int32_t* old_to_new,
                     uint32_t n_new, double* out, uint32_t out_stride);

/* The same for many variant sites at once, ON THE DEVICE, over the matrices a resident batch (below) holds there after
 * phmm_resident_run — for a caller that keeps the per-site haplotype -> allele maps of many regions (the genotyping step walks
 * the variant sites of a region and marginalises its read likelihoods once per site, germline_genotying_engine via
 * AlleleLikelihoods::marginalize).  Site k reads the matrix of region sites[k].region of the batch and writes
 * [n_new][n_reads of that region] doubles at out + out_off[k]: the first n_kept columns are the marginal likelihoods of the
 * surviving reads, the rest NaN like the matrix's own tail.  Needs an engine whose epilogue runs on the device
 * (PHMM_EPILOGUE_AUTO on a glibc host, or PHMM_EPILOGUE_DEVICE); otherwise PHMM_ERR_STATE: download and use phmm_marginalize. */
typedef struct phmm_marginal_site
{
    uint32_t region;              /* index of the region in the resident batch */
    uint32_t n_new;               /* alleles after marginalisation */
    const int32_t* old_to_new;    /* [n_haps of the region]: new allele of every haplotype, < 0 = maps to nothing */
} phmm_marginal_site;
struct phmm_resident;
int phmm_resident_marginalize(phmm_engine* e, struct phmm_resident* r, const phmm_marginal_site* sites, uint32_t n_sites,
                              const uint64_t* out_off, double* out);

/* Device-resident form for kernel-only timing: upload once, run the device path (prep + forward +
 * FP64 recompute kernels, no host<->device copies) any number of times, download when wanted. */
typedef struct phmm_resident phmm_resident;
int phmm_resident_upload(phmm_engine* e, const phmm_flat_batch* batch, phmm_resident** out);   /* on the engine's first device */
int phmm_resident_upload_on(phmm_engine* e, uint32_t device_index, const phmm_flat_batch* batch, phmm_resident** out);
int phmm_resident_run(phmm_engine* e, phmm_resident* r, float* gpu_ms, float* fwd_ms);
int phmm_resident_download(phmm_engine* e, phmm_resident* r, phmm_flat_result* result);
void phmm_resident_free(phmm_engine* e, phmm_resident* r);
uint32_t phmm_resident_launches(const phmm_resident* r);   /* kernels one phmm_resident_run launches */

/* ---- the step after the path: read -> best-haplotype realignment ----------------------------
 * Replaces the per-read call sw->avx_function(sw) (common/base/include/smithwaterman_common.h:95; sw_avx512_run /
 * sw_avx2_run, common/base/lib/sw_avx/sw_avx.c:492-504) that AlignmentUtils::create_read_aligned_to_ref makes
 * (genotype/utils/alignment_utils.cpp:574-579) for every read of a region — and the same call of the assembler
 * (apply/assemble/hc_assemble_cigar_cacl.c:427) — by one batched call.  Score, alignment offset and CIGAR are identical to
 * the reference's. */
typedef struct phmm_sw_params       /* sw_parameters, smithwaterman_common.h:52-58 */
{
    int32_t match, mismatch, gap_open, gap_extend;
} phmm_sw_params;
/* the reference's two parameter sets (sw_avx.c:13-14) */
#define PHMM_SW_PARAMS_REALIGN {10, -15, -30, -5}       /* sw_avx_init(0): read vs best haplotype */
#define PHMM_SW_PARAMS_ASSEMBLE {200, -150, -260, -11}  /* sw_avx_init(1): haplotype vs reference */

typedef struct phmm_sw_pair
{
    const uint8_t* ref;      /* lib_sw_avx.seq1 / len1: the haplotype (or reference) */
    uint32_t ref_len;
    const uint8_t* read;     /* lib_sw_avx.seq2 / len2: the read (or haplotype) */
    uint32_t read_len;
} phmm_sw_pair;

typedef struct phmm_sw_result
{
    uint32_t* cigar;             /* caller buffer; element = length << 4 | op, op 0 M, 1 I, 2 D, 4 S (lib_sw_avx.cigar) */
    uint32_t cigar_cap;
    uint32_t n_cigar;            /* out: lib_sw_avx.cigar_count */
    int32_t alignment_offset;    /* out: lib_sw_avx.alignment_offset */
    int32_t score;               /* out: lib_sw_avx.score */
    int32_t status;              /* out: 0 ok; 1 = a sequence of 2048 bases or more: the reference gives up with
                                    alignment_offset = 1 and no CIGAR (sw_avx.c:494-499), so does this; 2 = cigar_cap too small */
} phmm_sw_result;

/* All pairs are aligned on the device (fill kernel: one warp per alignment; trace kernel: one thread per alignment), in
 * sub-batches of 32 768 alignments on two streams so that packing, kernels and the hand-back of results overlap.  Empty
 * sequences are a PHMM_ERR_ARG.  Thread-safe (calls are serialised per engine). */
int phmm_sw_align(phmm_engine* e, const phmm_sw_pair* pairs, uint32_t n_pairs, const phmm_sw_params* params, phmm_sw_result* results);
/* Device time (CUDA events: the union of the sub-batches' fill + trace kernel spans) and cells (sum ref_len x read_len) of the
 * last phmm_sw_align call. */
int phmm_sw_last_timing(phmm_engine* e, float* gpu_ms, uint64_t* cells);

int phmm_get_stats(phmm_engine* e, phmm_stats* out);                                    /* summed over the engine's devices */
int phmm_get_device_stats(phmm_engine* e, uint32_t device_index, phmm_stats* out);      /* one device's share */
/* Static facts about the forward-kernel instantiation that serves reads of `read_len` bases (registers, spill bytes,
 * lanes per read, read rows per lane) for reports. */
int phmm_kernel_info(uint32_t read_len, int fma_mode, int* regs, int* local_bytes, int* lanes, int* rows_per_lane);
/* Measured FP32 rate of this GPU (issue-rate micro-benchmark: mode 0 scalar FFMA, 1 packed FFMA2, 2 FFMA2 + select,
 * 3 + shuffle) in FMA lane-operations per second — the empirical denominator of the FP32 roofline. */
int phmm_ubench(phmm_engine* e, int mode, uint32_t iters, float* ms, double* fma_per_s);
const char* phmm_last_error(void);   /* thread-local */
const char* phmm_version(void);

#ifdef __cplusplus
}
#endif
#endif
