// phmm_kernels.h — launch interface between the host engine (C++) and the CUDA kernels.
#ifndef PHMM_KERNELS_H_
#define PHMM_KERNELS_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "phmm_types.h"

// Device addresses of one batch (see phmm_types.h for the layout).
struct PhmmBatchDev
{
    const PhmmRead* reads;
    const PhmmHap* haps;
    const PhmmChunk* chunks;
    const PhmmCta* ctas;
    const uint32_t* task_reads;
    const uint8_t* read_bases;
    const uint8_t* read_qual;
    const uint8_t* read_ins;   // may be null
    const uint8_t* read_del;   // may be null
    const uint8_t* read_gcp;   // may be null
    const uint8_t* hap_bases;
    uint32_t* words;
    float* out_f32;
    double* out_f64;
    PhmmPair* f64_pairs;       // pairs whose FP32 sum fell below 1e-28f (filled by the forward kernels)
    uint32_t* f64_count;
    const PhmmPair* slow_pairs;   // out-of-envelope pairs routed to the per-pair kernel by the host
    const PhmmTables* tables;
    uint32_t n_reads;
    float rg_gapm, rg_ii;         // reference-gap variant: 1 - ph2pr[gcp] and ph2pr[gcp] of the launch-wide gcp (10)
};

// Device addresses of the epilogue's inputs and outputs (phmm_epilogue_kernel).
struct PhmmEpilogueDev
{
    const PhmmRegionDev* regions;
    const PhmmRead* reads;
    const float* raw;             // out_f32
    const double* hap_priority;   // per haplotype of the batch, or null (best-haplotype tie-break, allele_likelihoods.hpp:624-645)
    double* lik;                  // per region [n_haps][n_reads] at out_off: surviving reads in front, NaN tail
    uint32_t* kept_idx;           // per region at read_begin: original index of every surviving read
    uint32_t* n_kept;             // per region
    uint32_t* n_under;            // per region: pairs whose raw sum is below MIN_ACCEPTED (the FP64 branch, :578)
    uint32_t* best;               // per region at read_begin, or null: best / second-best haplotype of every surviving read
    uint32_t* second;
    uint32_t n_regions;
};
cudaError_t phmm_launch_epilogue(const PhmmEpilogueDev& e, cudaStream_t st);
// one variant site for phmm_marginalize_kernel: which region's matrix, how many alleles come out, where its haplotype -> allele
// map starts in the concatenated maps, where its [n_new][n_reads] output starts
struct PhmmMargSite
{
    uint32_t region, n_new, map_off, rsv;
    uint64_t out_off;
};
cudaError_t phmm_launch_marginalize(const PhmmEpilogueDev& e, const PhmmMargSite* sites, const int32_t* maps, uint32_t n_sites, double* out,
                                    cudaStream_t st);
// for tests: out[i] = (double)(log10f(x[i]) - LOG10_INITIAL_CONSTANT_F) as the device computes it
cudaError_t phmm_launch_log10_probe(const float* x, double* out, uint32_t n, cudaStream_t st);

// The forward kernel comes in one instantiation per rows-per-lane count RPT (and arithmetic variant): a read of `len` bases
// takes ceil((len + 1) / RPT) consecutive lanes of a warp (at least one boundary row on top), so any read up to 32 * RPT - 1
// bases fits, and a warp carries whatever mix of reads packs into its 32 lanes.  More rows per lane = fewer shuffles and
// shared-memory loads per cell ("kernel budget"); fewer rows per lane = more lanes per read = a shorter wavefront
// step, which is what a small batch on an idle device wants.  The planner picks RPT per region (phmm_plan.cpp).
#define PHMM_VAR_GENERIC_NOFMA 0
#define PHMM_VAR_GENERIC 1
#define PHMM_VAR_REFGAPS 2

// CTAs of PHMM_WARPS warps the register file of an SM holds of the instantiation with `rpt` rows per lane (the occupancy
// targets the kernels are compiled for, __launch_bounds__ in phmm_kernels.cu)
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr int phmm_fwd_blocks_per_sm(int rpt, bool refgaps)
{
    // (reference gaps, 4-8 rows: layouts of the latency policy only — a batch too small to fill the device — whose kernels keep
    // the next step's stream word and priors in registers: they get the registers rather than the occupancy)
    return refgaps ? (rpt <= 3 ? 6 : rpt <= 6 ? 4 : rpt <= 8 ? 3 : rpt == 9 ? 4 : rpt <= 12 ? 3 : 2)
                   : (rpt <= 3 ? 6 : rpt <= 5 ? 5 : (rpt == 6 || rpt == 8) ? 4 : rpt <= 10 ? 3 : 2);
}
static inline int phmm_max_rpt(bool refgaps) { return refgaps ? PHMM_MAX_RPT : PHMM_MAX_RPT_GENERIC; }
static inline uint32_t phmm_read_lanes(uint32_t len, int rpt) { return (len + (uint32_t)rpt) / (uint32_t)rpt; }   // ceil((len + 1) / rpt)
// longest read the wavefront kernel takes (longer ones go to the per-pair kernel)
static inline uint32_t phmm_max_read_len(bool refgaps) { return 32u * (uint32_t)phmm_max_rpt(refgaps) - 1u; }

// Prior tables in shared memory: [6 base classes][thread][ceil(RPT/4) float4].  The per-thread slice is padded to an odd number
// of 16-byte chunks so that the eight chunks a quarter warp fetches with one LDS.128 fall into eight different bank groups
// whatever classes the lanes ask for (a class offset is a multiple of 512 B).
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr uint32_t phmm_prior_table_stride(int rpt) { return (uint32_t)(((rpt + 3) / 4) | 1) * 16u; }

// Shared memory one CTA of the forward kernel needs for a chunk: the two column streams, the per-haplotype seam records and
// the per-thread prior tables (6 base classes x ceil(RPT/4) float4 per thread).
// (paired reads — two reads in the halves of every lane — keep a second set of tables)
static inline size_t phmm_fwd_smem_bytes(uint32_t stream_len, uint32_t n_haps, int rpt, uint32_t cta_warps, bool duo = false)
{
    const size_t streams = (size_t)(stream_len + 32) * 8;
    const size_t info = (((streams + (size_t)(n_haps + 1) * 2 * 12) + 15) & ~(size_t)15) - streams;
    const size_t tables = (size_t)(duo ? 12 : 6) * phmm_prior_table_stride(rpt) * cta_warps * 32;
    return streams + info + tables;
}
// Rows per lane of the paired-reads kernel (phmm_fwd_kernel<.., DUO>): the most efficient count of the throughput layouts; a
// region whose reads need more than 32 lanes at it keeps its odd haplotype in the two-stream chunks.
#define PHMM_DUO_RPT 12

cudaError_t phmm_launch_prep(const PhmmBatchDev& b, cudaStream_t st);
// The one fixed-lanes specialisation of the forward kernel: reads of 133..151 bases take 8 lanes at 19 rows per lane (the
// reference-gap variant); a launch in which every read does may use it (fixed_lanes = 8), else 0.
static inline uint32_t phmm_fixed_lanes(int rpt, bool refgaps) { return (refgaps && rpt == PHMM_MAX_RPT) ? 8u : 0u; }
// Forward wavefront kernel over ctas[cta_base .. cta_base + n_ctas) (all of one rows-per-lane count).
cudaError_t phmm_launch_fwd(const PhmmBatchDev& b, int rpt, bool refgaps, bool fma, uint32_t fixed_lanes, bool duo, uint32_t cta_base,
                            uint32_t n_ctas, uint32_t cta_warps, size_t smem_bytes, cudaStream_t st);
// Per-pair kernels.  count_dev != null: the pair count is read on the device (FP64 recompute list).
cudaError_t phmm_launch_pairs_f32(const PhmmBatchDev& b, bool fma, const PhmmPair* pairs, uint32_t count, float* scratch,
                                  uint32_t threads, uint32_t max_cols, cudaStream_t st);
// FP64 recompute of the pairs listed on the device.  redo / redo_count (a second list of the same capacity and its counter,
// zeroed by the caller): the fast kernel takes the common shapes and hands the rest to the general one; null: the general
// kernel takes everything.  gapm_d / gcp_d: 1 - ph2pr_d[gcp] and ph2pr_d[gcp] of the default gap-continuation quality.
// trip_hi: high word of the smallest M the fast kernel accepts (PHMM_PAIR64_TRIP_HI = 2^-830; tests raise it to force hand-backs).
#define PHMM_PAIR64_TRIP_HI 0x0C100000u
cudaError_t phmm_launch_pairs_f64(const PhmmBatchDev& b, const PhmmPair* pairs, const uint32_t* count_dev, double* scratch,
                                  uint32_t threads, uint32_t max_cols, PhmmPair* redo, uint32_t* redo_count, double gapm_d, double gcp_d,
                                  uint32_t trip_hi, cudaStream_t st);
// Static facts for reports (registers, spills) — filled from cudaFuncGetAttributes.
struct PhmmKernelInfo
{
    int regs;
    int local_bytes;
    int max_threads;
};
cudaError_t phmm_fwd_kernel_info(int rpt, bool refgaps, bool fma, PhmmKernelInfo* info);

// ---- batched Smith-Waterman (sw_kernels.cu) -------------------------------------------------------------------------
struct PhmmSwTask
{
    uint32_t ref_off, ref_len;     // haplotype (rows) in the byte blob
    uint32_t read_off, read_len;   // read (columns)
    uint64_t bt_off;               // first back-track word of this alignment, [step = row - 1 + lane][lane][ceil(cpt/8)] u32
    uint64_t aux_off;              // last row (read_len ints) then last column (ref_len ints) of H
    uint32_t cpt;                  // columns per lane of the fill kernel that serves it
    uint32_t narrow;               // 1: its scores fit 16 bits, the two-alignments-per-warp kernel takes it
};
struct PhmmSwOut
{
    uint32_t n_cigar;
    int32_t alignment_offset, score;
    int32_t status;                // 0 ok, 2 CIGAR longer than the caller's buffer, 3 corrupt back-track (never)
};
struct PhmmSwBatchDev
{
    const PhmmSwTask* tasks;
    const uint32_t* order;         // task ids grouped by cpt
    const uint8_t* bytes;
    uint32_t* bt;
    int32_t* aux;
    uint32_t* cigars;              // [task][cigar_cap]
    PhmmSwOut* outs;
    uint32_t cigar_cap;
    int32_t match, mismatch, gap_open, gap_extend;
};
static const int kPhmmSwCpt[] = {2, 5, 8, 16, 32, 64};     // columns per lane: reads up to 64, 160, 256, 512, 1024, 2048 bases
static const int kPhmmSwNumCpt = sizeof(kPhmmSwCpt) / sizeof(kPhmmSwCpt[0]);
cudaError_t phmm_launch_sw_fill(const PhmmSwBatchDev& b, int cpt, uint32_t first, uint32_t count, uint32_t warps, cudaStream_t st);
// the same for tasks whose scores fit 16 bits (PhmmSwTask.narrow): two alignments per warp in packed 16-bit halves
cudaError_t phmm_launch_sw_fill16(const PhmmSwBatchDev& b, int cpt, uint32_t first, uint32_t count, uint32_t warps, cudaStream_t st);
// a task may take the 16-bit kernel if min(rows, columns) x max(|match|, |mismatch|) + |open| + |extend| stays under this
static const int kPhmmSwNarrowBound = 12000;
cudaError_t phmm_launch_sw_trace(const PhmmSwBatchDev& b, uint32_t first, uint32_t count, cudaStream_t st);

// FP32 issue-rate micro-benchmark; each thread executes iters*32 FMA instructions on 2 lanes (see .cu).
cudaError_t phmm_launch_ubench(int mode, uint32_t blocks, uint32_t iters, float* out, cudaStream_t st);

#endif
