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
};

// Read-length bins of the wavefront kernel: a group of LANES lanes (a half warp or a whole warp) owns one read, lane l owns
// rows l*RPT .. l*RPT+RPT-1 (top-padded with at least one boundary row), so a bin holds reads of length <= LANES*RPT - 1.
// Half-warp bins come first: more rows per lane means fewer shuffles per cell (3 per RPT cells and stream), which is what
// the issue/register-read budget of the SM rewards ("kernel budget"); whole-warp bins take over where a
// half-warp read would need more than 12 rows per lane (register file).
struct PhmmBin
{
    int lanes;
    int rpt;
};
static const PhmmBin kPhmmBins[] = {{16, 1}, {16, 2}, {16, 3}, {16, 4}, {16, 5}, {16, 6}, {16, 7}, {16, 8}, {16, 9},
                                    {16, 10}, {16, 11}, {16, 12}, {32, 7}, {32, 8}, {32, 9}, {32, 10}, {32, 11}, {32, 12}};
static const int kPhmmNumBins = sizeof(kPhmmBins) / sizeof(kPhmmBins[0]);
static inline uint32_t phmm_bin_capacity(int bin) { return (uint32_t)(kPhmmBins[bin].lanes * kPhmmBins[bin].rpt) - 1u; }
static inline uint32_t phmm_bin_reads_per_cta(int bin) { return (uint32_t)(PHMM_WARPS * 32 / kPhmmBins[bin].lanes); }

// Shared memory one CTA of the forward kernel needs for a chunk: the two column streams, the per-haplotype seam records and
// the per-thread prior tables (6 base classes x ceil(RPT/4) float4 per thread).
static inline size_t phmm_fwd_smem_bytes(uint32_t stream_len, uint32_t n_haps, int bin)
{
    const size_t streams = (size_t)(stream_len + 32) * 8;
    const size_t info = (((streams + (size_t)(n_haps + 1) * 2 * 12) + 15) & ~(size_t)15) - streams;
    const size_t tables = (size_t)6 * ((kPhmmBins[bin].rpt + 3) / 4) * (PHMM_WARPS * 32) * 16;
    return streams + info + tables;
}

cudaError_t phmm_launch_prep(const PhmmBatchDev& b, cudaStream_t st);
// Forward wavefront kernel over ctas[cta_base .. cta_base + n_ctas) (all of one RPT bin).
cudaError_t phmm_launch_fwd(const PhmmBatchDev& b, int bin, bool fma, uint32_t cta_base, uint32_t n_ctas, size_t smem_bytes,
                            cudaStream_t st);
// Per-pair kernels.  count_dev != null: the pair count is read on the device (FP64 recompute list).
cudaError_t phmm_launch_pairs_f32(const PhmmBatchDev& b, bool fma, const PhmmPair* pairs, uint32_t count, float* scratch,
                                  uint32_t threads, uint32_t max_cols, cudaStream_t st);
cudaError_t phmm_launch_pairs_f64(const PhmmBatchDev& b, const PhmmPair* pairs, const uint32_t* count_dev, double* scratch,
                                  uint32_t threads, uint32_t max_cols, cudaStream_t st);
// Static facts for reports (registers, spills) — filled from cudaFuncGetAttributes.
struct PhmmKernelInfo
{
    int regs;
    int local_bytes;
    int max_threads;
};
cudaError_t phmm_fwd_kernel_info(int bin, bool fma, PhmmKernelInfo* info);

// FP32 issue-rate micro-benchmark; each thread executes iters*32 FMA instructions on 2 lanes (see .cu).
cudaError_t phmm_launch_ubench(int mode, uint32_t blocks, uint32_t iters, float* out, cudaStream_t st);

#endif
