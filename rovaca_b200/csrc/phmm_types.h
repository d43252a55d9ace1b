// phmm_types.h — records shared by the host engine and the CUDA kernels.
//
// HBM layout of one batch (one allocation, one H2D copy; offsets in PhmmBatchDev):
//   reads[]      PhmmRead, one per read of every region in the batch
//   haps[]       PhmmHap, one per haplotype
//   chunks[]     PhmmChunk: a region's haplotype list (or a slice of it) laid out as two column streams
//   ctas[]       PhmmCta: up to PHMM_WARPS warps of reads (packed lane by lane) that sweep one chunk together
//   task_reads[] read ids, ordered by (rows per lane, chunk, warp)
//   raw blobs    read bases / quals (/ ins / del / gcp when the caller supplies them), haplotype bases
//   words[]      written by the prep kernel: one packed u32 per read base (q | ins | del | gcp | one-hot)
//   out_f32[]    raw forward sums, [read][hap] per region  (pairhmm_internal.cpp:529 temp layout)
//   out_f64[]    FP64 recompute results, same indexing, only where out_f32 < 1e-28f
#ifndef PHMM_TYPES_H_
#define PHMM_TYPES_H_

#include <stdint.h>

#define PHMM_FLAG_BIT 0x80000000u   // stream word: first column of a haplotype (or the end sentinel)
#define PHMM_MAX_RPT 19              // rows per lane of the forward kernel: 1..19 for the reference's own gap qualities ...
#define PHMM_MAX_RPT_GENERIC 12      // ... 1..12 when a caller brings per-base deletion / continuation qualities (register file)
#define PHMM_WARPS 4                // reads per CTA (G = 32)

// read flags
#define PHMM_RD_HAS_INS 1u
#define PHMM_RD_HAS_DEL 2u
#define PHMM_RD_HAS_GCP 4u

struct PhmmRead
{
    uint32_t raw_off;    // byte offset of this read in every raw read blob
    uint32_t word_off;   // offset (in u32) of its packed words
    uint32_t len;
    uint32_t out_off;    // index of pair (read, hap 0) in out_f32 / out_f64
    uint32_t n_haps;     // haplotypes in its region (row length in out_*)
    uint32_t hap_begin;  // first haplotype of its region in haps[]
    int16_t thr;         // base-quality threshold (PairhmmArgument, haplotypecaller_args.h:35-39)
    uint8_t mapq;
    uint8_t model_flags; // pcr model (low nibble) | PHMM_RD_* << 4
    uint32_t rsv;
};

struct PhmmHap
{
    uint32_t byte_off;   // in the haplotype blob
    uint32_t len;
    uint32_t pos;        // first column position inside its stream (dummies included)
    uint32_t ord_stream; // ordinal within its stream << 1 | stream (0 = x, 1 = y)
    uint32_t out_col;    // haplotype index within its region
    uint32_t rsv;
};

struct PhmmChunk
{
    uint32_t hap_begin;  // haps[hap_begin .. hap_begin + n_haps)
    uint32_t n_haps;
    uint32_t stream_len; // positions per stream: 32 lead-in dummies + columns + sentinel, padded to the longer one
    uint32_t n_xy;       // haplotypes placed in stream x (low 16 bits) and in stream y (high 16 bits); haplotypes of the span
                         // that are in neither (empty, over-long: pos == 0) go to the per-pair kernel
};

struct PhmmCta
{
    uint32_t chunk;
    uint32_t read_begin; // into task_reads: the reads of warp 0, then of warp 1, ...
    uint32_t n_reads;    // all warps together
    uint32_t warp_reads; // reads of warp w in byte w (PHMM_WARPS = 4 warps at most); their lanes, ceil((len + 1) / RPT) each, sum to <= 32
};

// one (read, haplotype) pair for the per-pair kernels (FP64 recompute, out-of-envelope shapes)
struct PhmmPair
{
    uint32_t read;
    uint32_t hap;
    uint32_t out_idx;
};

// one region for the device epilogue (phmm_epilogue_kernel)
struct PhmmRegionDev
{
    uint32_t read_begin; // reads[read_begin .. read_begin + n_reads)
    uint32_t n_reads;
    uint32_t hap_begin;
    uint32_t n_haps;
    uint32_t out_off;    // index of pair (read 0, hap 0) in out_f32; the region's [n_haps][n_reads] matrix starts at the same index
    uint32_t rsv[3];
};

// lookup tables, built on the host with libm exactly as core/context.h does, uploaded once
struct PhmmTables
{
    float ph2pr[128];    // powf(10.f, -x/10.f)
    float gapm[128];     // 1.0f - ph2pr[x]
    float pm[128];       // 1.0f - ph2pr[q]        match prior
    float pmm[128];      // ph2pr[q] / 3.0f        mismatch prior
    float mm[8256];      // matchToMatchProb, triangular, max qual 127
    double ph2pr_d[128];
    double mm_d[8256];
};

#endif
