// phmm_kernels.cu — hand-written CUDA (sm_100a) for Rovaca's PairHMM likelihood path.
//
// What the reference computes per (read, haplotype) pair is the forward algorithm of
// rovaca/avx_512_float.cc:29-181 (equivalently core/avx512_s.cpp:259-397): an R x H table of three
// states M/I/D, 4 FMUL + 4 FFMA per cell in the exact order GCC gives the reference's AVX-512 TUs
// (SURVEY.md §8a), flush-to-zero, result = (sum_c M[R][c]) + (sum_c I[R][c]).
//
// Mapping (B200-first, not a port of the AVX lane layout; DESIGN.md §4-5 has the measurements behind each choice):
//   * a run of consecutive lanes of a warp owns one read: as many lanes as the read needs at RPT rows per lane,
//     ceil((len + 1) / RPT), so a warp carries reads of any mix of lengths packed lane by lane (8 lanes x 19 rows for a
//     150/151-bp read, 3 lanes for a 50-bp one beside it).  Lane l of a read owns rows l*RPT .. l*RPT+RPT-1 of its table; the
//     read is top-padded to a whole number of lanes with "boundary generator" rows (prior = 0, dd = 1) that reproduce the
//     reference's row -1 (M = I = 0, D = 2^120/H) for ANY garbage fed in from above — which is what lets the first lane of a
//     read sit right under the last lane of another read in the same shuffle — and the last read row is always the last
//     register row of the read's last lane.  Many rows per lane need the fewest shuffles and loads per cell;
//   * the lanes of a read sweep the columns as an anti-diagonal wavefront: at step s lane l is at stream position s - l and
//     receives the bottom row (M, I, D) of lane l-1 with three __shfl_up_sync per stream;
//   * every thread carries TWO haplotype streams at once in the two halves of packed f32x2 registers (FFMA2 / FMUL2 /
//     FADD2); the per-row transition probabilities stay scalar and enter the packed instructions as broadcast operands —
//     from registers, or, for the gap parameters the reference keeps constant, from uniform registers (no register-file
//     read: the register file's read bandwidth, not the FMA pipe, is what bounds this kernel);
//   * all haplotypes of a region are chained head to tail in the two streams, staged once per CTA in shared memory as
//     byte offsets of their base class's prior table, so the fill/drain is paid once per region, not once per pair; the
//     first column of each haplotype carries a flag bit that makes the lane emit the finished pair (last lane) and re-seed
//     its column -1 state;
//   * the match / mismatch prior of a cell is not selected (LOP3 + FSEL per cell) but fetched: every thread keeps a private
//     table prior[base class][row] in shared memory and reads four rows per LDS.128;
//   * read rows (packed q/ins/del/gcp/base words written by the prep kernel) are turned into the per-row floats once per
//     read through host-built libm tables and then live in registers for the whole sweep.
//
// Toolchain hazard, checked by tests/test_abi.py::test_sass_audit: ptxas 12.9 contracts a single-use mul.rn.f32x2 feeding an
// add.rn.f32x2 into FFMA2 although both carry .rn.  Every packed product below either feeds an FMA addend or has a second
// consumer; the un-contracted (AVX2) arithmetic variant is therefore written with scalar intrinsics, which ptxas leaves alone.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <atomic>

#include "phmm_kernels.h"
#include "phmm_types.h"

namespace
{

__device__ __forceinline__ uint32_t onehot(uint32_t c)   // pairhmm_internal.cpp:27-41
{
    uint32_t v = 0;
    v = (c == 'A') ? 0x1u : v;
    v = (c == 'C') ? 0x2u : v;
    v = (c == 'T') ? 0x4u : v;
    v = (c == 'G') ? 0x8u : v;
    v = (c == 'N') ? 0xFu : v;
    return v;
}

// ------------------------------------------------------------------------------------------------
// Read preparation: quality capping (pairhmm_internal.cpp:434-439) and the PCR-indel model
// (:65-159) as an integer pre-kernel, one warp per read, one lane per base.  Output: one u32 per
// base = q | ins<<7 | del<<14 | gcp<<21 | onehot<<28, every quality already masked & 127 (:466-469).
// ------------------------------------------------------------------------------------------------

__constant__ uint8_t c_pcr_cache[4][21] = {   // adapter.h:13-15
    {0},
    {40, 40, 39, 38, 37, 36, 34, 32, 28, 23, 17, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10},
    {40, 40, 40, 39, 39, 39, 38, 38, 37, 37, 36, 35, 34, 33, 32, 30, 28, 26, 23, 20, 17},
    {40, 40, 40, 40, 39, 39, 39, 39, 39, 38, 38, 38, 37, 37, 37, 36, 36, 35, 34, 33, 33}};

__device__ __forceinline__ bool unit_equal(const uint8_t* a, const uint8_t* b, int len)
{
    for (int i = 0; i < len; ++i)
        if (a[i] != b[i]) return false;
    return true;
}

// find_tandem_repeat_units(bases, offset) for offset < n - 1, the only offsets apply_pcr_error_model
// passes (:126-127).  There the backward unit (recorded at offset-str-1, :78) never equals the forward
// unit (at offset+1, :98), so the result is always: trailing repeats of the *forward* unit in
// bases[0..offset] plus its leading repeats in bases[offset+1..n) (:107-110), capped at 20.
__device__ int repeat_length(const uint8_t* b, int n, int offset)
{
    int fw_len = 1, max_fw = 0;
    const int test_len = n - offset - 1;
    const uint8_t* t = b + offset + 1;
    for (int str = 1; str <= 8; ++str) {
        if (offset + str + 1 > n) break;
        int cnt = 0;
        for (int s = 0; s <= test_len - str; s += str) {
            if (!unit_equal(t + s, t, str)) break;
            ++cnt;
        }
        max_fw = cnt;
        if (cnt > 1) { fw_len = str; break; }
    }
    int max_bw = 0;
    for (int s = offset + 1 - fw_len; s >= 0; s -= fw_len) {
        if (!unit_equal(b + s, t, fw_len)) break;
        ++max_bw;
    }
    const int rl = max_bw + max_fw;
    return rl > 20 ? 20 : rl;
}

// Bit-parallel form of the same scan for reads that fit a warp's shared-memory scratch: E[str-1] holds, for every
// position p, whether bases[p] == bases[p + str] (and p + str < n).  A unit of length str starting at t repeats at
// t + s (s a multiple of str) exactly when the str-periodicity holds on [t, t + s), i.e. when the run of set bits of
// E[str-1] starting at t is at least s long; the backward count is the run of set bits ending at `offset`.
constexpr int PREP_MAX = 1024;                 // bases per read on the fast path
constexpr int PREP_WORDS = PREP_MAX / 32;
constexpr uint32_t kPrepSplitReads = 2048;     // batches of at most this many reads get four warps per read in the prep kernel

__device__ __forceinline__ int run_up(const uint32_t* e, int nwords, int p0, int cap)     // consecutive ones at p0, p0+1, ...
{
    int run = 0, w = p0 >> 5, b = p0 & 31;
    while (w < nwords && run < cap) {
        const uint32_t inv = ~(e[w] >> b);         // zeros shifted in at the top read as "stop"
        const int avail = 32 - b;
        const int z = inv ? __ffs((int)inv) - 1 : 32;
        if (z < avail) return run + z;
        run += avail;
        ++w;
        b = 0;
    }
    return run;
}

__device__ __forceinline__ int run_down(const uint32_t* e, int p0, int cap)              // consecutive ones at p0, p0-1, ...
{
    int run = 0, w = p0 >> 5, b = p0 & 31;
    while (w >= 0 && run < cap) {
        const uint32_t inv = ~(e[w] << (31 - b));   // zeros shifted in at the bottom read as "stop"
        const int avail = b + 1;
        const int z = inv ? __clz((int)inv) : 32;
        if (z < avail) return run + z;
        run += avail;
        --w;
        b = 31;
    }
    return run;
}

// x / d for 0 <= x < 4096 and 1 <= d <= 8 (runs of at most 199 bits over unit lengths 1..8) without the integer-division
// sequence: (x + 0.5) / d lies at least 1/16 inside the integer interval of x / d, far beyond the approximate divide's error.
__device__ __forceinline__ int div_small(int x, int d) { return __float2int_rz(__fdividef((float)x + 0.5f, (float)d)); }

__device__ __forceinline__ int repeat_length_bits(const uint32_t (*E)[PREP_WORDS], int nwords, int n, int offset)
{
    const int test_len = n - offset - 1;
    const int p0 = offset + 1, w0 = p0 >> 5, b0 = p0 & 31;
    // The shortest unit that repeats at all (the next `str` periodicity bits all set), all eight lengths looked at without a
    // branch: the lanes of a warp find theirs at different lengths, and counting the whole run inside the loop made the warp
    // walk that (divergent) path at nearly every length.  Only then the run itself, once.
    int fw = 0;
#pragma unroll
    for (int str = 1; str <= 8; ++str) {
        const uint32_t lo = E[str - 1][w0], hi = (w0 + 1 < nwords) ? E[str - 1][w0 + 1] : 0u;
        const uint32_t need = (1u << str) - 1u;
        const bool hit = str <= test_len && (__funnelshift_r(lo, hi, b0) & need) == need;   // :126-127 offsets only; unit must fit
        fw = (fw == 0 && hit) ? str : fw;
    }
    int fw_len = 1, max_fw = 1;
    if (fw) {
        const int run = run_up(E[fw - 1], nwords, p0, 21 * fw);
        max_fw = div_small(run, fw) + 1;                                 // bounded by test_len / str through E's validity bit
        fw_len = fw;
    }
    const int max_bw = div_small(run_down(E[fw_len - 1], offset, 21 * fw_len), fw_len);
    const int rl = max_bw + max_fw;
    return rl > 20 ? 20 : rl;
}

// WPR warps per read: 1 for a batch that fills the device with reads; 4 for a small one (a synchronous caller's region), where
// the kernel is one warp's latency — the staging, the periodicity masks and the bases are then dealt to four warps that meet at
// a named barrier of their own.
template <int WPR>
__global__ void __launch_bounds__(256) phmm_prep_kernel(PhmmBatchDev p)
{
    constexpr int RPB = 8 / WPR;                            // reads per block
    constexpr int TS = 32 * WPR;                            // threads per read
    __shared__ uint8_t s_bases[RPB][PREP_MAX + 16];
    __shared__ uint32_t s_eq[RPB][8][PREP_WORDS];
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t warp = gw / WPR, part = gw % WPR;        // my read, and which of its warps I am
    const uint32_t lane = threadIdx.x & 31, wib = (threadIdx.x >> 5) / WPR;
    const int t0 = (int)(part * 32 + lane);                 // my first base
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.f64_count = 0u;   // the forward kernels' list of underflowing pairs starts empty
    if (warp >= p.n_reads) return;                          // (all warps of a read leave together)
    auto sync_read = [&] {
        if (WPR == 1)
            __syncwarp();
        else
            asm volatile("bar.sync %0, %1;" ::"r"(1 + wib), "r"(TS) : "memory");
    };
    const PhmmRead rd = p.reads[warp];
    const int n = (int)rd.len;
    const uint8_t* b = p.read_bases + rd.raw_off;
    const uint8_t* q = p.read_qual + rd.raw_off;
    const uint32_t flags = rd.model_flags >> 4;
    const int model = rd.model_flags & 15;
    const int thr = rd.thr;
    const bool fast = model != 0 && n <= PREP_MAX;
    const int nwords = (n + 31) >> 5;
    // the qualities of my first bases are fetched before the staging and the ballots below, not one global load per round of
    // the loop that needs them
    constexpr int kPre = 5;
    uint32_t qpre[kPre];
#pragma unroll
    for (int k = 0; k < kPre; ++k) {
        const int i = t0 + TS * k;
        qpre[k] = i < n ? q[i] : 0u;
    }
    if (fast) {
        uint8_t* sb = s_bases[wib];
#pragma unroll 4
        for (int i = t0; i < nwords * 32 + 8; i += TS) sb[i] = i < n ? b[i] : 0;
        sync_read();
        uint32_t turn = 0;
        for (int str = 1; str <= 8; ++str)
#pragma unroll 2
            for (int w = 0; w < nwords; ++w, ++turn) {
                if (WPR > 1 && (turn & (WPR - 1)) != part) continue;
                const int pos = w * 32 + (int)lane;
                const uint32_t m = __ballot_sync(0xffffffffu, pos + str < n && sb[pos] == sb[pos + str]);
                if (lane == 0) s_eq[wib][str - 1][w] = m;
            }
        sync_read();
    }
    auto one_base = [&](int i, uint32_t qq) {
        qq = qq < rd.mapq ? qq : rd.mapq;
        qq = ((int)qq < thr) ? 6u : qq;                                     // MIN_QUALITY, adapter.h:5
        uint32_t gi = (flags & PHMM_RD_HAS_INS) ? p.read_ins[rd.raw_off + i] : 45u;   // read_record.cpp:191-196
        if (model != 0 && i + 1 < n) {
            const int rl = fast ? repeat_length_bits(s_eq[wib], nwords, n, i) : repeat_length(b, n, i);
            const uint32_t cap = c_pcr_cache[model][rl];
            gi = gi < cap ? gi : cap;
        }
        const uint32_t gd = (flags & PHMM_RD_HAS_DEL) ? p.read_del[rd.raw_off + i] : gi;   // del aliases ins, :425,:559
        const uint32_t gc = (flags & PHMM_RD_HAS_GCP) ? p.read_gcp[rd.raw_off + i] : 10u; // read_record.cpp:205-210
        const uint32_t base = fast ? s_bases[wib][i] : b[i];
        p.words[rd.word_off + i] = (qq & 127u) | ((gi & 127u) << 7) | ((gd & 127u) << 14) | ((gc & 127u) << 21) | (onehot(base) << 28);
    };
#pragma unroll
    for (int k = 0; k < kPre; ++k) {
        const int i = t0 + TS * k;
        if (i < n) one_base(i, qpre[k]);
    }
    for (int i = t0 + TS * kPre; i < n; i += TS) one_base(i, q[i]);
}

// ------------------------------------------------------------------------------------------------
// Forward wavefront kernel.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }

// base class of a haplotype byte: index into the per-thread prior tables.  Classes follow the one-hot test of
// pairhmm_internal.cpp:27-41: A C T G match themselves and N, N matches every class but "other", other never matches.
__device__ __forceinline__ uint32_t base_class(uint32_t c)
{
    uint32_t v = 5u;
    v = (c == 'A') ? 0u : v;
    v = (c == 'C') ? 1u : v;
    v = (c == 'T') ? 2u : v;
    v = (c == 'G') ? 3u : v;
    v = (c == 'N') ? 4u : v;
    return v;
}
__device__ __forceinline__ uint32_t class_onehot(uint32_t b) { return b < 4u ? (1u << b) : (b == 4u ? 0xFu : 0u); }

// occupancy targets found with tools/phmm_ablate.cu: the kernel is bound by register-file reads and issue slots, not by
// latency, so it wants enough registers to hold RPT rows x 2 streams x 3 states plus the row parameters without spilling.
constexpr int fwd_min_blocks(int rpt) { return phmm_fwd_blocks_per_sm(rpt, false); }
constexpr int fwd_min_blocks_rg(int rpt) { return phmm_fwd_blocks_per_sm(rpt, true); }

// VAR: PHMM_VAR_GENERIC_NOFMA — any per-base qualities, un-contracted AVX2 arithmetic (core/avx2_s.cpp:207-216);
//      PHMM_VAR_GENERIC       — any per-base qualities, AVX-512 contraction;
//      PHMM_VAR_REFGAPS       — what the reference itself feeds the kernel: deletion qualities alias the insertion ones and
//                               the gap-continuation quality is one constant per read (read_record.cpp:191-210,
//                               pairhmm_internal.cpp:425,559), so md == mi and gapm / ii / dd are per-read scalars: half the
//                               parameter registers, which is what lets 19 rows per lane (8 lanes x 19 = 152 rows for 150/151-bp
//                               reads, 1 % padding instead of 6 %) fit the register file.
// FIXED: 0 = every read on the lanes it needs (the general kernel); else every read of the launch takes exactly FIXED lanes (a
// power of two) — the round-1 layout, kept as a specialisation for the commonest shape (150/151-bp reads on 8 lanes x 19 rows):
// read slot and lane are the thread's position, the shuffles are segmented per read; ptxas schedules this hot loop 1.5 % better.
// DUO: the two halves of every packed register hold two READS of the same lane count sweeping ONE haplotype stream (instead of
// one read sweeping two streams).  For the haplotype a region has left over when its count is odd: dealt into the two streams it
// would make one of them longer, and the other half of every lane idle for that long.  Each half has its own row parameters (as
// true packed operands), its own prior tables and pad rows; a seam is crossed by both at once.  Reference-gap variant only.
template <int RPT, int VAR, int FIXED, int DUO = 0>
__global__ void __launch_bounds__(PHMM_WARPS * 32, DUO ? 2 : VAR == PHMM_VAR_REFGAPS ? fwd_min_blocks_rg(RPT) : fwd_min_blocks(RPT))
phmm_fwd_kernel(PhmmBatchDev p, uint32_t cta_base)
{
    constexpr bool FMA = VAR != PHMM_VAR_GENERIC_NOFMA;
    constexpr bool RG = VAR == PHMM_VAR_REFGAPS;
    static_assert(!DUO || (RG && FIXED == 0), "paired reads: reference-gap variant, general lane layout");
    constexpr int NP = RG ? 1 : RPT;                      // per-row (generic) or per-read (reference gaps) gap parameters
    constexpr int G = (RPT + 3) / 4;                     // float4 groups of the prior table
    constexpr uint32_t TSTRIDE = phmm_prior_table_stride(RPT);   // bytes per thread and base class: G float4, padded to an odd count
    const uint32_t NT = blockDim.x;                       // 1..PHMM_WARPS warps per CTA, chosen by the host per launch
    const uint32_t CSTRIDE = NT * TSTRIDE;                // bytes between the tables of two base classes (a multiple of 512)
    extern __shared__ __align__(16) unsigned char smem[];
    const PhmmCta cta = p.ctas[cta_base + blockIdx.x];
    const PhmmChunk ch = p.chunks[cta.chunk];
    const uint32_t L = ch.stream_len;
    uint2* stream = reinterpret_cast<uint2*>(smem);                   // [L + 32] {x word, y word}
    const uint32_t NI = ch.n_haps + 1;
    float* info_init = reinterpret_cast<float*>(stream + L + 32);     // [2][NI]  2^120 / H
    uint32_t* info_col = reinterpret_cast<uint32_t*>(info_init + 2 * NI);   // [2][NI] output column
    uint32_t* info_hap = info_col + 2 * NI;                            // [2][NI] global haplotype id
    unsigned char* tbl = smem + ((((size_t)(L + 32) * 8 + (size_t)NI * 2 * 12) + 15) & ~(size_t)15);   // [6 classes][thread][G] float4

    // ---- stage the two haplotype streams of this chunk: word = class * CSTRIDE (| flag on a first column) ----
    for (uint32_t i = threadIdx.x; i < L + 32; i += NT) stream[i] = make_uint2(0u, 0u);
    for (uint32_t i = threadIdx.x; i < 2 * NI; i += NT) { info_init[i] = 0.f; info_col[i] = 0u; info_hap[i] = 0u; }
    __syncthreads();
    {
        uint32_t* sw = reinterpret_cast<uint32_t*>(stream);
        if (threadIdx.x == 0) {                      // empty streams still need their end sentinel
            if ((ch.n_xy & 0xffffu) == 0) sw[2 * 32 + 0] = PHMM_FLAG_BIT;
            if ((ch.n_xy >> 16) == 0) sw[2 * 32 + 1] = PHMM_FLAG_BIT;
        }
        for (uint32_t j = 0; j < ch.n_haps; ++j) {
            const PhmmHap hp = p.haps[ch.hap_begin + j];
            if (hp.pos == 0) continue;               // not in a stream (empty / over-long: per-pair kernel)
            const uint32_t e = hp.ord_stream & 1u, ord = hp.ord_stream >> 1;
            const uint32_t n_e = e ? (ch.n_xy >> 16) : (ch.n_xy & 0xffffu);
            const uint8_t* hb = p.hap_bases + hp.byte_off;
            for (uint32_t c = threadIdx.x; c < hp.len; c += NT) {
                uint32_t w = base_class(hb[c]) * CSTRIDE;
                if (c == 0) w |= PHMM_FLAG_BIT;
                sw[2 * (hp.pos + c) + e] = w;
            }
            if (threadIdx.x == 0) {
                info_init[e * NI + ord] = __fdiv_rn(__uint_as_float(0x7B800000u), (float)hp.len);   // 2^120 / H, avx_512_float.cc:33
                info_col[e * NI + ord] = hp.out_col;
                info_hap[e * NI + ord] = ch.hap_begin + j;
                if (ord + 1 == n_e) sw[2 * (hp.pos + hp.len) + e] = PHMM_FLAG_BIT;   // end sentinel
            }
        }
    }
    __syncthreads();

    // ---- which read is mine: this warp's reads, each on ceil((len + 1) / RPT) consecutive lanes ---------------
    const uint32_t wlane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wcnt = (cta.warp_reads >> (8 * warp)) & 0xffu;                  // reads of this warp
    if (wcnt == 0) return;                                                         // (the host packs warps so that this is rare)
    uint32_t wbeg = cta.read_begin;                                                // ... and where they start in the task list
    constexpr uint32_t TS = DUO ? 2u : 1u;                                         // (paired reads: two task entries per slot)
#pragma unroll
    for (uint32_t w = 0; w < PHMM_WARPS; ++w) wbeg += w < warp ? TS * ((cta.warp_reads >> (8 * w)) & 0xffu) : 0u;
    uint32_t rid, nl, lane, nl_max;
    uint32_t ridB = 0xffffffffu;                                                   // DUO: the read in the .y halves (none: ~0)
    bool live;
    if (FIXED) {
        live = wlane / (FIXED ? FIXED : 1) < wcnt;                                 // a dead group shadows the warp's last read, never writes
        rid = p.task_reads[wbeg + (live ? wlane / (FIXED ? FIXED : 1) : wcnt - 1)];
        lane = wlane & ((FIXED ? FIXED : 1) - 1);
        nl = nl_max = FIXED;
    } else {
        uint32_t my_rid = 0, my_nl = 0, my_ridB = 0xffffffffu;
        if (wlane < wcnt) {
            my_rid = p.task_reads[wbeg + TS * wlane];
            if (DUO) my_ridB = p.task_reads[wbeg + TS * wlane + 1];                // (same lane count as its partner: the host pairs them so)
            my_nl = (p.reads[my_rid].len + RPT) / RPT;                             // ceil((len + 1) / RPT): at least one pad row on top
        }
        uint32_t start = my_nl;                                                    // exclusive prefix sum: first lane of read i
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, start, d);
            if (wlane >= (uint32_t)d) start += v;
        }
        const uint32_t used = __shfl_sync(0xffffffffu, start, 31);                 // lanes with a read (<= 32 by the host's packing)
        start -= my_nl;
        const uint32_t starts = __reduce_or_sync(0xffffffffu, wlane < wcnt ? (1u << start) : 0u);
        nl_max = __reduce_max_sync(0xffffffffu, my_nl);
        live = wlane < used;                                                       // a lane without a read shadows lane 0 of the last read, never writes
        const uint32_t below = starts & (0xffffffffu >> (31 - (live ? wlane : used - 1)));
        const uint32_t ridx = (uint32_t)__popc(below) - 1u;                       // which of the warp's reads covers me
        const uint32_t first = 31u - (uint32_t)__clz((int)below);
        rid = __shfl_sync(0xffffffffu, my_rid, ridx);
        if (DUO) ridB = __shfl_sync(0xffffffffu, my_ridB, ridx);
        nl = __shfl_sync(0xffffffffu, my_nl, ridx);
        lane = live ? wlane - first : 0u;                                          // my lane within the read
    }
    constexpr int shfl_w = FIXED ? FIXED : 32;       // (segmenting the shuffles at run time where the lane counts allow it measured slower)
    const PhmmRead rd = p.reads[rid];
    const int npad = (int)(nl * RPT) - (int)rd.len;   // >= 1
    const bool hasB = DUO && ridB != 0xffffffffu;     // a slot with one read only: the .y halves shadow it and never write
    const PhmmRead rdB = DUO ? p.reads[hasB ? ridB : rid] : rd;
    const int npadB = (int)(nl * RPT) - (int)rdB.len;

    // ---- per-row parameters (pairhmm_internal.cpp:463-481) into registers, priors into my table ----
    float mm[RPT], mi[RPT], dd[RPT], gapm[NP], ii[NP], md[NP];
    float ii0 = 0.f;           // RG: gap-continuation probability of row 0 (0 on lane 0, whose row 0 is always a pad row)
    uint32_t padmask = 0u;     // bit k: row k is a pad row; bit RPT: the row above row 0 is one
    {
        const PhmmTables* __restrict__ T = p.tables;
        float pm[G * 4], pmm[G * 4];
        uint32_t rr[G * 4];
#pragma unroll
        for (int k = 0; k < G * 4; ++k) { pm[k] = 0.f; pmm[k] = 0.f; rr[k] = 0u; }
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = (int)lane * RPT + k - npad;
            if (r >= 0) {
                const uint32_t w = p.words[rd.word_off + r];
                const uint32_t q = w & 127u, qi = (w >> 7) & 127u, qd = (w >> 14) & 127u, qc = (w >> 21) & 127u;
                const uint32_t lo = qi < qd ? qi : qd, hi = qi < qd ? qd : qi;
                mm[k] = __ldg(&T->mm[((hi * (hi + 1)) >> 1) + lo]);
                mi[k] = __ldg(&T->ph2pr[qi]);
                dd[k] = __ldg(&T->ph2pr[qc]);
                if (!RG) {
                    gapm[k] = __ldg(&T->gapm[qc]);
                    ii[k] = dd[k];
                    md[k] = __ldg(&T->ph2pr[qd]);
                }
                pm[k] = __ldg(&T->pm[q]);
                pmm[k] = __ldg(&T->pmm[q]);
                rr[k] = w >> 28;
            } else {                                   // boundary-generator row: prior 0, I stays 0, D carried unchanged
                mm[k] = 0.f; mi[k] = 0.f; dd[k] = 1.f;
                if (!RG) { gapm[k] = 0.f; ii[k] = 0.f; md[k] = 0.f; }
                padmask |= 1u << k;
            }
        }
        if (RG) {
            // One gap-continuation quality for the whole launch (the reference's constant 10): gapm and ii are kernel
            // parameters, i.e. uniform-register operands of FFMA2 that cost no register-file read.  Pad rows keep mi = 0 and
            // dd = 1; with the shared ii their I is ii x (I of the pad row above) = 0, except on the very first row of the
            // group, which sees lane 0's own shuffled-back bottom row: that row multiplies by ii0 = 0 instead.  Their M is
            // (finite) x prior 0 = 0.
            gapm[0] = p.rg_gapm;
            ii[0] = p.rg_ii;
            ii0 = lane == 0 ? 0.f : p.rg_ii;
        }
        if ((int)lane * RPT - 1 < npad) padmask |= 1u << RPT;
#pragma unroll
        for (uint32_t b = 0; b < 6; ++b) {
            const uint32_t hb = class_onehot(b);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float4 v;
                v.x = (rr[4 * g] & hb) ? pm[4 * g] : pmm[4 * g];
                v.y = (rr[4 * g + 1] & hb) ? pm[4 * g + 1] : pmm[4 * g + 1];
                v.z = (rr[4 * g + 2] & hb) ? pm[4 * g + 2] : pmm[4 * g + 2];
                v.w = (rr[4 * g + 3] & hb) ? pm[4 * g + 3] : pmm[4 * g + 3];
                *reinterpret_cast<float4*>(tbl + b * CSTRIDE + threadIdx.x * TSTRIDE + g * 16) = v;
            }
        }
    }

    // ---- paired reads: the same for the read in the .y halves, with tables of its own behind the first six ----
    constexpr int RB = DUO ? RPT : 1;
    float mmB[RB], miB[RB], ddB[RB];
    uint32_t padmaskB = 0u;
    if constexpr (DUO != 0) {
        const PhmmTables* __restrict__ T = p.tables;
        float pm[G * 4], pmm[G * 4];
        uint32_t rr[G * 4];
#pragma unroll
        for (int k = 0; k < G * 4; ++k) { pm[k] = 0.f; pmm[k] = 0.f; rr[k] = 0u; }
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = (int)lane * RPT + k - npadB;
            if (r >= 0) {
                const uint32_t w = p.words[rdB.word_off + r];
                const uint32_t q = w & 127u, qi = (w >> 7) & 127u, qd = (w >> 14) & 127u, qc = (w >> 21) & 127u;
                const uint32_t lo = qi < qd ? qi : qd, hi = qi < qd ? qd : qi;
                mmB[k] = __ldg(&T->mm[((hi * (hi + 1)) >> 1) + lo]);
                miB[k] = __ldg(&T->ph2pr[qi]);
                ddB[k] = __ldg(&T->ph2pr[qc]);
                pm[k] = __ldg(&T->pm[q]);
                pmm[k] = __ldg(&T->pmm[q]);
                rr[k] = w >> 28;
            } else {
                mmB[k] = 0.f; miB[k] = 0.f; ddB[k] = 1.f;
                padmaskB |= 1u << k;
            }
        }
        if ((int)lane * RPT - 1 < npadB) padmaskB |= 1u << RPT;
#pragma unroll
        for (uint32_t b = 0; b < 6; ++b) {
            const uint32_t hb = class_onehot(b);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float4 v;
                v.x = (rr[4 * g] & hb) ? pm[4 * g] : pmm[4 * g];
                v.y = (rr[4 * g + 1] & hb) ? pm[4 * g + 1] : pmm[4 * g + 1];
                v.z = (rr[4 * g + 2] & hb) ? pm[4 * g + 2] : pmm[4 * g + 2];
                v.w = (rr[4 * g + 3] & hb) ? pm[4 * g + 3] : pmm[4 * g + 3];
                *reinterpret_cast<float4*>(tbl + (6 + b) * CSTRIDE + threadIdx.x * TSTRIDE + g * 16) = v;
            }
        }
    }

    // ---- state ------------------------------------------------------------------------------
    float2 Ml[RPT], Il[RPT], Dl[RPT];              // previous column, my rows
#pragma unroll
    for (int k = 0; k < RPT; ++k) { Ml[k] = bc(0.f); Il[k] = bc(0.f); Dl[k] = bc(0.f); }
    float2 upM = bc(0.f), upI = bc(0.f), upD = bc(0.f);   // row above my row 0, current column
    float2 dgM = bc(0.f), dgI = bc(0.f), dgD = bc(0.f);   // same row, previous column
    float2 sumM = bc(0.f), sumI = bc(0.f);
    uint32_t jx = 0, jy = 0;                       // haplotypes begun so far in each stream
    const uint2* sp = stream + 32 - lane;
    const unsigned char* tb = tbl + threadIdx.x * TSTRIDE;
    const uint32_t nsteps = L - 1 - (32 - nl_max); // the last lane of the warp's widest read reaches the end sentinel at step L - 33 + nl_max
    const bool writer = live && (lane == nl - 1);
    const bool writerB = writer && hasB;

    // paired reads: one stream, both halves cross a seam together
#define PHMM_BOUNDARY_DUO()                                                                             \
    {                                                                                                   \
        if (writer && jx > 0) {                                                                         \
            const uint32_t col = info_col[jx - 1], hap = info_hap[jx - 1];                              \
            const float resA = __fadd_rn(sumM.x, sumI.x);                                               \
            p.out_f32[rd.out_off + col] = resA;                                                         \
            if (resA < 1e-28f) {                                                                        \
                const uint32_t slot64 = atomicAdd(p.f64_count, 1u);                                     \
                PhmmPair pr; pr.read = rid; pr.hap = hap; pr.out_idx = rd.out_off + col;                \
                p.f64_pairs[slot64] = pr;                                                               \
            }                                                                                           \
            if (writerB) {                                                                              \
                const float resB = __fadd_rn(sumM.y, sumI.y);                                           \
                p.out_f32[rdB.out_off + col] = resB;                                                    \
                if (resB < 1e-28f) {                                                                    \
                    const uint32_t slot64 = atomicAdd(p.f64_count, 1u);                                 \
                    PhmmPair pr; pr.read = ridB; pr.hap = hap; pr.out_idx = rdB.out_off + col;          \
                    p.f64_pairs[slot64] = pr;                                                           \
                }                                                                                       \
            }                                                                                           \
        }                                                                                               \
        const float init = info_init[jx];                                                               \
        ++jx;                                                                                           \
        sumM = bc(0.f); sumI = bc(0.f);                                                                 \
        _Pragma("unroll") for (int k = 0; k < RPT; ++k) {                                               \
            Ml[k] = bc(0.f); Il[k] = bc(0.f);                                                           \
            Dl[k].x = ((padmask >> k) & 1u) ? init : 0.f;                                               \
            Dl[k].y = ((padmaskB >> k) & 1u) ? init : 0.f;                                              \
        }                                                                                               \
        dgM = bc(0.f); dgI = bc(0.f);                                                                   \
        dgD.x = ((padmask >> RPT) & 1u) ? init : 0.f;                                                   \
        dgD.y = ((padmaskB >> RPT) & 1u) ? init : 0.f;                                                  \
    }

#define PHMM_BOUNDARY(E, J, EI)                                                                         \
    {                                                                                                   \
        if (writer && J > 0) {                                                                          \
            const float res = __fadd_rn(sumM.E, sumI.E);                                                \
            const uint32_t oi = rd.out_off + info_col[EI * NI + J - 1];                                 \
            p.out_f32[oi] = res;                                                                        \
            if (res < 1e-28f) {                                                                         \
                const uint32_t slot64 = atomicAdd(p.f64_count, 1u);                                     \
                PhmmPair pr; pr.read = rid; pr.hap = info_hap[EI * NI + J - 1]; pr.out_idx = oi;        \
                p.f64_pairs[slot64] = pr;                                                               \
            }                                                                                           \
        }                                                                                               \
        const float init = info_init[EI * NI + J];                                                      \
        ++J;                                                                                            \
        sumM.E = 0.f; sumI.E = 0.f;                                                                     \
        _Pragma("unroll") for (int k = 0; k < RPT; ++k) {                                               \
            Ml[k].E = 0.f; Il[k].E = 0.f; Dl[k].E = ((padmask >> k) & 1u) ? init : 0.f;                 \
        }                                                                                               \
        dgM.E = 0.f; dgI.E = 0.f; dgD.E = ((padmask >> RPT) & 1u) ? init : 0.f;                         \
    }

    // Few rows per lane is the layout of a batch too small to fill the device (the latency policy of the planner): one warp per
    // scheduler, every step waiting for its own stream word (LDS) and then for its priors (LDS again).  There the word and the
    // priors of the next step are fetched one step ahead; the throughput layouts (more rows, other warps to issue from) keep the
    // shorter live ranges.
    constexpr bool PIPE = RG && RPT <= 8;
    auto load_priors = [&](uint32_t wx, uint32_t wy, float* px, float* py) {
        // priors of my RPT rows against a column's base class, four rows per LDS (pairhmm_internal.cpp:479-480 values)
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const unsigned char* ax = tb + wx + g * 16;
            const unsigned char* ay = tb + wy + g * 16;
            if (RPT - 4 * g >= 3) {
                const float4 vx = *reinterpret_cast<const float4*>(ax), vy = *reinterpret_cast<const float4*>(ay);
                px[4 * g] = vx.x; px[4 * g + 1] = vx.y; px[4 * g + 2] = vx.z; px[4 * g + 3] = vx.w;
                py[4 * g] = vy.x; py[4 * g + 1] = vy.y; py[4 * g + 2] = vy.z; py[4 * g + 3] = vy.w;
            } else if (RPT - 4 * g == 2) {
                const float2 vx = *reinterpret_cast<const float2*>(ax), vy = *reinterpret_cast<const float2*>(ay);
                px[4 * g] = vx.x; px[4 * g + 1] = vx.y;
                py[4 * g] = vy.x; py[4 * g + 1] = vy.y;
            } else {
                px[4 * g] = *reinterpret_cast<const float*>(ax);
                py[4 * g] = *reinterpret_cast<const float*>(ay);
            }
        }
    };
    // at the top of step s: wc = word(s) and pxn / pyn = its priors, wn = word(s + 1) — all in registers; the step asks for the
    // priors of wn and for word(s + 2), and nobody waits for either before the next step
    // The next step's stream word alone, fetched a step ahead: measured on every throughput instantiation (ptxas schedules each its own
    // way) — 16 rows +1.1 % (cfg4), 12 rows -3.4 %, the fixed-lanes 19-row kernel -1.0 % (profiles/r02_word_prefetch_ab.txt).
    constexpr bool PREW = !PIPE && !DUO && RG && FIXED == 0 && RPT == 16;
    uint2 wpre = make_uint2(0u, 0u);
    if (PREW && nsteps > 0) wpre = sp[0];
    uint2 wc = make_uint2(0u, 0u), wn = make_uint2(0u, 0u);
    float pxn[G * 4], pyn[G * 4];
#pragma unroll
    for (int k = 0; k < G * 4; ++k) { pxn[k] = 0.f; pyn[k] = 0.f; }
    if (PIPE && nsteps > 0) {
        wc = sp[0];
        wn = sp[1];
        load_priors(wc.x & ~PHMM_FLAG_BIT, wc.y & ~PHMM_FLAG_BIT, pxn, pyn);
    }

    // two steps per trip let the registers of a step's inputs and outputs alternate (one step per trip costs a third of the
    // throughput in moves); four pay for their code size only in the middle (measured: +1-3 % at 9-13 rows, -2-4 % below 9, 0 above;
    // at 9 rows and its 128-register cap four steps spill, so it stays at two)
#pragma unroll(RPT >= 10 && RPT <= 13 ? 4 : 2)
    for (uint32_t s = 0; s < nsteps; ++s) {
        uint2 w;
        float px[G * 4], py[G * 4];
        if (PIPE) {
            w = wc;
#pragma unroll
            for (int k = 0; k < G * 4; ++k) { px[k] = pxn[k]; py[k] = pyn[k]; }
            load_priors(wn.x & ~PHMM_FLAG_BIT, wn.y & ~PHMM_FLAG_BIT, pxn, pyn);
            wc = wn;
            wn = sp[s + 2];                        // (32 zero words follow the end sentinel; a word beyond them is never used)
        } else if (PREW) {
            w = wpre;
            wpre = sp[s + 1];
        } else {
            w = sp[s];
        }
        if constexpr (DUO != 0) {
            if ((int)w.x < 0) {
                PHMM_BOUNDARY_DUO()
                w.x &= ~PHMM_FLAG_BIT;
            }
            load_priors(w.x, w.x + 6u * CSTRIDE, px, py);       // the same column for both reads, each in its own tables
        } else {
            if ((int)(w.x | w.y) < 0) {
                if ((int)w.x < 0) PHMM_BOUNDARY(x, jx, 0)
                if ((int)w.y < 0) PHMM_BOUNDARY(y, jy, 1)
                w.x &= ~PHMM_FLAG_BIT;
                w.y &= ~PHMM_FLAG_BIT;
            }
            if (!PIPE) load_priors(w.x, w.y, px, py);
        }
        float2 pM = dgM, pI = dgI, pD = dgD;       // diagonal inputs of row 0
        float2 uM = upM, uI = upI;                 // inputs from above of row 0
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const float2 oM = Ml[k], oI = Il[k], oD = Dl[k];
            float2 nM, nI, nD;
            if constexpr (DUO != 0) {               // two reads: every row parameter is a true pair
                float2 t = __fmul2_rn(pM, make_float2(mm[k], mmB[k]));
                t = __ffma2_rn(pI, bc(p.rg_gapm), t);
                t = __ffma2_rn(pD, bc(p.rg_gapm), t);
                nM.x = __fmul_rn(t.x, px[k]);
                nM.y = __fmul_rn(t.y, py[k]);
                nI = __ffma2_rn(uI, bc(k == 0 ? ii0 : p.rg_ii), __fmul2_rn(uM, make_float2(mi[k], miB[k])));
                nD = __ffma2_rn(oD, make_float2(dd[k], ddB[k]), __fmul2_rn(oM, make_float2(mi[k], miB[k])));
            } else if (RG) {                        // same operations, md == mi and one gapm / ii per read
                float2 t = __fmul2_rn(pM, bc(mm[k]));
                t = __ffma2_rn(pI, bc(p.rg_gapm), t);
                t = __ffma2_rn(pD, bc(p.rg_gapm), t);
                nM.x = __fmul_rn(t.x, px[k]);
                nM.y = __fmul_rn(t.y, py[k]);
                nI = __ffma2_rn(uI, bc(k == 0 ? ii0 : p.rg_ii), __fmul2_rn(uM, bc(mi[k])));
                nD = __ffma2_rn(oD, bc(dd[k]), __fmul2_rn(oM, bc(mi[k])));
            } else if (FMA) {                       // SURVEY.md §8a, AVX-512 contraction
                float2 t = __fmul2_rn(pM, bc(mm[k]));
                t = __ffma2_rn(pI, bc(gapm[k]), t);
                t = __ffma2_rn(pD, bc(gapm[k]), t);
                nM.x = __fmul_rn(t.x, px[k]);       // scalar on purpose: the priors arrive un-paired from the table
                nM.y = __fmul_rn(t.y, py[k]);
                nI = __ffma2_rn(uI, bc(ii[k]), __fmul2_rn(uM, bc(mi[k])));
                nD = __ffma2_rn(oD, bc(dd[k]), __fmul2_rn(oM, bc(md[k])));
            } else {                                // core/avx2_s.cpp:207-216, scalar on purpose (see header)
                float tx = __fadd_rn(__fmul_rn(pM.x, mm[k]), __fmul_rn(pI.x, gapm[k]));
                float ty = __fadd_rn(__fmul_rn(pM.y, mm[k]), __fmul_rn(pI.y, gapm[k]));
                tx = __fadd_rn(tx, __fmul_rn(pD.x, gapm[k]));
                ty = __fadd_rn(ty, __fmul_rn(pD.y, gapm[k]));
                nM.x = __fmul_rn(tx, px[k]);
                nM.y = __fmul_rn(ty, py[k]);
                nI.x = __fadd_rn(__fmul_rn(uM.x, mi[k]), __fmul_rn(uI.x, ii[k]));
                nI.y = __fadd_rn(__fmul_rn(uM.y, mi[k]), __fmul_rn(uI.y, ii[k]));
                nD.x = __fadd_rn(__fmul_rn(oM.x, md[k]), __fmul_rn(oD.x, dd[k]));
                nD.y = __fadd_rn(__fmul_rn(oM.y, md[k]), __fmul_rn(oD.y, dd[k]));
            }
            Ml[k] = nM; Il[k] = nI; Dl[k] = nD;
            pM = oM; pI = oI; pD = oD;
            uM = nM; uI = nI;
        }
        dgM = upM; dgI = upI; dgD = upD;
        if (FMA) {
            sumM = __fadd2_rn(sumM, Ml[RPT - 1]);   // avx_512_float.cc:166-171, ascending columns
            sumI = __fadd2_rn(sumI, Il[RPT - 1]);
        } else {
            sumM.x = __fadd_rn(sumM.x, Ml[RPT - 1].x); sumM.y = __fadd_rn(sumM.y, Ml[RPT - 1].y);
            sumI.x = __fadd_rn(sumI.x, Il[RPT - 1].x); sumI.y = __fadd_rn(sumI.y, Il[RPT - 1].y);
        }
        upM.x = __shfl_up_sync(0xffffffffu, Ml[RPT - 1].x, 1, shfl_w);
        upM.y = __shfl_up_sync(0xffffffffu, Ml[RPT - 1].y, 1, shfl_w);
        upI.x = __shfl_up_sync(0xffffffffu, Il[RPT - 1].x, 1, shfl_w);
        upI.y = __shfl_up_sync(0xffffffffu, Il[RPT - 1].y, 1, shfl_w);
        upD.x = __shfl_up_sync(0xffffffffu, Dl[RPT - 1].x, 1, shfl_w);
        upD.y = __shfl_up_sync(0xffffffffu, Dl[RPT - 1].y, 1, shfl_w);
    }
#undef PHMM_BOUNDARY
#undef PHMM_BOUNDARY_DUO
}

// ------------------------------------------------------------------------------------------------
// Per-pair wavefront kernel: one warp per (read, haplotype) pair, any read and haplotype length, FP32 or FP64.  Used for
// the FP64 recompute of pairs whose FP32 sum underflowed (pairhmm_internal.cpp:578-581; init 2^1020/H, double tables —
// core/context.h:99-111) and, in FP32, for shapes outside the batched kernel's envelope.  Same recurrence, same operation
// order.  The read is cut into strips of 32 x PAIR_RPT rows (top-padded with boundary-generator rows like the batched
// kernel); a strip sweeps the haplotype as a wavefront, its bottom row goes through a per-warp global scratch to the next
// strip.  The pair list and its length live on the device (the forward kernels append to it), so the grid is fixed and
// every warp strides over the list.
// ------------------------------------------------------------------------------------------------

template <typename T> struct Num;
template <> struct Num<float>
{
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float init(uint32_t H) { return __fdiv_rn(__uint_as_float(0x7B800000u), (float)H); }
};
template <> struct Num<double>
{
    // The reference runs its FP64 kernels on a thread whose MXCSR has flush-to-zero set (init_native,
    // pairhmm_internal.cpp:21-25), which also flushes tiny *double* results; the GPU has no FP64 FTZ mode, so every
    // result below DBL_MIN is flushed here (sign kept), as the SSE unit does.
    static __device__ __forceinline__ double ftz(double x) { return fabs(x) < 2.2250738585072014e-308 ? copysign(0.0, x) : x; }
    static __device__ __forceinline__ double mul(double a, double b) { return ftz(__dmul_rn(a, b)); }
    static __device__ __forceinline__ double add(double a, double b) { return ftz(__dadd_rn(a, b)); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ftz(__fma_rn(a, b, c)); }
    static __device__ __forceinline__ double init(uint32_t H) { return __ddiv_rn(__longlong_as_double(0x7FB0000000000000LL), (double)H); }   // 2^1020 / H
};

constexpr int PAIR_RPT = 5;

template <typename T> struct RowParams;
template <> struct RowParams<float>
{
    static __device__ __forceinline__ void load(const PhmmTables* tb, uint32_t w, float& mm, float& gapm, float& mi, float& ii, float& md,
                                                float& pm, float& pmm)
    {
        const uint32_t q = w & 127u, qi = (w >> 7) & 127u, qd = (w >> 14) & 127u, qc = (w >> 21) & 127u;
        const uint32_t lo = qi < qd ? qi : qd, hi = qi < qd ? qd : qi;
        mm = tb->mm[((hi * (hi + 1)) >> 1) + lo]; gapm = tb->gapm[qc]; mi = tb->ph2pr[qi]; ii = tb->ph2pr[qc];
        md = tb->ph2pr[qd]; pm = tb->pm[q]; pmm = tb->pmm[q];
    }
};
template <> struct RowParams<double>
{
    static __device__ __forceinline__ void load(const PhmmTables* tb, uint32_t w, double& mm, double& gapm, double& mi, double& ii,
                                                double& md, double& pm, double& pmm)
    {
        const uint32_t q = w & 127u, qi = (w >> 7) & 127u, qd = (w >> 14) & 127u, qc = (w >> 21) & 127u;
        const uint32_t lo = qi < qd ? qi : qd, hi = qi < qd ? qd : qi;
        mm = tb->mm_d[((hi * (hi + 1)) >> 1) + lo]; gapm = __dsub_rn(1.0, tb->ph2pr_d[qc]); mi = tb->ph2pr_d[qi];
        ii = tb->ph2pr_d[qc]; md = tb->ph2pr_d[qd]; pm = __dsub_rn(1.0, tb->ph2pr_d[q]); pmm = __ddiv_rn(tb->ph2pr_d[q], 3.0);
    }
};

template <typename T, bool FMA>
__global__ void __launch_bounds__(128) phmm_pair_kernel(PhmmBatchDev p, const PhmmPair* pairs, const uint32_t* count_dev,
                                                        uint32_t count, T* scratch, uint32_t max_cols, T* out)
{
    constexpr int RPT = PAIR_RPT, ROWS = 32 * RPT;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n = count_dev ? *count_dev : count;
    T* sM = scratch + (size_t)wid * 3 * (max_cols + 1);        // bottom row of the previous strip, per column
    T* sI = sM + (max_cols + 1);
    T* sD = sI + (max_cols + 1);
    const PhmmTables* __restrict__ tb = p.tables;
    for (uint32_t pi = wid; pi < n; pi += nwarps) {
        const PhmmPair pr = pairs[pi];
        const PhmmRead rd = p.reads[pr.read];
        const PhmmHap hp = p.haps[pr.hap];
        const int H = (int)hp.len, R = (int)rd.len;
        const uint8_t* hb = p.hap_bases + hp.byte_off;
        const T init = Num<T>::init((uint32_t)H);
        const int nstrips = (R + 1 + ROWS - 1) / ROWS;         // at least one generator row on top
        const int npad = nstrips * ROWS - R;
        T sumM = (T)0, sumI = (T)0;
        for (int strip = 0; strip < nstrips; ++strip) {
            T mm[RPT], gapm[RPT], mi[RPT], ii[RPT], md[RPT], dd[RPT], pm[RPT], pmm[RPT];
            uint32_t rb[RPT];
            bool pad[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int r = strip * ROWS + (int)lane * RPT + k - npad;
                pad[k] = r < 0;
                if (r >= 0) {
                    const uint32_t w = p.words[rd.word_off + r];
                    RowParams<T>::load(tb, w, mm[k], gapm[k], mi[k], ii[k], md[k], pm[k], pmm[k]);
                    dd[k] = ii[k];
                    rb[k] = w >> 28;
                } else {                                       // boundary-generator row
                    mm[k] = gapm[k] = mi[k] = ii[k] = md[k] = pm[k] = pmm[k] = (T)0;
                    dd[k] = (T)1;
                    rb[k] = 0u;
                }
            }
            const bool above_pad = strip * ROWS + (int)lane * RPT - 1 < npad;
            const bool from_scratch = strip > 0 && lane == 0;
            const bool to_scratch = strip + 1 < nstrips && lane == 31;
            const bool last = strip + 1 == nstrips && lane == 31;
            T Ml[RPT], Il[RPT], Dl[RPT];
            T upM = (T)0, upI = (T)0, upD = (T)0, dgM = (T)0, dgI = (T)0, dgD = (T)0;
#pragma unroll
            for (int k = 0; k < RPT; ++k) { Ml[k] = (T)0; Il[k] = (T)0; Dl[k] = (T)0; }
            for (int s = 0; s < H + 31; ++s) {
                const int c = s - (int)lane;                   // my column this step; outside [0, H): idle work, never stored
                if (c == 0) {                                  // column -1 state: M = I = D = 0, generator rows hold D = init
#pragma unroll
                    for (int k = 0; k < RPT; ++k) { Ml[k] = (T)0; Il[k] = (T)0; Dl[k] = pad[k] ? init : (T)0; }
                    dgM = (T)0; dgI = (T)0; dgD = above_pad ? init : (T)0;
                }
                const bool valid = c >= 0 && c < H;
                if (from_scratch && valid) { upM = sM[c]; upI = sI[c]; upD = sD[c]; }
                const uint32_t hbase = onehot(valid ? hb[c] : 0);
                T pM = dgM, pI = dgI, pD = dgD, uM = upM, uI = upI;
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const T oM = Ml[k], oI = Il[k], oD = Dl[k];
                    const T prior = (rb[k] & hbase) ? pm[k] : pmm[k];
                    T m, i, d;
                    if (FMA) {
                        T t = Num<T>::mul(pM, mm[k]);
                        t = Num<T>::fma(pI, gapm[k], t);
                        t = Num<T>::fma(pD, gapm[k], t);
                        m = Num<T>::mul(t, prior);
                        i = Num<T>::fma(uI, ii[k], Num<T>::mul(uM, mi[k]));
                        d = Num<T>::fma(oD, dd[k], Num<T>::mul(oM, md[k]));
                    } else {
                        T t = Num<T>::add(Num<T>::mul(pM, mm[k]), Num<T>::mul(pI, gapm[k]));
                        t = Num<T>::add(t, Num<T>::mul(pD, gapm[k]));
                        m = Num<T>::mul(t, prior);
                        i = Num<T>::add(Num<T>::mul(uM, mi[k]), Num<T>::mul(uI, ii[k]));
                        d = Num<T>::add(Num<T>::mul(oM, md[k]), Num<T>::mul(oD, dd[k]));
                    }
                    Ml[k] = m; Il[k] = i; Dl[k] = d;
                    pM = oM; pI = oI; pD = oD;
                    uM = m; uI = i;
                }
                dgM = upM; dgI = upI; dgD = upD;
                if (valid) {
                    if (last) { sumM = Num<T>::add(sumM, Ml[RPT - 1]); sumI = Num<T>::add(sumI, Il[RPT - 1]); }
                    if (to_scratch) { sM[c] = Ml[RPT - 1]; sI[c] = Il[RPT - 1]; sD[c] = Dl[RPT - 1]; }
                }
                const T nuM = __shfl_up_sync(0xffffffffu, Ml[RPT - 1], 1);
                const T nuI = __shfl_up_sync(0xffffffffu, Il[RPT - 1], 1);
                const T nuD = __shfl_up_sync(0xffffffffu, Dl[RPT - 1], 1);
                if (lane != 0) { upM = nuM; upI = nuI; upD = nuD; }
            }
            __syncwarp();                                      // scratch written by lane 31 is read by lane 0 of the next strip
        }
        if (lane == 31) {
            const T res = Num<T>::add(sumM, sumI);
            out[pr.out_idx] = res;
            if (sizeof(T) == 4 && res < (T)1e-28f) {            // pairhmm_internal.cpp:578: queue the FP64 recompute
                const uint32_t slot = atomicAdd(p.f64_count, 1u);
                p.f64_pairs[slot] = pr;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// phmm_pair64_kernel — the FP64 recompute (pairhmm_internal.cpp:578-581) for the common shapes, fast: one warp per pair, the
// read on all 32 lanes at RPT = ceil((R + 1) / 32) <= 5 rows per lane (one strip, nothing through memory), the haplotype's base
// classes staged once in shared memory, priors fetched from a per-lane table instead of selected, gap-continuation parameters
// as kernel-wide constants.  Same recurrence, same operation order and the same FP64 tables as phmm_pair_kernel<double>.
//
// Flush-to-zero without a test after every operation.  The reference runs its FP64 kernels with MXCSR.FTZ set, so a result
// below DBL_MIN becomes 0; the GPU has no FP64 FTZ mode, and phmm_pair_kernel<double> tests after each of the 8 operations of a
// cell (which costs more than the arithmetic).  Here: every value is >= 0, so only a PRODUCT can come out tiny, and a product
// of a state value v with a table constant c (>= 6.6e-14, the smallest prior) is tiny only if v < 3.4e-295.  State values are
// bounded below by their M term: I[r][c] >= M[r-1][c] * mi and D[r][c] >= M[r][c-1] * md with mi, md >= 2e-13 (quality 127).
// So as long as every nonzero M the sweep produces is >= 2^-830 (1.4e-250), every operand of every product is 0 or
// >= 2.8e-263, no result is tiny, and flushing would have changed nothing.  The kernel checks exactly that — two integer
// instructions per cell on the high word of M — and a pair that trips the check is handed to phmm_pair_kernel<double>, which
// flushes operation by operation.  (With the default gap-continuation quality a read of <= 159 bases cannot get there: one
// long insertion costs 10^-(R + 5), so every nonzero value stays above 10^-180 of the 2^1020 scale.  The check is what makes
// that argument unnecessary.)  Pairs whose read brings its own deletion / continuation qualities, reads over 159 bases and
// haplotypes over PAIR64_MAX_H go to the general kernel too.
// ------------------------------------------------------------------------------------------------
constexpr int PAIR64_MAX_RPT = 5;
constexpr uint32_t PAIR64_MAX_H = 2048;
constexpr uint32_t PAIR64_TSTRIDE = 48;            // bytes per lane and class: 5 doubles, padded to an odd number of 16-byte chunks

template <int RPT>
__device__ __forceinline__ bool pair64_sweep(const PhmmBatchDev& p, const PhmmRead& rd, const PhmmHap& hp, uint32_t lane, const uint8_t* s_cls,
                                             unsigned char* s_tbl, double gapm, double gcp_p, uint32_t trip_hi, double* result)
{
    const PhmmTables* __restrict__ tb = p.tables;
    const int H = (int)hp.len, R = (int)rd.len;
    const int npad = 32 * RPT - R;                               // >= 1
    double mm[RPT], mi[RPT], dd[RPT];
    {
        double pm[RPT], pmm[RPT];
        uint32_t rb[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = (int)lane * RPT + k - npad;
            if (r >= 0) {
                const uint32_t w = p.words[rd.word_off + r];
                const uint32_t q = w & 127u, qi = (w >> 7) & 127u;       // deletion quality == insertion quality here (no explicit ones)
                mm[k] = tb->mm_d[((qi * (qi + 1)) >> 1) + qi];
                mi[k] = tb->ph2pr_d[qi];
                dd[k] = gcp_p;
                pm[k] = __dsub_rn(1.0, tb->ph2pr_d[q]);
                pmm[k] = __ddiv_rn(tb->ph2pr_d[q], 3.0);
                rb[k] = w >> 28;
            } else {                                             // boundary-generator row
                mm[k] = 0.0; mi[k] = 0.0; dd[k] = 1.0; pm[k] = 0.0; pmm[k] = 0.0; rb[k] = 0u;
            }
        }
#pragma unroll
        for (uint32_t b = 0; b < 6; ++b) {
            const uint32_t hb = class_onehot(b);
            double* t = reinterpret_cast<double*>(s_tbl + b * (32 * PAIR64_TSTRIDE) + lane * PAIR64_TSTRIDE);
#pragma unroll
            for (int k = 0; k < RPT; ++k) t[k] = (rb[k] & hb) ? pm[k] : pmm[k];
        }
    }
    __syncwarp();
    const double init = __ddiv_rn(__longlong_as_double(0x7FB0000000000000LL), (double)H);      // 2^1020 / H
    const bool above_pad = (int)lane * RPT - 1 < npad;
    // column -1 state: M = I = D = 0, generator rows hold D = init.  A lane starts on column 0 at step `lane` and does nothing
    // before (so this state is still intact then); what it computes past column H - 1 is never kept.
    double Ml[RPT], Il[RPT], Dl[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) { Ml[k] = 0.0; Il[k] = 0.0; Dl[k] = (((int)lane * RPT + k) < npad) ? init : 0.0; }
    double upM = 0.0, upI = 0.0, upD = 0.0, dgM = 0.0, dgI = 0.0, dgD = above_pad ? init : 0.0, sumM = 0.0, sumI = 0.0;
    const unsigned char* tl = s_tbl + lane * PAIR64_TSTRIDE;
    bool small = false;
#pragma unroll 2
    for (int s = 0; s < H + 31; ++s) {
        const int c = s - (int)lane;                             // my column this step
        if (c >= 0) {
            const bool valid = c < H;
            const double* pr = reinterpret_cast<const double*>(tl + (valid ? (uint32_t)s_cls[c] : 5u) * (32 * PAIR64_TSTRIDE));
            double prior[RPT];
            if (RPT >= 2) {
                const double2 v = *reinterpret_cast<const double2*>(pr);
                prior[0] = v.x;
                prior[RPT >= 2 ? 1 : 0] = v.y;
            } else {
                prior[0] = pr[0];
            }
            if (RPT >= 4) {
                const double2 v = *reinterpret_cast<const double2*>(pr + 2);
                prior[RPT >= 4 ? 2 : 0] = v.x;
                prior[RPT >= 4 ? 3 : 0] = v.y;
            } else if (RPT == 3) {
                prior[RPT >= 3 ? 2 : 0] = pr[2];
            }
            if (RPT == 5) prior[RPT >= 5 ? 4 : 0] = pr[4];
            double pM = dgM, pI = dgI, pD = dgD, uM = upM, uI = upI;
            bool step_small = false;
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const double oM = Ml[k], oI = Il[k], oD = Dl[k];
                double t = __dmul_rn(pM, mm[k]);
                t = __fma_rn(pI, gapm, t);
                t = __fma_rn(pD, gapm, t);
                const double m = __dmul_rn(t, prior[k]);
                const double i = __fma_rn(uI, gcp_p, __dmul_rn(uM, mi[k]));
                const double d = __fma_rn(oD, dd[k], __dmul_rn(oM, mi[k]));
                step_small |= (uint32_t)(__double2hiint(m) - 1) < trip_hi - 1u;         // 0 < m < 2^-830 (trip_hi = 0x0C100000)
                Ml[k] = m; Il[k] = i; Dl[k] = d;
                pM = oM; pI = oI; pD = oD;
                uM = m; uI = i;
            }
            dgM = upM; dgI = upI; dgD = upD;
            if (valid) {
                small |= step_small;
                sumM = __dadd_rn(sumM, Ml[RPT - 1]);             // only the last lane's sums are read
                sumI = __dadd_rn(sumI, Il[RPT - 1]);
            }
        }
        const double nuM = __shfl_up_sync(0xffffffffu, Ml[RPT - 1], 1);
        const double nuI = __shfl_up_sync(0xffffffffu, Il[RPT - 1], 1);
        const double nuD = __shfl_up_sync(0xffffffffu, Dl[RPT - 1], 1);
        if (lane != 0) { upM = nuM; upI = nuI; upD = nuD; }
    }
    *result = __dadd_rn(sumM, sumI);
    return __any_sync(0xffffffffu, small);
}

__global__ void __launch_bounds__(128) phmm_pair64_kernel(PhmmBatchDev p, const PhmmPair* pairs, const uint32_t* count_dev, double* out,
                                                          PhmmPair* redo, uint32_t* redo_count, double gapm, double gcp_p, uint32_t trip_hi)
{
    __shared__ __align__(16) unsigned char s_tbl_all[4][6 * 32 * PAIR64_TSTRIDE];
    __shared__ uint8_t s_cls_all[4][PAIR64_MAX_H];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n = *count_dev;
    unsigned char* s_tbl = s_tbl_all[wib];
    uint8_t* s_cls = s_cls_all[wib];
    for (uint32_t pi = wid; pi < n; pi += nwarps) {
        const PhmmPair pr = pairs[pi];
        const PhmmRead rd = p.reads[pr.read];
        const PhmmHap hp = p.haps[pr.hap];
        const uint32_t flags = rd.model_flags >> 4;
        bool redo_it = rd.len == 0 || rd.len > 32u * PAIR64_MAX_RPT - 1u || hp.len == 0 || hp.len > PAIR64_MAX_H ||
                       (flags & (PHMM_RD_HAS_DEL | PHMM_RD_HAS_GCP)) != 0;
        double res = 0.0;
        if (!redo_it) {
            const uint8_t* hb = p.hap_bases + hp.byte_off;
            __syncwarp();                                        // the previous pair's tables and classes are no longer read
            for (uint32_t c = lane; c < hp.len; c += 32) s_cls[c] = (uint8_t)base_class(hb[c]);
            const uint32_t rpt = (rd.len + 32u) / 32u;           // ceil((R + 1) / 32)
            switch (rpt) {
                case 1: redo_it = pair64_sweep<1>(p, rd, hp, lane, s_cls, s_tbl, gapm, gcp_p, trip_hi, &res); break;
                case 2: redo_it = pair64_sweep<2>(p, rd, hp, lane, s_cls, s_tbl, gapm, gcp_p, trip_hi, &res); break;
                case 3: redo_it = pair64_sweep<3>(p, rd, hp, lane, s_cls, s_tbl, gapm, gcp_p, trip_hi, &res); break;
                case 4: redo_it = pair64_sweep<4>(p, rd, hp, lane, s_cls, s_tbl, gapm, gcp_p, trip_hi, &res); break;
                default: redo_it = pair64_sweep<5>(p, rd, hp, lane, s_cls, s_tbl, gapm, gcp_p, trip_hi, &res); break;
            }
        }
        if (lane == 31) {
            if (redo_it) redo[atomicAdd(redo_count, 1u)] = pr;
            else out[pr.out_idx] = res;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// phmm_epilogue_kernel — the reference's host side after the kernels, on the device: raw sums ->
// log10 (pairhmm_internal.cpp:578-584), normalize_likelihoods (:53-63), filter_poorly_modelled_evidence (:161-182),
// transpose_likelihood_matrix (:184-193) with its NaN tail, and the per-read best / second-best haplotype search the next
// step starts with (AlleleLikelihoods::search_best_allele, genotype/allele_likelihoods.hpp:589-657, as called by
// best_alleles_breaking_ties :364-383 from realign_reads_to_their_best_haplotype).
//
// log10f on the device: the reference's value is whatever the HOST's libm returns, so this is glibc's algorithm restated
// operation by operation (the fdlibm log10f formula around the table-driven logf; phmm_log10.cpp has the same restatement
// for the host, and that one is tested bit for bit against the host's libm before an engine is allowed to use either).
// Integer steps, IEEE double fma/mul/add, IEEE float mul/add/sub, one double->float rounding: the same bits on any IEEE
// machine.  A pair under MIN_ACCEPTED stands at -inf, which is what the reference as built computes and can never reach
// the matrix.
// One CTA per region; threads own reads (strided), a block scan in read order gives each surviving read its column.
// ------------------------------------------------------------------------------------------------
__constant__ double c_logf_invc[16] = {0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010b0p+0, 0x1.3c995b0b80385p+0,
                                       0x1.30d190c8864a5p+0, 0x1.25e227b0b8ea0p+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
                                       0x1.0953f419900a7p+0, 0x1p+0,               0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aa0p-1,
                                       0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
__constant__ double c_logf_logc[16] = {-0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,
                                       -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c8100p-3, -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,
                                       -0x1.252f438e10c1ep-5, 0x0p+0,                0x1.aa5aa5df25984p-5,  0x1.c5e53aa362eb4p-4,
                                       0x1.526e57720db08p-3,  0x1.bc2860d224770p-3,  0x1.1058bc8a07ee1p-2,  0x1.4043057b6ee09p-2};

__device__ __forceinline__ double log10_of_sum(float f)
{
    if (f < 1e-28f) return -CUDART_INF;                     // MIN_ACCEPTED (rovaca/common.h:9): the FP64 branch, see above
    int hx = __float_as_int(f);
    const int k = (hx >> 23) - 127;                          // the fdlibm split: f = 2^(k+neg) * m, m in [0.5, 2)
    const int neg = (int)((unsigned)k >> 31);
    hx = (hx & 0x007fffff) | ((0x7f - neg) << 23);
    const float y = __int2float_rn(k + neg);
    const unsigned tmp = (unsigned)hx - 0x3f330000u;          // logf(m): m = 2^kk * z, z in [OFF, 2*OFF)
    const int idx = (int)((tmp >> 19) & 15u);
    const int kk = (int)tmp >> 23;
    const unsigned iz = (unsigned)hx - (tmp & 0xff800000u);
    const double z = (double)__uint_as_float(iz);
    const double r = __fma_rn(z, c_logf_invc[idx], -1.0);
    const double y0 = __fma_rn((double)kk, 0x1.62e42fefa39efp-1, c_logf_logc[idx]);
    const double r2 = __dmul_rn(r, r);
    double p = __fma_rn(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    p = __fma_rn(-0x1.00ea348b88334p-2, r2, p);
    p = __fma_rn(p, r2, __dadd_rn(y0, r));
    const float lf = __double2float_rn(p);
    const float zz = __fadd_rn(__fmul_rn(y, 7.9034151668e-07f), __fmul_rn(4.3429449201e-01f, lf));
    const float res = __fadd_rn(zz, __fmul_rn(y, 3.0102920532e-01f));
    return (double)__fsub_rn(res, 36.1236000061f);          // LOG10_INITIAL_CONSTANT_F (adapter.h:25): float subtract, then widen
}

__global__ void __launch_bounds__(128) phmm_log10_probe_kernel(const float* x, double* out, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = log10_of_sum(x[i]);
}

__global__ void __launch_bounds__(128) phmm_epilogue_kernel(PhmmEpilogueDev e)
{
    __shared__ uint32_t warp_kept[4];
    __shared__ uint32_t base_kept, under_sum;
    const PhmmRegionDev rg = e.regions[blockIdx.x];
    const uint32_t nr = rg.n_reads, nh = rg.n_haps;
    if (nr == 0 || nh == 0) {
        if (threadIdx.x == 0) e.n_kept[blockIdx.x] = e.n_under[blockIdx.x] = 0;
        return;
    }
    uint32_t under = 0;                                      // pairs below MIN_ACCEPTED (:578), reported like the host path does
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* out = e.lik + rg.out_off;
    if (threadIdx.x == 0) base_kept = under_sum = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < nr; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const float* row = e.raw + rg.out_off + (size_t)i * nh;
        bool keep = false;
        double cmp = 0.0;
        if (i < nr) {
            // std::max_element semantics of :56,:170 — the first element, replaced only by a strictly greater one
            double best = log10_of_sum(row[0]);
            under += row[0] < 1e-28f;
            for (uint32_t j = 1; j < nh; ++j) {
                const double v = log10_of_sum(row[j]);
                under += row[j] < 1e-28f;
                if (best < v) best = v;
            }
            cmp = __dadd_rn(best, -4.5);                     // normalize_likelihoods caps at best - 4.5 (adapter.h:23)
            const int len = (int)e.reads[rg.read_begin + i].len;
            const double lim = fmin(2.0, ceil(__dmul_rn((double)len, 0.02))) * (-4.0);   // adapter.h:20-22
            keep = !(best < lim);
        }
        // column of every surviving read: exclusive scan of the keep flags in read order
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_kept[warp] = (uint32_t)__popc(ballot);
        __syncthreads();
        uint32_t pos = base_kept + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
        for (uint32_t w = 0; w < warp; ++w) pos += warp_kept[w];
        if (keep) {
            uint32_t bi = 0, si = 0;
            double bl = 0.0, sl = -CUDART_INF;
            for (uint32_t j = 0; j < nh; ++j) {
                const double v = log10_of_sum(row[j]);
                const double c = v < cmp ? cmp : v;          // :58-61
                out[(size_t)j * nr + pos] = c;               // transpose :184-193
                // search_best_allele, first pass (allele_likelihoods.hpp:609-622), can_be_reference = true
                if (j == 0) bl = c;
                else if (c > bl) { si = bi; bi = j; sl = bl; bl = c; }
                else if (c > sl) { si = j; sl = c; }
            }
            e.kept_idx[rg.read_begin + pos] = i;
            if (e.best) {
                if (e.hap_priority && bl - sl < 0.2) {       // ties broken by priority (:624-645); 0.2 = s_log_10_informative_threshold
                    const double* pri = e.hap_priority + rg.hap_begin;
                    double bp = pri[bi], sp = pri[si];
                    for (uint32_t a = 0; a < nh; ++a) {
                        const double c = out[(size_t)a * nr + pos];
                        if (a == bi || bl - c > 0.2) continue;
                        const double cp = pri[a];
                        if (cp > bp) { si = bi; bi = a; sp = bp; bp = cp; }
                        else if (cp > sp) { si = a; sp = cp; }
                    }
                }
                e.best[rg.read_begin + pos] = bi;
                e.second[rg.read_begin + pos] = si;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) base_kept += warp_kept[0] + warp_kept[1] + warp_kept[2] + warp_kept[3];
        __syncthreads();
    }
    const uint32_t kept = base_kept;
    if (under) atomicAdd(&under_sum, under);
    __syncthreads();
    if (threadIdx.x == 0) {
        e.n_kept[blockIdx.x] = kept;
        e.n_under[blockIdx.x] = under_sum;
    }
    // the NaN tail of the filtered reads (the matrix is allocated for every read before the filter runs, :516)
    const uint32_t tail = nr - kept;
    for (uint32_t t = threadIdx.x; t < tail * nh; t += blockDim.x) out[(size_t)(t / tail) * nr + kept + t % tail] = __longlong_as_double(0x7FF8000000000000LL);   // the host's (double)NAN, sign bit clear
}

// ------------------------------------------------------------------------------------------------
// phmm_marginalize_kernel — AlleleLikelihoods::marginalize -> marginal_likelihoods_direct (genotype/allele_likelihoods.hpp:189-216,
// 702-732) for many variant sites at once, on the matrices the epilogue left on the device: site s of
// region g maps the region's haplotypes onto n_new alleles; out[n][e] = the largest lik[a][e] over the haplotypes a with
// map[a] == n, -inf when there is none (:714), for the region's surviving reads e (columns past n_kept are the matrix's NaN
// tail and stay NaN).  One CTA per site, threads own evidence columns, so the running maximum needs no atomics; the compare
// is the reference's `if (in > out) out = in` (:723-726).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) phmm_marginalize_kernel(PhmmEpilogueDev e, const PhmmMargSite* sites, const int32_t* maps, double* out)
{
    const PhmmMargSite st = sites[blockIdx.x];
    const PhmmRegionDev rg = e.regions[st.region];
    const uint32_t nr = rg.n_reads, nh = rg.n_haps, kept = e.n_kept[st.region];
    const double* lik = e.lik + rg.out_off;
    const int32_t* map = maps + st.map_off;
    double* o = out + st.out_off;
    for (uint32_t col = threadIdx.x; col < nr; col += blockDim.x) {
        if (col >= kept) {
            for (uint32_t n = 0; n < st.n_new; ++n) o[(size_t)n * nr + col] = __longlong_as_double(0x7FF8000000000000LL);
            continue;
        }
        for (uint32_t n = 0; n < st.n_new; ++n) o[(size_t)n * nr + col] = -CUDART_INF;
        for (uint32_t a = 0; a < nh; ++a) {
            const int32_t n = map[a];
            if (n < 0 || (uint32_t)n >= st.n_new) continue;          // s_missing_index (:717); the host rejects n >= n_new
            const double v = lik[(size_t)a * nr + col];
            double* cell = o + (size_t)n * nr + col;
            if (v > *cell) *cell = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Issue-rate micro-benchmarks (bench.py's measured FP32 peak; DESIGN.md's pipe model).  Every thread runs
// independent dependency chains long enough to hide the 4-cycle FMA latency.
//   mode 0: scalar FFMA                     mode 1: packed FFMA2 (two FMAs per instruction)
//   mode 2: FFMA2 + (LOP3->P, FSEL) 2:1     mode 3: mode 2 + one SHFL.UP per 8 FFMA2
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) phmm_ubench_kernel(float* out, uint32_t iters, float a, float b, uint32_t m)
{
    const float t = (float)threadIdx.x * 1e-6f;
    float2 x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = make_float2(t + (float)k, t - (float)k);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    uint32_t w = m + threadIdx.x;
    float sel = 0.f;
    for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) {
                    x[k].x = __fmaf_rn(x[k].x, a, b);
                    x[k].y = __fmaf_rn(x[k].y, a, b);
                } else {
                    x[k] = __ffma2_rn(x[k], aa, bb);
                }
                if (MODE >= 2 && (k & 1)) {
                    sel += ((w >> (k + r)) & m) ? a : b;     // LOP3 -> predicate, FSEL (+ the FADD, FP32 pipe)
                }
            }
            if (MODE == 3) x[r].x = __shfl_up_sync(0xffffffffu, x[r].x, 1);
            w = w * 1664525u + 1013904223u;
        }
    }
    float acc = sel;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += x[k].x + x[k].y;
    if (acc == 123.456f) out[0] = acc;
}

template <int RPT, int VAR, int FIXED = 0, int DUO = 0>
cudaError_t launch_fwd_t(const PhmmBatchDev& b, uint32_t cta_base, uint32_t n_ctas, uint32_t cta_warps, size_t smem, cudaStream_t st)
{
    auto kern = phmm_fwd_kernel<RPT, VAR, FIXED, DUO>;
    // Raise the dynamic shared-memory limit of this instantiation once, to the device maximum: several host workers launch
    // concurrently, and a per-launch "set to what I need" lets one worker lower the limit under another's launch.
    static std::atomic<uint64_t> limit_set{0};            // one bit per device ordinal
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = 1ull << (dev & 63);
    if (!(limit_set.load(std::memory_order_acquire) & bit)) {
        int optin = 0;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (e != cudaSuccess) return e;
        limit_set.fetch_or(bit, std::memory_order_release);
    }
    kern<<<n_ctas, cta_warps * 32, smem, st>>>(b, cta_base);
    return cudaGetLastError();
}

template <int RPT, int VAR>
cudaError_t info_t(PhmmKernelInfo* info)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, phmm_fwd_kernel<RPT, VAR, 0>);
    if (e != cudaSuccess) return e;
    info->regs = a.numRegs;
    info->local_bytes = (int)a.localSizeBytes;
    info->max_threads = a.maxThreadsPerBlock;
    return cudaSuccess;
}

}  // namespace

cudaError_t phmm_launch_prep(const PhmmBatchDev& b, cudaStream_t st)
{
    if (b.n_reads == 0) return cudaSuccess;
    const uint32_t warps_per_block = 8;
    if (b.n_reads <= kPrepSplitReads) {                   // a small batch: four warps per read (latency)
        const uint32_t blocks = (b.n_reads * 4 + warps_per_block - 1) / warps_per_block;
        phmm_prep_kernel<4><<<blocks, warps_per_block * 32, 0, st>>>(b);
    } else {
        const uint32_t blocks = (b.n_reads + warps_per_block - 1) / warps_per_block;
        phmm_prep_kernel<1><<<blocks, warps_per_block * 32, 0, st>>>(b);
    }
    return cudaGetLastError();
}

// rows per lane 1..PHMM_MAX_RPT_GENERIC exist in both arithmetic flavours for any per-base gap qualities, 1..PHMM_MAX_RPT in the
// (default, contracted) flavour for the reference's own gap qualities
#define PHMM_CASE(R, FN, ...)                                                                                        \
    if (rpt == R) {                                                                                                  \
        if (refgaps) return fma ? FN<R, PHMM_VAR_REFGAPS>(__VA_ARGS__) : cudaErrorInvalidValue;                      \
        if (R <= PHMM_MAX_RPT_GENERIC)                                                                               \
            return fma ? FN<(R <= PHMM_MAX_RPT_GENERIC ? R : 1), PHMM_VAR_GENERIC>(__VA_ARGS__)                      \
                       : FN<(R <= PHMM_MAX_RPT_GENERIC ? R : 1), PHMM_VAR_GENERIC_NOFMA>(__VA_ARGS__);               \
        return cudaErrorInvalidValue;                                                                                \
    }
#define PHMM_DISPATCH_RPT(FN, ...)                                                                                   \
    PHMM_CASE(1, FN, __VA_ARGS__) PHMM_CASE(2, FN, __VA_ARGS__) PHMM_CASE(3, FN, __VA_ARGS__) PHMM_CASE(4, FN, __VA_ARGS__)     \
    PHMM_CASE(5, FN, __VA_ARGS__) PHMM_CASE(6, FN, __VA_ARGS__) PHMM_CASE(7, FN, __VA_ARGS__) PHMM_CASE(8, FN, __VA_ARGS__)     \
    PHMM_CASE(9, FN, __VA_ARGS__) PHMM_CASE(10, FN, __VA_ARGS__) PHMM_CASE(11, FN, __VA_ARGS__) PHMM_CASE(12, FN, __VA_ARGS__)  \
    PHMM_CASE(13, FN, __VA_ARGS__) PHMM_CASE(14, FN, __VA_ARGS__) PHMM_CASE(15, FN, __VA_ARGS__) PHMM_CASE(16, FN, __VA_ARGS__) \
    PHMM_CASE(17, FN, __VA_ARGS__) PHMM_CASE(18, FN, __VA_ARGS__) PHMM_CASE(19, FN, __VA_ARGS__)                                \
    return cudaErrorInvalidValue;

cudaError_t phmm_launch_fwd(const PhmmBatchDev& b, int rpt, bool refgaps, bool fma, uint32_t fixed_lanes, bool duo, uint32_t cta_base,
                            uint32_t n_ctas, uint32_t cta_warps, size_t smem_bytes, cudaStream_t st)
{
    if (n_ctas == 0) return cudaSuccess;
    if (rpt < 1 || rpt > phmm_max_rpt(refgaps) || cta_warps < 1 || cta_warps > PHMM_WARPS) return cudaErrorInvalidValue;
    if (duo) {                                       // paired reads: one instantiation, see PHMM_DUO_RPT
        if (rpt != PHMM_DUO_RPT || !refgaps || !fma || fixed_lanes) return cudaErrorInvalidValue;
        return launch_fwd_t<PHMM_DUO_RPT, PHMM_VAR_REFGAPS, 0, 1>(b, cta_base, n_ctas, cta_warps, smem_bytes, st);
    }
    if (fixed_lanes) {                               // the one specialisation: see phmm_fixed_lanes()
        if (fixed_lanes != phmm_fixed_lanes(rpt, refgaps) || !fma) return cudaErrorInvalidValue;
        return launch_fwd_t<PHMM_MAX_RPT, PHMM_VAR_REFGAPS, 8>(b, cta_base, n_ctas, cta_warps, smem_bytes, st);
    }
    PHMM_DISPATCH_RPT(launch_fwd_t, b, cta_base, n_ctas, cta_warps, smem_bytes, st)
}

cudaError_t phmm_fwd_kernel_info(int rpt, bool refgaps, bool fma, PhmmKernelInfo* info)
{
    if (rpt < 1 || rpt > phmm_max_rpt(refgaps)) return cudaErrorInvalidValue;
    PHMM_DISPATCH_RPT(info_t, info)
}

cudaError_t phmm_launch_pairs_f32(const PhmmBatchDev& b, bool fma, const PhmmPair* pairs, uint32_t count, float* scratch,
                                  uint32_t threads, uint32_t max_cols, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    const uint32_t blocks = threads / 128;
    if (fma)
        phmm_pair_kernel<float, true><<<blocks, 128, 0, st>>>(b, pairs, nullptr, count, scratch, max_cols, b.out_f32);
    else
        phmm_pair_kernel<float, false><<<blocks, 128, 0, st>>>(b, pairs, nullptr, count, scratch, max_cols, b.out_f32);
    return cudaGetLastError();
}

cudaError_t phmm_launch_pairs_f64(const PhmmBatchDev& b, const PhmmPair* pairs, const uint32_t* count_dev, double* scratch,
                                  uint32_t threads, uint32_t max_cols, PhmmPair* redo, uint32_t* redo_count, double gapm_d, double gcp_d,
                                  uint32_t trip_hi, cudaStream_t st)
{
    const uint32_t blocks = threads / 128;
    // the common shapes one warp per pair without a flush test per operation; what that kernel hands back (and every pair when
    // there is no redo list) through the general kernel
    if (redo) {
        phmm_pair64_kernel<<<blocks * 2, 128, 0, st>>>(b, pairs, count_dev, b.out_f64, redo, redo_count, gapm_d, gcp_d, trip_hi);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        phmm_pair_kernel<double, true><<<blocks, 128, 0, st>>>(b, redo, redo_count, 0u, scratch, max_cols, b.out_f64);
    } else {
        phmm_pair_kernel<double, true><<<blocks, 128, 0, st>>>(b, pairs, count_dev, 0u, scratch, max_cols, b.out_f64);
    }
    return cudaGetLastError();
}

cudaError_t phmm_launch_ubench(int mode, uint32_t blocks, uint32_t iters, float* out, cudaStream_t st)
{
    switch (mode) {
        case 0: phmm_ubench_kernel<0><<<blocks, 256, 0, st>>>(out, iters, 0.999f, 1e-3f, 0x1u); break;
        case 1: phmm_ubench_kernel<1><<<blocks, 256, 0, st>>>(out, iters, 0.999f, 1e-3f, 0x1u); break;
        case 2: phmm_ubench_kernel<2><<<blocks, 256, 0, st>>>(out, iters, 0.999f, 1e-3f, 0x1u); break;
        case 3: phmm_ubench_kernel<3><<<blocks, 256, 0, st>>>(out, iters, 0.999f, 1e-3f, 0x1u); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t phmm_launch_epilogue(const PhmmEpilogueDev& e, cudaStream_t st)
{
    if (e.n_regions == 0) return cudaSuccess;
    phmm_epilogue_kernel<<<e.n_regions, 128, 0, st>>>(e);
    return cudaGetLastError();
}

cudaError_t phmm_launch_marginalize(const PhmmEpilogueDev& e, const PhmmMargSite* sites, const int32_t* maps, uint32_t n_sites, double* out,
                                    cudaStream_t st)
{
    if (n_sites == 0) return cudaSuccess;
    phmm_marginalize_kernel<<<n_sites, 128, 0, st>>>(e, sites, maps, out);
    return cudaGetLastError();
}

cudaError_t phmm_launch_log10_probe(const float* x, double* out, uint32_t n, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    phmm_log10_probe_kernel<<<(n + 127) / 128, 128, 0, st>>>(x, out, n);
    return cudaGetLastError();
}
