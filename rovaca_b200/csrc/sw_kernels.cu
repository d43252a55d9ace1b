// sw_kernels.cu — batched Smith-Waterman for the step right after the PairHMM path: the reference
// realigns every read to its best haplotype with sw->avx_function(sw) (common/base/include/smithwaterman_common.h:95, called
// at genotype/utils/alignment_utils.cpp:574-579) = sw_avx512_run / sw_avx2_run (common/base/lib/sw_avx/sw_avx.c, sw_avx2.c),
// one alignment per call on a CPU thread.  Here: many alignments per launch, results identical to the reference's
// (score, alignment offset, CIGAR).
//
//   sw_fill_kernel<CPT>  one warp per alignment.  Lane l owns columns l*CPT+1 .. l*CPT+CPT of the read, the warp sweeps the
//                        rows of the haplotype as an anti-diagonal wavefront (lane l is at row s - l + 1 in step s), H and E
//                        of a lane's last column reach the next lane through two __shfl_up_sync per step.  Per cell the
//                        reference's Gotoh recurrence with its 4-bit back-track code (sw_avx.c:40-76); the codes of a lane's
//                        CPT cells are packed into words and stored [step][lane] (the lanes of a step write one contiguous line),
//                        the last row and last column of H are kept for the end-cell search.
//   sw_trace_kernel      one thread per alignment: the reference's end-cell choice with its tie rules in anti-diagonal order
//                        (sw_avx.c:196-217), its trace-back state machine and CIGAR compaction (:252-361).
// Integer arithmetic only; nothing here depends on floating point.
#include <cuda_runtime.h>
#include <stdint.h>

#include "phmm_kernels.h"

namespace
{

constexpr int32_t SW_MIN_CUTOFF = -100000000;              // MATRIX_MIN_CUTOFF, sw_avx.c:18
constexpr int32_t SW_LOW_INIT = INT32_MIN / 2;             // LOW_INIT_VALUE, :19
enum { BT_MATCH = 0, BT_INS = 1, BT_DEL = 2, BT_INS_EXT = 4, BT_DEL_EXT = 8 };   // smithwaterman_common.h:40-50

template <int CPT>
__global__ void __launch_bounds__(128) sw_fill_kernel(PhmmSwBatchDev b, uint32_t first, uint32_t count)
{
    constexpr int W = (CPT + 7) / 8;                       // back-track words per lane and row
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = wid; t < count; t += nwarps) {
        const PhmmSwTask task = b.tasks[b.order[first + t]];
        const int nrow = (int)task.ref_len, ncol = (int)task.read_len;
        const uint8_t* seq1 = b.bytes + task.ref_off;
        const uint8_t* seq2 = b.bytes + task.read_off;
        uint32_t* bt = b.bt + task.bt_off;
        int32_t* hrow = b.aux + task.aux_off;              // H(nrow, 1..ncol)
        int32_t* hcol = hrow + ncol;                       // H(1..nrow, ncol)
        const int open = b.gap_open, ext = b.gap_extend, match = b.match, mismatch = b.mismatch;
        const int j0 = (int)lane * CPT;                    // my columns are j0+1 .. j0+CPT
        uint32_t c2[CPT];
        int32_t Hp[CPT], Fv[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            c2[k] = (j0 + k < ncol) ? seq2[j0 + k] : 0x100u;   // never equal to a byte of seq1
            Hp[k] = 0;                                     // H(0, j) = 0
            Fv[k] = SW_LOW_INIT;                           // F(0, j)
        }
        const int kcol = ncol - 1 - j0;                    // my register that holds column ncol, if 0 <= kcol < CPT
        int32_t lastH = 0, lastE = SW_LOW_INIT, diagH = 0;
        for (int s = 0; s < nrow + 31; ++s) {
            int32_t inH = __shfl_up_sync(0xffffffffu, lastH, 1);
            int32_t inE = __shfl_up_sync(0xffffffffu, lastE, 1);
            if (lane == 0) { inH = 0; inE = SW_LOW_INIT; }      // H(i, 0) = 0, E(i, 0) = LOW
            const int i = s - (int)lane + 1;
            if (i >= 1 && i <= nrow) {
                const uint32_t c1 = seq1[i - 1];
                int32_t hl = inH, e = inE, hd = diagH;
                uint32_t words[W];
#pragma unroll
                for (int w = 0; w < W; ++w) words[w] = 0u;
                int32_t hc = 0;
#pragma unroll
                for (int k = 0; k < CPT; ++k) {            // MAIN_CODE, sw_avx.c:40-76
                    const int32_t ext_h = e + ext, open_h = hl + open;
                    const int32_t e11 = max(open_h, ext_h);
                    uint32_t code = (open_h > ext_h) ? 0u : (uint32_t)BT_INS_EXT;
                    const int32_t ext_v = Fv[k] + ext, open_v = Hp[k] + open;
                    const int32_t f11 = max(ext_v, open_v);
                    code |= (open_v > ext_v) ? 0u : (uint32_t)BT_DEL_EXT;
                    const int32_t m11 = hd + (c1 == c2[k] ? match : mismatch);
                    int32_t h11 = max(SW_MIN_CUTOFF, m11);
                    uint32_t dir = (e11 > h11) ? (uint32_t)BT_INS : (uint32_t)BT_MATCH;
                    h11 = max(h11, e11);
                    if (f11 > h11) dir = (uint32_t)BT_DEL;
                    h11 = max(h11, f11);
                    words[k / 8] |= (code | dir) << (4 * (k % 8));
                    hd = Hp[k];
                    Hp[k] = h11;
                    Fv[k] = f11;
                    hl = h11;
                    e = e11;
                    if (k == kcol) hc = h11;
                }
                lastH = hl;
                lastE = e;
                diagH = inH;
#pragma unroll
                for (int w = 0; w < W; ++w) bt[((size_t)s * 32 + lane) * W + w] = words[w];   // [step][lane]: one coalesced line per step
                if (kcol >= 0 && kcol < CPT) hcol[i - 1] = hc;
                if (i == nrow) {
#pragma unroll
                    for (int k = 0; k < CPT; ++k)
                        if (j0 + k < ncol) hrow[j0 + k] = Hp[k];
                }
            }
        }
    }
}

// The same fill for TWO alignments per warp in the 16-bit halves of packed registers (VIADD.16x2, VIMNMX.S16x2 with its two
// predicate outputs: the back-track bits are the predicates of the maxima).  For alignments whose scores fit 16 bits with
// room to spare — the host sends a task here only if min(rows, columns) x max(|match|, |mismatch|) + |open| + |extend| <= 12000
// (H is bounded by a diagonal run of matches above and of mismatches below; E and F by H + open), which holds for the
// realignment parameter set up to ~800-base reads and never for the assembler's.  LOW is -16384 here: like the 32-bit
// LOW_INIT_VALUE it only ever loses the comparison it enters (LOW + extend < any H + open), so every predicate — every
// back-track bit — and every H is the 32-bit kernel's.  Results per alignment go where sw_fill_kernel puts them.
template <int CPT>
__global__ void __launch_bounds__(128) sw_fill16_kernel(PhmmSwBatchDev b, uint32_t first, uint32_t count)
{
    constexpr int W = (CPT + 7) / 8;
    constexpr uint32_t LOW2 = 0xC000C000u;                 // -16384 in both halves
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t npair = (count + 1) / 2;
    auto pack = [](int v) { return ((uint32_t)v & 0xffffu) * 0x00010001u; };
    const uint32_t open2 = pack(b.gap_open), ext2 = pack(b.gap_extend), match2 = pack(b.match);
    const uint32_t drop = (uint32_t)(b.match - b.mismatch);            // match - mismatch > 0: what a mismatch loses
    for (uint32_t t = wid; t < npair; t += nwarps) {
        const bool hasB = 2 * t + 1 < count;
        const PhmmSwTask ta = b.tasks[b.order[first + 2 * t]];
        const PhmmSwTask tb = b.tasks[b.order[first + 2 * t + (hasB ? 1 : 0)]];
        const int nrowA = (int)ta.ref_len, ncolA = (int)ta.read_len, nrowB = (int)tb.ref_len, ncolB = (int)tb.read_len;
        const int nrow = max(nrowA, nrowB);
        const uint8_t* seq1A = b.bytes + ta.ref_off;
        const uint8_t* seq1B = b.bytes + tb.ref_off;
        const uint8_t* seq2A = b.bytes + ta.read_off;
        const uint8_t* seq2B = b.bytes + tb.read_off;
        uint32_t* btA = b.bt + ta.bt_off;
        uint32_t* btB = b.bt + tb.bt_off;
        int32_t* hrowA = b.aux + ta.aux_off;
        int32_t* hcolA = hrowA + ncolA;
        int32_t* hrowB = b.aux + tb.aux_off;
        int32_t* hcolB = hrowB + ncolB;
        const int j0 = (int)lane * CPT;
        uint32_t c2[CPT], Hp[CPT], Fv[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const uint32_t a = (j0 + k < ncolA) ? seq2A[j0 + k] : 0x100u;      // never equal to a byte of seq1
            const uint32_t bb = (j0 + k < ncolB) ? seq2B[j0 + k] : 0x100u;
            c2[k] = a | (bb << 16);
            Hp[k] = 0u;
            Fv[k] = LOW2;
        }
        const int kcolA = ncolA - 1 - j0, kcolB = ncolB - 1 - j0;
        uint32_t lastH = 0u, lastE = LOW2, diagH = 0u;
        for (int s = 0; s < nrow + 31; ++s) {
            uint32_t inH = __shfl_up_sync(0xffffffffu, lastH, 1);
            uint32_t inE = __shfl_up_sync(0xffffffffu, lastE, 1);
            if (lane == 0) { inH = 0u; inE = LOW2; }
            const int i = s - (int)lane + 1;
            if (i >= 1 && i <= nrow) {
                const bool rowA = i <= nrowA, rowB = i <= nrowB;
                const uint32_t c1 = (rowA ? (uint32_t)seq1A[i - 1] : 0x200u) | ((rowB ? (uint32_t)seq1B[i - 1] : 0x200u) << 16);
                uint32_t hl = inH, e = inE, hd = diagH;
                uint32_t wa[W], wb[W];
#pragma unroll
                for (int w = 0; w < W; ++w) { wa[w] = 0u; wb[w] = 0u; }
                uint32_t hc = 0u;
#pragma unroll
                for (int k = 0; k < CPT; ++k) {            // MAIN_CODE, sw_avx.c:40-76, two alignments per instruction
                    bool eh_b, eh_a, ev_b, ev_a, kh_b, kh_a, kf_b, kf_a;          // (high half = alignment B, low half = A)
                    const uint32_t e11 = __vibmax_s16x2(__vadd2(e, ext2), __vadd2(hl, open2), &eh_b, &eh_a);       // extension wins or ties
                    const uint32_t f11 = __vibmax_s16x2(__vadd2(Fv[k], ext2), __vadd2(Hp[k], open2), &ev_b, &ev_a);
                    const uint32_t ne = __vminu2(c1 ^ c2[k], 0x00010001u);        // 1 where the bases differ
                    const uint32_t m11 = __vsub2(__vadd2(hd, match2), ne * drop); // diagonal + (match or mismatch); no carry: ne is 0 / 1 per half
                    uint32_t h11 = __vibmax_s16x2(m11, e11, &kh_b, &kh_a);        // diagonal >= E (the cut-off at -1e8 cannot bind here)
                    h11 = __vibmax_s16x2(h11, f11, &kf_b, &kf_a);                 // max so far >= F
                    const uint32_t ca = (eh_a ? (uint32_t)BT_INS_EXT : 0u) | (ev_a ? (uint32_t)BT_DEL_EXT : 0u) |
                                        (kf_a ? (kh_a ? (uint32_t)BT_MATCH : (uint32_t)BT_INS) : (uint32_t)BT_DEL);
                    const uint32_t cb = (eh_b ? (uint32_t)BT_INS_EXT : 0u) | (ev_b ? (uint32_t)BT_DEL_EXT : 0u) |
                                        (kf_b ? (kh_b ? (uint32_t)BT_MATCH : (uint32_t)BT_INS) : (uint32_t)BT_DEL);
                    wa[k / 8] += ca << (4 * (k % 8));
                    wb[k / 8] += cb << (4 * (k % 8));
                    hd = Hp[k];
                    Hp[k] = h11;
                    Fv[k] = f11;
                    hl = h11;
                    e = e11;
                    if (k == kcolA) hc = (hc & 0xffff0000u) | (h11 & 0xffffu);
                    if (k == kcolB) hc = (hc & 0x0000ffffu) | (h11 & 0xffff0000u);
                }
                lastH = hl;
                lastE = e;
                diagH = inH;
                if (rowA) {
#pragma unroll
                    for (int w = 0; w < W; ++w) btA[((size_t)s * 32 + lane) * W + w] = wa[w];
                    if (kcolA >= 0 && kcolA < CPT) hcolA[i - 1] = (int32_t)(int16_t)(hc & 0xffffu);
                    if (i == nrowA) {
#pragma unroll
                        for (int k = 0; k < CPT; ++k)
                            if (j0 + k < ncolA) hrowA[j0 + k] = (int32_t)(int16_t)(Hp[k] & 0xffffu);
                    }
                }
                if (rowB && hasB) {
#pragma unroll
                    for (int w = 0; w < W; ++w) btB[((size_t)s * 32 + lane) * W + w] = wb[w];
                    if (kcolB >= 0 && kcolB < CPT) hcolB[i - 1] = (int32_t)(int16_t)(hc >> 16);
                    if (i == nrowB) {
#pragma unroll
                        for (int k = 0; k < CPT; ++k)
                            if (j0 + k < ncolB) hrowB[j0 + k] = (int32_t)(int16_t)(Hp[k] >> 16);
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ uint32_t bt_code(const uint32_t* bt, int cpt, int i, int j)
{
    const int w = (cpt + 7) / 8, lane = (j - 1) / cpt, k = (j - 1) % cpt;
    const int step = i - 1 + lane;                        // the fill kernel's lane `lane` was at row i in this step
    return (bt[((size_t)step * 32 + lane) * w + k / 8] >> (4 * (k % 8))) & 15u;
}

__global__ void __launch_bounds__(128) sw_trace_kernel(PhmmSwBatchDev b, uint32_t first, uint32_t count)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint32_t id = b.order[first + t];
    const PhmmSwTask task = b.tasks[id];
    const int nrow = (int)task.ref_len, ncol = (int)task.read_len, cpt = (int)task.cpt;
    const uint32_t* bt = b.bt + task.bt_off;
    const int32_t* hrow = b.aux + task.aux_off;
    const int32_t* hcol = hrow + ncol;
    PhmmSwOut out;
    out.status = 0;
    // end cell: last row and last column as the anti-diagonals pass them, row candidate first (sw_avx.c:196-217)
    int32_t max_score = INT32_MIN, max_i = 0, max_j = 0;
    for (int ad = 1; ad <= nrow + ncol; ++ad) {
        const int ilo = min(ad, nrow + 1), jhi = min(ad, ncol + 1);
        const int ihi = ad - jhi, jlo = ad - ilo;
        if (ilo == nrow + 1) {
            const int32_t s = hrow[jlo];                  // H(nrow, jlo + 1)
            if (max_score < s || (max_score == s && abs(ilo - jlo - 2) < abs(max_i - max_j))) {
                max_score = s; max_i = ilo - 1; max_j = jlo + 1;
            }
        }
        if (jhi == ncol + 1) {
            const int32_t s = hcol[ihi];                  // H(ihi + 1, ncol)
            if (max_score < s || (max_score == s && (max_j == ncol || abs(ihi - jhi + 2) <= abs(max_i - max_j)))) {
                max_score = s; max_i = ihi + 1; max_j = jhi - 1;
            }
        }
    }
    out.score = max_score;
    // trace-back (sw_avx_get_cigar, :252-361); neighbours of one kind are merged as they are produced, which is what the
    // reference's second pass (:337-350) does to its raw elements; the elements come out last-to-first
    uint32_t* cig = b.cigars + (size_t)id * b.cigar_cap;
    uint32_t n = 0;
    int cur_op = -1;
    uint32_t cur_len = 0;
    bool overflow = false;
    auto emit = [&](int op, uint32_t len) {
        if (op == cur_op) { cur_len += len; return; }
        if (cur_op >= 0) {
            if (n < b.cigar_cap) cig[n++] = (cur_len << 4) | (uint32_t)cur_op; else overflow = true;
        }
        cur_op = op;
        cur_len = len;
    };
    int i = max_i, j = max_j, state = 0;
    if (j < ncol) emit(4, (uint32_t)(ncol - j));          // soft clip, hc_marco.h:26
    while (i > 0 && j > 0) {
        const uint32_t btr = bt_code(bt, cpt, i, j);
        if (state == BT_INS_EXT) {
            --j; emit(1, 1u); state = (int)(btr & BT_INS_EXT);
        } else if (state == BT_DEL_EXT) {
            --i; emit(2, 1u); state = (int)(btr & BT_DEL_EXT);
        } else {
            switch (btr & 3u) {
                case BT_MATCH: --i; --j; emit(0, 1u); state = 0; break;
                case BT_INS: --j; emit(1, 1u); state = (int)(btr & BT_INS_EXT); break;
                case BT_DEL: --i; emit(2, 1u); state = (int)(btr & BT_DEL_EXT); break;
                default: i = 0; out.status = 3; break;     // cannot happen: a DEL replaces an INS, never joins it
            }
        }
    }
    if (j > 0) emit(4, (uint32_t)j);
    emit(-2, 0u);                                          // flush the last element
    out.alignment_offset = i;
    for (uint32_t k = 0; k < n / 2; ++k) {                 // first-to-last (:352-359)
        const uint32_t x = cig[k];
        cig[k] = cig[n - 1 - k];
        cig[n - 1 - k] = x;
    }
    out.n_cigar = n;
    if (overflow) out.status = 2;
    b.outs[id] = out;
}

}  // namespace

cudaError_t phmm_launch_sw_fill(const PhmmSwBatchDev& b, int cpt, uint32_t first, uint32_t count, uint32_t warps, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    const uint32_t blocks = (warps + 3) / 4;
    switch (cpt) {
        case 2: sw_fill_kernel<2><<<blocks, 128, 0, st>>>(b, first, count); break;
        case 5: sw_fill_kernel<5><<<blocks, 128, 0, st>>>(b, first, count); break;
        case 8: sw_fill_kernel<8><<<blocks, 128, 0, st>>>(b, first, count); break;
        case 16: sw_fill_kernel<16><<<blocks, 128, 0, st>>>(b, first, count); break;
        case 32: sw_fill_kernel<32><<<blocks, 128, 0, st>>>(b, first, count); break;
        case 64: sw_fill_kernel<64><<<blocks, 128, 0, st>>>(b, first, count); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t phmm_launch_sw_fill16(const PhmmSwBatchDev& b, int cpt, uint32_t first, uint32_t count, uint32_t warps, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    const uint32_t blocks = (warps + 3) / 4;
    switch (cpt) {
        case 2: sw_fill16_kernel<2><<<blocks, 128, 0, st>>>(b, first, count); break;
        case 5: sw_fill16_kernel<5><<<blocks, 128, 0, st>>>(b, first, count); break;
        case 8: sw_fill16_kernel<8><<<blocks, 128, 0, st>>>(b, first, count); break;
        case 16: sw_fill16_kernel<16><<<blocks, 128, 0, st>>>(b, first, count); break;
        case 32: sw_fill16_kernel<32><<<blocks, 128, 0, st>>>(b, first, count); break;
        default: return cudaErrorInvalidValue;                 // 64 columns per lane: reads over 1024 bases never qualify
    }
    return cudaGetLastError();
}

cudaError_t phmm_launch_sw_trace(const PhmmSwBatchDev& b, uint32_t first, uint32_t count, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    sw_trace_kernel<<<(count + 127) / 128, 128, 0, st>>>(b, first, count);
    return cudaGetLastError();
}
