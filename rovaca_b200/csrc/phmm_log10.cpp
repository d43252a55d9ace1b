// phmm_log10.cpp — the reference's conversion of a raw FP32 sum to a log10 likelihood (pairhmm_internal.cpp:583):
//     (double)(log10f(f) - LOG10_INITIAL_CONSTANT_F)
// for a whole region at a time.  log10f is the host's biggest cost per pair, and its bits are whatever the host's libm
// says — so libm stays the authority and everything faster is a restatement that is TESTED against it on this host when an
// engine is created, never assumed:
//   mode 0  call log10f.
//   mode 1  glibc (to 2.39 at least) computes log10f the fdlibm way: split off the exponent k, then
//           y*log10_2lo + ivln10*logf(m) + y*log10_2hi in float arithmetic with m in [0.5, 2).  The same formula inlined
//           around the host's own logf: same bits, one call less.
//   mode 2  that formula with logf itself restated (glibc's table-driven logf: 16 intervals, a degree-3 polynomial in
//           double, rounded to float once), eight pairs per AVX-512 instruction.  The formula only ever hands logf an m in
//           [0.5, 2) — 2^24 inputs, so mode 2 is checked against the host's logf EXHAUSTIVELY (tests/test_abi.py; a
//           strided sweep at creation), plus a million inputs through the whole formula.
// Any mismatch drops to the next lower mode.
#include <immintrin.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "phmm_log10.h"

namespace
{

constexpr float kMinAccepted = 1e-28f;                     // MIN_ACCEPTED, rovaca/common.h:9
constexpr float kLog10InitF = 36.1236000061f;              // LOG10_INITIAL_CONSTANT_F, adapter.h:25
constexpr float kIvLn10 = 4.3429449201e-01f, kLog10_2hi = 3.0102920532e-01f, kLog10_2lo = 7.9034151668e-07f;

// glibc's logf data (sysdeps/ieee754/flt-32/e_logf_data.c, LOGF_TABLE_BITS = 4): 1/c and log(c) for the centre c of each of
// the 16 sub-intervals of [OFF, 2*OFF), OFF = 0x3f330000; then ln2 and the polynomial.
alignas(64) const double kInvC[16] = {0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010b0p+0, 0x1.3c995b0b80385p+0,
                                      0x1.30d190c8864a5p+0, 0x1.25e227b0b8ea0p+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
                                      0x1.0953f419900a7p+0, 0x1p+0,               0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aa0p-1,
                                      0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
alignas(64) const double kLogC[16] = {-0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,
                                      -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c8100p-3, -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,
                                      -0x1.252f438e10c1ep-5, 0x0p+0,                0x1.aa5aa5df25984p-5,  0x1.c5e53aa362eb4p-4,
                                      0x1.526e57720db08p-3,  0x1.bc2860d224770p-3,  0x1.1058bc8a07ee1p-2,  0x1.4043057b6ee09p-2};
constexpr double kLn2 = 0x1.62e42fefa39efp-1, kA0 = -0x1.00ea348b88334p-2, kA1 = 0x1.5575b0be00b6ap-2, kA2 = -0x1.ffffef20a4123p-2;
constexpr uint32_t kOff = 0x3f330000u;

inline float log10f_around_logf(float x)
{
    uint32_t ux;
    std::memcpy(&ux, &x, 4);
    int32_t hx = (int32_t)ux, k = 0;
    if (hx < 0x00800000) return log10f(x);                   // zero, negative or subnormal: libm's business
    if (hx >= 0x7f800000) return x + x;
    k += (hx >> 23) - 127;
    const int32_t i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
    hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
    const float y = (float)(k + i);
    ux = (uint32_t)hx;
    std::memcpy(&x, &ux, 4);
    const float z = y * kLog10_2lo + kIvLn10 * logf(x);
    return z + y * kLog10_2hi;
}

// The restatement one value at a time, written with the very operations the device epilogue uses (phmm_kernels.cu,
// log10_of_sum): what the device computes is what this computes, and this is what gets compared with libm.
inline float log10f_restated(float x)
{
    uint32_t ux;
    std::memcpy(&ux, &x, 4);
    int32_t hx = (int32_t)ux;
    const int32_t k = (hx >> 23) - 127;
    const int32_t neg = (int32_t)((uint32_t)k >> 31);
    hx = (hx & 0x007fffff) | ((0x7f - neg) << 23);
    const float y = (float)(k + neg);
    const uint32_t tmp = (uint32_t)hx - kOff;
    const int idx = (int)((tmp >> 19) & 15u);
    const int32_t kk = (int32_t)tmp >> 23;
    const uint32_t iz = (uint32_t)hx - (tmp & 0xff800000u);
    float zf;
    std::memcpy(&zf, &iz, 4);
    const double z = zf;
    const double r = std::fma(z, kInvC[idx], -1.0);
    const double y0 = std::fma((double)kk, kLn2, kLogC[idx]);
    const double r2 = r * r;
    double p = std::fma(kA1, r, kA2);
    p = std::fma(kA0, r2, p);
    p = std::fma(p, r2, y0 + r);
    const float lf = (float)p;
    const float zz = y * kLog10_2lo + kIvLn10 * lf;
    return zz + y * kLog10_2hi;
}

// Eight sums per step.  Lanes below MIN_ACCEPTED (and the lanes past n) come out as garbage and are counted; the caller
// overwrites them.  `logf_only`: v = (double)logf(f) for f in [0.5, 2), the self-test's view of the middle of the formula.
__attribute__((target("avx512f,avx512dq,avx512vl,fma"))) size_t block_avx512(const float* f, double* v, size_t n, bool logf_only)
{
    const __m512d invc_lo = _mm512_load_pd(kInvC), invc_hi = _mm512_load_pd(kInvC + 8);
    const __m512d logc_lo = _mm512_load_pd(kLogC), logc_hi = _mm512_load_pd(kLogC + 8);
    const __m512d ln2 = _mm512_set1_pd(kLn2), a0 = _mm512_set1_pd(kA0), a1 = _mm512_set1_pd(kA1), a2 = _mm512_set1_pd(kA2);
    const __m512d minus1 = _mm512_set1_pd(-1.0);
    const __m256 lo = _mm256_set1_ps(kLog10_2lo), hi = _mm256_set1_ps(kLog10_2hi), iv = _mm256_set1_ps(kIvLn10);
    const __m256 init = _mm256_set1_ps(kLog10InitF), tiny = _mm256_set1_ps(kMinAccepted);
    size_t under = 0;
    for (size_t i = 0; i < n; i += 8) {
        const __mmask8 m = n - i >= 8 ? (__mmask8)0xff : (__mmask8)((1u << (n - i)) - 1u);
        const __m256 x = _mm256_maskz_loadu_ps(m, f + i);
        under += (size_t)__builtin_popcount((unsigned)_mm256_mask_cmp_ps_mask(m, x, tiny, _CMP_LT_OQ));
        __m256i hx = _mm256_castps_si256(x);
        __m256 y = _mm256_setzero_ps();
        if (!logf_only) {                                    // the fdlibm split: x = 2^(k+i) * m, m in [0.5, 2)
            const __m256i k = _mm256_sub_epi32(_mm256_srai_epi32(hx, 23), _mm256_set1_epi32(127));
            const __m256i neg = _mm256_srli_epi32(k, 31);
            hx = _mm256_or_si256(_mm256_and_si256(hx, _mm256_set1_epi32(0x007fffff)),
                                 _mm256_slli_epi32(_mm256_sub_epi32(_mm256_set1_epi32(0x7f), neg), 23));
            y = _mm256_cvtepi32_ps(_mm256_add_epi32(k, neg));
        }
        // logf(m): m = 2^kk * z with z in [OFF, 2*OFF); log(m) = log1p(z/c - 1) + log(c) + kk*ln2
        const __m256i tmp = _mm256_sub_epi32(hx, _mm256_set1_epi32((int)kOff));
        const __m256i idx = _mm256_and_si256(_mm256_srli_epi32(tmp, 19), _mm256_set1_epi32(15));
        const __m256i kk = _mm256_srai_epi32(tmp, 23);
        const __m256i iz = _mm256_sub_epi32(hx, _mm256_and_si256(tmp, _mm256_set1_epi32((int)0xff800000u)));
        const __m512d z = _mm512_cvtps_pd(_mm256_castsi256_ps(iz));
        const __m512i idx64 = _mm512_cvtepi32_epi64(idx);
        const __m512d invc = _mm512_permutex2var_pd(invc_lo, idx64, invc_hi);
        const __m512d logc = _mm512_permutex2var_pd(logc_lo, idx64, logc_hi);
        const __m512d r = _mm512_fmadd_pd(z, invc, minus1);
        const __m512d y0 = _mm512_fmadd_pd(_mm512_cvtepi32_pd(kk), ln2, logc);
        const __m512d r2 = _mm512_mul_pd(r, r);
        __m512d p = _mm512_fmadd_pd(a1, r, a2);
        p = _mm512_fmadd_pd(a0, r2, p);
        p = _mm512_fmadd_pd(p, r2, _mm512_add_pd(y0, r));
        const __m256 lf = _mm512_cvtpd_ps(p);
        if (logf_only) {
            _mm512_mask_storeu_pd(v + i, m, _mm512_cvtps_pd(lf));
            continue;
        }
        const __m256 zz = _mm256_add_ps(_mm256_mul_ps(y, lo), _mm256_mul_ps(iv, lf));      // float arithmetic, nothing contracted
        const __m256 res = _mm256_add_ps(zz, _mm256_mul_ps(y, hi));
        _mm512_mask_storeu_pd(v + i, m, _mm512_cvtps_pd(_mm256_sub_ps(res, init)));          // float subtract, then widen (:583)
    }
    return under;
}

size_t block_scalar(const float* f, double* v, size_t n, int mode)
{
    size_t under = 0;
    for (size_t i = 0; i < n; ++i) {
        const float x = f[i];
        under += x < kMinAccepted;
        const float l = mode == 3 ? log10f_restated(x) : mode == 1 ? log10f_around_logf(x) : log10f(x);
        v[i] = (double)(l - kLog10InitF);
    }
    return under;
}

// normalize (:53-63) + filter (:161-182) + transpose (:184-193) of one region, eight reads per step: lane r of a vector is
// read i0 + r, the haplotypes are walked one gather (stride nh) at a time.  A block with a NaN in it goes the scalar way.
__attribute__((target("avx512f,avx512dq,avx512vl,fma"))) uint32_t postprocess_avx512(const double* v, uint32_t nr, uint32_t nh,
                                                                                      const uint32_t* read_len, double* out,
                                                                                      uint32_t* kept_idx)
{
    const __m512i lane_off = _mm512_mullo_epi64(_mm512_set_epi64(7, 6, 5, 4, 3, 2, 1, 0), _mm512_set1_epi64((long long)nh));
    const __m256i iota = _mm256_set_epi32(7, 6, 5, 4, 3, 2, 1, 0);
    uint32_t kept = 0;
    for (uint32_t i0 = 0; i0 < nr; i0 += 8) {
        const uint32_t nb = nr - i0 < 8 ? nr - i0 : 8;
        const __mmask8 m = (__mmask8)((1u << nb) - 1u);
        const double* base = v + (size_t)i0 * nh;
        __m512d best = _mm512_mask_i64gather_pd(_mm512_setzero_pd(), m, lane_off, base, 8);
        __mmask8 nan = _mm512_mask_cmp_pd_mask(m, best, best, _CMP_UNORD_Q);
        for (uint32_t j = 1; j < nh; ++j) {
            const __m512d t = _mm512_mask_i64gather_pd(_mm512_setzero_pd(), m, lane_off, base + j, 8);
            nan |= _mm512_mask_cmp_pd_mask(m, t, t, _CMP_UNORD_Q);
            best = _mm512_max_pd(best, t);
        }
        alignas(64) double lim[8];
        for (uint32_t r = 0; r < 8; ++r)
            lim[r] = r < nb ? std::min(2.0, std::ceil((int32_t)read_len[i0 + r] * 0.02)) * (-4.0) : 0.0;      // adapter.h:20-22
        if (nan) {                                           // cannot come out of the kernels; keep the scalar semantics anyway
            for (uint32_t r = 0; r < nb; ++r) {
                const double* row = base + (size_t)r * nh;
                double b = row[0];
                if (b == b)
                    for (uint32_t j = 1; j < nh; ++j)
                        if (b < row[j]) b = row[j];
                const double cmp = b + (-4.5);
                if (b < lim[r]) continue;
                for (uint32_t j = 0; j < nh; ++j) out[(size_t)j * nr + kept] = row[j] < cmp ? cmp : row[j];
                kept_idx[kept++] = i0 + r;
            }
            continue;
        }
        const __mmask8 keep = _mm512_mask_cmp_pd_mask(m, best, _mm512_load_pd(lim), _CMP_NLT_UQ);   // dropped: best < lim
        const uint32_t nk = (uint32_t)__builtin_popcount((unsigned)keep);
        if (nk == 0) continue;
        const __mmask8 km = (__mmask8)((1u << nk) - 1u);
        const __m512d cmp = _mm512_add_pd(best, _mm512_set1_pd(-4.5));
        for (uint32_t j = 0; j < nh; ++j) {
            const __m512d t = _mm512_mask_i64gather_pd(_mm512_setzero_pd(), m, lane_off, base + j, 8);
            const __m512d c = _mm512_max_pd(t, cmp);                                                // row[j] < cmp ? cmp : row[j]
            _mm512_mask_storeu_pd(out + (size_t)j * nr + kept, km, _mm512_maskz_compress_pd(keep, c));
        }
        _mm256_mask_storeu_epi32(kept_idx + kept, km, _mm256_maskz_compress_epi32(keep, _mm256_add_epi32(iota, _mm256_set1_epi32((int)i0))));
        kept += nk;
    }
    return kept;
}

uint32_t postprocess_scalar(const double* v, uint32_t nr, uint32_t nh, const uint32_t* read_len, double* out, uint32_t* kept_idx)
{
    uint32_t kept = 0;
    for (uint32_t i = 0; i < nr; ++i) {
        const double* row = v + (size_t)i * nh;
        // std::max_element semantics of :56,:170 — the first element, replaced only by a strictly greater one
        double best = row[0];
        if (best == best)
            for (uint32_t j = 1; j < nh; ++j)
                if (best < row[j]) best = row[j];
        // normalize_likelihoods :53-63 caps at best - 4.5 (adapter.h:23); the maximum itself is unchanged, so
        // filter_poorly_modelled_evidence :161-182 sees the same `best`
        const double cmp = best + (-4.5);
        const double lim = std::min(2.0, std::ceil((int32_t)read_len[i] * 0.02)) * (-4.0);          // adapter.h:20-22
        if (best < lim) continue;
        for (uint32_t j = 0; j < nh; ++j) out[(size_t)j * nr + kept] = row[j] < cmp ? cmp : row[j];   // transpose :184-193
        kept_idx[kept++] = i;
    }
    return kept;
}

bool cpu_has_avx512()
{
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
           __builtin_cpu_supports("fma");
}

// a million inputs from ~1e-30 up to 2^127 in prime-ish strides: does `mode` give libm's bits?
bool mode_matches_libm(int mode)
{
    constexpr int kN = 1 << 20;
    static thread_local float x[4096];
    static thread_local double got[4096];
    uint32_t bits = 0x0d800000u;
    for (int base = 0; base < kN; base += 4096) {
        for (int n = 0; n < 4096; ++n, bits += 1913u + (uint32_t)(n & 7)) {
            if (bits >= 0x7f000000u) bits = 0x0d800000u + (uint32_t)(base + n);
            std::memcpy(&x[n], &bits, 4);
        }
        phmm_log10_block(x, got, 4096, mode);
        for (int n = 0; n < 4096; ++n) {
            const double want = (double)(log10f(x[n]) - kLog10InitF);
            if (std::memcmp(&want, &got[n], 8) != 0) return false;
        }
    }
    return true;
}

}  // namespace

size_t phmm_log10_block(const float* f, double* v, size_t n, int mode)
{
    if (mode == 2) return block_avx512(f, v, n, false);
    return block_scalar(f, v, n, mode);
}

uint64_t phmm_log10_logf_mismatches(uint32_t stride)
{
    if (!cpu_has_avx512()) return UINT64_MAX;
    if (stride == 0) stride = 1;
    uint64_t bad = 0;
    float x[4096];
    double got[4096];
    uint32_t bits = 0x3f000000u;                            // [0.5, 2): everything the formula hands to logf
    while (bits < 0x40000000u) {
        int n = 0;
        for (; n < 4096 && bits < 0x40000000u; ++n, bits += stride) std::memcpy(&x[n], &bits, 4);
        block_avx512(x, got, (size_t)n, true);
        for (int k = 0; k < n; ++k) {
            const double want = (double)logf(x[k]);
            bad += std::memcmp(&want, &got[k], 8) != 0;
        }
    }
    return bad;
}

bool phmm_log10_restated_matches_libm()
{
    // the whole domain the formula hands to logf, every 61st float (tests/test_abi.py walks all of it), then the formula itself
    for (uint32_t bits = 0x3f000000u; bits < 0x40000000u; bits += 61u) {
        float x;
        std::memcpy(&x, &bits, 4);
        const float a = log10f(x), b = log10f_restated(x);
        if (std::memcmp(&a, &b, 4) != 0) return false;
    }
    return mode_matches_libm(3);
}

uint32_t phmm_best_haplotypes(const double* lik, uint32_t nr, uint32_t nh, uint32_t kept, const double* priority, uint32_t* best,
                              uint32_t* second)
{
    constexpr double kInformative = 0.2;                     // s_log_10_informative_threshold, allele_likelihoods.hpp:27
    for (uint32_t pos = 0; pos < kept; ++pos) {
        // search_best_allele with can_be_reference = true (allele_likelihoods.hpp:609-622)
        uint32_t bi = 0, si = 0;
        double bl = lik[pos], sl = -INFINITY;
        for (uint32_t a = 1; a < nh; ++a) {
            const double c = lik[(size_t)a * nr + pos];
            if (c > bl) { si = bi; bi = a; sl = bl; bl = c; }
            else if (c > sl) { si = a; sl = c; }
        }
        if (priority && bl - sl < kInformative) {            // :624-645
            double bp = priority[bi], sp = priority[si];
            for (uint32_t a = 0; a < nh; ++a) {
                const double c = lik[(size_t)a * nr + pos];
                if (a == bi || bl - c > kInformative) continue;
                const double cp = priority[a];
                if (cp > bp) { si = bi; bi = a; sp = bp; bp = cp; }
                else if (cp > sp) { si = a; sp = cp; }
            }
        }
        best[pos] = bi;
        second[pos] = si;
    }
    return kept;
}

int phmm_log10_pick_mode()
{
    if (std::getenv("PHMM_B200_LIBM_LOG10F")) return 0;      // deployment override: libm's log10f, nothing restated
    if (cpu_has_avx512() && phmm_log10_logf_mismatches(61) == 0 && mode_matches_libm(2)) return 2;
    if (mode_matches_libm(1)) return 1;
    return 0;
}

uint32_t phmm_postprocess_region(const double* v, uint32_t nr, uint32_t nh, const uint32_t* read_len, double* out, uint32_t* kept_idx,
                                 int mode)
{
    const uint32_t kept = mode == 2 ? postprocess_avx512(v, nr, nh, read_len, out, kept_idx) : postprocess_scalar(v, nr, nh, read_len, out, kept_idx);
    if (kept < nr)                                                                   // the NaN tail of the filtered reads
        for (uint32_t j = 0; j < nh; ++j)
            for (uint32_t c = kept; c < nr; ++c) out[(size_t)j * nr + c] = (double)NAN;
    return kept;
}
