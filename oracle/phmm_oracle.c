/*
 * phmm_oracle.c — scalar CPU restatement of Rovaca's PairHMM path.  TEST INFRASTRUCTURE ONLY
 * (see phmm_oracle.h).  Plain C, one pair at a time, explicit operation order; compiled with
 * -ffp-contract=off so the only fused operations are the fmaf()/fma() calls written below.
 *
 * Citations: /root/reference/src/haplotypecaller/pairhmm/<file>:<line>.
 */
#include "phmm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>

/* ---------------------------------------------------------------- tables (core/context.h) -- */

#define ORC_NQ 128                                  /* quals are masked & 127 before any lookup */
#define ORC_TRI (ORC_NQ * (ORC_NQ + 1) / 2)         /* 8256 */
#define JAC_TOL 8.0                                 /* MAX_JACOBIAN_TOLERANCE      context.h:8  */
#define JAC_STEP 0.0001                             /* JACOBIAN_LOG_TABLE_STEP     context.h:9  */
#define JAC_INV_STEP (1.0 / JAC_STEP)               /*                             context.h:10 */
#define JAC_SIZE ((int)(JAC_TOL / JAC_STEP) + 1)    /*                             context.h:11 */

static float g_ph2pr_f[ORC_NQ];
static double g_ph2pr_d[ORC_NQ];
static float g_mm_f[ORC_TRI];
static double g_mm_d[ORC_TRI];
static float g_pm_f[ORC_NQ];    /* 1.0f - ph2pr[q]   pairhmm_internal.cpp:479 */
static float g_pmm_f[ORC_NQ];   /* ph2pr[q] / 3.0f   pairhmm_internal.cpp:480 */
static volatile int g_ready = 0;

/* approximateLog10SumLog10, context.h:64-86, in float and in double arithmetic. */
static float approx_log10_sum_f(float small, float big, const float* jac)
{
    if (small > big) { float t = big; big = small; small = t; }
    if (isinf(small) == 1 || isinf(big) == 1) return big;
    float diff = big - small;
    if (diff >= (float)JAC_TOL) return big;
    float scaled = (float)(diff * ((float)JAC_INV_STEP));
    int ind = (scaled > 0.0f) ? (int)(scaled + 0.5f) : (int)(scaled - 0.5f);   /* fastRound :62 */
    return big + jac[ind];
}

static double approx_log10_sum_d(double small, double big, const double* jac)
{
    if (small > big) { double t = big; big = small; small = t; }
    if (isinf(small) == 1 || isinf(big) == 1) return big;
    double diff = big - small;
    if (diff >= JAC_TOL) return big;
    double scaled = diff * JAC_INV_STEP;
    int ind = (scaled > 0.0) ? (int)(scaled + 0.5) : (int)(scaled - 0.5);
    return big + jac[ind];
}

void orc_init(void)
{
    if (g_ready) return;
    float* jac_f = (float*)malloc(sizeof(float) * JAC_SIZE);
    double* jac_d = (double*)malloc(sizeof(double) * JAC_SIZE);
    for (int k = 0; k < JAC_SIZE; k++) {                       /* context.h:41-46 */
        double v = log10(1.0 + pow(10.0, -((double)k) * JAC_STEP));
        jac_f[k] = (float)v;
        jac_d[k] = v;
    }
    double inv_ln10 = 1.0 / log(10);                            /* context.h:51-52 */
    for (int i = 0, offset = 0; i < ORC_NQ; offset += ++i) {   /* context.h:53-58 (rows 0..127 of 0..254) */
        for (int j = 0; j <= i; j++) {
            {
                float s = approx_log10_sum_f((float)(-0.1 * i), (float)(-0.1 * j), jac_f);
                double m2m = log1p(-fmin(1.0, pow(10, (double)s))) * inv_ln10;
                g_mm_f[offset + j] = (float)pow(10, m2m);
            }
            {
                double s = approx_log10_sum_d(-0.1 * i, -0.1 * j, jac_d);
                double m2m = log1p(-fmin(1.0, pow(10, s))) * inv_ln10;
                g_mm_d[offset + j] = pow(10, m2m);
            }
        }
    }
    for (int x = 0; x < ORC_NQ; x++) {
        g_ph2pr_f[x] = powf(10.f, -((float)x) / 10.f);          /* context.h:144-146 */
        g_ph2pr_d[x] = pow(10.0, -((double)x) / 10.0);          /* context.h:102-104 */
        g_pm_f[x] = 1.0f - g_ph2pr_f[x];
        g_pmm_f[x] = g_ph2pr_f[x] / 3.0f;
    }
    free(jac_f);
    free(jac_d);
    __sync_synchronize();
    g_ready = 1;
}

const float* orc_ph2pr_f32(void) { orc_init(); return g_ph2pr_f; }
const double* orc_ph2pr_f64(void) { orc_init(); return g_ph2pr_d; }
const float* orc_mm_tri_f32(void) { orc_init(); return g_mm_f; }
const double* orc_mm_tri_f64(void) { orc_init(); return g_mm_d; }

/* set_mm_prob, context.h:120-131 / :175-186 (quals <= 127 so the table branch always applies). */
static inline int mm_index(int ins_q, int del_q)
{
    int lo = del_q, hi = ins_q;
    if (ins_q <= del_q) { lo = ins_q; hi = del_q; }
    return ((hi * (hi + 1)) >> 1) + lo;
}

/* one-hot base code, pairhmm_internal.cpp:27-41 */
static inline unsigned onehot(uint8_t c)
{
    switch (c) {
        case 'A': return 0x1;
        case 'C': return 0x2;
        case 'T': return 0x4;
        case 'G': return 0x8;
        case 'N': return 0xF;
        default: return 0;
    }
}

/* init_native, pairhmm_internal.cpp:21-25 — the reference leaves FTZ on; we restore on exit. */
static inline unsigned ftz_on(void)
{
    unsigned old = _mm_getcsr();
    _mm_setcsr(old | 0x8000u);
    return old;
}
static inline void ftz_restore(unsigned old) { _mm_setcsr(old); }

/* ------------------------------------------------------------------ forward recurrence ---- */

/* FP32, one pair.  Boundary and cell follow rovaca/avx_512_float.cc:29-181 (equivalently
 * core/avx512_s.cpp:259-397); the fused/unfused pattern is the one GCC 13 -O3 emits for the
 * reference TUs, selected by fma_mode. */
float orc_pair_f32(const uint8_t* hap, uint32_t H, const uint8_t* rs, const uint8_t* q,
                   const uint8_t* ins, const uint8_t* del, const uint8_t* gcp, uint32_t R,
                   int fma_mode)
{
    orc_init();
    unsigned csr = ftz_on();
    float* buf = (float*)malloc(sizeof(float) * 6 * (H + 1));
    float *Mp = buf, *Ip = Mp + (H + 1), *Dp = Ip + (H + 1);
    float *Mc = Dp + (H + 1), *Ic = Mc + (H + 1), *Dc = Ic + (H + 1);

    const float init_d = ldexpf(1.f, 120) / (float)H;           /* avx_512_float.cc:33 */
    for (uint32_t c = 0; c <= H; c++) { Mp[c] = 0.f; Ip[c] = 0.f; Dp[c] = init_d; }   /* :35-42 */

    for (uint32_t r = 0; r < R; r++) {
        const int qi = ins[r] & 127, qd = del[r] & 127, qc = gcp[r] & 127, qq = q[r] & 127;
        const float p_mm = g_mm_f[mm_index(qi, qd)];             /* pairhmm_internal.cpp:471 */
        const float p_gapm = 1.0f - g_ph2pr_f[qc];               /* :472 */
        const float p_mx = g_ph2pr_f[qi];                        /* :473 */
        const float p_xx = g_ph2pr_f[qc];                        /* :474 */
        const float p_my = g_ph2pr_f[qd];                        /* :475 */
        const float p_yy = g_ph2pr_f[qc];                        /* :476 */
        const float prior_match = g_pm_f[qq];                    /* :479 */
        const float prior_mismatch = g_pmm_f[qq];                /* :480 */
        const unsigned rb = onehot(rs[r]);
        Mc[0] = 0.f; Ic[0] = 0.f; Dc[0] = 0.f;                   /* avx_512_float.cc:66 */
        for (uint32_t c = 1; c <= H; c++) {
            const float prior = (rb & onehot(hap[c - 1])) ? prior_match : prior_mismatch;   /* :81-82 */
            float m, i, d;
            if (fma_mode == ORC_FMA_AVX512) {
                float t = Mp[c - 1] * p_mm;
                t = fmaf(Ip[c - 1], p_gapm, t);
                t = fmaf(Dp[c - 1], p_gapm, t);
                m = t * prior;                                   /* :84-86 */
                i = fmaf(Ip[c], p_xx, Mp[c] * p_mx);             /* :87 */
                d = fmaf(Dc[c - 1], p_yy, Mc[c - 1] * p_my);     /* :88 */
            } else {                                             /* core/avx2_s.cpp:207-216 */
                float t = Mp[c - 1] * p_mm + Ip[c - 1] * p_gapm;
                t = t + Dp[c - 1] * p_gapm;
                m = t * prior;
                i = Mp[c] * p_mx + Ip[c] * p_xx;
                d = Mc[c - 1] * p_my + Dc[c - 1] * p_yy;
            }
            Mc[c] = m; Ic[c] = i; Dc[c] = d;
        }
        float* t;
        t = Mp; Mp = Mc; Mc = t;
        t = Ip; Ip = Ic; Ic = t;
        t = Dp; Dp = Dc; Dc = t;
    }
    float sum_m = 0.f, sum_i = 0.f;                             /* avx_512_float.cc:166-171 */
    for (uint32_t c = 1; c <= H; c++) { sum_m = sum_m + Mp[c]; sum_i = sum_i + Ip[c]; }
    float res = sum_m + sum_i;                                  /* :178-180 */
    free(buf);
    ftz_restore(csr);
    return res;
}

/* FP64, one pair: the same recurrence in double with init 2^1020/H and double tables — what
 * core/avx512_d.cpp / avx2_d.cpp are meant to compute (GATK semantics).  The reference's two
 * FP64 TUs are defective as built and return 0.0 for multi-stripe reads, so
 * for this branch the scalar restatement *is* the oracle; the as-built value is reported
 * alongside by oracle/_ref in the tests. */
double orc_pair_f64(const uint8_t* hap, uint32_t H, const uint8_t* rs, const uint8_t* q,
                    const uint8_t* ins, const uint8_t* del, const uint8_t* gcp, uint32_t R)
{
    orc_init();
    unsigned csr = ftz_on();
    double* buf = (double*)malloc(sizeof(double) * 6 * (H + 1));
    double *Mp = buf, *Ip = Mp + (H + 1), *Dp = Ip + (H + 1);
    double *Mc = Dp + (H + 1), *Ic = Mc + (H + 1), *Dc = Ic + (H + 1);
    const double init_d = ldexp(1.0, 1020) / (double)H;         /* core/context.h:106 */
    for (uint32_t c = 0; c <= H; c++) { Mp[c] = 0.0; Ip[c] = 0.0; Dp[c] = init_d; }
    for (uint32_t r = 0; r < R; r++) {
        const int qi = ins[r] & 127, qd = del[r] & 127, qc = gcp[r] & 127, qq = q[r] & 127;
        const double p_mm = g_mm_d[mm_index(qi, qd)];
        const double p_gapm = 1.0 - g_ph2pr_d[qc];
        const double p_mx = g_ph2pr_d[qi], p_xx = g_ph2pr_d[qc];
        const double p_my = g_ph2pr_d[qd], p_yy = g_ph2pr_d[qc];
        const double eps = g_ph2pr_d[qq];
        const double prior_match = 1.0 - eps, prior_mismatch = eps / 3.0;
        const unsigned rb = onehot(rs[r]);
        Mc[0] = 0.0; Ic[0] = 0.0; Dc[0] = 0.0;
        for (uint32_t c = 1; c <= H; c++) {
            const double prior = (rb & onehot(hap[c - 1])) ? prior_match : prior_mismatch;
            double t = Mp[c - 1] * p_mm;
            t = fma(Ip[c - 1], p_gapm, t);
            t = fma(Dp[c - 1], p_gapm, t);
            Mc[c] = t * prior;
            Ic[c] = fma(Ip[c], p_xx, Mp[c] * p_mx);
            Dc[c] = fma(Dc[c - 1], p_yy, Mc[c - 1] * p_my);
        }
        double* t;
        t = Mp; Mp = Mc; Mc = t;
        t = Ip; Ip = Ic; Ic = t;
        t = Dp; Dp = Dc; Dc = t;
    }
    double sum_m = 0.0, sum_i = 0.0;
    for (uint32_t c = 1; c <= H; c++) { sum_m = sum_m + Mp[c]; sum_i = sum_i + Ip[c]; }
    double res = sum_m + sum_i;
    free(buf);
    ftz_restore(csr);
    return res;
}

/* pairhmm_internal.cpp:578-584; constants adapter.h:25-26, rovaca/common.h:9.  The float branch
 * subtracts in float and only then widens. */
double orc_log10_likelihood(float raw_f32, double raw_f64)
{
    const float k_log10_init_f = 36.1236000061f;
    const double k_log10_init_d = 307.050595577260822;
    if (raw_f32 < 1e-28f) return log10(raw_f64) - k_log10_init_d;
    return (double)(log10f(raw_f32) - k_log10_init_f);
}

/* ------------------------------------------------------------------ read preparation ------ */

/* modify_read_qualities, pairhmm_internal.cpp:434-439 (MIN_QUALITY = 6, adapter.h:5) */
void orc_cap_quals(uint8_t* qual, uint32_t n, uint8_t mapq, int32_t thr)
{
    for (uint32_t x = 0; x < n; x++) {
        uint8_t v = qual[x] < mapq ? qual[x] : mapq;
        qual[x] = ((int32_t)v < thr) ? 6 : v;
    }
}

/* find_number_of_repetitions, pairhmm_internal.cpp:195-229 (unit and test string share `bases`) */
static int32_t count_repeats(const uint8_t* bases, int32_t unit_off, int32_t unit_len, int32_t test_off,
                             int32_t test_len, int leading)
{
    if (test_len == 0) return 0;
    int32_t n = 0;
    const int32_t diff = test_len - unit_len;
    if (leading) {
        for (int32_t s = 0; s <= diff; s += unit_len) {
            if (memcmp(bases + test_off + s, bases + unit_off, (size_t)unit_len) != 0) return n;
            ++n;
        }
    } else {
        for (int32_t s = diff; s >= 0; s -= unit_len) {
            if (memcmp(bases + test_off + s, bases + unit_off, (size_t)unit_len) != 0) return n;
            ++n;
        }
    }
    return n;
}

/* find_tandem_repeat_units, pairhmm_internal.cpp:65-119.  Kept quirk for quirk: the backward
 * unit is recorded at (offset - str - 1), so it can never equal the forward unit at (offset + 1)
 * and whenever a forward scan runs the backward count is redone with the forward unit (:107-110). */
int32_t orc_find_tandem_repeat_units(const uint8_t* bases, uint32_t n_u, uint32_t offset_u)
{
    const int32_t offset = (int32_t)offset_u, n = (int32_t)n_u;
    int32_t max_bw = 0, bw_first = offset, bw_len = 1;
    for (int32_t str = 1; str <= 8; str++) {                    /* MAX_STR_UNIT_LENGTH adapter.h:7 */
        if (offset + 1 - str < 0) break;
        max_bw = count_repeats(bases, offset - str + 1, str, 0, offset + 1, 0);
        if (max_bw > 1) { bw_first = offset - str - 1; bw_len = str; break; }
    }
    int32_t max_rl = max_bw;
    if (offset < n - 1) {
        int32_t fw_first = offset + 1, fw_len = 1, max_fw = 0;
        for (int32_t str = 1; str <= 8; str++) {
            if (offset + str + 1 > n) break;
            max_fw = count_repeats(bases, offset + 1, str, offset + 1, n - offset - 1, 1);
            if (max_fw > 1) { fw_first = offset + 1; fw_len = str; break; }
        }
        if (fw_first == bw_first && fw_len == bw_len) {
            max_rl = max_bw + max_fw;
        } else {
            max_bw = count_repeats(bases, fw_first, fw_len, 0, offset + 1, 0);
            max_rl = max_bw + max_fw;
        }
    }
    if (max_rl > 20) max_rl = 20;                               /* MAX_REPEAT_LENGTH adapter.h:6 */
    return max_rl;
}

/* adapter.h:13-15 */
static const uint8_t k_pcr_cache[4][21] = {
    {0},
    {40, 40, 39, 38, 37, 36, 34, 32, 28, 23, 17, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10},
    {40, 40, 40, 39, 39, 39, 38, 38, 37, 37, 36, 35, 34, 33, 32, 30, 28, 26, 23, 20, 17},
    {40, 40, 40, 40, 39, 39, 39, 39, 39, 38, 38, 38, 37, 37, 37, 36, 36, 35, 34, 33, 33}};

/* apply_pcr_error_model, pairhmm_internal.cpp:121-159: positions 0..n-2 only. */
void orc_apply_pcr_error_model(const uint8_t* bases, uint8_t* gap_qual, uint32_t n, int model)
{
    if (model < ORC_PCR_HOSTILE || model > ORC_PCR_CONSERVATIVE) return;
    for (uint32_t i = 1; i < n; i++) {
        int32_t rl = orc_find_tandem_repeat_units(bases, n, i - 1);
        uint8_t cap = k_pcr_cache[model][rl];
        if (!(gap_qual[i - 1] < cap)) gap_qual[i - 1] = cap;
    }
}

/* ------------------------------------------------------------------ region driver --------- */

int64_t orc_region(const orc_read* reads, uint32_t n_reads, const orc_hap* haps, uint32_t n_haps,
                   int32_t thr, int pcr_model, int fma_mode, double* out, uint32_t* kept_idx,
                   uint32_t* n_kept, float* raw_f32, double* raw_f64)
{
    orc_init();
    *n_kept = 0;
    if (n_reads == 0 || n_haps == 0) return 0;                  /* pairhmm_internal.cpp:508-510 */
    unsigned csr = ftz_on();
    int64_t n_fp64 = 0;
    uint32_t max_len = 0;
    for (uint32_t i = 0; i < n_reads; i++) if (reads[i].len > max_len) max_len = reads[i].len;
    uint8_t* q = (uint8_t*)malloc(4 * (size_t)(max_len + 1));
    uint8_t *ins = q + (max_len + 1), *del_buf = ins + (max_len + 1), *gcp = del_buf + (max_len + 1);
    double* lk = (double*)malloc(sizeof(double) * (size_t)n_reads * n_haps);   /* [read][hap] :529 */

    for (uint32_t i = 0; i < n_reads; i++) {
        const orc_read* rd = &reads[i];
        const uint32_t n = rd->len;
        memcpy(q, rd->qual, n);
        orc_cap_quals(q, n, rd->mapq, thr);                     /* :545-550 */
        if (rd->ins) memcpy(ins, rd->ins, n); else memset(ins, '-', n);   /* :552-554 */
        if (pcr_model != ORC_PCR_NONE) orc_apply_pcr_error_model(rd->bases, ins, n, pcr_model);   /* :555-557 */
        const uint8_t* del = ins;                               /* :559 */
        if (rd->del) { memcpy(del_buf, rd->del, n); del = del_buf; }
        if (rd->gcp) memcpy(gcp, rd->gcp, n); else memset(gcp, '+' - 33, n);   /* :561-563 */
        for (uint32_t j = 0; j < n_haps; j++) {
            float f = orc_pair_f32(haps[j].bases, haps[j].len, rd->bases, q, ins, del, gcp, n, fma_mode);
            double d = NAN;
            if (f < 1e-28f) {                                   /* :578 */
                d = orc_pair_f64(haps[j].bases, haps[j].len, rd->bases, q, ins, del, gcp, n);
                ++n_fp64;
            }
            lk[(size_t)i * n_haps + j] = orc_log10_likelihood(f, d);
            if (raw_f32) raw_f32[(size_t)i * n_haps + j] = f;
            if (raw_f64) raw_f64[(size_t)i * n_haps + j] = d;
        }
    }

    /* normalize_likelihoods, :53-63 (MAXIMUM_BEST_ALT_LIKELIHOOD_DIFFERENCE = -4.5, adapter.h:23) */
    for (uint32_t i = 0; i < n_reads; i++) {
        double* row = lk + (size_t)i * n_haps;
        double best = row[0];
        for (uint32_t j = 1; j < n_haps; j++) if (best < row[j]) best = row[j];   /* std::max_element */
        const double cmp = best + (-4.5);
        for (uint32_t j = 0; j < n_haps; j++) row[j] = row[j] < cmp ? cmp : row[j];
    }
    /* filter_poorly_modelled_evidence, :161-182 (adapter.h:20-22) */
    uint32_t kept = 0;
    for (uint32_t i = 0; i < n_reads; i++) {
        const double* row = lk + (size_t)i * n_haps;
        double best = row[0];
        for (uint32_t j = 1; j < n_haps; j++) if (best < row[j]) best = row[j];
        double lim = fmin(2.0, ceil((int32_t)reads[i].len * 0.02)) * (-4.0);
        if (best < lim) continue;
        kept_idx[kept++] = i;
    }
    /* transpose into the NaN-initialised [hap][orig_read_count] matrix, :184-193, :516 */
    for (size_t k = 0; k < (size_t)n_haps * n_reads; k++) out[k] = NAN;
    for (uint32_t c = 0; c < kept; c++)
        for (uint32_t j = 0; j < n_haps; j++) out[(size_t)j * n_reads + c] = lk[(size_t)kept_idx[c] * n_haps + j];
    *n_kept = kept;
    free(lk);
    free(q);
    ftz_restore(csr);
    return n_fp64;
}

/* Restatement of the best-allele search (allele_likelihoods.hpp:589-657) for can_be_reference = true.
 *   pass 1 (:609-622): scan the alleles in order keeping a leader and a runner-up; a candidate takes the lead only when
 *                      strictly greater, the runner-up slot likewise (so among equals the earliest allele leads);
 *   pass 2 (:624-645): only when priorities exist and the lead over the runner-up is under 0.2: every allele within 0.2 of
 *                      the pass-1 leader's likelihood competes again on priority alone, strictly-greater again; the
 *                      likelihood bar stays the pass-1 leader's even after the lead changes hands;
 *   report (:647-656): likelihoods are re-read for the final pair, the runner-up's is -inf when it is the leader itself,
 *                      confidence is their difference, or exactly 0 when that difference is below DBL_EPSILON. */
typedef struct { int32_t lead, runner; } orc_rank;

static orc_rank orc_rank_by_likelihood(const double* col, size_t step, uint32_t n, double* lead_lik, double* runner_lik)
{
    orc_rank r = {0, 0};
    double top = col[0], next = -INFINITY;
    for (uint32_t a = 1; a < n; a++) {
        const double v = col[(size_t)a * step];
        if (v > top) { r.runner = r.lead; next = top; r.lead = (int32_t)a; top = v; }
        else if (v > next) { r.runner = (int32_t)a; next = v; }
    }
    *lead_lik = top;
    *runner_lik = next;
    return r;
}

void orc_best_alleles(const double* values, uint32_t n_alleles, uint32_t stride, uint32_t n_evidence, const double* priorities,
                      uint32_t* best, uint32_t* second, double* best_lik, double* second_lik, double* confidence)
{
    const double window = 0.2;                               /* s_log_10_informative_threshold, :27 */
    for (uint32_t e = 0; e < n_evidence; e++) {
        const double* col = values + e;
        double bar, below;
        orc_rank r = orc_rank_by_likelihood(col, stride, n_alleles, &bar, &below);
        if (priorities != NULL && bar - below < window) {
            double lead_pri = priorities[r.lead], runner_pri = priorities[r.runner];
            uint32_t a = 0;
            while (a < n_alleles) {
                const int contender = (int32_t)a != r.lead && !(bar - col[(size_t)a * stride] > window);
                if (contender) {
                    const double p = priorities[a];
                    if (p > lead_pri) { r.runner = r.lead; runner_pri = lead_pri; r.lead = (int32_t)a; lead_pri = p; }
                    else if (p > runner_pri) { r.runner = (int32_t)a; runner_pri = p; }
                }
                a++;
            }
        }
        const double lik1 = col[(size_t)r.lead * stride];
        const double lik2 = r.runner == r.lead ? -INFINITY : col[(size_t)r.runner * stride];
        best[e] = (uint32_t)r.lead;
        second[e] = (uint32_t)r.runner;
        if (best_lik) best_lik[e] = lik1;
        if (second_lik) second_lik[e] = lik2;
        if (confidence) confidence[e] = fabs(lik1 - lik2) < 2.220446049250313e-16 ? 0.0 : lik1 - lik2;   /* BestAllele, :52-62 */
    }
}

/* Restatement of marginal_likelihoods_direct (allele_likelihoods.hpp:702-732): every new allele's row starts at -inf (:714) and
 * is raised, evidence by evidence, to any strictly greater value found in the rows of the old alleles mapped onto it
 * (:723-726); old alleles mapped to nothing (s_missing_index) contribute nowhere (:717). */
void orc_marginalize(const double* old_values, uint32_t old_allele_count, uint32_t stride, uint32_t evidence_count,
                     const int32_t* old2new_idx_map, uint32_t new_allele_count, double* new_values)
{
    for (uint32_t n = 0; n < new_allele_count; n++) {
        double* dst = new_values + (size_t)n * evidence_count;
        for (uint32_t e = 0; e < evidence_count; e++) {
            double m = -INFINITY;
            for (uint32_t a = 0; a < old_allele_count; a++)
                if (old2new_idx_map[a] == (int32_t)n && old_values[(size_t)a * stride + e] > m) m = old_values[(size_t)a * stride + e];
            dst[e] = m;
        }
    }
}
