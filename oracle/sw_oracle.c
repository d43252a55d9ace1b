/* sw_oracle.c — see sw_oracle.h.  TEST INFRASTRUCTURE ONLY. */
#include "sw_oracle.h"

#include <stdlib.h>
#include <string.h>

#define SW_MAX_SEQ_LEN 2048            /* MAX_SEQ_LEN, sw_avx.c:17 */
#define SW_MIN_CUTOFF (-100000000)     /* MATRIX_MIN_CUTOFF, :18 */
#define SW_LOW_INIT (INT32_MIN / 2)    /* LOW_INIT_VALUE, :19 */
enum { BT_MATCH = 0, BT_INS = 1, BT_DEL = 2, BT_INS_EXT = 4, BT_DEL_EXT = 8, ST_SOFT_CLIP = 9 };   /* smithwaterman_common.h:40-50 */

static int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

int32_t orc_sw_align(const uint8_t* seq1, int32_t len1, const uint8_t* seq2, int32_t len2, const orc_sw_params* p, uint32_t* cigar,
                     int32_t cigar_cap, int32_t* alignment_offset, int32_t* score_out)
{
    const int32_t nrow = len1, ncol = len2;
    if (imax(nrow, ncol) >= SW_MAX_SEQ_LEN) {       /* sw_avx512_run, :494-499 */
        *alignment_offset = 1;
        return -1;
    }
    /* H(i,0) = H(0,j) = 0, E(i,0) = F(0,j) = LOW (:107-114 and the per-diagonal border resets :190-193) */
    int32_t* Hprev = (int32_t*)calloc((size_t)ncol + 1, sizeof(int32_t));
    int32_t* Hcur = (int32_t*)calloc((size_t)ncol + 1, sizeof(int32_t));
    int32_t* F = (int32_t*)malloc(((size_t)ncol + 1) * sizeof(int32_t));
    int32_t* last_col = (int32_t*)calloc((size_t)nrow + 1, sizeof(int32_t));
    uint8_t* bt = (uint8_t*)calloc(((size_t)nrow + 1) * ((size_t)ncol + 1), 1);
    for (int32_t j = 0; j <= ncol; ++j) F[j] = SW_LOW_INIT;
    for (int32_t i = 1; i <= nrow; ++i) {
        int32_t E = SW_LOW_INIT;
        Hcur[0] = 0;
        for (int32_t j = 1; j <= ncol; ++j) {       /* MAIN_CODE, :40-76 */
            const int32_t ext_h = E + p->gap_extend, open_h = Hcur[j - 1] + p->gap_open;
            const int32_t e11 = imax(open_h, ext_h);
            int32_t ext = (open_h > ext_h) ? 0 : BT_INS_EXT;
            E = e11;
            const int32_t ext_v = F[j] + p->gap_extend, open_v = Hprev[j] + p->gap_open;
            const int32_t f11 = imax(ext_v, open_v);
            ext |= (open_v > ext_v) ? 0 : BT_DEL_EXT;
            F[j] = f11;
            const int32_t m11 = Hprev[j - 1] + (seq1[i - 1] == seq2[j - 1] ? p->match : p->mismatch);
            int32_t h11 = imax(SW_MIN_CUTOFF, m11);
            int32_t b = (e11 > h11) ? BT_INS : BT_MATCH;
            h11 = imax(h11, e11);
            if (f11 > h11) b = BT_DEL;
            h11 = imax(h11, f11);
            Hcur[j] = h11;
            bt[(size_t)i * (ncol + 1) + j] = (uint8_t)(b | ext);
        }
        last_col[i] = Hcur[ncol];
        int32_t* t = Hprev; Hprev = Hcur; Hcur = t;
    }
    /* Hprev now holds row nrow.  Best cell: the reference looks at the last row and the last column as the anti-diagonals
     * pass them, row candidate first (:196-217). */
    int32_t max_score = INT32_MIN, max_i = 0, max_j = 0;
    for (int32_t ad = 1; ad <= nrow + ncol; ++ad) {
        const int32_t ilo = ad < nrow + 1 ? ad : nrow + 1, jhi = ad < ncol + 1 ? ad : ncol + 1;
        const int32_t ihi = ad - jhi, jlo = ad - ilo;
        if (ilo == nrow + 1) {                      /* cell (nrow, jlo + 1) */
            const int32_t s = Hprev[jlo + 1];
            if (max_score < s || (max_score == s && abs(ilo - jlo - 2) < abs(max_i - max_j))) {
                max_score = s; max_i = ilo - 1; max_j = jlo + 1;
            }
        }
        if (jhi == ncol + 1) {                      /* cell (ihi + 1, ncol) */
            const int32_t s = last_col[ihi + 1];
            if (max_score < s || (max_score == s && (max_j == ncol || abs(ihi - jhi + 2) <= abs(max_i - max_j)))) {
                max_score = s; max_i = ihi + 1; max_j = jhi - 1;
            }
        }
    }
    *score_out = max_score;
    /* trace-back, sw_avx_get_cigar :252-361 */
    const int32_t cap = 2 * (nrow + ncol) + 8;
    int32_t* st = (int32_t*)malloc((size_t)cap * 2 * sizeof(int32_t));
    int32_t n = 0, i = max_i, j = max_j, state = 0;
    if (j < ncol) { st[2 * n] = ST_SOFT_CLIP; st[2 * n + 1] = ncol - j; ++n; }
    while (i > 0 && j > 0) {
        const int32_t btr = bt[(size_t)i * (ncol + 1) + j];
        if (state == BT_INS_EXT) {
            --j; st[2 * n - 1]++; state = btr & BT_INS_EXT;
        } else if (state == BT_DEL_EXT) {
            --i; st[2 * n - 1]++; state = btr & BT_DEL_EXT;
        } else {
            switch (btr & 3) {
                case BT_MATCH: --i; --j; st[2 * n] = BT_MATCH; st[2 * n + 1] = 1; state = 0; ++n; break;
                case BT_INS: --j; st[2 * n] = BT_INS; st[2 * n + 1] = 1; state = btr & BT_INS_EXT; ++n; break;
                case BT_DEL: --i; st[2 * n] = BT_DEL; st[2 * n + 1] = 1; state = btr & BT_DEL_EXT; ++n; break;
                default: break;                      /* 3 cannot occur: DEL replaces INS, never joins it */
            }
        }
    }
    if (j > 0) { st[2 * n] = ST_SOFT_CLIP; st[2 * n + 1] = j; ++n; }
    *alignment_offset = i;
    int32_t out = 0;
    if (n > 0) {
        int32_t m = 0;                               /* merge neighbours of one kind (:337-350) */
        for (int32_t k = 1; k < n; ++k) {
            if (st[2 * k] == st[2 * m]) st[2 * m + 1] += st[2 * k + 1];
            else { ++m; st[2 * m] = st[2 * k]; st[2 * m + 1] = st[2 * k + 1]; }
        }
        for (int32_t k = m; k >= 0; --k) {          /* reverse into CIGAR elements (:352-359, sw_avx_make_element :234-244) */
            if (st[2 * k + 1] <= 0 || out >= cigar_cap) continue;
            const uint32_t op = st[2 * k] == BT_MATCH ? 0u : st[2 * k] == BT_INS ? 1u : st[2 * k] == BT_DEL ? 2u : 4u;
            cigar[out++] = ((uint32_t)st[2 * k + 1] << 4) | op;
        }
    }
    free(st); free(bt); free(last_col); free(F); free(Hprev); free(Hcur);
    return out;
}
