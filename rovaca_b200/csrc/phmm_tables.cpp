// phmm_tables.cpp — the lookup tables of the reference's Context<float>/Context<double>
// (pairhmm/core/context.h:13-190 == pairhmm/rovaca/context.h), built on the host with glibc's libm so
// that the GPU multiplies by the very same bits the reference's CPU path does.  Only the first 128
// quality rows of matchToMatchProb are kept: every quality is masked & 127 before it indexes a table
// (pairhmm_internal.cpp:466-469).
#include <cmath>
#include <vector>

#include "phmm_types.h"

namespace
{
constexpr double kJacTol = 8.0;        // MAX_JACOBIAN_TOLERANCE      context.h:8
constexpr double kJacStep = 0.0001;    // JACOBIAN_LOG_TABLE_STEP     context.h:9
constexpr double kJacInvStep = 1.0 / kJacStep;
constexpr int kJacSize = (int)(kJacTol / kJacStep) + 1;

template <typename N>
N approx_log10_sum(N small, N big, const std::vector<N>& jac)   // context.h:64-86
{
    if (small > big) std::swap(small, big);
    if (std::isinf(small) == 1 || std::isinf(big) == 1) return big;
    const N diff = big - small;
    if (diff >= (N)kJacTol) return big;
    const N scaled = (N)(diff * ((N)kJacInvStep));
    const int ind = (scaled > (N)0.0) ? (int)(scaled + (N)0.5) : (int)(scaled - (N)0.5);   // fastRound :62
    return big + jac[ind];
}

template <typename N>
void build_mm(N* out)   // initializeJacobianLogTable + initializeMatchToMatchProb, context.h:41-59
{
    std::vector<N> jac(kJacSize);
    for (int k = 0; k < kJacSize; k++) jac[k] = (N)(log10(1.0 + pow(10.0, -((double)k) * kJacStep)));
    const double LN10 = log(10);
    const double INV_LN10 = 1.0 / LN10;
    for (int i = 0, offset = 0; i < 128; offset += ++i)
        for (int j = 0; j <= i; j++) {
            const double log10Sum = approx_log10_sum<N>((N)(-0.1 * i), (N)(-0.1 * j), jac);
            const double matchToMatchLog10 = log1p(-std::min(1.0, pow(10, log10Sum))) * INV_LN10;
            out[offset + j] = (N)(pow(10, matchToMatchLog10));
        }
}
}  // namespace

void phmm_build_tables(PhmmTables* t)
{
    build_mm<float>(t->mm);
    build_mm<double>(t->mm_d);
    for (int x = 0; x < 128; x++) {
        t->ph2pr[x] = powf(10.f, -((float)x) / 10.f);       // context.h:144-146
        t->ph2pr_d[x] = pow(10.0, -((double)x) / 10.0);     // context.h:102-104
        t->gapm[x] = 1.0f - t->ph2pr[x];                    // pairhmm_internal.cpp:472
        t->pm[x] = 1.0f - t->ph2pr[x];                      // pairhmm_internal.cpp:479
        t->pmm[x] = t->ph2pr[x] / 3.0f;                     // pairhmm_internal.cpp:480
    }
}
