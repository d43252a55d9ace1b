// realign_chain.cpp — the two calls a maintainer makes per region once Rovaca's PairHMM path is bound to the B200 engine:
// phmm_call_region (= call_pairhmm + best_alleles_breaking_ties, haplotypecaller_engine.cpp:250 and
// genotype/utils/assembly_based_caller_utils.cpp:146-160) and phmm_sw_align (= the sw->avx_function(sw) of every read against
// its best haplotype, genotype/utils/alignment_utils.cpp:574-579).  One synthetic region; prints what came back.
//   g++ -O2 -std=c++17 -Iinclude examples/realign_chain.cpp -Lrovaca_b200 -lphmm_b200 -Wl,-rpath,'$ORIGIN/../rovaca_b200' -o examples/realign_chain
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "phmm_b200.h"

int main()
{
    std::mt19937_64 rng(7);
    const char acgt[] = "ACGT";
    const uint32_t n_haps = 4, n_reads = 12, H = 200, R = 100;
    std::vector<uint8_t> hap_bases(n_haps * H), read_bases(n_reads * R), read_qual(n_reads * R, 35);
    for (uint32_t c = 0; c < H; ++c) hap_bases[c] = acgt[rng() & 3];
    for (uint32_t j = 1; j < n_haps; ++j) {                       // the reference haplotype with one substitution each
        std::memcpy(&hap_bases[j * H], &hap_bases[0], H);
        hap_bases[j * H + 40 * j] = acgt[rng() & 3];
    }
    std::vector<phmm_hap> haps(n_haps);
    for (uint32_t j = 0; j < n_haps; ++j) haps[j] = phmm_hap{&hap_bases[j * H], H, /*priority=*/j == 0 ? 1.0 : 0.0};
    std::vector<phmm_read> reads(n_reads);
    for (uint32_t k = 0; k < n_reads; ++k) {
        std::memcpy(&read_bases[k * R], &hap_bases[(k % n_haps) * H + (rng() % (H - R))], R);
        reads[k] = phmm_read{&read_bases[k * R], &read_qual[k * R], nullptr, nullptr, nullptr, R, 60};
    }

    phmm_engine* engine = nullptr;
    if (phmm_create(nullptr, &engine) != PHMM_OK) {
        std::fprintf(stderr, "phmm_create: %s\n", phmm_last_error());
        return 1;
    }
    std::vector<double> lik(n_haps * n_reads);
    std::vector<uint32_t> kept(n_reads), best(n_reads), second(n_reads);
    phmm_region region{reads.data(), n_reads, haps.data(), n_haps, 18, PHMM_PCR_CONSERVATIVE};
    phmm_result res{};
    res.likelihoods = lik.data();
    res.kept_idx = kept.data();
    res.best_hap = best.data();
    res.second_hap = second.data();
    if (phmm_call_region(engine, &region, &res) != PHMM_OK) {
        std::fprintf(stderr, "phmm_call_region: %s\n", phmm_last_error());
        return 1;
    }

    std::vector<phmm_sw_pair> pairs(res.n_kept);
    std::vector<phmm_sw_result> aligned(res.n_kept);
    std::vector<uint32_t> cigars(res.n_kept * 16);
    for (uint32_t k = 0; k < res.n_kept; ++k) {
        const phmm_hap& h = haps[best[k]];
        const phmm_read& r = reads[kept[k]];
        pairs[k] = phmm_sw_pair{h.bases, h.len, r.bases, r.len};
        aligned[k] = phmm_sw_result{};
        aligned[k].cigar = &cigars[k * 16];
        aligned[k].cigar_cap = 16;
    }
    const phmm_sw_params prm = PHMM_SW_PARAMS_REALIGN;
    if (phmm_sw_align(engine, pairs.data(), res.n_kept, &prm, aligned.data()) != PHMM_OK) {
        std::fprintf(stderr, "phmm_sw_align: %s\n", phmm_last_error());
        return 1;
    }
    for (uint32_t k = 0; k < res.n_kept; ++k) {
        const double bl = lik[best[k] * n_reads + k], sl = second[k] != best[k] ? lik[second[k] * n_reads + k] : -1.0 / 0.0;
        std::printf("read %2u: best haplotype %u (%.3f), second %u, %s; aligned at %d, %u CIGAR element(s), first %u%c\n", kept[k], best[k], bl,
                    second[k], bl - sl > 0.2 ? "informative" : "uninformative", aligned[k].alignment_offset, aligned[k].n_cigar,
                    aligned[k].n_cigar ? aligned[k].cigar[0] >> 4 : 0, aligned[k].n_cigar ? "MIDNS"[aligned[k].cigar[0] & 15] : '-');
    }
    phmm_destroy(engine);
    return 0;
}
