// Times the sections of PhmmDevice::plan on the CPU (host-only slot, no GPU), cfg2- or cfg5-shaped synthetic regions (NOT product).
//   nvcc -O2 -std=c++17 -x cu -Irovaca_b200/csrc -Iinclude -o build/plan_timing tools/plan_timing.cpp -Lrovaca_b200 -lphmm_b200 -Xlinker -rpath=$PWD/rovaca_b200
#include <chrono>
static double g_t[12];
#define PT_NOW() std::chrono::steady_clock::now()
#define PT_ADD(i, a, b) g_t[i] += std::chrono::duration<double, std::milli>(b - a).count()
#define PHMM_PLAN_TIMING 1
#include "../rovaca_b200/csrc/phmm_plan.cpp"
#include <random>
#include <string>
int main(int argc, char** argv)
{
    const bool trim = argc > 1 && std::string(argv[1]) == "cfg5t";     // post-trim shapes: H 41-120, reads 10..min(H, 100)
    const bool mixed = trim || (argc > 1 && std::string(argv[1]) == "cfg5");     // cfg5-like: Poisson(41) reads of 60-151 bp, 2-24 haplotypes of 201-499
    const int G = trim ? 32768 : mixed ? 16384 : 4096;
    std::mt19937_64 rng(1);
    std::vector<uint32_t> nr(G), nh(G);
    std::vector<std::vector<uint32_t>> rl(G), hl(G);
    size_t n_reads = 0, n_haps = 0, rbytes = 0, hbytes = 0, n_pairs = 0;
    std::poisson_distribution<int> pr(41), ph(7.8);
    for (int g = 0; g < G; ++g) {
        nr[g] = mixed ? std::max(1, pr(rng)) : 64;
        nh[g] = mixed ? std::min(24, std::max(2, ph(rng) + 1)) : 8;
        const uint32_t H = trim ? 41 + rng() % 80 : mixed ? 201 + rng() % 299 : 300;
        for (uint32_t k = 0; k < nr[g]; ++k) { rl[g].push_back(trim ? 10 + rng() % (std::min(H, 100u) - 9) : mixed ? 60 + rng() % 92 : 150); rbytes += rl[g].back(); }
        std::sort(rl[g].begin(), rl[g].end());
        for (uint32_t j = 0; j < nh[g]; ++j) { hl[g].push_back(H); hbytes += H; }
        n_reads += nr[g]; n_haps += nh[g]; n_pairs += (size_t)nr[g] * nh[g];
    }
    std::vector<uint8_t> rb(rbytes), rq(rbytes), hb(hbytes);
    for (auto& x : rb) x = "ACGT"[rng() & 3];
    for (auto& x : rq) x = 30 + rng() % 10;
    for (auto& x : hb) x = "ACGT"[rng() & 3];
    // a flat batch over the arrays above: what phmm_run_flat hands the planner (reads are not materialised as view structs)
    std::vector<uint64_t> rro(G + 1, 0), rho(G + 1, 0), roff(n_reads), hoff(n_haps);
    std::vector<uint32_t> rlen(n_reads), hlen(n_haps);
    std::vector<uint8_t> mapq(n_reads, 60);
    {
        size_t ri = 0, hi = 0;
        uint64_t ro = 0, ho = 0;
        for (int g = 0; g < G; ++g) {
            for (uint32_t k = 0; k < nr[g]; ++k) { rlen[ri] = rl[g][k]; roff[ri] = ro; ro += rl[g][k]; ++ri; }
            for (uint32_t j = 0; j < nh[g]; ++j) { hlen[hi] = hl[g][j]; hoff[hi] = ho; ho += hl[g][j]; ++hi; }
            rro[g + 1] = ri;
            rho[g + 1] = hi;
        }
    }
    phmm_flat_batch fb{};
    fb.n_regions = (uint32_t)G; fb.reg_read_off = rro.data(); fb.reg_hap_off = rho.data(); fb.read_len = rlen.data(); fb.read_mapq = mapq.data();
    fb.read_off = roff.data(); fb.read_bases = rb.data(); fb.read_qual = rq.data(); fb.hap_len = hlen.data(); fb.hap_off = hoff.data();
    fb.hap_bases = hb.data(); fb.min_quality_threshold = 18; fb.pcr_model = 3;
    std::vector<phmm_hap> haps;
    std::vector<RegionIn> regs;
    const auto tf = PT_NOW();
    for (int it = 0; it < 5; ++it) flat_to_regions(&fb, nullptr, 0, (uint32_t)G, haps, regs);
    const double flat_ms = std::chrono::duration<double, std::milli>(PT_NOW() - tf).count() / 5;
    PhmmSettings cfg;
    PhmmDevice e;
    e.cfg = &cfg;
    e.busy_batches.store(16);
    Slot s;
    s.set_host_only();
    for (int it = 0; it < 6; ++it) {
        if (it == 1) for (auto& x : g_t) x = 0;
        const auto a = PT_NOW();
        if (e.plan(s, regs) != PHMM_OK) return 1;
        PT_ADD(7, a, PT_NOW());
    }
    std::printf("flat_to_regions %.2f ms; ", flat_ms);
    std::printf("%d regions, %zu reads, %zu pairs: plan total %.2f ms/call (%.2f us/region): count %.2f, regions loop %.2f, ctas %.2f, records copy %.2f, gather %.2f (inside the regions loop: read-length scan %.2f, haplotype chunks %.2f, read records %.2f, paired-reads packing %.2f, rows per lane + packing %.2f)\n", G,
                n_reads, n_pairs, g_t[7] / 5, g_t[7] / 5 * 1e3 / G, g_t[0] / 5, g_t[1] / 5, g_t[2] / 5, g_t[3] / 5, g_t[4] / 5, g_t[5] / 5, g_t[8] / 5, g_t[9] / 5, g_t[6] / 5, g_t[10] / 5);
    return 0;
}
