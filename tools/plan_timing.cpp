// Times the sections of phmm_engine::plan on the CPU (host-only slot, no GPU), cfg2-shaped synthetic regions (NOT product).
//   nvcc -O2 -std=c++17 -x cu -Irovaca_b200/csrc -Iinclude -o /tmp/plan_timing tools/plan_timing.cpp rovaca_b200/csrc/phmm_tables.cpp -Lrovaca_b200 -lphmm_b200 -Xlinker -rpath=$PWD/rovaca_b200
#include <chrono>
static double g_t[8];
#define PT_NOW() std::chrono::steady_clock::now()
#define PT_ADD(i, a, b) g_t[i] += std::chrono::duration<double, std::milli>(b - a).count()
#define PHMM_PLAN_TIMING 1
#include "../rovaca_b200/csrc/phmm_engine.cpp"
#include <random>
int main()
{
    const int G = 4096, NR = 64, NH = 8, R = 150, H = 300;
    std::vector<uint8_t> rb((size_t)G * NR * R), rq(rb.size()), hb((size_t)G * NH * H);
    std::mt19937_64 rng(1);
    for (auto& x : rb) x = "ACGT"[rng() & 3];
    for (auto& x : rq) x = 30 + rng() % 10;
    for (auto& x : hb) x = "ACGT"[rng() & 3];
    std::vector<phmm_read> reads((size_t)G * NR);
    std::vector<phmm_hap> haps((size_t)G * NH);
    for (size_t k = 0; k < reads.size(); ++k) reads[k] = phmm_read{rb.data() + k * R, rq.data() + k * R, nullptr, nullptr, nullptr, R, 60};
    for (size_t j = 0; j < haps.size(); ++j) haps[j] = phmm_hap{hb.data() + j * H, H};
    std::vector<double> out((size_t)G * NR * NH);
    std::vector<uint32_t> kept((size_t)G * NR);
    std::vector<RegionIn> regs(G);
    for (int g = 0; g < G; ++g) {
        regs[g] = RegionIn{};
        regs[g].reads = reads.data() + (size_t)g * NR; regs[g].n_reads = NR; regs[g].haps = haps.data() + (size_t)g * NH; regs[g].n_haps = NH;
        regs[g].thr = 18; regs[g].pcr = 3; regs[g].out.likelihoods = out.data() + (size_t)g * NR * NH; regs[g].out.kept_idx = kept.data() + (size_t)g * NR;
    }
    phmm_engine e;
    e.busy_batches.store(16);
    Slot s;
    s.set_host_only();
    for (int it = 0; it < 6; ++it) {
        if (it == 1) for (auto& x : g_t) x = 0;
        const auto a = PT_NOW();
        if (e.plan(s, regs) != PHMM_OK) return 1;
        PT_ADD(7, a, PT_NOW());
    }
    std::printf("plan total %.2f ms/call: count %.2f, regions loop %.2f, ctas %.2f, records copy %.2f, gather %.2f\n", g_t[7] / 5, g_t[0] / 5, g_t[1] / 5, g_t[2] / 5,
                g_t[3] / 5, g_t[4] / 5);
    return 0;
}
