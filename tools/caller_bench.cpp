// caller_bench.cpp — the reference's calling pattern against the C ABI: T caller threads (the reference's pool workers,
// haplotypecaller.cpp:52-53,97), each walking its own regions one at a time (haplotypecaller_engine.cpp:145-284).
//   mode sync : phmm_call_region per region, as the drop-in adapter does (one region in flight per thread)
//   mode async: phmm_submit for a window of W regions, then phmm_wait for each — a caller that owns a RegionSource
// Reports regions/s, GCUPS and the mean per-call latency.  Regions are cfg2-shaped (64 x 150-bp reads, 8 x 300-bp
// haplotypes), generated here with mt19937_64.
//   g++ -O2 -std=c++17 -Iinclude tools/caller_bench.cpp -Lrovaca_b200 -lphmm_b200 -Wl,-rpath,'$ORIGIN/../rovaca_b200' -lpthread -o tools/caller_bench
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "phmm_b200.h"

struct Region
{
    std::vector<uint8_t> bytes;                 // read bases | read quals | haplotype bases
    std::vector<phmm_read> reads;
    std::vector<phmm_hap> haps;
    std::vector<double> out;
    std::vector<uint32_t> kept;
    phmm_region in{};
    phmm_result res{};
};

static void make_region(Region& r, std::mt19937_64& rng, int n_reads, int n_haps, int R, int H)
{
    static const char acgt[] = "ACGT";
    r.bytes.resize((size_t)n_reads * R * 2 + (size_t)n_haps * H);
    uint8_t* rb = r.bytes.data();
    uint8_t* rq = rb + (size_t)n_reads * R;
    uint8_t* hb = rq + (size_t)n_reads * R;
    for (int c = 0; c < H; ++c) hb[c] = acgt[rng() & 3];
    for (int j = 1; j < n_haps; ++j) {
        std::memcpy(hb + (size_t)j * H, hb, H);
        for (int s = 0, n = 1 + (int)(rng() % 3); s < n; ++s) hb[(size_t)j * H + rng() % H] = acgt[rng() & 3];
    }
    r.reads.resize(n_reads);
    for (int k = 0; k < n_reads; ++k) {
        const uint8_t* src = hb + (size_t)(rng() % n_haps) * H + rng() % (H - R + 1);
        for (int i = 0; i < R; ++i) {
            rb[(size_t)k * R + i] = (rng() % 100 == 0) ? acgt[rng() & 3] : src[i];
            const unsigned u = rng() % 100;
            rq[(size_t)k * R + i] = (uint8_t)(u < 80 ? 30 + rng() % 11 : u < 95 ? 20 + rng() % 10 : 2 + rng() % 18);
        }
        r.reads[k] = phmm_read{rb + (size_t)k * R, rq + (size_t)k * R, nullptr, nullptr, nullptr, (uint32_t)R, 60};
    }
    r.haps.resize(n_haps);
    for (int j = 0; j < n_haps; ++j) r.haps[j] = phmm_hap{hb + (size_t)j * H, (uint32_t)H};
    r.out.resize((size_t)n_reads * n_haps);
    r.kept.resize(n_reads);
    r.in = phmm_region{r.reads.data(), (uint32_t)n_reads, r.haps.data(), (uint32_t)n_haps, 18, PHMM_PCR_CONSERVATIVE};
    r.res = phmm_result{};
    r.res.likelihoods = r.out.data();
    r.res.kept_idx = r.kept.data();
}

int main(int argc, char** argv)
{
    int threads = 64, per_thread = 64, window = 1, n_reads = 64, n_haps = 8, R = 150, H = 300;
    uint32_t linger = 0, workers = 0, depth = 0, lanes = 0;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        const int v = std::atoi(argv[i + 1]);
        if (k == "--threads") threads = v;
        else if (k == "--regions") per_thread = v;
        else if (k == "--window") window = v;
        else if (k == "--linger") linger = (uint32_t)v;
        else if (k == "--workers") workers = (uint32_t)v;
        else if (k == "--depth") depth = (uint32_t)v;
        else if (k == "--lanes") lanes = (uint32_t)v;             // phmm_config.lanes_per_read: pins the layout policy (experiments)
        else if (k == "--reads") n_reads = v;
        else if (k == "--haps") n_haps = v;
    }
    phmm_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.device = -1;
    cfg.linger_us = linger;
    cfg.n_workers = workers;
    cfg.pipeline_depth = depth;
    cfg.lanes_per_read = lanes;
    phmm_engine* eng = nullptr;
    if (phmm_create(&cfg, &eng) != PHMM_OK) {
        std::fprintf(stderr, "phmm_create: %s\n", phmm_last_error());
        return 1;
    }
    std::vector<std::vector<Region>> regs(threads);
    {
        std::vector<std::thread> gen;
        for (int t = 0; t < threads; ++t)
            gen.emplace_back([&, t] {
                std::mt19937_64 rng(1000 + t);
                regs[t].resize(per_thread);
                for (auto& r : regs[t]) make_region(r, rng, n_reads, n_haps, R, H);
            });
        for (auto& g : gen) g.join();
    }
    std::atomic<int> failures{0};
    std::atomic<uint64_t> lat_ns{0}, calls{0};
    auto body = [&](int t, int count) {
        for (int base = 0; base < count; base += window) {
            const int n = std::min(window, count - base);
            const auto t0 = std::chrono::steady_clock::now();
            if (window == 1) {
                if (phmm_call_region(eng, &regs[t][base].in, &regs[t][base].res) != PHMM_OK) ++failures;
            } else {
                std::vector<uint64_t> tick(n);
                for (int k = 0; k < n; ++k)
                    if (phmm_submit(eng, &regs[t][base + k].in, &regs[t][base + k].res, &tick[k]) != PHMM_OK) ++failures;
                for (int k = 0; k < n; ++k)
                    if (phmm_wait(eng, tick[k]) != PHMM_OK) ++failures;
            }
            lat_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
            ++calls;
        }
    };
    for (int pass = -1; pass < 2; ++pass) {                      // two passes warm the arenas of every slot up
        const int count = per_thread;
        lat_ns = 0;
        calls = 0;
        phmm_stats st0;
        phmm_get_stats(eng, &st0);
        const auto w0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back(body, t, count);
        for (auto& x : th) x.join();
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
        if (pass == 1) {
            const double regions = (double)threads * count;
            const double cells = regions * n_reads * n_haps * (double)R * H;
            phmm_stats st;
            phmm_get_stats(eng, &st);
            st.batches -= st0.batches; st.regions -= st0.regions; st.host_plan_ms -= st0.host_plan_ms; st.gpu_ms -= st0.gpu_ms;
            st.host_epilogue_ms -= st0.host_epilogue_ms;
            std::printf("{\"mode\": \"%s\", \"threads\": %d, \"window\": %d, \"regions\": %.0f, \"seconds\": %.4f, \"regions_per_s\": %.0f, "
                        "\"gcups\": %.1f, \"mean_call_latency_us\": %.1f, \"batches\": %llu, \"regions_per_batch\": %.1f, \"per_batch_us\": {\"plan\": %.0f, "
                        "\"gpu\": %.0f, \"epilogue\": %.0f}, \"failures\": %d}\n",
                        window == 1 ? "sync" : "async", threads, window, regions, secs, regions / secs, cells / secs / 1e9,
                        (double)lat_ns / 1e3 / (double)calls, (unsigned long long)st.batches, (double)st.regions / (double)st.batches,
                        st.host_plan_ms * 1e3 / (double)st.batches, st.gpu_ms * 1e3 / (double)st.batches, st.host_epilogue_ms * 1e3 / (double)st.batches,
                        failures.load());
        }
    }
    phmm_destroy(eng);
    return failures ? 1 : 0;
}
