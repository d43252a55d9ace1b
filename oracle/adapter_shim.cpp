/*
 * adapter_shim.cpp — drives rovaca_b200/adapter/call_pairhmm_b200.cpp (the drop-in with the reference's own signature)
 * through the reference's real types, for tests.  TEST INFRASTRUCTURE ONLY (see phmm_oracle.h).
 *
 * Built by oracle/Makefile into oracle/_ref/libphmm_adapter_test.so together with the adapter; links against
 * oracle/_ref/libphmm_ref.so (ReadRecord, RovacaMemoryPool — the reference's own objects) and
 * rovaca_b200/libphmm_b200.so (the product).  It defines the reference's global seam rovaca::call_pairhmm
 * (pairhmm/pairhmm_engine.cpp:16-17; that file itself needs the spdlog logger) so that the test can check the pointer is
 * re-bound, and builds ReadRecord / Haplotype inputs exactly like ref_region() in ref_shim.cpp does for the CPU path.
 */
#include <memory>

#include "ref_glue.h"

#include "../rovaca_b200/adapter/call_pairhmm_b200.h"

namespace rovaca
{
DoubleVector2D (*call_pairhmm)(const HaplotypeVector& hs, ReadVector& rs, int32_t min_quality_threshold, PcrIndelModel pcr_option,
                               pMemoryPool pool) = nullptr;
}

namespace
{
thread_local std::string g_err;
}

extern "C" {

const char* adapter_last_error(void) { return g_err.c_str(); }

/* init_pairhmm_ptr_b200(device); returns 0, or -1 with adapter_last_error() set (no GPU, library missing). */
int adapter_init(int device)
{
    try {
        rovaca::init_pairhmm_ptr_b200(device);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    return rovaca::call_pairhmm == rovaca::call_pairhmm_b200 ? 0 : -2;
}

void adapter_shutdown(void) { rovaca::shutdown_pairhmm_b200(); }

/* One region through the global rovaca::call_pairhmm pointer, same flat inputs/outputs as ref_region(). */
int adapter_region(uint32_t n_reads, const uint8_t* rbase, const uint8_t* rqual, const uint64_t* roff, const uint32_t* rlen,
                   const uint8_t* rmapq, uint32_t n_haps, const uint8_t* hbase, const uint64_t* hoff, const uint32_t* hlen,
                   int32_t min_quality_threshold, int pcr_model, double* out, uint32_t* kept_idx, uint32_t* n_kept)
{
    if (!rovaca::call_pairhmm) {
        g_err = "adapter_init has not been called";
        return -1;
    }
    std::vector<uint8_t> buffer((size_t)64 << 20);
    rovaca::RovacaMemoryPool pool{buffer.data(), buffer.size()};
    std::vector<ref_glue::FakeBam> bams(n_reads);
    rovaca::ReadVector rs{&pool};
    std::unordered_map<rovaca::pReadRecord, uint32_t> index_of;
    for (uint32_t k = 0; k < n_reads; ++k) {
        bams[k].build(rbase + roff[k], rqual + roff[k], rlen[k], rmapq[k]);
        rovaca::pReadRecord r = rovaca::ReadRecord::create(&pool, nullptr, &bams[k].b);
        rs.push_back(r);
        index_of[r] = k;
    }
    rovaca::HaplotypeVector hs{&pool};
    for (uint32_t j = 0; j < n_haps; ++j) hs.push_back(ref_glue::make_haplotype(&pool, hbase + hoff[j], hlen[j]));
    try {
        rovaca::DoubleVector2D m = rovaca::call_pairhmm(hs, rs, min_quality_threshold, static_cast<PcrIndelModel>(pcr_model), &pool);
        if (m.size() != n_haps && !(m.empty() && (n_haps == 0 || n_reads == 0))) return -4;
        for (uint32_t j = 0; j < m.size(); ++j) {
            if (m[j].size() != n_reads) return -5;
            std::memcpy(out + (size_t)j * n_reads, m[j].data(), sizeof(double) * n_reads);
        }
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    *n_kept = (uint32_t)rs.size();
    for (uint32_t c = 0; c < rs.size(); ++c) kept_idx[c] = index_of[rs[c]];
    return 0;
}

/* init_pairhmm_ptr_b200_devices(devices, n): one engine over several GPUs (n = 0: all visible). */
int adapter_init_devices(const int* devices, int n)
{
    try {
        rovaca::init_pairhmm_ptr_b200_devices(devices, n);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    return rovaca::call_pairhmm == rovaca::call_pairhmm_b200 ? 0 : -2;
}

/*
 * A window of regions through call_pairhmm_b200_window (submit all, then collect in order): the asynchronous use of the
 * boundary a region loop makes (haplotypecaller_engine.cpp:145-284).  Flat inputs as ref_regions_flat(); every region gets
 * its own RovacaMemoryPool, as the window form requires.  Region g's matrix at out + lik_off[g], its kept indices at
 * kept_idx + reg_read_off[g].
 */
int adapter_window(uint32_t n_regions, const uint64_t* reg_read_off, const uint64_t* reg_hap_off, const uint8_t* rbase,
                   const uint8_t* rqual, const uint64_t* roff, const uint32_t* rlen, const uint8_t* rmapq, const uint8_t* hbase,
                   const uint64_t* hoff, const uint32_t* hlen, int32_t min_quality_threshold, int pcr_model, const uint64_t* lik_off,
                   double* out, uint32_t* kept_idx, uint32_t* n_kept)
{
    if (!rovaca::call_pairhmm) {
        g_err = "adapter_init has not been called";
        return -1;
    }
    struct Region
    {
        std::unique_ptr<uint8_t[]> buffer;
        std::unique_ptr<rovaca::RovacaMemoryPool> pool;
        std::vector<ref_glue::FakeBam> bams;
        std::unique_ptr<rovaca::ReadVector> rs;
        std::unique_ptr<rovaca::HaplotypeVector> hs;
        std::unordered_map<rovaca::pReadRecord, uint32_t> index_of;
    };
    std::vector<Region> regs(n_regions);
    std::vector<rovaca::PairhmmRegionB200> win(n_regions);
    for (uint32_t g = 0; g < n_regions; ++g) {
        Region& R = regs[g];
        const uint64_t r0 = reg_read_off[g], h0 = reg_hap_off[g];
        const uint32_t nr = (uint32_t)(reg_read_off[g + 1] - r0), nh = (uint32_t)(reg_hap_off[g + 1] - h0);
        size_t need = (size_t)4 << 20;
        for (uint32_t k = 0; k < nr; ++k) need += (size_t)rlen[r0 + k] * 8 + 512;
        for (uint32_t j = 0; j < nh; ++j) need += (size_t)hlen[h0 + j] + 256;
        need += (size_t)nr * nh * 48;
        R.buffer.reset(new uint8_t[need]);
        R.pool.reset(new rovaca::RovacaMemoryPool{R.buffer.get(), need});
        R.bams.resize(nr);
        R.rs.reset(new rovaca::ReadVector{R.pool.get()});
        R.hs.reset(new rovaca::HaplotypeVector{R.pool.get()});
        for (uint32_t k = 0; k < nr; ++k) {
            R.bams[k].build(rbase + roff[r0 + k], rqual + roff[r0 + k], rlen[r0 + k], rmapq[r0 + k]);
            rovaca::pReadRecord r = rovaca::ReadRecord::create(R.pool.get(), nullptr, &R.bams[k].b);
            R.rs->push_back(r);
            R.index_of[r] = k;
        }
        for (uint32_t j = 0; j < nh; ++j) R.hs->push_back(ref_glue::make_haplotype(R.pool.get(), hbase + hoff[h0 + j], hlen[h0 + j]));
        win[g].hs = R.hs.get();
        win[g].rs = R.rs.get();
        win[g].pool = R.pool.get();
    }
    try {
        rovaca::call_pairhmm_b200_window(win.data(), n_regions, min_quality_threshold, static_cast<PcrIndelModel>(pcr_model));
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
    for (uint32_t g = 0; g < n_regions; ++g) {
        const uint32_t nr = (uint32_t)(reg_read_off[g + 1] - reg_read_off[g]), nh = (uint32_t)(reg_hap_off[g + 1] - reg_hap_off[g]);
        const rovaca::DoubleVector2D& m = win[g].likelihoods;
        if (m.size() != nh && !(m.empty() && (nh == 0 || nr == 0))) return -4;
        for (uint32_t j = 0; j < m.size(); ++j) {
            if (m[j].size() != nr) return -5;
            std::memcpy(out + lik_off[g] + (size_t)j * nr, m[j].data(), sizeof(double) * nr);
        }
        n_kept[g] = (uint32_t)regs[g].rs->size();
        for (uint32_t c = 0; c < regs[g].rs->size(); ++c) kept_idx[reg_read_off[g] + c] = regs[g].index_of[(*regs[g].rs)[c]];
        win[g].likelihoods = rovaca::DoubleVector2D{};        /* release before its pool goes */
    }
    return 0;
}

}  // extern "C"
