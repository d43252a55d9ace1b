// call_pairhmm_b200.cpp — see call_pairhmm_b200.h.  Host glue only: it hands the reference's objects to the C ABI
// (include/phmm_b200.h) as plain pointers and rebuilds the reference's return value; no likelihood arithmetic here.
#include "call_pairhmm_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include <xmmintrin.h>

#include "common/utils/rovaca_memory_pool.h"
#include "genotype/genotype_struct.h"
#include "genotype/haplotype.h"
#include "genotype/read_record.h"
#include "phmm_b200.h"

namespace rovaca
{

// the reference's seam (pairhmm/pairhmm_engine.cpp:16-17); defined there, or by the test shim
extern DoubleVector2D (*call_pairhmm)(const HaplotypeVector& hs, ReadVector& rs, int32_t min_quality_threshold, PcrIndelModel pcr_option,
                                      pMemoryPool pool);

namespace
{
phmm_engine* g_engine = nullptr;
std::mutex g_engine_mutex;

// Optional capture of what the call site hands over: when ROVACA_B200_DUMP names a file, every region
// is appended to it as a PHMMREG1 record (format: rovaca_b200/dump.py) before it is computed, so that real-data batches can
// be replayed wherever only this repository is available.
std::mutex g_dump_mutex;
void dump_region(const phmm_region& region)
{
    const char* path = std::getenv("ROVACA_B200_DUMP");   // read per call: cheap next to a region, and tools may switch it on late
    if (!path || !*path) return;
    std::lock_guard<std::mutex> lk(g_dump_mutex);
    FILE* f = std::fopen(path, "ab");
    if (!f) return;
    const int32_t head[2] = {region.min_quality_threshold, region.pcr_model};
    std::fwrite("PHMMREG1", 1, 8, f);
    std::fwrite(&region.n_reads, 4, 1, f);
    std::fwrite(&region.n_haps, 4, 1, f);
    std::fwrite(head, 4, 2, f);
    for (uint32_t i = 0; i < region.n_reads; ++i) {
        const phmm_read& r = region.reads[i];
        std::fwrite(&r.len, 4, 1, f);
        std::fwrite(&r.mapq, 1, 1, f);
        std::fwrite(r.bases, 1, r.len, f);
        std::fwrite(r.qual, 1, r.len, f);
    }
    for (uint32_t j = 0; j < region.n_haps; ++j) {
        std::fwrite(&region.haps[j].len, 4, 1, f);
        std::fwrite(region.haps[j].bases, 1, region.haps[j].len, f);
    }
    std::fclose(f);
}
}  // namespace

// ROVACA_B200_DEVICES=all | 0,1,2,...: which GPUs the one engine of this process drives
static bool devices_from_env(std::vector<int32_t>& out, bool& all)
{
    const char* env = std::getenv("ROVACA_B200_DEVICES");
    if (!env || !*env) return false;
    all = std::strcmp(env, "all") == 0;
    if (all) return true;
    for (const char* p = env; *p;) {
        char* end = nullptr;
        const long v = std::strtol(p, &end, 10);
        if (end == p) break;
        out.push_back((int32_t)v);
        p = *end == ',' ? end + 1 : end;
    }
    return !out.empty();
}

static void create_engine(const int32_t* devices, uint32_t n)
{
    phmm_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    if (phmm_create_multi(&cfg, devices, n, &g_engine) != PHMM_OK)
        throw std::runtime_error(std::string("init_pairhmm_ptr_b200: ") + phmm_last_error());
}

void init_pairhmm_ptr_b200(int device)
{
    std::lock_guard<std::mutex> lk(g_engine_mutex);
    if (!g_engine) {
        std::vector<int32_t> ords;
        bool all = false;
        if (devices_from_env(ords, all)) {
            create_engine(all ? nullptr : ords.data(), all ? 0u : (uint32_t)ords.size());
        } else {
            const int32_t d = device;
            create_engine(&d, 1);
        }
    }
    call_pairhmm = call_pairhmm_b200;
}

void init_pairhmm_ptr_b200_devices(const int* devices, int n_devices)
{
    std::lock_guard<std::mutex> lk(g_engine_mutex);
    if (!g_engine) {
        std::vector<int32_t> ords(devices, devices + (n_devices > 0 ? n_devices : 0));
        create_engine(ords.empty() ? nullptr : ords.data(), (uint32_t)ords.size());
    }
    call_pairhmm = call_pairhmm_b200;
}

void shutdown_pairhmm_b200()
{
    std::lock_guard<std::mutex> lk(g_engine_mutex);
    if (g_engine) phmm_destroy(g_engine);
    g_engine = nullptr;
}

// One call in flight: everything lives in the region's pool.
struct PairhmmCallB200
{
    uint64_t ticket;
    phmm_result result;
    size_t n_reads, n_haps;
    double* flat;
    uint32_t* kept;
    pMemoryPool pool;
};

// The reference's objects -> the C ABI's plain views (staged in `pool`), and the submit.
static PairhmmCallB200* submit_region(const HaplotypeVector& hs, const ReadVector& rs, int32_t min_quality_threshold,
                                      PcrIndelModel pcr_option, pMemoryPool pool)
{
    if (!g_engine) throw std::runtime_error("call_pairhmm_b200: init_pairhmm_ptr_b200 has not been called");
    const size_t n_reads = rs.size(), n_haps = hs.size();
    size_t read_bytes = 0;
    for (pReadRecord r : rs) read_bytes += (size_t)r->seq_length();
    auto* call = static_cast<PairhmmCallB200*>(pool->allocate(sizeof(PairhmmCallB200), alignof(PairhmmCallB200)));
    auto* reads = static_cast<phmm_read*>(pool->allocate(sizeof(phmm_read) * n_reads, alignof(phmm_read)));
    auto* haps = static_cast<phmm_hap*>(pool->allocate(sizeof(phmm_hap) * n_haps, alignof(phmm_hap)));
    auto* bases = static_cast<uint8_t*>(pool->allocate(read_bytes + 1, 8));
    auto* region = static_cast<phmm_region*>(pool->allocate(sizeof(phmm_region), alignof(phmm_region)));
    call->flat = static_cast<double*>(pool->allocate(sizeof(double) * n_reads * n_haps, alignof(double)));
    call->kept = static_cast<uint32_t*>(pool->allocate(sizeof(uint32_t) * n_reads, alignof(uint32_t)));
    call->n_reads = n_reads;
    call->n_haps = n_haps;
    call->pool = pool;

    size_t at = 0;
    for (size_t i = 0; i < n_reads; ++i) {
        pReadRecord r = rs[i];
        const uint32_t len = (uint32_t)r->seq_length();
        for (uint32_t k = 0; k < len; ++k) bases[at + k] = r->seq_i(k);   // decode_base, read_record.cpp:176-182
        reads[i].bases = bases + at;
        reads[i].qual = r->qual();                                      // :521-523
        reads[i].ins = nullptr;                                         // ins_gops(): '-' everywhere, read_record.cpp:191-196
        reads[i].del = nullptr;                                         // del aliases ins, pairhmm_internal.cpp:559
        reads[i].gcp = nullptr;                                         // gap_conts(): '+'-33, read_record.cpp:205-210
        reads[i].len = len;
        reads[i].mapq = r->mapping_quality();
        at += len;
    }
    for (size_t j = 0; j < n_haps; ++j) {
        pBases b = hs[j]->get_display_string();                        // genotype/allele.h:104
        haps[j].bases = b->data;
        haps[j].len = b->num;
        haps[j].priority = 0.0;
    }
    region->reads = reads;
    region->n_reads = (uint32_t)n_reads;
    region->haps = haps;
    region->n_haps = (uint32_t)n_haps;
    region->min_quality_threshold = min_quality_threshold;
    region->pcr_model = static_cast<int32_t>(pcr_option);              // common/enum.h:8, same numeric values
    dump_region(*region);
    std::memset(&call->result, 0, sizeof(call->result));
    call->result.likelihoods = call->flat;
    call->result.kept_idx = call->kept;
    if (phmm_submit(g_engine, region, &call->result, &call->ticket) != PHMM_OK)
        throw std::runtime_error(std::string("call_pairhmm_b200: ") + phmm_last_error());   // caught at rovaca_main.cpp:39-42
    return call;
}

// Wait, then the reference's return value and its side effect on `rs`.
static void finish_region(PairhmmCallB200* call, ReadVector& rs, DoubleVector2D& final_likelihood)
{
    if (phmm_wait(g_engine, call->ticket) != PHMM_OK)
        throw std::runtime_error(std::string("call_pairhmm_b200: ") + phmm_last_error());
    const size_t n_reads = call->n_reads, n_haps = call->n_haps;
    for (size_t j = 0; j < n_haps; ++j) std::memcpy(final_likelihood[j].data(), call->flat + j * n_reads, sizeof(double) * n_reads);
    // filter_poorly_modelled_evidence's side effect on `rs` (:178-181): erase, back to front, every read not kept
    size_t k = call->result.n_kept;
    for (size_t i = n_reads; i-- > 0;) {
        if (k > 0 && call->kept[k - 1] == i) {
            --k;
            continue;
        }
        rs.erase(rs.begin() + (std::ptrdiff_t)i);
    }
}

DoubleVector2D call_pairhmm_b200(const HaplotypeVector& hs, ReadVector& rs, int32_t min_quality_threshold, PcrIndelModel pcr_option,
                                 pMemoryPool pool)
{
    if (hs.empty() || rs.empty()) return {};                           // pairhmm_internal.cpp:508-510
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);                        // init_native, :21-25: stays set on the caller thread
    DoubleVector2D final_likelihood{hs.size(), DoubleVector{rs.size(), NAN, pool}, pool};   // :516, allocated before the guard
    auto* lp = static_cast<RovacaMemoryPool*>(pool);
    MemoryPoolGuard scratch_guard{lp};                                  // :517-518: everything below is rewound on return
    PairhmmCallB200* call = submit_region(hs, rs, min_quality_threshold, pcr_option, pool);
    finish_region(call, rs, final_likelihood);
    return final_likelihood;
}

PairhmmCallB200* call_pairhmm_b200_submit(const HaplotypeVector& hs, const ReadVector& rs, int32_t min_quality_threshold,
                                          PcrIndelModel pcr_option, pMemoryPool pool)
{
    if (hs.empty() || rs.empty()) return nullptr;                      // wait() returns the reference's empty matrix
    return submit_region(hs, rs, min_quality_threshold, pcr_option, pool);
}

DoubleVector2D call_pairhmm_b200_wait(PairhmmCallB200* call, ReadVector& rs)
{
    if (!call) return {};
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    DoubleVector2D final_likelihood{call->n_haps, DoubleVector{call->n_reads, NAN, call->pool}, call->pool};
    finish_region(call, rs, final_likelihood);
    return final_likelihood;
}

void call_pairhmm_b200_window(PairhmmRegionB200* regions, size_t n_regions, int32_t min_quality_threshold, PcrIndelModel pcr_option)
{
    std::vector<PairhmmCallB200*> calls(n_regions, nullptr);
    for (size_t k = 0; k < n_regions; ++k)
        calls[k] = call_pairhmm_b200_submit(*regions[k].hs, *regions[k].rs, min_quality_threshold, pcr_option, regions[k].pool);
    if (g_engine) phmm_flush(g_engine);                                // the window is complete: do not wait for more callers
    for (size_t k = 0; k < n_regions; ++k) regions[k].likelihoods = call_pairhmm_b200_wait(calls[k], *regions[k].rs);
}

}  // namespace rovaca
