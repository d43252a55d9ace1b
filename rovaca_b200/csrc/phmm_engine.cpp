// phmm_engine.cpp — host side of the B200 PairHMM engine and its C ABI (include/phmm_b200.h).
//
// Mirrors what the reference's host wrapper does around its kernels (pairhmm/pairhmm_internal.cpp
// call_old_pairhmm :505-594 / call_new_pairhmm002 :596-753), re-cut for a GPU:
//   * regions from many callers are gathered into batches (the reference runs one region per call on a pool thread,
//     haplotypecaller_engine.cpp:119,250 — far too little work for a B200): tickets (phmm_submit / phmm_wait) with a
//     batch-while-busy sealing policy for its synchronous callers, phmm_run_flat for whole streams;
//   * per batch the host only *plans*: reads are binned by length (lanes per read x rows per lane), remainders of sparse bins
//     are promoted so that warps run full, each region's haplotypes are dealt into chunks of two balanced column streams
//     (shorter chunks when the device is idle: a chunk's length is a call's latency), CTAs of 1-4 warps are formed; the
//     records go through one pinned arena, the read / haplotype bytes too, or straight from the caller's page-locked buffers;
//   * quality capping, the PCR-indel model, the table look-ups, the forward recurrence and the FP64 underflow recompute all
//     run on the device (phmm_kernels.cu), the per-bin launches side by side on four streams;
//   * the raw sums come back and the host finishes exactly as the reference does: log10f/log10 with glibc (so the bits match,
//     pairhmm_internal.cpp:578-584), normalize_likelihoods (:53-63), filter_poorly_modelled_evidence (:161-182), transpose
//     into the NaN-tailed matrix (:184-193);
//   * the step after the path, Smith-Waterman realignment of reads to their best haplotype, is phmm_sw_align (sw_kernels.cu).
// There is no CPU implementation of the recurrence in this library and no fallback.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <xmmintrin.h>

#include <cuda_runtime.h>

#include "../../include/phmm_b200.h"
#include "phmm_kernels.h"
#include "phmm_log10.h"
#include "phmm_types.h"

void phmm_build_tables(PhmmTables* t);

namespace
{

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t _e = (call);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            return fail(PHMM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));               \
    } while (0)

constexpr uint32_t kTargetStreamCols = 2400;    // per stream and chunk: keeps streams + prior tables under ~64 KB per CTA (3+ CTAs/SM)
constexpr uint32_t kWavesWanted = 16;          // warps per warp slot a batch should have before its chunks may be full length
constexpr uint32_t kMinStreamCols = 200;        // small batches are cut finer (more, shorter warps), but never below this
constexpr uint32_t kMaxStreamCols = 22000;      // hard limit of one haplotype in the wavefront kernel (176 KB of streams + 37 KB of prior tables)
constexpr uint32_t kPairThreads = 148 * 3 * 128;   // per-pair wavefront kernel: one warp per pair, 12 warps per SM
constexpr float kMinAccepted = 1e-28f;                     // MIN_ACCEPTED, rovaca/common.h:9
constexpr float kLog10InitF = 36.1236000061f;              // LOG10_INITIAL_CONSTANT_F, adapter.h:25
constexpr double kLog10InitD = 307.050595577260822;        // LOG10_INITIAL_CONSTANT_D, adapter.h:26

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct RegionIn
{
    const phmm_read* reads;
    uint32_t n_reads;
    const phmm_hap* haps;
    uint32_t n_haps;
    int32_t thr;
    int32_t pcr;
    phmm_result out;        // caller buffers
    uint32_t* n_kept_out;   // where to store n_kept
    uint64_t* n_fp64_out;   // accumulates (may be null)
    uint32_t* n_fp64_out32; // per-region (may be null)
    bool pinned_src;        // the read / haplotype bytes live in page-locked memory the device can read directly
};

struct RegionPlan
{
    uint32_t read_begin, n_reads, hap_begin, n_haps;
    uint64_t out_off;
};

// Growing an arena means cudaFree + cudaMalloc (a device-wide sync) or re-pinning host pages (milliseconds), so arenas grow
// in big steps and straight to the biggest size any worker of the engine has needed for that buffer (the high-water mark):
// after a change of workload every slot reallocates at most once.
inline size_t grow_to(size_t bytes, std::atomic<size_t>* high_water)
{
    size_t want = align_up(std::max(bytes + bytes / 2, (size_t)4 << 20), 1 << 20);
    if (high_water) {
        size_t seen = high_water->load(std::memory_order_relaxed);
        while (seen < want && !high_water->compare_exchange_weak(seen, want, std::memory_order_relaxed)) {}
        want = std::max(want, seen);
    }
    return want;
}

struct DevBuf
{
    void* p = nullptr;
    size_t cap = 0;
    bool host_only = false;     // planner self-check (phmm_debug_plan): no device behind it, nothing is allocated
    int ensure(size_t bytes, std::atomic<size_t>* high_water = nullptr)
    {
        if (bytes <= cap || host_only) return PHMM_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = grow_to(bytes, high_water);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(PHMM_ERR_MEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        cap = want;
        return PHMM_OK;
    }
    void release()
    {
        if (p && !host_only) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostBuf
{
    void* p = nullptr;
    size_t cap = 0;
    bool host_only = false;     // planner self-check: plain malloc instead of pinned memory
    int ensure(size_t bytes, std::atomic<size_t>* high_water = nullptr)
    {
        if (bytes <= cap) return PHMM_OK;
        release();
        size_t want = grow_to(bytes, high_water);
        if (host_only) {
            p = std::malloc(want);
            if (!p) return fail(PHMM_ERR_MEM, "malloc failed");
        } else {
            cudaError_t e = cudaMallocHost(&p, want);
            if (e != cudaSuccess) return fail(PHMM_ERR_MEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
        }
        cap = want;
        return PHMM_OK;
    }
    void release()
    {
        if (p) {
            if (host_only) std::free(p); else cudaFreeHost(p);
        }
        p = nullptr;
        cap = 0;
    }
};

struct FwdLaunch
{
    int bin;
    uint32_t cta_base, n_ctas, cta_warps;
    size_t smem;
};

// One batch in flight: its plan, its pinned staging, its device memory, its stream.
struct Slot
{
    static constexpr int kAux = 3;
    cudaStream_t stream = nullptr;
    cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr};       // the per-bin forward launches of one batch run side by side
    cudaEvent_t aux_done[kAux] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // kernels begin, forward begin, forward end, kernels end
    HostBuf h_in, h_f32, h_f64, h_cnt;
    DevBuf d_in, d_words, d_f32, d_f64, d_pairs, d_cnt, d_scr64, d_scr32;
    PhmmBatchDev dev{};
    std::vector<RegionPlan> regions;
    std::vector<FwdLaunch> launches;
    std::vector<uint32_t> read_len;      // per read (for the filter)
    // plan()'s working storage: kept between batches so that a plan allocates nothing once its slot has seen a batch its size
    struct Group { uint32_t chunk; uint32_t stream_len; uint32_t first; uint32_t n; };
    struct PlanScratch
    {
        std::vector<PhmmRead> reads;
        std::vector<PhmmHap> haps;
        std::vector<PhmmChunk> chunks;
        std::vector<PhmmPair> slow;
        std::vector<PhmmCta> ctas;
        std::vector<PhmmRegionDev> regions_dev;
        std::vector<double> pri;
        std::vector<uint32_t> task_reads, order, region_chunks, idx, bucket;
        std::vector<std::vector<Group>> groups;
        std::vector<Group> sorted;
        std::vector<std::vector<uint32_t>> group_reads, by_bin;
    } ps;
    size_t in_bytes = 0;
    size_t staged_bytes = 0;             // leading part of the arena that goes through the pinned staging copy
    struct DirectCopy { size_t dst; const uint8_t* src; size_t len; };
    std::vector<DirectCopy> direct;      // byte blobs copied straight from the caller's page-locked buffers
    uint64_t n_pairs = 0, n_cells = 0, n_words = 0;
    uint32_t n_reads = 0, n_haps = 0, n_slow = 0, max_cols = 0;
    uint32_t n_launches = 0;
    // device epilogue: one buffer on each side, [lik | kept_idx | best | second | n_kept | n_under]
    DevBuf d_epi;
    HostBuf h_epi;
    PhmmEpilogueDev epi{};
    size_t e_kept = 0, e_best = 0, e_second = 0, e_nkept = 0, e_under = 0, epi_bytes = 0;
    size_t o_regions = 0, o_pri = 0;
    bool dev_epi = false;                // this batch's epilogue runs on the device
    bool want_best = false;              // some region of the batch asks for best / second-best haplotypes
    bool want_raw = false;               // ... for the raw FP32 sums
    bool want_f64 = false;               // the FP64 recompute goes out with the batch (someone will read its values)
    bool f64_done = false;               // ... and has been launched for the sums now on the device
    float gpu_ms = 0.f, fwd_ms = 0.f;
    size_t o_reads = 0, o_haps = 0, o_chunks = 0, o_ctas = 0, o_tasks = 0, o_slow = 0;   // arena offsets of the record arrays
    size_t n_chunks = 0, n_ctas = 0, n_tasks = 0;

    void set_host_only()
    {
        for (auto* b : {&h_in, &h_f32, &h_f64, &h_cnt, &h_epi}) b->host_only = true;
        for (auto* b : {&d_in, &d_words, &d_f32, &d_f64, &d_pairs, &d_cnt, &d_scr64, &d_scr32, &d_epi}) b->host_only = true;
    }

    int init()
    {
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (auto& a : aux) CU(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        for (auto& e : aux_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ev) CU(cudaEventCreate(&e));
        return PHMM_OK;
    }
    void destroy()
    {
        for (auto* b : {&h_in, &h_f32, &h_f64, &h_cnt, &h_epi}) b->release();
        for (auto* b : {&d_in, &d_words, &d_f32, &d_f64, &d_pairs, &d_cnt, &d_scr64, &d_scr32, &d_epi}) b->release();
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : aux_done)
            if (e) cudaEventDestroy(e);
        for (auto& a : aux)
            if (a) cudaStreamDestroy(a);
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};

struct Stats
{
    std::atomic<uint64_t> regions{0}, pairs{0}, cells{0}, fp64_pairs{0}, slow_pairs{0}, h2d{0}, d2h{0}, batches{0}, launches{0}, dev_epilogues{0};
    std::mutex m;
    double gpu_ms = 0.0, plan_ms = 0.0, epilogue_ms = 0.0;
};

// Fill the read slots of a region's CTAs.  A warp of the forward kernel holds 32 / LANES reads of ONE bin; a region whose
// reads scatter over many length bins would run many part-empty warps.  A read may run in any higher bin of its lane family
// (more top padding, same result), so the remainder of a bin can be carried up into the next occupied one.  Cost of a warp
// = rows per lane of its bin; the small DP below picks, per family, which remainders to carry (f[i][c]: best cost of bins
// i.. with c reads carried in).
void promote_remainders(std::vector<std::vector<uint32_t>>& by_bin, int first, int last)
{
    constexpr int kMaxBins = 24, kMaxS = 4;
    const int S = 32 / kPhmmBins[first].lanes;            // reads per warp in this family
    const int nb = last - first + 1;
    if (S <= 1 || nb < 2 || nb > kMaxBins || S > kMaxS) return;
    int occupied = 0;
    bool ragged = false;
    for (int i = 0; i < nb; ++i) {
        const size_t n = by_bin[first + i].size();
        occupied += n != 0;
        ragged |= n % S != 0;
    }
    if (occupied < 2 || !ragged) return;                   // nothing to merge
    uint32_t f[kMaxBins + 1][kMaxS] = {};
    uint8_t carry[kMaxBins][kMaxS] = {};
    uint32_t own[kMaxBins];
    for (int i = nb - 1; i >= 0; --i) {
        const uint32_t rpt = (uint32_t)kPhmmBins[first + i].rpt;
        own[i] = (uint32_t)by_bin[first + i].size();
        for (int c = 0; c < S; ++c) {
            const uint32_t t = (uint32_t)c + own[i];
            uint32_t best = (t + S - 1) / S * rpt + f[i + 1][0];
            uint8_t how = 0;
            if (i + 1 < nb && t % S != 0) {
                const uint32_t alt = t / S * rpt + f[i + 1][t % S];
                if (alt < best) { best = alt; how = 1; }
            }
            f[i][c] = best;
            carry[i][c] = how;
        }
    }
    uint32_t c = 0;                                        // reads carried into bin i (already appended to its vector)
    for (int i = 0; i + 1 < nb; ++i) {
        auto& cur = by_bin[first + i];
        const uint32_t t = c + own[i];                    // == cur.size()
        if (carry[i][c]) {
            const uint32_t rem = t % S;
            auto& up = by_bin[first + i + 1];
            up.insert(up.end(), cur.end() - rem, cur.end());
            cur.resize(t - rem);
            c = rem;
        } else {
            c = 0;
        }
    }
}

}  // namespace

struct phmm_resident
{
    Slot slot;
    std::vector<RegionIn> regs;
    std::vector<phmm_read> reads;
    std::vector<phmm_hap> haps;
};

struct phmm_engine
{
    int device = 0;
    bool fma = true;
    uint32_t n_workers = 6;
    uint64_t batch_cells = 32000000000ull;   // mixed-length batches under ~8e9 cells run their per-bin kernels too small to fill the device
    uint32_t linger_us = 200;
    uint32_t cta_warps = 0;      // warps per forward-kernel CTA; 0 = chosen per launch from the fill of the read slots
    uint32_t lanes_pref = 0;     // lanes per read of the reference-gap kernels: 0 = by batch size, else 8 / 16 / 32
    bool fp64_always = false;    // PHMM_FP64_ALWAYS: recompute every underflowing pair even when nobody can see the value
    uint32_t warp_slots = 148 * 8; // warps of the widest forward kernel the device holds at once (8 per SM at 231 registers)
    bool device_epilogue = false;  // log10 / normalize / filter / transpose / best haplotype in one more kernel (PHMM_EPILOGUE_*)
    uint64_t device_epilogue_min_pairs = 0;   // ... for batches of at least this many pairs (AUTO: small batches stay on the host)
    int log10_mode = 0;            // how raw sums become log10 doubles: phmm_log10.cpp, picked (and tested against libm) at creation
    std::atomic<uint32_t> busy_batches{0};   // batches somewhere between plan and the end of their epilogue
    std::atomic<size_t> high_water[8] = {};   // per arena kind, see grow_to()
    float default_gapm = 0.f, default_ii = 0.f;   // table values of the default gap-continuation quality

    // batched Smith-Waterman (phmm_sw_align): its own stream and arenas, one call at a time
    struct SwSlot                    // one sub-batch in flight: the host packs the next one while this one runs
    {
        cudaStream_t stream = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        HostBuf h_in, h_out;
        DevBuf d_in, d_bt, d_aux, d_cig, d_out;
        std::vector<PhmmSwTask> tasks;
        std::vector<uint32_t> src, order;       // task -> pair of the call; tasks ordered by kernel instantiation
        uint32_t n = 0, cigar_cap = 0;
        uint64_t cells = 0;
        size_t in_bytes = 0;
        bool busy = false;
    };
    struct SwContext
    {
        std::mutex m;
        SwSlot slot[2];
        cudaEvent_t base = nullptr;      // start of the call on the device's clock: sub-batches overlap, their kernel time is a union
        float covered_ms = 0.f;
        float last_ms = 0.f;
        uint64_t last_cells = 0;
    } sw;
    DevBuf d_tables;
    Stats stats;

    // workers + job queue
    std::vector<std::thread> workers;
    std::vector<std::unique_ptr<Slot>> slots;
    std::mutex qm;
    std::condition_variable qcv;
    std::deque<std::function<void(Slot&)>> jobs;
    bool stopping = false;

    // region (ticket) API
    struct OpenBatch
    {
        uint64_t id = 0;
        std::vector<RegionIn> regs;
        uint64_t cells = 0;
        uint32_t waiters = 0;          // threads blocked in phmm_wait on this batch
        std::chrono::steady_clock::time_point first;
    };
    struct BatchSync            // completion of one sealed batch; waiters block on the batch's own condition variable
    {
        std::mutex m;
        std::condition_variable cv;
        bool done = false;
        int rc = PHMM_OK;
        std::string err;
    };
    struct Pending
    {
        std::shared_ptr<BatchSync> sync;
        uint32_t waits = 0;     // regions submitted under this ticket that have not been waited for yet
    };
    std::mutex bm;
    std::condition_variable open_cv;   // signalled when the open batch is sealed
    OpenBatch open;
    uint64_t next_batch_id = 1;
    std::unordered_map<uint64_t, Pending> batch_state;
    // Sealing policy (the callers are the reference's synchronous pool workers, haplotypecaller_engine.cpp:250): batch while
    // busy.  A batch with a waiter goes out at once while fewer than kPipelineDepth batches are in flight (the pipeline is
    // starving: plan / GPU / epilogue of different batches overlap); otherwise it keeps collecting the callers that return
    // from earlier batches and goes out when one of those completes, when it is big (batch_cells), or after linger_us.
    uint32_t in_flight = 0;            // sealed batches that have not completed
    uint32_t pipeline_depth = 3;       // batches the ticket path keeps in flight before it starts batching up

    int plan(Slot& s, const std::vector<RegionIn>& regs);
    int launch(Slot& s, bool copy_in);
    int finish(Slot& s, const std::vector<RegionIn>& regs, bool copy_out);
    void epilogue(const Slot& s, const std::vector<RegionIn>& regs, const float* f32, const double* f64, bool have64) const;
    void epilogue_from_device(const Slot& s, const std::vector<RegionIn>& regs, bool have64) const;
    int execute(Slot& s, const std::vector<RegionIn>& regs);
    void worker_main(uint32_t idx);
    void enqueue(std::function<void(Slot&)> fn);
    void seal_locked();
};

// ------------------------------------------------------------------------------------------------
// Planning + packing
// ------------------------------------------------------------------------------------------------

#ifndef PHMM_PLAN_TIMING      // section timers for tools/plan_timing.cpp; compiled out of the library
#define PT_NOW() 0
#define PT_ADD(i, a, b) (void)0
#endif

int phmm_engine::plan(Slot& s, const std::vector<RegionIn>& regs)
{
    [[maybe_unused]] const auto pt0 = PT_NOW();
    uint64_t n_reads = 0, n_haps = 0, read_bytes = 0, hap_bytes = 0, n_pairs = 0, n_cells = 0;
    bool any_ins = false, any_del = false, any_gcp = false;
    s.want_f64 = fp64_always;
    s.want_best = s.want_raw = false;
    for (const RegionIn& r : regs) {
        if (r.n_reads == 0 || r.n_haps == 0) continue;           // pairhmm_internal.cpp:508-510
        s.want_f64 |= r.out.raw_f64 != nullptr;
        s.want_raw |= r.out.raw_f32 != nullptr || r.out.raw_f64 != nullptr;
        s.want_best |= r.out.best_hap != nullptr;
        uint64_t rb = 0, hb = 0;
        for (uint32_t k = 0; k < r.n_reads; ++k) {
            rb += r.reads[k].len;
            any_ins |= r.reads[k].ins != nullptr;
            any_del |= r.reads[k].del != nullptr;
            any_gcp |= r.reads[k].gcp != nullptr;
        }
        for (uint32_t j = 0; j < r.n_haps; ++j) hb += r.haps[j].len;
        n_reads += r.n_reads;
        n_haps += r.n_haps;
        read_bytes += rb;
        hap_bytes += hb;
        n_pairs += (uint64_t)r.n_reads * r.n_haps;
        n_cells += rb * hb;
    }
    // One more kernel and a D2H twice the size pay off once the host would otherwise spend longer on the pairs than a launch
    // costs; the few-region batches of synchronous callers keep the (vector) host epilogue.
    s.dev_epi = device_epilogue && n_pairs >= device_epilogue_min_pairs;
    if (n_pairs >= (1ull << 31) || read_bytes >= (1ull << 31) || hap_bytes >= (1ull << 31))
        return fail(PHMM_ERR_ARG, "batch too large for 32-bit device offsets; lower batch_cells");
    // What the reference itself feeds its kernels — deletion qualities aliasing the insertion ones, one gap-continuation
    // quality per read (read_record.cpp:191-210, pairhmm_internal.cpp:425,559) — is served by the leaner kernel variant.
    const bool refgaps = fma && !any_del && !any_gcp;
    // A warp sweeps its chunk's columns one after the other (~0.2 us per column), so the chunk length is the latency of a
    // batch and chunks x reads its parallelism.  A big batch fills the GPU with long chunks (fewest seams and staging); a
    // small one (the synchronous per-region callers) is cut finer until it has about two warps per warp slot of the device.
    uint32_t target_cols = kTargetStreamCols;
    int min_lanes = 8;             // lanes per read of the reference-gap kernels
    {
        uint64_t est_warps = 0;
        for (const RegionIn& r : regs) {
            uint64_t hb = 0;
            for (uint32_t j = 0; j < r.n_haps; ++j) hb += r.haps[j].len;
            est_warps += (uint64_t)((r.n_reads + 3) / 4) * std::max<uint64_t>(1, (hb + 2ull * kTargetStreamCols - 1) / (2ull * kTargetStreamCols));
        }
        // batches already on the device fill it too: share the target among them
        const uint64_t want = 2ull * warp_slots / std::max<uint32_t>(1, busy_batches.load(std::memory_order_relaxed));
        // First spread each read over more lanes (fewer rows per lane = a cheaper step: the latency bins), then cut the chunks
        // shorter, until the batch has about `want` warps.
        // (only while even the widened batch leaves half of its share of the device empty: the wide layouts do 10-15 % less
        // work per issue slot, so a busy device wants the throughput layout)
        if (lanes_pref)
            min_lanes = (int)lanes_pref;
        else if (est_warps * 8 <= want)
            min_lanes = 32;
        else if (est_warps * 4 <= want)
            min_lanes = 16;
        est_warps *= (uint64_t)min_lanes / 8;
        // ... and a batch runs at 1 / (1 + slots / warps) of the device's rate while its last warps finish alone: below
        // kWavesWanted waves the chunks are cut shorter (a 300-column sweep costs a warp ~2 % in prologue, tools/chunk_curve.py)
        const uint64_t want_chunks = want * (kWavesWanted / 2);
        if (est_warps > 0 && est_warps < want_chunks)
            target_cols = (uint32_t)std::max<uint64_t>(kMinStreamCols, (uint64_t)kTargetStreamCols * est_warps / want_chunks);
    }
    Slot::PlanScratch& ps = s.ps;
    std::vector<PhmmRead>& reads = ps.reads;
    std::vector<PhmmHap>& haps = ps.haps;
    std::vector<PhmmChunk>& chunks = ps.chunks;
    std::vector<PhmmPair>& slow = ps.slow;
    reads.clear();
    haps.clear();
    chunks.clear();
    slow.clear();
    ps.pri.clear();
    reads.reserve(n_reads);
    haps.reserve(n_haps);
    s.regions.clear();
    s.regions.reserve(regs.size());
    s.read_len.clear();
    s.read_len.reserve(n_reads);
    // per bin: (chunk, reads...) groups
    using Group = Slot::Group;
    std::vector<std::vector<Group>>& groups = ps.groups;
    std::vector<std::vector<uint32_t>>& group_reads = ps.group_reads;
    std::vector<std::vector<uint32_t>>& by_bin = ps.by_bin;
    groups.resize(kPhmmNumBins);
    group_reads.resize(kPhmmNumBins);
    by_bin.resize(kPhmmNumBins);
    for (int b = 0; b < kPhmmNumBins; ++b) {
        groups[b].clear();
        group_reads[b].clear();
        by_bin[b].clear();
    }
    std::vector<uint32_t>&order = ps.order, &region_chunks = ps.region_chunks, &idx = ps.idx;
    // read length -> bin, for the lengths the wavefront kernels take (a table instead of a search per read)
    uint8_t bin_lut[640];
    int bin_lo = kPhmmNumBins, bin_hi = -1;        // the bins this batch can use (one variant, lane families from min_lanes up)
    for (uint32_t len = 0; len < 640; ++len) {
        const int b = phmm_bin_of(len, refgaps, min_lanes);
        bin_lut[len] = b < 0 ? 255 : (uint8_t)b;
        if (b >= 0) {
            bin_lo = std::min(bin_lo, b);
            bin_hi = std::max(bin_hi, b);
        }
    }

    uint64_t raw_off = 0, hap_off = 0, out_off = 0;
    uint32_t max_cols = 0;
    [[maybe_unused]] const auto pt1 = PT_NOW();
    PT_ADD(0, pt0, pt1);
    for (const RegionIn& r : regs) {
        RegionPlan rp{};
        rp.read_begin = (uint32_t)reads.size();
        rp.hap_begin = (uint32_t)haps.size();
        rp.out_off = out_off;
        if (r.n_reads == 0 || r.n_haps == 0) {
            rp.n_reads = rp.n_haps = 0;
            s.regions.push_back(rp);
            continue;
        }
        rp.n_reads = r.n_reads;
        rp.n_haps = r.n_haps;
        s.regions.push_back(rp);
        if (s.want_best)
            for (uint32_t j = 0; j < r.n_haps; ++j) ps.pri.push_back(r.haps[j].priority);
        // haplotypes -> chunks of two balanced streams
        for (uint32_t j = 0; j < r.n_haps; ++j) {
            PhmmHap h{};
            h.byte_off = (uint32_t)hap_off;
            h.len = r.haps[j].len;
            h.out_col = j;
            hap_off += h.len;
            max_cols = std::max(max_cols, h.len);
            haps.push_back(h);
        }
        order.clear();
        uint64_t total = 0;
        for (uint32_t j = 0; j < r.n_haps; ++j)
            if (r.haps[j].len > 0 && r.haps[j].len <= kMaxStreamCols) {
                order.push_back(j);
                total += r.haps[j].len;
            }
        const bool all_haps_in_stream = order.size() == r.n_haps;
        region_chunks.clear();
        if (!order.empty()) {
            // never more chunks than pairs of haplotypes: a chunk with a single haplotype leaves its second stream empty
            const uint32_t n_chunks = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1, order.size() / 2),
                                                                 std::max<uint64_t>(1, (total + 2ull * target_cols - 1) / (2ull * target_cols)));
            const uint64_t per_chunk = (total + n_chunks - 1) / n_chunks;
            size_t at = 0;
            while (at < order.size()) {
                // take haplotypes (in region order) until this chunk holds its share
                size_t end = at;
                uint64_t acc = 0;
                while (end < order.size() && (end == at || acc + r.haps[order[end]].len <= per_chunk + target_cols / 4)) {
                    acc += r.haps[order[end]].len;
                    ++end;
                    if (acc >= per_chunk) break;
                }
                // longest first onto the lighter stream
                idx.assign(order.begin() + at, order.begin() + end);
                for (size_t a = 1; a < idx.size(); ++a) {                  // stable insertion sort, longest first (a handful of entries)
                    const uint32_t v = idx[a];
                    size_t at2 = a;
                    while (at2 > 0 && r.haps[idx[at2 - 1]].len < r.haps[v].len) {
                        idx[at2] = idx[at2 - 1];
                        --at2;
                    }
                    idx[at2] = v;
                }
                uint32_t sum[2] = {0, 0}, cnt[2] = {0, 0};
                for (uint32_t j : idx) {
                    const uint32_t e = (sum[1] < sum[0]) ? 1u : 0u;
                    PhmmHap& h = haps[rp.hap_begin + j];
                    h.pos = 32 + sum[e];
                    h.ord_stream = (cnt[e] << 1) | e;
                    sum[e] += h.len;
                    ++cnt[e];
                }
                // the device walks haps[hap_begin .. +n_haps); empty or over-long haplotypes inside that range
                // keep pos == 0, which the kernel skips (they are served by the per-pair kernel)
                PhmmChunk c{};
                c.hap_begin = rp.hap_begin + order[at];
                c.n_haps = order[end - 1] - order[at] + 1;
                c.stream_len = 32 + std::max(sum[0], sum[1]) + 1;
                c.n_x = cnt[0];
                region_chunks.push_back((uint32_t)chunks.size());
                chunks.push_back(c);
                at = end;
            }
        }
        // reads
        for (int b = bin_lo; b <= bin_hi; ++b) by_bin[b].clear();
        for (uint32_t k = 0; k < r.n_reads; ++k) {
            const phmm_read& in = r.reads[k];
            PhmmRead d{};
            d.raw_off = (uint32_t)raw_off;
            d.word_off = (uint32_t)raw_off;
            d.len = in.len;
            d.out_off = (uint32_t)(out_off + (uint64_t)k * r.n_haps);
            d.n_haps = r.n_haps;
            d.hap_begin = rp.hap_begin;
            d.thr = (int16_t)std::max(-32768, std::min(32767, r.thr));
            d.mapq = in.mapq;
            const uint32_t fl = (in.ins ? PHMM_RD_HAS_INS : 0u) | (in.del ? PHMM_RD_HAS_DEL : 0u) | (in.gcp ? PHMM_RD_HAS_GCP : 0u);
            d.model_flags = (uint8_t)((r.pcr & 15) | (fl << 4));
            raw_off += in.len;
            const uint32_t rid = (uint32_t)reads.size();
            reads.push_back(d);
            s.read_len.push_back(in.len);
            const int b = in.len < 640 ? (bin_lut[in.len] == 255 ? -1 : (int)bin_lut[in.len]) : -1;
            if (b >= 0) by_bin[b].push_back(rid);
            if (b < 0 || !all_haps_in_stream)
                for (uint32_t j = 0; j < r.n_haps; ++j) {
                const bool hap_ok = r.haps[j].len > 0 && r.haps[j].len <= kMaxStreamCols;
                if (b < 0 || !hap_ok) {
                    PhmmPair p{};
                    p.read = rid;
                    p.hap = rp.hap_begin + j;
                    p.out_idx = d.out_off + j;
                    slow.push_back(p);
                }
            }
        }
        for (int b0 = bin_lo; b0 <= bin_hi;) {           // one lane family at a time (bins of a family are contiguous)
            int b1 = b0;
            while (b1 + 1 <= bin_hi && kPhmmBins[b1 + 1].lanes == kPhmmBins[b0].lanes && kPhmmBins[b1 + 1].refgaps == kPhmmBins[b0].refgaps) ++b1;
            promote_remainders(by_bin, b0, b1);
            b0 = b1 + 1;
        }
        for (int b = bin_lo; b <= bin_hi; ++b) {
            if (by_bin[b].empty()) continue;
            // every chunk of the region sweeps the same reads: one copy of the list, one group per chunk
            const uint32_t first = (uint32_t)group_reads[b].size();
            group_reads[b].insert(group_reads[b].end(), by_bin[b].begin(), by_bin[b].end());
            for (uint32_t ci : region_chunks) groups[b].push_back(Group{ci, chunks[ci].stream_len, first, (uint32_t)by_bin[b].size()});
        }
        out_off += (uint64_t)r.n_reads * r.n_haps;
    }

    [[maybe_unused]] const auto pt2 = PT_NOW();
    PT_ADD(1, pt1, pt2);
    // CTAs per bin, longest streams first
    std::vector<PhmmCta>& ctas = ps.ctas;
    std::vector<uint32_t>& task_reads = ps.task_reads;
    ctas.clear();
    task_reads.clear();
    s.launches.clear();
    for (int b = std::max(bin_lo, 0); b <= bin_hi; ++b) {
        if (groups[b].empty()) continue;
        {
            // stable, longest stream first: a counting sort on the stream length (at most kMaxStreamCols + 33 values)
            uint32_t longest = 0;
            for (const Group& g : groups[b]) longest = std::max(longest, g.stream_len);
            ps.bucket.assign((size_t)longest + 2, 0u);
            for (const Group& g : groups[b]) ++ps.bucket[longest - g.stream_len + 1];
            for (uint32_t k = 1; k <= longest + 1; ++k) ps.bucket[k] += ps.bucket[k - 1];
            ps.sorted.resize(groups[b].size());
            for (const Group& g : groups[b]) ps.sorted[ps.bucket[longest - g.stream_len]++] = g;
            groups[b].swap(ps.sorted);
        }
        FwdLaunch fl{};
        fl.bin = b;
        fl.cta_base = (uint32_t)ctas.size();
        size_t smem = 0;
        // warps per CTA of this launch: bigger CTAs stage each chunk's streams fewer times and run a little faster, but their
        // read slots come in bigger lumps; take the biggest size that wastes at most 8 % more slots than single-warp CTAs
        uint32_t cta_warps = this->cta_warps;
        if (cta_warps == 0) {
            const uint32_t per_warp = phmm_bin_reads_per_cta(b, 1);
            uint64_t slots[3] = {0, 0, 0};
            for (const Group& g : groups[b])
                for (int k = 0; k < 3; ++k) {
                    const uint32_t per = per_warp << k;
                    slots[k] += (uint64_t)((g.n + per - 1) / per) * per;
                }
            cta_warps = 1;
            for (int k = 1; k < 3; ++k)
                if (slots[k] * 100 <= slots[0] * 108) cta_warps = 1u << k;
        }
        for (const Group& g : groups[b]) {
            smem = std::max(smem, phmm_fwd_smem_bytes(chunks[g.chunk].stream_len, chunks[g.chunk].n_haps, b, cta_warps));
            const uint32_t per_cta = phmm_bin_reads_per_cta(b, cta_warps);
            for (uint32_t at = 0; at < g.n; at += per_cta) {
                PhmmCta c{};
                c.chunk = g.chunk;
                c.read_begin = (uint32_t)task_reads.size();
                c.n_reads = std::min<uint32_t>(per_cta, g.n - at);
                for (uint32_t k = 0; k < c.n_reads; ++k) task_reads.push_back(group_reads[b][g.first + at + k]);
                ctas.push_back(c);
            }
        }
        fl.n_ctas = (uint32_t)ctas.size() - fl.cta_base;
        fl.cta_warps = cta_warps;
        fl.smem = smem;
        s.launches.push_back(fl);
    }

    // biggest launches first: the small ones fill the tails
    std::stable_sort(s.launches.begin(), s.launches.end(), [](const FwdLaunch& x, const FwdLaunch& y) {
        return (uint64_t)x.n_ctas * x.cta_warps * kPhmmBins[x.bin].rpt > (uint64_t)y.n_ctas * y.cta_warps * kPhmmBins[y.bin].rpt;
    });

    [[maybe_unused]] const auto pt3 = PT_NOW();
    PT_ADD(2, pt2, pt3);
    // ---- arena layout ---------------------------------------------------------------------------
    size_t off = 0;
    auto place = [&](size_t bytes) { size_t at = off; off = align_up(off + bytes, 256); return at; };
    const size_t o_reads = place(reads.size() * sizeof(PhmmRead));
    const size_t o_haps = place(haps.size() * sizeof(PhmmHap));
    const size_t o_chunks = place(chunks.size() * sizeof(PhmmChunk));
    const size_t o_ctas = place(ctas.size() * sizeof(PhmmCta));
    const size_t o_tasks = place(task_reads.size() * sizeof(uint32_t));
    const size_t o_slow = place(slow.size() * sizeof(PhmmPair));
    ps.regions_dev.clear();
    if (s.dev_epi)
        for (const RegionPlan& rp : s.regions)
            ps.regions_dev.push_back(PhmmRegionDev{rp.read_begin, rp.n_reads, rp.hap_begin, rp.n_haps, (uint32_t)rp.out_off, {0, 0, 0}});
    const size_t o_regions = place(ps.regions_dev.size() * sizeof(PhmmRegionDev));
    const size_t o_pri = place((s.dev_epi ? ps.pri.size() : 0) * sizeof(double));
    const size_t o_rb = place(read_bytes + 16);
    const size_t o_rq = place(read_bytes + 16);
    const size_t o_ri = any_ins ? place(read_bytes + 16) : 0;
    const size_t o_rd = any_del ? place(read_bytes + 16) : 0;
    const size_t o_rg = any_gcp ? place(read_bytes + 16) : 0;
    const size_t o_hb = place(hap_bytes + 16);
    s.in_bytes = off;
    s.o_reads = o_reads; s.o_haps = o_haps; s.o_chunks = o_chunks; s.o_ctas = o_ctas; s.o_tasks = o_tasks; s.o_slow = o_slow;
    s.n_chunks = chunks.size(); s.n_ctas = ctas.size(); s.n_tasks = task_reads.size();
    int rc;
    if ((rc = s.h_in.ensure(off, &high_water[0])) != PHMM_OK) return rc;
    if ((rc = s.d_in.ensure(off, &high_water[0])) != PHMM_OK) return rc;
    uint8_t* h = static_cast<uint8_t*>(s.h_in.p);
    std::memcpy(h + o_reads, reads.data(), reads.size() * sizeof(PhmmRead));
    std::memcpy(h + o_haps, haps.data(), haps.size() * sizeof(PhmmHap));
    std::memcpy(h + o_chunks, chunks.data(), chunks.size() * sizeof(PhmmChunk));
    std::memcpy(h + o_ctas, ctas.data(), ctas.size() * sizeof(PhmmCta));
    std::memcpy(h + o_tasks, task_reads.data(), task_reads.size() * sizeof(uint32_t));
    std::memcpy(h + o_slow, slow.data(), slow.size() * sizeof(PhmmPair));
    std::memcpy(h + o_regions, ps.regions_dev.data(), ps.regions_dev.size() * sizeof(PhmmRegionDev));
    if (s.dev_epi) std::memcpy(h + o_pri, ps.pri.data(), ps.pri.size() * sizeof(double));
    s.o_regions = o_regions;
    s.o_pri = o_pri;
    [[maybe_unused]] const auto pt4 = PT_NOW();
    PT_ADD(3, pt3, pt4);
    {
        // gather the raw bytes; callers that keep their reads back to back (flat batches, BAM-like buffers) get one
        // memcpy per contiguous run instead of one per read.  When the caller says its buffers are page-locked and the runs
        // are few, nothing is gathered at all: the runs become host-to-device copies of their own (launch()).
        struct Run
        {
            const uint8_t* src = nullptr;
            size_t len = 0, dst = 0;
            std::vector<Slot::DirectCopy>* runs;
            void add(const uint8_t* p, size_t n, size_t at)
            {
                if (src && p == src + len && at == dst + len) {
                    len += n;
                    return;
                }
                flush();
                src = p;
                len = n;
                dst = at;
            }
            void flush()
            {
                if (src && len) runs->push_back(Slot::DirectCopy{dst, src, len});
                src = nullptr;
            }
        };
        std::vector<Slot::DirectCopy> runs;
        Run rb_run{nullptr, 0, 0, &runs}, rq_run = rb_run, ri_run = rb_run, rd_run = rb_run, rg_run = rb_run, hb_run = rb_run;
        bool pinned = true;
        size_t ro = 0, ho = 0;
        for (const RegionIn& r : regs) {
            if (r.n_reads == 0 || r.n_haps == 0) continue;
            pinned &= r.pinned_src;
            for (uint32_t k = 0; k < r.n_reads; ++k) {
                const phmm_read& in = r.reads[k];
                rb_run.add(in.bases, in.len, o_rb + ro);
                rq_run.add(in.qual, in.len, o_rq + ro);
                if (in.ins) ri_run.add(in.ins, in.len, o_ri + ro);
                if (in.del) rd_run.add(in.del, in.len, o_rd + ro);
                if (in.gcp) rg_run.add(in.gcp, in.len, o_rg + ro);
                ro += in.len;
            }
            for (uint32_t j = 0; j < r.n_haps; ++j) {
                hb_run.add(r.haps[j].bases, r.haps[j].len, o_hb + ho);
                ho += r.haps[j].len;
            }
        }
        rb_run.flush(); rq_run.flush(); ri_run.flush(); rd_run.flush(); rg_run.flush(); hb_run.flush();
        s.direct.clear();
        s.staged_bytes = s.in_bytes;
        if (pinned && !s.h_in.host_only && runs.size() <= 24) {
            s.direct = std::move(runs);
            s.staged_bytes = o_rb;                   // records only; the byte blobs follow them in the arena
        } else {
            for (const Slot::DirectCopy& c : runs) std::memcpy(h + c.dst, c.src, c.len);
        }
    }
    PT_ADD(4, pt4, PT_NOW());
    // ---- device-only buffers --------------------------------------------------------------------
    const size_t pairs_alloc = std::max<uint64_t>(n_pairs, 1);
    if ((rc = s.d_words.ensure((read_bytes + 16) * sizeof(uint32_t), &high_water[1])) != PHMM_OK) return rc;
    if ((rc = s.d_f32.ensure(pairs_alloc * sizeof(float), &high_water[2])) != PHMM_OK) return rc;
    if ((rc = s.d_f64.ensure(pairs_alloc * sizeof(double), &high_water[3])) != PHMM_OK) return rc;
    if ((rc = s.d_pairs.ensure(pairs_alloc * sizeof(PhmmPair), &high_water[4])) != PHMM_OK) return rc;
    if ((rc = s.d_cnt.ensure(256)) != PHMM_OK) return rc;
    if ((rc = s.d_scr64.ensure((size_t)(kPairThreads / 32) * (max_cols + 1) * 3 * sizeof(double), &high_water[5])) != PHMM_OK) return rc;
    if (!slow.empty())
        if ((rc = s.d_scr32.ensure((size_t)(kPairThreads / 32) * (max_cols + 1) * 3 * sizeof(float), &high_water[6])) != PHMM_OK) return rc;
    if ((rc = s.h_f32.ensure(pairs_alloc * sizeof(float), &high_water[2])) != PHMM_OK) return rc;
    if ((rc = s.h_f64.ensure(pairs_alloc * sizeof(double), &high_water[3])) != PHMM_OK) return rc;
    if ((rc = s.h_cnt.ensure(256)) != PHMM_OK) return rc;

    if (s.dev_epi) {
        // [lik: one double per pair | kept_idx, best, second: one u32 per read | n_kept, n_under: one u32 per region]
        size_t eo = 0;
        auto eplace = [&](size_t bytes) { size_t at = eo; eo = align_up(eo + bytes, 256); return at; };
        eplace(pairs_alloc * sizeof(double));
        s.e_kept = eplace(std::max<uint64_t>(n_reads, 1) * sizeof(uint32_t));
        s.e_best = s.want_best ? eplace(std::max<uint64_t>(n_reads, 1) * sizeof(uint32_t)) : 0;
        s.e_second = s.want_best ? eplace(std::max<uint64_t>(n_reads, 1) * sizeof(uint32_t)) : 0;
        s.e_nkept = eplace(std::max<size_t>(regs.size(), 1) * sizeof(uint32_t));
        s.e_under = eplace(std::max<size_t>(regs.size(), 1) * sizeof(uint32_t));
        s.epi_bytes = eo;
        if ((rc = s.d_epi.ensure(eo, &high_water[7])) != PHMM_OK) return rc;
        if ((rc = s.h_epi.ensure(eo, &high_water[7])) != PHMM_OK) return rc;
    }

    uint8_t* d = static_cast<uint8_t*>(s.d_in.p);
    PhmmBatchDev& b = s.dev;
    b.reads = reinterpret_cast<const PhmmRead*>(d + o_reads);
    b.haps = reinterpret_cast<const PhmmHap*>(d + o_haps);
    b.chunks = reinterpret_cast<const PhmmChunk*>(d + o_chunks);
    b.ctas = reinterpret_cast<const PhmmCta*>(d + o_ctas);
    b.task_reads = reinterpret_cast<const uint32_t*>(d + o_tasks);
    b.slow_pairs = reinterpret_cast<const PhmmPair*>(d + o_slow);
    b.read_bases = d + o_rb;
    b.read_qual = d + o_rq;
    b.read_ins = any_ins ? d + o_ri : nullptr;
    b.read_del = any_del ? d + o_rd : nullptr;
    b.read_gcp = any_gcp ? d + o_rg : nullptr;
    b.hap_bases = d + o_hb;
    b.words = static_cast<uint32_t*>(s.d_words.p);
    b.out_f32 = static_cast<float*>(s.d_f32.p);
    b.out_f64 = static_cast<double*>(s.d_f64.p);
    b.f64_pairs = static_cast<PhmmPair*>(s.d_pairs.p);
    b.f64_count = static_cast<uint32_t*>(s.d_cnt.p);
    b.tables = static_cast<const PhmmTables*>(d_tables.p);
    b.n_reads = (uint32_t)reads.size();
    b.rg_gapm = default_gapm;      // gap_conts() is the constant '+' - 33 = 10 for every read, read_record.cpp:205-210
    b.rg_ii = default_ii;
    if (s.dev_epi) {
        uint8_t* de = static_cast<uint8_t*>(s.d_epi.p);
        PhmmEpilogueDev& ep = s.epi;
        ep.regions = reinterpret_cast<const PhmmRegionDev*>(d + o_regions);
        ep.reads = b.reads;
        ep.raw = b.out_f32;
        ep.hap_priority = s.want_best ? reinterpret_cast<const double*>(d + o_pri) : nullptr;
        ep.lik = reinterpret_cast<double*>(de);
        ep.kept_idx = reinterpret_cast<uint32_t*>(de + s.e_kept);
        ep.best = s.want_best ? reinterpret_cast<uint32_t*>(de + s.e_best) : nullptr;
        ep.second = s.want_best ? reinterpret_cast<uint32_t*>(de + s.e_second) : nullptr;
        ep.n_kept = reinterpret_cast<uint32_t*>(de + s.e_nkept);
        ep.n_under = reinterpret_cast<uint32_t*>(de + s.e_under);
        ep.n_regions = (uint32_t)regs.size();
    }

    s.n_pairs = n_pairs;
    s.n_cells = n_cells;
    s.n_words = read_bytes;
    s.n_reads = (uint32_t)reads.size();
    s.n_haps = (uint32_t)haps.size();
    s.n_slow = (uint32_t)slow.size();
    s.max_cols = max_cols;
    s.n_launches = 1 + (uint32_t)s.launches.size() + (s.n_slow ? 1 : 0) + (s.want_f64 ? 1 : 0) + (s.dev_epi ? 1 : 0);
    return PHMM_OK;
}

// ------------------------------------------------------------------------------------------------
// Device side of one batch
// ------------------------------------------------------------------------------------------------

int phmm_engine::launch(Slot& s, bool copy_in)
{
    if (copy_in && s.in_bytes) {
        CU(cudaMemcpyAsync(s.d_in.p, s.h_in.p, s.staged_bytes, cudaMemcpyHostToDevice, s.stream));
        for (const Slot::DirectCopy& c : s.direct)
            CU(cudaMemcpyAsync(static_cast<uint8_t*>(s.d_in.p) + c.dst, c.src, c.len, cudaMemcpyHostToDevice, s.stream));
        stats.h2d += s.in_bytes;
    }
    CU(cudaMemsetAsync(s.d_cnt.p, 0, 4, s.stream));
    CU(cudaEventRecord(s.ev[0], s.stream));
    CU(phmm_launch_prep(s.dev, s.stream));
    CU(cudaEventRecord(s.ev[1], s.stream));
    // One launch per read-length bin.  A batch of mixed lengths has many of them, none big enough to fill the GPU to the
    // end of its last wave, so they go out on up to four streams (forked after the prep kernel, joined before the FP64 pass).
    const size_t n_side = std::min<size_t>(Slot::kAux, s.launches.size() > 1 ? s.launches.size() - 1 : 0);
    for (size_t a = 0; a < n_side; ++a) CU(cudaStreamWaitEvent(s.aux[a], s.ev[1], 0));
    for (size_t i = 0; i < s.launches.size(); ++i) {
        const FwdLaunch& fl = s.launches[i];
        const size_t lane = i % (n_side + 1);
        cudaStream_t st = lane == 0 ? s.stream : s.aux[lane - 1];
        const cudaError_t le = phmm_launch_fwd(s.dev, fl.bin, fma, fl.cta_base, fl.n_ctas, fl.cta_warps, fl.smem, st);
        if (le != cudaSuccess)
            return fail(PHMM_ERR_CUDA, std::string("forward kernel launch (") + std::to_string(kPhmmBins[fl.bin].lanes) + " lanes x " +
                                           std::to_string(kPhmmBins[fl.bin].rpt) + " rows, " + std::to_string(fl.n_ctas) + " CTAs, " +
                                           std::to_string(fl.smem) + " B shared): " + cudaGetErrorString(le));
    }
    for (size_t a = 0; a < n_side; ++a) {
        CU(cudaEventRecord(s.aux_done[a], s.aux[a]));
        CU(cudaStreamWaitEvent(s.stream, s.aux_done[a], 0));
    }
    CU(cudaEventRecord(s.ev[2], s.stream));
    if (s.n_slow)
        CU(phmm_launch_pairs_f32(s.dev, fma, s.dev.slow_pairs, s.n_slow, static_cast<float*>(s.d_scr32.p), kPairThreads, s.max_cols,
                                 s.stream));
    s.f64_done = s.want_f64 && s.n_pairs;
    if (s.f64_done)
        CU(phmm_launch_pairs_f64(s.dev, s.dev.f64_pairs, s.dev.f64_count, static_cast<double*>(s.d_scr64.p), kPairThreads, s.max_cols,
                                 s.stream));
    if (s.dev_epi && s.n_pairs) CU(phmm_launch_epilogue(s.epi, s.stream));
    CU(cudaEventRecord(s.ev[3], s.stream));
    stats.launches += s.n_launches;
    return PHMM_OK;
}

// D2H + the reference's host epilogue.
int phmm_engine::finish(Slot& s, const std::vector<RegionIn>& regs, bool copy_out)
{
    uint32_t n64 = 0;
    bool asks_best = false, asks_raw = false, asks_f64 = false;
    for (const RegionIn& r : regs) {
        if (r.n_reads == 0 || r.n_haps == 0) continue;
        asks_best |= r.out.best_hap != nullptr;
        asks_raw |= r.out.raw_f32 != nullptr || r.out.raw_f64 != nullptr;
        asks_f64 |= r.out.raw_f64 != nullptr;
    }
    // a resident batch was planned before anyone said what to download: without best / second-best buffers on the device
    // the host epilogue serves that call
    const bool dev = s.dev_epi && (!asks_best || s.want_best);
    const bool need_f32 = !dev || asks_raw;
    if (copy_out && s.n_pairs) {
        if (need_f32) {
            CU(cudaMemcpyAsync(s.h_f32.p, s.d_f32.p, s.n_pairs * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
            stats.d2h += s.n_pairs * sizeof(float);
        }
        CU(cudaMemcpyAsync(s.h_cnt.p, s.d_cnt.p, 4, cudaMemcpyDeviceToHost, s.stream));
        if (dev) {
            CU(cudaMemcpyAsync(s.h_epi.p, s.d_epi.p, s.epi_bytes, cudaMemcpyDeviceToHost, s.stream));
            stats.d2h += s.epi_bytes;
        }
        stats.d2h += 4;
    }
    {
        // short batches (the synchronous per-region callers) finish in tens of microseconds: poll before paying a blocking
        // sync's wake-up latency
        cudaError_t q = cudaErrorNotReady;
        if (s.n_cells < 1500000000ull) {            // < ~0.5 ms of device work; big batches sleep instead of burning a core
            const auto t_poll = std::chrono::steady_clock::now();
            while ((q = cudaStreamQuery(s.stream)) == cudaErrorNotReady) {
                if (std::chrono::steady_clock::now() - t_poll > std::chrono::microseconds(400)) break;
                _mm_pause();
            }
            if (q != cudaSuccess && q != cudaErrorNotReady) return fail(PHMM_ERR_CUDA, std::string("cudaStreamQuery: ") + cudaGetErrorString(q));
        }
        if (q == cudaErrorNotReady) CU(cudaStreamSynchronize(s.stream));
    }
    // The FP64 sums are read only by a caller that asked for them (raw_f64) or under PHMM_FP64_ALWAYS; see `have64` below.
    bool have64 = s.f64_done;
    if (s.n_pairs) {
        n64 = *static_cast<uint32_t*>(s.h_cnt.p);
        if (n64 > 0 && !have64) {
            if (asks_f64) {                                     // a resident batch learns at download time who wants raw sums
                CU(phmm_launch_pairs_f64(s.dev, s.dev.f64_pairs, s.dev.f64_count, static_cast<double*>(s.d_scr64.p), kPairThreads,
                                         s.max_cols, s.stream));
                s.f64_done = have64 = true;
            }
        }
        if (n64 > 0 && have64) {
            CU(cudaMemcpyAsync(s.h_f64.p, s.d_f64.p, s.n_pairs * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
            CU(cudaStreamSynchronize(s.stream));
            stats.d2h += s.n_pairs * sizeof(double);
        }
    }
    CU(cudaEventElapsedTime(&s.gpu_ms, s.ev[0], s.ev[3]));
    CU(cudaEventElapsedTime(&s.fwd_ms, s.ev[1], s.ev[2]));

    const auto t_epi = std::chrono::steady_clock::now();
    if (dev) {
        epilogue_from_device(s, regs, have64);
        stats.dev_epilogues += 1;
    }
    else
        epilogue(s, regs, static_cast<const float*>(s.h_f32.p), static_cast<const double*>(s.h_f64.p), have64);
    stats.regions += regs.size();
    stats.pairs += s.n_pairs;
    stats.cells += s.n_cells;
    stats.fp64_pairs += n64;
    stats.slow_pairs += s.n_slow;
    stats.batches += 1;
    {
        std::lock_guard<std::mutex> lk2(stats.m);
        stats.gpu_ms += s.gpu_ms;
        stats.epilogue_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_epi).count();
    }
    return PHMM_OK;
}

// The epilogue ran on the device (phmm_epilogue_kernel): hand every region its matrix, surviving reads and best haplotypes.
void phmm_engine::epilogue_from_device(const Slot& s, const std::vector<RegionIn>& regs, bool have64) const
{
    const uint8_t* he = static_cast<const uint8_t*>(s.h_epi.p);
    const double* lik = reinterpret_cast<const double*>(he);
    const uint32_t* kept_idx = reinterpret_cast<const uint32_t*>(he + s.e_kept);
    const uint32_t* best = reinterpret_cast<const uint32_t*>(he + s.e_best);
    const uint32_t* second = reinterpret_cast<const uint32_t*>(he + s.e_second);
    const uint32_t* n_kept = reinterpret_cast<const uint32_t*>(he + s.e_nkept);
    const uint32_t* n_under = reinterpret_cast<const uint32_t*>(he + s.e_under);
    const float* f32 = static_cast<const float*>(s.h_f32.p);
    const double* f64 = static_cast<const double*>(s.h_f64.p);
    for (size_t g = 0; g < regs.size(); ++g) {
        const RegionIn& r = regs[g];
        const RegionPlan& rp = s.regions[g];
        if (r.n_kept_out) *r.n_kept_out = 0;
        if (r.n_fp64_out32) *r.n_fp64_out32 = 0;
        if (rp.n_reads == 0 || rp.n_haps == 0) continue;
        const size_t np = (size_t)rp.n_reads * rp.n_haps;
        const uint32_t kept = n_kept[g];
        std::memcpy(r.out.likelihoods, lik + rp.out_off, np * sizeof(double));
        std::memcpy(r.out.kept_idx, kept_idx + rp.read_begin, kept * sizeof(uint32_t));
        if (r.out.best_hap) {
            std::memcpy(r.out.best_hap, best + rp.read_begin, kept * sizeof(uint32_t));
            std::memcpy(r.out.second_hap, second + rp.read_begin, kept * sizeof(uint32_t));
        }
        if (r.out.raw_f32) std::memcpy(r.out.raw_f32, f32 + rp.out_off, np * sizeof(float));
        if (r.out.raw_f64)
            for (size_t q = 0; q < np; ++q)
                r.out.raw_f64[q] = (f32[rp.out_off + q] < kMinAccepted && have64) ? f64[rp.out_off + q] : (double)NAN;
        if (r.n_kept_out) *r.n_kept_out = kept;
        if (r.n_fp64_out32) *r.n_fp64_out32 = n_under[g];
        if (r.n_fp64_out) __atomic_fetch_add(r.n_fp64_out, (uint64_t)n_under[g], __ATOMIC_RELAXED);
    }
}

// The reference's host side after the kernels (pairhmm_internal.cpp:578-584 conversion, :53-63 normalize, :161-182 filter,
// :184-193 transpose), region by region, from the raw sums in batch order.
void phmm_engine::epilogue(const Slot& s, const std::vector<RegionIn>& regs, const float* f32, const double* f64, bool have64) const
{
    const unsigned old_csr = _mm_getcsr();
    _mm_setcsr(old_csr | 0x8000u);                 // init_native, pairhmm_internal.cpp:21-25
    std::vector<double> lk;
    for (size_t g = 0; g < regs.size(); ++g) {
        const RegionIn& r = regs[g];
        const RegionPlan& rp = s.regions[g];
        if (r.n_kept_out) *r.n_kept_out = 0;
        if (r.n_fp64_out32) *r.n_fp64_out32 = 0;
        if (rp.n_reads == 0 || rp.n_haps == 0) continue;
        const uint32_t nr = rp.n_reads, nh = rp.n_haps;
        const size_t np = (size_t)nr * nh;
        const float* fr = f32 + rp.out_off;
        if (lk.size() < np + 8) lk.resize(np + 8);
        // :583 for the whole region, [read][hap]; the pairs under MIN_ACCEPTED (:578) are redone below
        const uint32_t reg64 = (uint32_t)phmm_log10_block(fr, lk.data(), np, log10_mode);
        if (reg64)
            for (size_t q = 0; q < np; ++q)
                if (fr[q] < kMinAccepted)
                    // :580.  Without the FP64 sum the pair stands at -inf — which is also what the reference as built
                    // computes here (its FP64 kernels return 0).  Either way the value is below -64 and
                    // cannot reach the matrix: the read is dropped (best < -8 / -4) or the entry is capped at best - 4.5.
                    lk[q] = have64 ? log10(f64[rp.out_off + q]) - kLog10InitD : -INFINITY;
        if (r.out.raw_f32) std::memcpy(r.out.raw_f32, fr, np * sizeof(float));
        if (r.out.raw_f64)
            for (size_t q = 0; q < np; ++q) r.out.raw_f64[q] = (fr[q] < kMinAccepted && have64) ? f64[rp.out_off + q] : (double)NAN;
        // normalize :53-63, filter :161-182, transpose :184-193 into the caller's [nh][nr] matrix (:516) with its NaN tail
        const uint32_t kept = phmm_postprocess_region(lk.data(), nr, nh, s.read_len.data() + rp.read_begin, r.out.likelihoods,
                                                      r.out.kept_idx, log10_mode);
        if (r.out.best_hap) {
            // the haplotypes' priorities, contiguous for the search
            double pri[128];
            std::vector<double> big;
            double* pp = pri;
            if (nh > 128) {
                big.resize(nh);
                pp = big.data();
            }
            for (uint32_t j = 0; j < nh; ++j) pp[j] = r.haps[j].priority;
            phmm_best_haplotypes(r.out.likelihoods, nr, nh, kept, pp, r.out.best_hap, r.out.second_hap);
        }
        if (r.n_kept_out) *r.n_kept_out = kept;
        if (r.n_fp64_out32) *r.n_fp64_out32 = reg64;
        if (r.n_fp64_out) __atomic_fetch_add(r.n_fp64_out, (uint64_t)reg64, __ATOMIC_RELAXED);
    }
    _mm_setcsr(old_csr);
}

int phmm_engine::execute(Slot& s, const std::vector<RegionIn>& regs)
{
    int rc;
    const auto t0 = std::chrono::steady_clock::now();
    busy_batches.fetch_add(1, std::memory_order_relaxed);
    if ((rc = plan(s, regs)) != PHMM_OK) {
        busy_batches.fetch_sub(1, std::memory_order_relaxed);
        return rc;
    }
    const auto t1 = std::chrono::steady_clock::now();
    if ((rc = launch(s, true)) != PHMM_OK) {
        busy_batches.fetch_sub(1, std::memory_order_relaxed);
        return rc;
    }
    rc = finish(s, regs, true);
    busy_batches.fetch_sub(1, std::memory_order_relaxed);
    {
        std::lock_guard<std::mutex> lk(stats.m);
        stats.plan_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Workers
// ------------------------------------------------------------------------------------------------

void phmm_engine::worker_main(uint32_t idx)
{
    cudaSetDevice(device);
    Slot& s = *slots[idx];
    for (;;) {
        std::function<void(Slot&)> job;
        {
            std::unique_lock<std::mutex> lk(qm);
            qcv.wait(lk, [&] { return stopping || !jobs.empty(); });
            if (jobs.empty()) return;
            job = std::move(jobs.front());
            jobs.pop_front();
        }
        job(s);
    }
}

void phmm_engine::enqueue(std::function<void(Slot&)> fn)
{
    {
        std::lock_guard<std::mutex> lk(qm);
        jobs.push_back(std::move(fn));
    }
    qcv.notify_one();
}

// bm must be held
void phmm_engine::seal_locked()
{
    if (open.regs.empty()) return;
    auto regs = std::make_shared<std::vector<RegionIn>>(std::move(open.regs));
    std::shared_ptr<BatchSync> sync = batch_state[open.id].sync;
    open = OpenBatch{};
    ++in_flight;
    open_cv.notify_all();
    enqueue([this, regs, sync](Slot& s) {
        const int rc = execute(s, *regs);
        const std::string err = rc != PHMM_OK ? g_err : std::string();
        {
            std::lock_guard<std::mutex> lk(sync->m);
            sync->done = true;
            sync->rc = rc;
            sync->err = err;
        }
        sync->cv.notify_all();
        std::lock_guard<std::mutex> lk(bm);
        --in_flight;
        if (open.waiters > 0 && in_flight < pipeline_depth) seal_locked();   // a stage is free: send what has gathered meanwhile
    });
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------

// Regions [g0, g1) of a flat batch as RegionIn records (pointers into the caller's blobs).  reads/haps are resized to the
// range; region g's reads start at reads[reg_read_off[g] - reg_read_off[g0]].
static void flat_to_regions(const phmm_flat_batch* fb, phmm_flat_result* fr, uint32_t g0, uint32_t g1, std::vector<phmm_read>& reads,
                            std::vector<phmm_hap>& haps, std::vector<RegionIn>& regs)
{
    const uint64_t rbeg = fb->reg_read_off[g0], hbeg = fb->reg_hap_off[g0];
    const uint64_t nr = fb->reg_read_off[g1] - rbeg, nh = fb->reg_hap_off[g1] - hbeg;
    reads.resize(nr);
    haps.resize(nh);
    for (uint64_t k = 0; k < nr; ++k) {
        phmm_read& r = reads[k];
        const uint64_t o = fb->read_off[rbeg + k];
        r.bases = fb->read_bases + o;
        r.qual = fb->read_qual + o;
        r.ins = fb->read_ins ? fb->read_ins + o : nullptr;
        r.del = fb->read_del ? fb->read_del + o : nullptr;
        r.gcp = fb->read_gcp ? fb->read_gcp + o : nullptr;
        r.len = fb->read_len[rbeg + k];
        r.mapq = fb->read_mapq[rbeg + k];
    }
    for (uint64_t j = 0; j < nh; ++j) {
        haps[j].bases = fb->hap_bases + fb->hap_off[hbeg + j];
        haps[j].len = fb->hap_len[hbeg + j];
        haps[j].priority = fb->hap_priority ? fb->hap_priority[hbeg + j] : 0.0;
    }
    regs.resize(g1 - g0);
    for (uint32_t g = g0; g < g1; ++g) {
        RegionIn& r = regs[g - g0];
        const uint64_t r0 = fb->reg_read_off[g], h0 = fb->reg_hap_off[g];
        r.reads = reads.data() + (r0 - rbeg);
        r.n_reads = (uint32_t)(fb->reg_read_off[g + 1] - r0);
        r.haps = haps.data() + (h0 - hbeg);
        r.n_haps = (uint32_t)(fb->reg_hap_off[g + 1] - h0);
        r.thr = fb->min_quality_threshold;
        r.pcr = fb->pcr_model;
        r.pinned_src = (fb->flags & PHMM_FLAT_PINNED) != 0;
        r.out = phmm_result{};
        r.n_kept_out = nullptr;
        r.n_fp64_out = nullptr;
        r.n_fp64_out32 = nullptr;
        if (fr) {
            r.out.likelihoods = fr->likelihoods + fr->lik_off[g];
            r.out.kept_idx = fr->kept_idx + r0;
            r.out.raw_f32 = fr->raw_f32 ? fr->raw_f32 + fr->lik_off[g] : nullptr;
            r.out.raw_f64 = fr->raw_f64 ? fr->raw_f64 + fr->lik_off[g] : nullptr;
            r.out.best_hap = fr->best_hap && fr->second_hap ? fr->best_hap + r0 : nullptr;
            r.out.second_hap = fr->best_hap && fr->second_hap ? fr->second_hap + r0 : nullptr;
            r.n_kept_out = fr->n_kept + g;
            r.n_fp64_out = &fr->n_fp64;
        }
    }
}

// Work of regions [0, n) of a flat batch as a running sum, in cell units: the cells of a region (device time) plus a
// per-pair charge for the host's share (planning + epilogue cost about as much per pair as 16 K cells cost the GPU), so that
// streams of small regions are cut into enough sub-batches to keep the host workers and the GPU overlapped.
constexpr uint64_t kPairWeight = 16384;
static std::vector<uint64_t> flat_work_prefix(const phmm_flat_batch* fb)
{
    std::vector<uint64_t> pre(fb->n_regions + 1, 0);
    for (uint32_t g = 0; g < fb->n_regions; ++g) {
        uint64_t rb = 0, hb = 0;
        for (uint64_t k = fb->reg_read_off[g]; k < fb->reg_read_off[g + 1]; ++k) rb += fb->read_len[k];
        for (uint64_t j = fb->reg_hap_off[g]; j < fb->reg_hap_off[g + 1]; ++j) hb += fb->hap_len[j];
        const uint64_t pairs = (fb->reg_read_off[g + 1] - fb->reg_read_off[g]) * (fb->reg_hap_off[g + 1] - fb->reg_hap_off[g]);
        pre[g + 1] = pre[g] + rb * hb + pairs * kPairWeight;
    }
    return pre;
}

// phmm_run_flat's cut of a stream of regions (pre = prefix sums of their work, flat_work_prefix) into sub-batches [first, end).
static std::vector<std::pair<uint32_t, uint32_t>> flat_partition(const std::vector<uint64_t>& pre, uint32_t n_regions, uint64_t batch_cells,
                                                                 uint32_t n_workers)
{
    const uint64_t total = pre[n_regions];
    // sub-batch size: batch_cells at most, and small enough that every worker gets about two of them (host-bound streams of
    // small regions need the planning and the epilogues spread over all workers), but not below a quarter of batch_cells
    // (mixed-length batches much smaller than that run their per-bin kernels too small to fill the device)
    const uint64_t full = std::max<uint64_t>(1, std::min<uint64_t>(batch_cells, std::max<uint64_t>(total / (2ull * std::max(n_workers, 1u)), batch_cells >> 2)));
    std::vector<std::pair<uint32_t, uint32_t>> parts;
    uint32_t g = 0;
    int ramp = 2;                                    // next sizes: full/4, full/2, then full
    while (g < n_regions) {
        const uint64_t left = total - pre[g];
        uint64_t want = ramp > 0 ? (full >> ramp) : full;
        if (ramp > 0) --ramp;
        // ramp down: never take more than half of what is left while more than full/4 remains
        if (left > (full >> 2)) want = std::min<uint64_t>(want, std::max<uint64_t>(left / 2, full >> 2));
        uint32_t end = g + 1;
        while (end < n_regions && pre[end] - pre[g] < want) ++end;
        parts.emplace_back(g, end);
        g = end;
    }
    return parts;
}

static uint64_t region_cells(const RegionIn& r)
{
    uint64_t rb = 0, hb = 0;
    for (uint32_t k = 0; k < r.n_reads; ++k) rb += r.reads[k].len;
    for (uint32_t j = 0; j < r.n_haps; ++j) hb += r.haps[j].len;
    return rb * hb;
}

extern "C" {

const char* phmm_last_error(void) { return g_err.c_str(); }
const char* phmm_version(void) { return "rovaca_b200 pairhmm 0.1 (sm_100a)"; }

int phmm_create(const phmm_config* cfg, phmm_engine** out)
{
    if (!out) return fail(PHMM_ERR_ARG, "phmm_create: out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(PHMM_ERR_CUDA, std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") +
                                       " (this engine has no CPU fallback)");
    std::unique_ptr<phmm_engine> e(new phmm_engine());
    int dev = cfg ? cfg->device : -1;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(PHMM_ERR_ARG, "phmm_create: device ordinal out of range");
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(PHMM_ERR_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                       "; this library carries sm_100a code only");
    e->device = dev;
    e->warp_slots = (uint32_t)prop.multiProcessorCount * 8;
    // host workers: one rank per GPU is the deployment, so a rank may count on its share of the cores
    e->n_workers = std::max(2u, std::min(6u, std::thread::hardware_concurrency() / (unsigned)std::max(ndev, 1)));
    if (const char* env = std::getenv("PHMM_B200_WORKERS"))          // deployment override of the default only
        if (std::atoi(env) > 0) e->n_workers = std::min(32, std::atoi(env));
    if (cfg) {
        e->fma = cfg->fma_mode != PHMM_FMA_NONE;
        if (cfg->n_workers) e->n_workers = std::min<uint32_t>(cfg->n_workers, 32);
        if (cfg->batch_cells) e->batch_cells = cfg->batch_cells;
        if (cfg->linger_us) e->linger_us = cfg->linger_us;
        if (cfg->cta_warps) e->cta_warps = std::min<uint32_t>(cfg->cta_warps, PHMM_WARPS);
        if (cfg->pipeline_depth) e->pipeline_depth = std::min<uint32_t>(cfg->pipeline_depth, 32);
        if (cfg->lanes_per_read == 8 || cfg->lanes_per_read == 16 || cfg->lanes_per_read == 32) e->lanes_pref = cfg->lanes_per_read;
        e->fp64_always = cfg->fp64_mode == PHMM_FP64_ALWAYS;
    }
    std::unique_ptr<PhmmTables> t(new PhmmTables());
    phmm_build_tables(t.get());
    int rc = e->d_tables.ensure(sizeof(PhmmTables));
    if (rc != PHMM_OK) return rc;
    CU(cudaMemcpy(e->d_tables.p, t.get(), sizeof(PhmmTables), cudaMemcpyHostToDevice));
    e->log10_mode = phmm_log10_pick_mode();
    {
        const uint32_t want = cfg ? cfg->epilogue : PHMM_EPILOGUE_AUTO;
        const bool pinned_libm = std::getenv("PHMM_B200_LIBM_LOG10F") != nullptr;
        const bool ok = want != PHMM_EPILOGUE_HOST && !pinned_libm && phmm_log10_restated_matches_libm();
        if (want == PHMM_EPILOGUE_DEVICE && !ok)
            return fail(PHMM_ERR_ARG, "PHMM_EPILOGUE_DEVICE: the restated log10f does not give this host's libm bits (or PHMM_B200_LIBM_LOG10F is set)");
        e->device_epilogue = ok;
        e->device_epilogue_min_pairs = want == PHMM_EPILOGUE_DEVICE ? 0 : 32768;
    }
    e->default_gapm = t->gapm[10];
    e->default_ii = t->ph2pr[10];
    for (uint32_t i = 0; i < e->n_workers; ++i) {
        e->slots.emplace_back(new Slot());
        if ((rc = e->slots.back()->init()) != PHMM_OK) return rc;
    }
    for (uint32_t i = 0; i < e->n_workers; ++i) e->workers.emplace_back([p = e.get(), i] { p->worker_main(i); });
    *out = e.release();
    return PHMM_OK;
}

void phmm_destroy(phmm_engine* e)
{
    if (!e) return;
    {
        std::lock_guard<std::mutex> lk(e->bm);
        e->seal_locked();
    }
    {
        std::lock_guard<std::mutex> lk(e->qm);
        e->stopping = true;
    }
    e->qcv.notify_all();
    for (auto& t : e->workers) t.join();
    cudaSetDevice(e->device);
    for (auto& s : e->slots) s->destroy();
    for (auto& ss : e->sw.slot) {
        for (auto* b : {&ss.h_in, &ss.h_out}) b->release();
        for (auto* b : {&ss.d_in, &ss.d_bt, &ss.d_aux, &ss.d_cig, &ss.d_out}) b->release();
        for (auto& ev : ss.ev)
            if (ev) cudaEventDestroy(ev);
        if (ss.stream) cudaStreamDestroy(ss.stream);
    }
    if (e->sw.base) cudaEventDestroy(e->sw.base);
    e->d_tables.release();
    delete e;
}

int phmm_submit(phmm_engine* e, const phmm_region* region, phmm_result* result, uint64_t* ticket)
{
    if (!e || !region || !result || !ticket) return fail(PHMM_ERR_ARG, "phmm_submit: null argument");
    if (region->n_reads && region->n_haps && (!result->likelihoods || !result->kept_idx))
        return fail(PHMM_ERR_ARG, "phmm_submit: result buffers are null");
    if ((result->best_hap != nullptr) != (result->second_hap != nullptr))
        return fail(PHMM_ERR_ARG, "phmm_submit: best_hap and second_hap go together");
    RegionIn r{};
    r.reads = region->reads;
    r.n_reads = region->n_reads;
    r.haps = region->haps;
    r.n_haps = region->n_haps;
    r.thr = region->min_quality_threshold;
    r.pcr = region->pcr_model;
    r.out = *result;
    r.n_kept_out = &result->n_kept;
    r.n_fp64_out32 = &result->n_fp64;
    const uint64_t cells = region_cells(r);
    std::lock_guard<std::mutex> lk(e->bm);
    if (e->open.regs.empty()) {
        e->open.id = e->next_batch_id++;
        e->open.first = std::chrono::steady_clock::now();
        e->batch_state[e->open.id].sync = std::make_shared<phmm_engine::BatchSync>();
    }
    *ticket = e->open.id;
    e->batch_state[e->open.id].waits += 1;
    e->open.regs.push_back(r);
    e->open.cells += cells;
    if (e->open.cells >= e->batch_cells) e->seal_locked();
    return PHMM_OK;
}

int phmm_flush(phmm_engine* e)
{
    if (!e) return fail(PHMM_ERR_ARG, "phmm_flush: null engine");
    std::lock_guard<std::mutex> lk(e->bm);
    e->seal_locked();
    return PHMM_OK;
}

int phmm_wait(phmm_engine* e, uint64_t ticket)
{
    if (!e) return fail(PHMM_ERR_ARG, "phmm_wait: null engine");
    std::shared_ptr<phmm_engine::BatchSync> sync;
    {
        std::unique_lock<std::mutex> lk(e->bm);
        auto it = e->batch_state.find(ticket);
        if (it == e->batch_state.end()) return fail(PHMM_ERR_STATE, "phmm_wait: unknown ticket");
        sync = it->second.sync;
        if (e->open.id == ticket && !e->open.regs.empty()) {
            if (e->in_flight < e->pipeline_depth) {
                e->seal_locked();                                   // the pipeline is starving: go now
            } else {
                // busy: keep collecting returning callers; a completing batch, the size limit or the deadline seals it
                ++e->open.waiters;
                const auto deadline = e->open.first + std::chrono::microseconds(e->linger_us);
                while (e->open.id == ticket && !e->open.regs.empty()) {
                    if (std::chrono::steady_clock::now() >= deadline) {
                        e->seal_locked();
                        break;
                    }
                    e->open_cv.wait_until(lk, deadline);
                }
            }
        }
    }
    {
        std::unique_lock<std::mutex> lk(sync->m);
        sync->cv.wait(lk, [&] { return sync->done; });
    }
    const int rc = sync->rc;
    if (rc != PHMM_OK) g_err = sync->err;
    std::lock_guard<std::mutex> lk(e->bm);
    auto it = e->batch_state.find(ticket);
    if (it != e->batch_state.end()) {
        if (it->second.waits <= 1) e->batch_state.erase(it); else it->second.waits -= 1;
    }
    return rc;
}

int phmm_call_region(phmm_engine* e, const phmm_region* region, phmm_result* result)
{
    uint64_t t = 0;
    int rc = phmm_submit(e, region, result, &t);
    if (rc != PHMM_OK) return rc;
    return phmm_wait(e, t);
}

int phmm_run_flat(phmm_engine* e, const phmm_flat_batch* fb, phmm_flat_result* fr)
{
    if (!e || !fb || !fr) return fail(PHMM_ERR_ARG, "phmm_run_flat: null argument");
    fr->n_fp64 = 0;
    // Cut the stream into sub-batches that the workers plan, ship, run and finish independently, so that packing, H2D,
    // kernels, D2H and the host epilogue of different sub-batches overlap.  Sizes ramp up from batch_cells / 8 and down
    // again at the end: the GPU starts after one small plan and only a small epilogue is left when it finishes.
    const std::vector<std::pair<uint32_t, uint32_t>> parts = flat_partition(flat_work_prefix(fb), fb->n_regions, e->batch_cells, e->n_workers);
    std::mutex m;
    std::condition_variable cv;
    size_t remaining = parts.size();
    int rc_all = PHMM_OK;
    std::string err_all;
    for (const auto& part : parts) {
        e->enqueue([&, part](Slot& s) {
            std::vector<phmm_read> reads;
            std::vector<phmm_hap> haps;
            std::vector<RegionIn> regs;
            flat_to_regions(fb, fr, part.first, part.second, reads, haps, regs);
            const int rc = e->execute(s, regs);
            std::lock_guard<std::mutex> lk(m);
            if (rc != PHMM_OK && rc_all == PHMM_OK) {
                rc_all = rc;
                err_all = g_err;
            }
            if (--remaining == 0) cv.notify_all();
        });
    }
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return remaining == 0; });
    }
    if (rc_all != PHMM_OK) g_err = err_all;
    return rc_all;
}

int phmm_marginalize(const double* values, uint32_t n_old, uint32_t stride, uint32_t n_evidence, const int32_t* old_to_new,
                     uint32_t n_new, double* out, uint32_t out_stride)
{
    if ((n_old && (!values || !old_to_new)) || (n_new && !out)) return fail(PHMM_ERR_ARG, "phmm_marginalize: null argument");
    if (n_evidence > stride || n_evidence > out_stride) return fail(PHMM_ERR_ARG, "phmm_marginalize: more evidence than columns");
    for (uint32_t n = 0; n < n_new; ++n)
        for (uint32_t e = 0; e < n_evidence; ++e) out[(size_t)n * out_stride + e] = -INFINITY;     // :714
    for (uint32_t a = 0; a < n_old; ++a) {
        const int32_t n = old_to_new[a];
        if (n < 0) continue;                                                                       // s_missing_index, :717
        if ((uint32_t)n >= n_new) return fail(PHMM_ERR_ARG, "phmm_marginalize: old_to_new names an allele past n_new");
        const double* in = values + (size_t)a * stride;
        double* o = out + (size_t)n * out_stride;
        for (uint32_t e = 0; e < n_evidence; ++e)
            if (in[e] > o[e]) o[e] = in[e];                                                        // :723-726
    }
    return PHMM_OK;
}

int phmm_resident_upload(phmm_engine* e, const phmm_flat_batch* fb, phmm_resident** out)
{
    if (!e || !fb || !out) return fail(PHMM_ERR_ARG, "phmm_resident_upload: null argument");
    CU(cudaSetDevice(e->device));
    std::unique_ptr<phmm_resident> r(new phmm_resident());
    int rc = r->slot.init();
    if (rc != PHMM_OK) return rc;
    flat_to_regions(fb, nullptr, 0, fb->n_regions, r->reads, r->haps, r->regs);
    if ((rc = e->plan(r->slot, r->regs)) != PHMM_OK) return rc;
    CU(cudaMemcpyAsync(r->slot.d_in.p, r->slot.h_in.p, r->slot.staged_bytes, cudaMemcpyHostToDevice, r->slot.stream));
    for (const Slot::DirectCopy& c : r->slot.direct)
        CU(cudaMemcpyAsync(static_cast<uint8_t*>(r->slot.d_in.p) + c.dst, c.src, c.len, cudaMemcpyHostToDevice, r->slot.stream));
    CU(cudaStreamSynchronize(r->slot.stream));
    e->stats.h2d += r->slot.in_bytes;
    *out = r.release();
    return PHMM_OK;
}

int phmm_resident_run(phmm_engine* e, phmm_resident* r, float* gpu_ms, float* fwd_ms)
{
    if (!e || !r) return fail(PHMM_ERR_ARG, "phmm_resident_run: null argument");
    CU(cudaSetDevice(e->device));
    int rc = e->launch(r->slot, false);
    if (rc != PHMM_OK) return rc;
    CU(cudaStreamSynchronize(r->slot.stream));
    CU(cudaEventElapsedTime(&r->slot.gpu_ms, r->slot.ev[0], r->slot.ev[3]));
    CU(cudaEventElapsedTime(&r->slot.fwd_ms, r->slot.ev[1], r->slot.ev[2]));
    if (gpu_ms) *gpu_ms = r->slot.gpu_ms;
    if (fwd_ms) *fwd_ms = r->slot.fwd_ms;
    return PHMM_OK;
}

int phmm_resident_download(phmm_engine* e, phmm_resident* r, phmm_flat_result* fr)
{
    if (!e || !r || !fr) return fail(PHMM_ERR_ARG, "phmm_resident_download: null argument");
    CU(cudaSetDevice(e->device));
    fr->n_fp64 = 0;
    for (size_t g = 0; g < r->regs.size(); ++g) {
        RegionIn& x = r->regs[g];
        const uint64_t r0 = (uint64_t)(x.reads - r->reads.data());
        x.out = phmm_result{};
        x.out.likelihoods = fr->likelihoods + fr->lik_off[g];
        x.out.kept_idx = fr->kept_idx + r0;
        x.out.raw_f32 = fr->raw_f32 ? fr->raw_f32 + fr->lik_off[g] : nullptr;
        x.out.raw_f64 = fr->raw_f64 ? fr->raw_f64 + fr->lik_off[g] : nullptr;
        x.out.best_hap = fr->best_hap && fr->second_hap ? fr->best_hap + r0 : nullptr;
        x.out.second_hap = fr->best_hap && fr->second_hap ? fr->second_hap + r0 : nullptr;
        x.n_kept_out = fr->n_kept + g;
        x.n_fp64_out = &fr->n_fp64;
    }
    return e->finish(r->slot, r->regs, true);
}

void phmm_resident_free(phmm_engine* e, phmm_resident* r)
{
    if (!r) return;
    if (e) cudaSetDevice(e->device);
    r->slot.destroy();
    delete r;
}

uint32_t phmm_resident_launches(const phmm_resident* r) { return r ? r->slot.n_launches : 0; }

// ------------------------------------------------------------------------------------------------
// Batched Smith-Waterman (include/phmm_b200.h; kernels in sw_kernels.cu)
// ------------------------------------------------------------------------------------------------

int phmm_sw_align(phmm_engine* e, const phmm_sw_pair* pairs, uint32_t n_pairs, const phmm_sw_params* params, phmm_sw_result* results)
{
    if (!e || !params || (n_pairs && (!pairs || !results))) return fail(PHMM_ERR_ARG, "phmm_sw_align: null argument");
    if (n_pairs == 0) return PHMM_OK;
    constexpr uint32_t kGiveUp = 2048;                                  // MAX_SEQ_LEN, sw_avx.c:17,494-499
    uint32_t cigar_cap = 0;
    for (uint32_t k = 0; k < n_pairs; ++k) {
        if (pairs[k].ref_len == 0 || pairs[k].read_len == 0 || !pairs[k].ref || !pairs[k].read)
            return fail(PHMM_ERR_ARG, "phmm_sw_align: empty sequence");
        if (results[k].cigar_cap && !results[k].cigar) return fail(PHMM_ERR_ARG, "phmm_sw_align: cigar buffer is null");
        cigar_cap = std::max(cigar_cap, results[k].cigar_cap);
    }
    cigar_cap = std::max(cigar_cap, 1u);
    phmm_engine::SwContext& c = e->sw;
    std::lock_guard<std::mutex> lk(c.m);
    CU(cudaSetDevice(e->device));
    for (auto& ss : c.slot)
        if (!ss.stream) {
            CU(cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking));
            for (auto& ev : ss.ev) CU(cudaEventCreate(&ev));
        }
    if (!c.base) CU(cudaEventCreate(&c.base));
    CU(cudaEventRecord(c.base, c.slot[0].stream));
    c.last_ms = 0.f;
    c.covered_ms = 0.f;
    c.last_cells = 0;
    constexpr size_t kBtBudget = (size_t)6 << 30;                        // back-track words one sub-batch may hold on the device
    constexpr uint32_t kSubBatch = 32768;                                // alignments per sub-batch: several per big call, so that packing,
                                                                         // kernels and the hand-back of results overlap

    // results of a finished sub-batch -> the caller's buffers
    auto finish = [&](phmm_engine::SwSlot& ss) -> int {
        if (!ss.busy) return PHMM_OK;
        ss.busy = false;
        CU(cudaStreamSynchronize(ss.stream));
        // kernel time of the call = the union of the sub-batches' kernel spans (two of them can be on the device at once)
        float t0 = 0.f, t1 = 0.f;
        CU(cudaEventElapsedTime(&t0, c.base, ss.ev[0]));
        CU(cudaEventElapsedTime(&t1, c.base, ss.ev[1]));
        if (t1 > c.covered_ms) {
            c.last_ms += t1 - std::max(t0, c.covered_ms);
            c.covered_ms = t1;
        }
        c.last_cells += ss.cells;
        const uint32_t n = ss.n;
        const uint8_t* ho = static_cast<const uint8_t*>(ss.h_out.p);
        const PhmmSwOut* outs = reinterpret_cast<const PhmmSwOut*>(ho);
        const uint32_t* cig = reinterpret_cast<const uint32_t*>(ho + (size_t)n * sizeof(PhmmSwOut));
        for (uint32_t k = 0; k < n; ++k) {
            phmm_sw_result& r = results[ss.src[k]];
            r.alignment_offset = outs[k].alignment_offset;
            r.score = outs[k].score;
            r.status = outs[k].status;
            r.n_cigar = outs[k].n_cigar;
            if (outs[k].n_cigar > r.cigar_cap) {                          // the device buffer was sized for the largest caller buffer
                r.status = 2;
                r.n_cigar = r.cigar_cap;
            }
            if (r.n_cigar) std::memcpy(r.cigar, cig + (size_t)k * ss.cigar_cap, (size_t)r.n_cigar * 4);
        }
        e->stats.launches += 2;
        e->stats.h2d += ss.in_bytes;
        e->stats.d2h += (size_t)n * (sizeof(PhmmSwOut) + (size_t)ss.cigar_cap * 4);
        return PHMM_OK;
    };

    uint32_t at = 0, turn = 0;
    int rc = PHMM_OK;
    while (at < n_pairs && rc == PHMM_OK) {
        phmm_engine::SwSlot& ss = c.slot[turn++ & 1];
        if ((rc = finish(ss)) != PHMM_OK) break;                          // the sub-batch before last: its kernels ran while the last one was packed
        std::vector<PhmmSwTask>& tasks = ss.tasks;
        std::vector<uint32_t>& src = ss.src;
        tasks.clear();
        src.clear();
        struct SwCopy { size_t dst; const uint8_t* src; size_t len; };
        std::vector<SwCopy> copies;                                     // what to gather into the byte blob
        constexpr int kRefCache = 16;
        struct RefSeen { const uint8_t* ptr; uint32_t len, off; } ref_cache[kRefCache] = {};
        int ref_next = 0;
        uint64_t cells = 0;
        size_t bytes = 0, bt_words = 0, aux_ints = 0;
        while (at < n_pairs && tasks.size() < kSubBatch) {
            const phmm_sw_pair& p = pairs[at];
            phmm_sw_result& r = results[at];
            if (std::max(p.ref_len, p.read_len) >= kGiveUp) {            // the reference returns alignment_offset = 1, no CIGAR
                r.n_cigar = 0; r.alignment_offset = 1; r.score = 0; r.status = 1;
                ++at;
                continue;
            }
            int cpt = 0;
            for (int q = 0; q < kPhmmSwNumCpt; ++q)
                if (p.read_len <= 32u * (uint32_t)kPhmmSwCpt[q]) { cpt = kPhmmSwCpt[q]; break; }
            const size_t words = ((size_t)p.ref_len + 31) * 32 * ((cpt + 7) / 8);     // one line per wavefront step
            if (!tasks.empty() && (bt_words + words) * 4 > kBtBudget) break;
            PhmmSwTask t{};
            // a region's reads meet the same few haplotypes again and again: ship each distinct one once (a small cache of
            // the latest haplotype buffers is enough, the callers walk region by region)
            int hit = -1;
            for (int q = 0; q < kRefCache; ++q)
                if (ref_cache[q].ptr == p.ref && ref_cache[q].len == p.ref_len) { hit = q; break; }
            if (hit >= 0) {
                t.ref_off = ref_cache[hit].off;
            } else {
                t.ref_off = (uint32_t)bytes;
                copies.push_back(SwCopy{bytes, p.ref, p.ref_len});
                bytes += p.ref_len;
                ref_cache[ref_next] = RefSeen{p.ref, p.ref_len, t.ref_off};
                ref_next = (ref_next + 1) % kRefCache;
            }
            t.ref_len = p.ref_len;
            t.read_off = (uint32_t)bytes; t.read_len = p.read_len;
            if (!copies.empty() && copies.back().src + copies.back().len == p.read && copies.back().dst + copies.back().len == bytes)
                copies.back().len += p.read_len;                         // reads that sit back to back in the caller's memory
            else
                copies.push_back(SwCopy{bytes, p.read, p.read_len});
            bytes += p.read_len;
            t.bt_off = bt_words; bt_words += words;
            t.aux_off = aux_ints; aux_ints += (size_t)p.ref_len + p.read_len;
            t.cpt = (uint32_t)cpt;
            if (bytes >= ((size_t)1 << 32)) { rc = fail(PHMM_ERR_ARG, "phmm_sw_align: more than 4 GiB of sequence in one sub-batch"); break; }
            tasks.push_back(t);
            src.push_back(at);
            cells += (uint64_t)p.ref_len * p.read_len;
            ++at;
        }
        if (rc != PHMM_OK) break;
        const uint32_t n = (uint32_t)tasks.size();
        if (n == 0) continue;
        // order by cpt (one fill launch per kernel instantiation): a counting sort, input order kept inside a group
        std::vector<uint32_t>& order = ss.order;
        order.resize(n);
        {
            uint32_t count[kPhmmSwNumCpt + 1] = {};
            auto group_of = [&](uint32_t cpt) { int q = 0; while (kPhmmSwCpt[q] != (int)cpt) ++q; return q; };
            for (uint32_t k = 0; k < n; ++k) ++count[group_of(tasks[k].cpt) + 1];
            for (int q = 0; q < kPhmmSwNumCpt; ++q) count[q + 1] += count[q];
            for (uint32_t k = 0; k < n; ++k) order[count[group_of(tasks[k].cpt)]++] = k;
        }
        size_t off = 0;
        auto place = [&](size_t nbytes) { size_t o = off; off = align_up(off + nbytes, 256); return o; };
        const size_t o_tasks = place(n * sizeof(PhmmSwTask)), o_order = place(n * sizeof(uint32_t)), o_bytes = place(bytes + 16);
        if ((rc = ss.h_in.ensure(off)) != PHMM_OK) break;
        if ((rc = ss.d_in.ensure(off)) != PHMM_OK) break;
        if ((rc = ss.d_bt.ensure(std::max<size_t>(bt_words, 1) * 4)) != PHMM_OK) break;
        if ((rc = ss.d_aux.ensure(std::max<size_t>(aux_ints, 1) * 4)) != PHMM_OK) break;
        if ((rc = ss.d_cig.ensure((size_t)n * cigar_cap * 4)) != PHMM_OK) break;
        if ((rc = ss.d_out.ensure((size_t)n * sizeof(PhmmSwOut))) != PHMM_OK) break;
        if ((rc = ss.h_out.ensure((size_t)n * (sizeof(PhmmSwOut) + (size_t)cigar_cap * 4))) != PHMM_OK) break;
        uint8_t* h = static_cast<uint8_t*>(ss.h_in.p);
        std::memcpy(h + o_tasks, tasks.data(), n * sizeof(PhmmSwTask));
        std::memcpy(h + o_order, order.data(), n * sizeof(uint32_t));
        for (const SwCopy& cp : copies) std::memcpy(h + o_bytes + cp.dst, cp.src, cp.len);
        uint8_t* d = static_cast<uint8_t*>(ss.d_in.p);
        PhmmSwBatchDev b{};
        b.tasks = reinterpret_cast<const PhmmSwTask*>(d + o_tasks);
        b.order = reinterpret_cast<const uint32_t*>(d + o_order);
        b.bytes = d + o_bytes;
        b.bt = static_cast<uint32_t*>(ss.d_bt.p);
        b.aux = static_cast<int32_t*>(ss.d_aux.p);
        b.cigars = static_cast<uint32_t*>(ss.d_cig.p);
        b.outs = static_cast<PhmmSwOut*>(ss.d_out.p);
        b.cigar_cap = cigar_cap;
        b.match = params->match; b.mismatch = params->mismatch; b.gap_open = params->gap_open; b.gap_extend = params->gap_extend;
        ss.n = n;
        ss.cigar_cap = cigar_cap;
        ss.cells = cells;
        ss.in_bytes = off;
        ss.busy = true;                                                  // from here on the slot must be drained, also on an error
        auto launch = [&]() -> int {
            CU(cudaMemcpyAsync(ss.d_in.p, ss.h_in.p, off, cudaMemcpyHostToDevice, ss.stream));
            CU(cudaEventRecord(ss.ev[0], ss.stream));
            for (uint32_t g0 = 0; g0 < n;) {                             // one fill launch per cpt group
                uint32_t g1 = g0;
                while (g1 < n && tasks[order[g1]].cpt == tasks[order[g0]].cpt) ++g1;
                const uint32_t warps = std::min<uint32_t>(g1 - g0, e->warp_slots * 4);
                CU(phmm_launch_sw_fill(b, (int)tasks[order[g0]].cpt, g0, g1 - g0, warps, ss.stream));
                g0 = g1;
            }
            CU(phmm_launch_sw_trace(b, 0, n, ss.stream));
            CU(cudaEventRecord(ss.ev[1], ss.stream));
            uint8_t* ho = static_cast<uint8_t*>(ss.h_out.p);
            CU(cudaMemcpyAsync(ho, ss.d_out.p, (size_t)n * sizeof(PhmmSwOut), cudaMemcpyDeviceToHost, ss.stream));
            CU(cudaMemcpyAsync(ho + (size_t)n * sizeof(PhmmSwOut), ss.d_cig.p, (size_t)n * cigar_cap * 4, cudaMemcpyDeviceToHost, ss.stream));
            return PHMM_OK;
        };
        rc = launch();
    }
    if (rc != PHMM_OK) {                                                 // leave no work in flight behind an error
        for (auto& ss : c.slot) {
            if (ss.busy) cudaStreamSynchronize(ss.stream);
            ss.busy = false;
        }
        return rc;
    }
    // drain: oldest first
    for (int k = 0; k < 2; ++k)
        if ((rc = finish(c.slot[(turn + k) & 1])) != PHMM_OK) return rc;
    return PHMM_OK;
}

int phmm_sw_last_timing(phmm_engine* e, float* gpu_ms, uint64_t* cells)
{
    if (!e) return fail(PHMM_ERR_ARG, "phmm_sw_last_timing: null engine");
    std::lock_guard<std::mutex> lk(e->sw.m);
    if (gpu_ms) *gpu_ms = e->sw.last_ms;
    if (cells) *cells = e->sw.last_cells;
    return PHMM_OK;
}

int phmm_get_stats(phmm_engine* e, phmm_stats* out)
{
    if (!e || !out) return fail(PHMM_ERR_ARG, "phmm_get_stats: null argument");
    out->regions = e->stats.regions;
    out->pairs = e->stats.pairs;
    out->cells = e->stats.cells;
    out->fp64_pairs = e->stats.fp64_pairs;
    out->slow_pairs = e->stats.slow_pairs;
    out->h2d_bytes = e->stats.h2d;
    out->d2h_bytes = e->stats.d2h;
    out->batches = e->stats.batches;
    out->kernel_launches = e->stats.launches;
    {
        std::lock_guard<std::mutex> lk(e->stats.m);
        out->gpu_ms = e->stats.gpu_ms;
        out->host_plan_ms = e->stats.plan_ms;
        out->host_epilogue_ms = e->stats.epilogue_ms;
    }
    out->device_epilogue_batches = e->stats.dev_epilogues;
    return PHMM_OK;
}

int phmm_kernel_info(uint32_t read_len, int fma_mode, int* regs, int* local_bytes, int* lanes, int* rows_per_lane)
{
    const int bin = phmm_bin_of(read_len, fma_mode != PHMM_FMA_NONE);
    if (bin < 0) return fail(PHMM_ERR_ARG, "phmm_kernel_info: no wavefront bin holds reads of this length (per-pair kernel)");
    PhmmKernelInfo ki{};
    CU(phmm_fwd_kernel_info(bin, fma_mode != PHMM_FMA_NONE, &ki));
    if (regs) *regs = ki.regs;
    if (local_bytes) *local_bytes = ki.local_bytes;
    if (lanes) *lanes = kPhmmBins[bin].lanes;
    if (rows_per_lane) *rows_per_lane = kPhmmBins[bin].rpt;
    return PHMM_OK;
}

// Measured FP32 peak for the roofline denominator: runs the issue-rate micro-benchmark (phmm_kernels.cu) on every SM
// and returns the sustained rate in FMA lane-operations per second (one FFMA2 counts 2 per lane).
int phmm_ubench(phmm_engine* e, int mode, uint32_t iters, float* ms_out, double* fma_per_s)
{
    if (!e) return fail(PHMM_ERR_ARG, "phmm_ubench: null engine");
    CU(cudaSetDevice(e->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, e->device));
    const uint32_t blocks = (uint32_t)prop.multiProcessorCount * 8;
    Slot& s = *e->slots[0];
    int rc = s.d_cnt.ensure(256);
    if (rc != PHMM_OK) return rc;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(s.ev[0], s.stream));
        CU(phmm_launch_ubench(mode, blocks, iters, static_cast<float*>(s.d_cnt.p), s.stream));
        CU(cudaEventRecord(s.ev[3], s.stream));
        CU(cudaStreamSynchronize(s.stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s.ev[0], s.ev[3]));
        if (rep > 0 && ms < best) best = ms;
    }
    if (ms_out) *ms_out = best;
    if (fma_per_s) *fma_per_s = (double)blocks * 256.0 * (double)iters * 32.0 * 2.0 / ((double)best * 1e-3);
    return PHMM_OK;
}

// For tests (no GPU needed): run the planner alone on a flat batch and audit what it produced.  Every (read, haplotype) pair
// must be computed exactly once — by exactly one forward-kernel warp slot whose chunk holds the haplotype, or by the per-pair
// kernel — every read must sit in a bin that holds it, every CTA must respect its launch's shape.  out[0..15]:
//   0 pairs  1 pairs covered by the wavefront kernel  2 per-pair-kernel pairs  3 pairs covered more than once  4 pairs missed
//   5 warps launched  6 read slots  7 read slots filled  8 reads in a bin too small for them  9 launches  10 chunks
//   11 lane-steps (sum over warps of 32 x RPT x steps)  12 cells (sum R x H over wavefront pairs)  13 CTA shape violations
int phmm_debug_plan(const phmm_flat_batch* fb, int fma_mode, uint32_t cta_warps, uint32_t busy_batches, uint32_t lanes_per_read, uint64_t* out)
{
    if (!fb || !out) return fail(PHMM_ERR_ARG, "phmm_debug_plan: null argument");
    phmm_engine e;
    e.fma = fma_mode != PHMM_FMA_NONE;
    e.cta_warps = std::min<uint32_t>(cta_warps, PHMM_WARPS);
    e.busy_batches.store(busy_batches);
    e.lanes_pref = lanes_per_read;
    Slot s;
    s.set_host_only();
    std::vector<phmm_read> reads;
    std::vector<phmm_hap> haps;
    std::vector<RegionIn> regs;
    flat_to_regions(fb, nullptr, 0, fb->n_regions, reads, haps, regs);
    int rc = e.plan(s, regs);
    if (rc != PHMM_OK) return rc;
    const uint8_t* h = static_cast<const uint8_t*>(s.h_in.p);
    const PhmmRead* R = reinterpret_cast<const PhmmRead*>(h + s.o_reads);
    const PhmmHap* H = reinterpret_cast<const PhmmHap*>(h + s.o_haps);
    const PhmmChunk* C = reinterpret_cast<const PhmmChunk*>(h + s.o_chunks);
    const PhmmCta* T = reinterpret_cast<const PhmmCta*>(h + s.o_ctas);
    const uint32_t* tasks = reinterpret_cast<const uint32_t*>(h + s.o_tasks);
    const PhmmPair* slow = reinterpret_cast<const PhmmPair*>(h + s.o_slow);
    for (int k = 0; k < 16; ++k) out[k] = 0;
    std::vector<uint8_t> seen(std::max<uint64_t>(s.n_pairs, 1), 0);
    out[0] = s.n_pairs;
    out[9] = s.launches.size();
    out[10] = s.n_chunks;
    for (const FwdLaunch& fl : s.launches) {
        const PhmmBin bin = kPhmmBins[fl.bin];
        const uint32_t per_warp = 32u / (uint32_t)bin.lanes, per_cta = per_warp * fl.cta_warps;
        for (uint32_t c = fl.cta_base; c < fl.cta_base + fl.n_ctas; ++c) {
            const PhmmCta& cta = T[c];
            const PhmmChunk& ch = C[cta.chunk];
            if (cta.n_reads == 0 || cta.n_reads > per_cta || cta.read_begin + cta.n_reads > s.n_tasks) ++out[13];
            const uint32_t warps = (cta.n_reads + per_warp - 1) / per_warp;
            const uint64_t steps = ch.stream_len - 1 - (32 - (uint32_t)bin.lanes);
            out[5] += warps;
            out[6] += (uint64_t)warps * per_warp;
            out[7] += cta.n_reads;
            out[11] += (uint64_t)warps * 32u * (uint32_t)bin.rpt * steps;
            for (uint32_t k = 0; k < cta.n_reads; ++k) {
                const PhmmRead& rd = R[tasks[cta.read_begin + k]];
                if (rd.len + 1 > (uint32_t)(bin.lanes * bin.rpt)) ++out[8];
                for (uint32_t j = 0; j < ch.n_haps; ++j) {
                    const PhmmHap& hp = H[ch.hap_begin + j];
                    if (hp.pos == 0) continue;
                    const uint64_t idx = (uint64_t)rd.out_off + hp.out_col;
                    if (idx >= s.n_pairs || hp.out_col >= rd.n_haps) { ++out[13]; continue; }
                    if (seen[idx]++) ++out[3];
                    ++out[1];
                    out[12] += (uint64_t)rd.len * hp.len;
                }
            }
        }
    }
    for (uint32_t k = 0; k < s.n_slow; ++k) {
        const uint64_t idx = slow[k].out_idx;
        if (idx >= s.n_pairs) { ++out[13]; continue; }
        if (seen[idx]++) ++out[3];
        ++out[2];
    }
    for (uint64_t k = 0; k < s.n_pairs; ++k)
        if (!seen[k]) ++out[4];
    for (auto* b : {&s.h_in, &s.h_f32, &s.h_f64, &s.h_cnt}) b->release();
    return PHMM_OK;
}

// For tests: the same conversion on the host in `mode` (phmm_log10.h; 0 = libm itself); values under MIN_ACCEPTED come out as -inf
// like the device's.
int phmm_debug_host_log10(const float* x, uint32_t n, int mode, double* out)
{
    if (!x || !out) return fail(PHMM_ERR_ARG, "phmm_debug_host_log10: null argument");
    phmm_log10_block(x, out, n, mode);
    for (uint32_t i = 0; i < n; ++i)
        if (x[i] < kMinAccepted) out[i] = -INFINITY;
    return PHMM_OK;
}

// For tests (GPU): out[i] = (double)(log10f(x[i]) - LOG10_INITIAL_CONSTANT_F) as the DEVICE epilogue computes it (-inf under
// MIN_ACCEPTED), to be compared with the host's libm.  *in_use says whether this engine runs its epilogue on the device.
int phmm_debug_device_log10(phmm_engine* e, const float* x, uint32_t n, double* out, int* in_use)
{
    if (!e || !x || !out) return fail(PHMM_ERR_ARG, "phmm_debug_device_log10: null argument");
    if (in_use) *in_use = e->device_epilogue ? 1 : 0;
    CU(cudaSetDevice(e->device));
    DevBuf dx, dy;
    int rc;
    if ((rc = dx.ensure((size_t)n * sizeof(float))) != PHMM_OK) return rc;
    if ((rc = dy.ensure((size_t)n * sizeof(double))) != PHMM_OK) return rc;
    CU(cudaMemcpy(dx.p, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    CU(phmm_launch_log10_probe(static_cast<const float*>(dx.p), static_cast<double*>(dy.p), n, nullptr));
    CU(cudaMemcpy(out, dy.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    dx.release();
    dy.release();
    return PHMM_OK;
}

// For tests and host timing (no GPU needed): the engine's host epilogue on raw sums the caller brings (the oracle's, say), laid
// out like phmm_flat_result.raw_f32 / raw_f64 (raw_f64 may be null: underflowing pairs then stand at -inf).  mode as in
// phmm_debug_log10f_check; runs `reps` times and reports the mean wall time of one run.
int phmm_debug_epilogue(const phmm_flat_batch* fb, const float* raw_f32, const double* raw_f64, phmm_flat_result* fr, int mode,
                        uint32_t reps, double* ms)
{
    if (!fb || !raw_f32 || !fr) return fail(PHMM_ERR_ARG, "phmm_debug_epilogue: null argument");
    phmm_engine e;
    e.log10_mode = mode < 0 ? phmm_log10_pick_mode() : mode;
    Slot s;
    s.set_host_only();
    std::vector<phmm_read> reads;
    std::vector<phmm_hap> haps;
    std::vector<RegionIn> regs;
    flat_to_regions(fb, fr, 0, fb->n_regions, reads, haps, regs);
    int rc = e.plan(s, regs);
    if (rc != PHMM_OK) return rc;
    for (size_t g = 0; g < regs.size(); ++g)
        if (s.regions[g].n_reads && s.regions[g].n_haps && s.regions[g].out_off != fr->lik_off[g])
            return fail(PHMM_ERR_ARG, "phmm_debug_epilogue: lik_off is not the running sum of reads x haplotypes");
    fr->n_fp64 = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t k = 0; k < std::max(reps, 1u); ++k) {
        fr->n_fp64 = 0;
        e.epilogue(s, regs, raw_f32, raw_f64, raw_f64 != nullptr);
    }
    if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / std::max(reps, 1u);
    for (auto* b : {&s.h_in, &s.h_f32, &s.h_f64, &s.h_cnt}) b->release();
    return PHMM_OK;
}

// For tests (no GPU needed): the raw-sum -> log10 conversion in `mode` (phmm_log10.h; -1 = the mode phmm_create picks on this
// host) against this host's libm over every `stride`-th normal float.  Returns the number of inputs whose bits differ (0 for
// every mode on glibc <= 2.39); *picked says what phmm_create would use.
uint64_t phmm_debug_log10f_check(uint32_t stride, int mode, int* picked)
{
    const int auto_mode = phmm_log10_pick_mode();
    if (picked) *picked = auto_mode;
    if (mode < 0) mode = auto_mode;
    if (mode == 2 && phmm_log10_logf_mismatches(1u << 20) == UINT64_MAX) return UINT64_MAX;    // no AVX-512 here
    uint64_t bad = 0;
    if (stride == 0) stride = 1;
    float x[4096];
    double got[4096];
    uint64_t b = 0x00800000u;
    while (b < 0x7f800000u) {
        int n = 0;
        for (; n < 4096 && b < 0x7f800000u; ++n, b += stride) {
            const uint32_t bits = (uint32_t)b;
            std::memcpy(&x[n], &bits, 4);
        }
        phmm_log10_block(x, got, (size_t)n, mode);
        for (int k = 0; k < n; ++k) {
            const double want = (double)(log10f(x[k]) - kLog10InitF);
            bad += std::memcmp(&want, &got[k], 8) != 0;
        }
    }
    return bad;
}

// For tests (no GPU needed): how phmm_run_flat would cut `fb` with `n_workers` workers and `batch_cells` (0 = the default);
// bounds[2k], bounds[2k+1] = first and one-past-last region of sub-batch k (at most cap of them), work[k] = its weighted work.
int phmm_debug_partition(const phmm_flat_batch* fb, uint32_t n_workers, uint64_t batch_cells, uint32_t* bounds, uint64_t* work,
                         uint32_t cap, uint32_t* n_parts)
{
    if (!fb || !bounds || !n_parts) return fail(PHMM_ERR_ARG, "phmm_debug_partition: null argument");
    phmm_engine e;
    const std::vector<uint64_t> pre = flat_work_prefix(fb);
    const auto parts = flat_partition(pre, fb->n_regions, batch_cells ? batch_cells : e.batch_cells, n_workers ? n_workers : e.n_workers);
    *n_parts = (uint32_t)parts.size();
    for (uint32_t k = 0; k < parts.size() && k < cap; ++k) {
        bounds[2 * k] = parts[k].first;
        bounds[2 * k + 1] = parts[k].second;
        if (work) work[k] = pre[parts[k].second] - pre[parts[k].first];
    }
    return PHMM_OK;
}

// For tests: would phmm_create (PHMM_EPILOGUE_AUTO) put the epilogue on the device on this host?
int phmm_debug_device_epilogue_allowed(void) { return std::getenv("PHMM_B200_LIBM_LOG10F") == nullptr && phmm_log10_restated_matches_libm() ? 1 : 0; }

// For tests: the restated logf against libm's on [0.5, 2), everything the log10f formula hands it (stride 1 = all 2^24).
uint64_t phmm_debug_logf_check(uint32_t stride) { return phmm_log10_logf_mismatches(stride); }

// For tests: the host-built tables, to be compared with the reference's Context<float> (core/context.h).
void phmm_debug_tables(float* ph2pr128, float* mm8256, double* ph2pr128_d, double* mm8256_d)
{
    std::unique_ptr<PhmmTables> t(new PhmmTables());
    phmm_build_tables(t.get());
    if (ph2pr128) std::memcpy(ph2pr128, t->ph2pr, sizeof(t->ph2pr));
    if (mm8256) std::memcpy(mm8256, t->mm, sizeof(t->mm));
    if (ph2pr128_d) std::memcpy(ph2pr128_d, t->ph2pr_d, sizeof(t->ph2pr_d));
    if (mm8256_d) std::memcpy(mm8256_d, t->mm_d, sizeof(t->mm_d));
}

}  // extern "C"
