// phmm_host.h — internal declarations shared by the host-side translation units of libphmm_b200.so:
//   phmm_plan.cpp     the planner (regions -> bins, chunks, CTAs, one arena)
//   phmm_device.cpp   one GPU: launches, D2H, epilogue hand-over, worker threads
//   phmm_front.cpp    the engine in front of 1..N GPUs (tickets, flat streams, dispatch) and the C ABI
//   phmm_sw_host.cpp  host side of the batched Smith-Waterman
//   phmm_debug.cpp    phmm_debug_* exports for the tests
// Nothing here is part of the public interface (include/phmm_b200.h).
#ifndef PHMM_HOST_H_
#define PHMM_HOST_H_

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

extern thread_local std::string g_phmm_err;      // phmm_last_error(), per calling thread

inline int fail(int code, const std::string& msg)
{
    g_phmm_err = msg;
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t _e = (call);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            return fail(PHMM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));               \
    } while (0)

// NVTX ranges around plan / launch / finish: resolved at run time from libnvToolsExt if a profiler brought it,
// otherwise two predictable branches (phmm_front.cpp).
struct PhmmRange
{
    explicit PhmmRange(const char* name);
    ~PhmmRange();
    bool on;
};

constexpr uint32_t kTargetStreamCols = 2400;    // per stream and chunk: keeps streams + prior tables under ~64 KB per CTA (3+ CTAs/SM)
constexpr uint32_t kMixedStreamCols = 1600;     // ... for regions whose reads differ in length (phmm_plan.cpp)
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
    const phmm_read* reads;     // null for a region of a flat batch: its reads are read straight from the caller's flat arrays
    const phmm_flat_batch* flat;   // ... this batch, reads flat_r0 .. flat_r0 + n_reads (no per-read view structs are built)
    uint64_t flat_r0;
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

// one read of a region, wherever it lives
struct ReadView
{
    const uint8_t *bases, *qual, *ins, *del, *gcp;
    uint32_t len;
    uint8_t mapq;
};
inline uint32_t region_read_len(const RegionIn& r, uint32_t k) { return r.flat ? r.flat->read_len[r.flat_r0 + k] : r.reads[k].len; }
inline ReadView region_read(const RegionIn& r, uint32_t k)
{
    if (r.flat) {
        const phmm_flat_batch* fb = r.flat;
        const uint64_t i = r.flat_r0 + k, o = fb->read_off[i];
        return ReadView{fb->read_bases + o, fb->read_qual + o, fb->read_ins ? fb->read_ins + o : nullptr, fb->read_del ? fb->read_del + o : nullptr,
                        fb->read_gcp ? fb->read_gcp + o : nullptr, fb->read_len[i], fb->read_mapq[i]};
    }
    const phmm_read& in = r.reads[k];
    return ReadView{in.bases, in.qual, in.ins, in.del, in.gcp, in.len, in.mapq};
}

struct RegionPlan
{
    uint32_t read_begin, n_reads, hap_begin, n_haps;
    uint64_t out_off;
};

// Growing an arena means a stream-ordered free + allocation (DevBuf) or re-pinning host pages (milliseconds), so arenas grow
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

// Device arenas come from the stream-ordered allocator on a stream of their own (one per device, phmm_device.cpp): growing one
// slot's arena must not stall the batches other slots have in flight, which cudaFree's device-wide synchronisation did.
cudaStream_t phmm_alloc_stream();   // of the calling thread's current device

struct DevBuf
{
    void* p = nullptr;
    size_t cap = 0;
    bool host_only = false;     // planner self-check (phmm_debug_plan): no device behind it, nothing is allocated
    int ensure(size_t bytes, std::atomic<size_t>* high_water = nullptr)
    {
        if (bytes <= cap || host_only) return PHMM_OK;
        static const bool sync_alloc = std::getenv("PHMM_B200_SYNC_ALLOC") != nullptr;   // A/B: plain cudaMalloc / cudaFree
        cudaStream_t st = phmm_alloc_stream();
        if (p) sync_alloc ? cudaFree(p) : cudaFreeAsync(p, st);   // (the slot is idle: nothing in flight reads the old arena)
        p = nullptr;
        cap = 0;
        size_t want = grow_to(bytes, high_water);
        cudaError_t e = sync_alloc ? cudaMalloc(&p, want) : cudaMallocAsync(&p, want, st);
        if (e == cudaSuccess && !sync_alloc) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(PHMM_ERR_MEM, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
        }
        cap = want;
        return PHMM_OK;
    }
    void release()
    {
        static const bool sync_alloc = std::getenv("PHMM_B200_SYNC_ALLOC") != nullptr;
        if (p && !host_only) sync_alloc ? cudaFree(p) : cudaFreeAsync(p, phmm_alloc_stream());
        p = nullptr;
        cap = 0;
    }
    // bring an idle slot's arena up to the device-wide mark of its kind (contents are not kept; errors are left to the next ensure)
    void catch_up(const std::atomic<size_t>& high_water)
    {
        const size_t mark = high_water.load(std::memory_order_relaxed);
        if (mark <= cap || host_only) return;
        release();
        if (ensure(mark * 2 / 3) != PHMM_OK) g_phmm_err.clear();      // (grow_to gives 1.5 x what is asked for: the mark)
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
    // grow, keeping the first `keep` bytes (the planner writes the read records into the arena as it makes them)
    int ensure_keep(size_t bytes, size_t keep, std::atomic<size_t>* high_water = nullptr)
    {
        if (bytes <= cap) return PHMM_OK;
        HostBuf old = *this;
        p = nullptr;
        cap = 0;
        const int rc = ensure(bytes, high_water);
        if (rc == PHMM_OK && old.p && keep) std::memcpy(p, old.p, std::min(keep, old.cap));
        old.release();
        return rc;
    }
    void release()
    {
        if (p) {
            if (host_only) std::free(p); else cudaFreeHost(p);
        }
        p = nullptr;
        cap = 0;
    }
    void catch_up(const std::atomic<size_t>& high_water)      // as DevBuf::catch_up
    {
        const size_t mark = high_water.load(std::memory_order_relaxed);
        if (mark <= cap || host_only) return;
        release();
        if (ensure(mark * 2 / 3) != PHMM_OK) g_phmm_err.clear();      // (grow_to gives 1.5 x what is asked for: the mark)
    }
};

struct FwdLaunch
{
    int rpt;                 // rows per lane: which instantiation of the forward kernel
    bool refgaps;            // ... and which variant (the reference's own gap qualities, or per-base ones)
    uint32_t fixed_lanes;    // every read of the launch takes exactly this many lanes: the fixed-lanes specialisation (0 = general kernel)
    bool duo;                // paired reads against one stream (two task entries per read slot), PHMM_DUO_RPT rows per lane
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
    struct Group { uint32_t chunk; uint32_t stream_len; uint32_t first_read; uint32_t first_warp; uint32_t n_warps; };
    struct PlanScratch
    {
        std::vector<PhmmHap> haps;
        std::vector<PhmmChunk> chunks;
        std::vector<PhmmPair> slow;
        std::vector<PhmmCta> ctas;
        std::vector<PhmmRegionDev> regions_dev;
        std::vector<double> pri;
        std::vector<uint32_t> task_reads, order, region_chunks, idx, bucket, rec_of, hap_boff, stream_sum, stream_of, stream_idx;
        std::vector<uint32_t> duo_a, duo_b, duo_order;          // paired reads: the two reads of every slot, slots warp after warp
        std::vector<std::vector<Group>> groups;
        std::vector<Group> sorted;
        std::vector<std::vector<uint32_t>> group_reads;
        std::vector<std::vector<uint8_t>> group_warps;
        std::vector<uint16_t> fast_len;
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
    bool force_dev_epi = false;          // ... whatever its size (resident batches: the matrices stay on the device for
                                         // phmm_resident_marginalize), if the host's libm allows the device epilogue at all
    bool want_best = false;              // some region of the batch asks for best / second-best haplotypes
    bool want_raw = false;               // ... for the raw FP32 sums
    bool want_f64 = false;               // the FP64 recompute goes out with the batch (someone will read its values)
    bool f64_done = false;               // ... and has been launched for the sums now on the device
    bool small_ok = false;               // set by the ticket path before plan(): a small batch may take the short device path
    bool host_out = false;               // this batch's raw sums are written straight into h_f32 (mapped pinned memory): no D2H
    bool lite = false;                   // one forward launch, no side streams: no forward-only timing events either
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

// What phmm_config asked for, shared by every device of an engine.
struct PhmmSettings
{
    bool fma = true;
    uint32_t n_workers = 6;      // host worker threads = batches in flight, PER DEVICE
    uint64_t batch_cells = 32000000000ull;   // mixed-length batches under ~8e9 cells run their per-bin kernels too small to fill the device
    uint32_t linger_us = 200;
    uint32_t cta_warps = 0;      // warps per forward-kernel CTA; 0 = chosen per launch from the fill of the read slots
    uint32_t lanes_pref = 0;     // lanes per read of the reference-gap kernels: 0 = by batch size, else 8 / 16 / 32
    bool fp64_always = false;    // PHMM_FP64_ALWAYS: recompute every underflowing pair even when nobody can see the value
    bool device_epilogue = false;  // log10 / normalize / filter / transpose / best haplotype in one more kernel (PHMM_EPILOGUE_*)
    uint64_t device_epilogue_min_pairs = 0;   // ... for batches of at least this many pairs (AUTO: small batches stay on the host)
    int log10_mode = 0;            // how raw sums become log10 doubles: phmm_log10.cpp, picked (and tested against libm) at creation
    float default_gapm = 0.f, default_ii = 0.f;   // table values of the default gap-continuation quality
    double default_gapm_d = 0.0, default_gcp_d = 0.0;   // ... in FP64 (the recompute kernel)
    uint32_t pipeline_depth = 6;   // batches per device the ticket path keeps in flight before it starts batching up (16-64 callers: +3-10 % over 3)
};

// batched Smith-Waterman (phmm_sw_align): per device its own streams and arenas
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
    SwSlot slot[2];
    cudaEvent_t base = nullptr;      // start of the call on the device's clock: sub-batches overlap, their kernel time is a union
    float covered_ms = 0.f;
    float last_ms = 0.f;
    uint64_t last_cells = 0;
};

// One GPU of an engine: its tables, its batch slots (one per worker thread), its job queue.  A job is one batch: the worker
// that takes it plans it, ships it, launches it, waits for it and hands the results to the callers' buffers.
struct PhmmDevice
{
    const PhmmSettings* cfg = nullptr;
    int ordinal = 0;             // CUDA device ordinal
    uint32_t index = 0;          // position in phmm_engine::devs
    uint32_t warp_slots = 148 * 8; // warps of the widest forward kernel the device holds at once (8 per SM at 231 registers)
    size_t smem_optin = 227 * 1024;  // dynamic shared memory one CTA may ask for (cudaDevAttrMaxSharedMemoryPerBlockOptin)
    size_t smem_sm = 228 * 1024;     // shared memory of one SM (cudaDevAttrMaxSharedMemoryPerMultiprocessor)
    std::atomic<uint32_t> busy_batches{0};   // batches somewhere between plan and the end of their epilogue
    std::atomic<uint64_t> outstanding{0};    // work (weighted cells) queued on or running on this device: the dispatch key
    std::atomic<size_t> high_water[8] = {};   // per arena kind, see grow_to()
    DevBuf d_tables;
    Stats stats;
    SwContext sw;
    // kernel-span union for phmm_stats.gpu_ms: spans of concurrent batches overlap on the device
    cudaEvent_t epoch = nullptr;             // the device time up to which batch spans have been counted (add_span)
    std::mutex span_m;
    // phmm_ubench's own stream, events and buffer (never a worker's)
    cudaStream_t ub_stream = nullptr;
    cudaEvent_t ub_ev[2] = {nullptr, nullptr};
    DevBuf ub_buf;
    std::mutex ub_m;

    std::vector<std::thread> workers;
    std::vector<std::unique_ptr<Slot>> slots;
    std::mutex qm;
    std::condition_variable qcv;
    std::deque<std::function<void(Slot&)>> jobs;
    bool stopping = false;
    // Slots for batches that run on their caller's own thread (phmm_wait on an idle pipeline: no hand-over to a worker and
    // back, which is two thread wake-ups per call).  Taken and returned with one atomic each.
    static constexpr uint32_t kCallerSlots = 4;
    std::unique_ptr<Slot> caller_slots[kCallerSlots];
    std::atomic<bool> caller_slot_busy[kCallerSlots] = {};
    void catch_up(Slot& s);                  // an idle worker brings its slot's arenas up to the device's high-water marks
    Slot* take_caller_slot();
    void return_caller_slot(Slot* s);

    int init(const PhmmSettings* settings, int cuda_ordinal, uint32_t idx, const PhmmTables& tables);
    void shutdown();
    int plan(Slot& s, const std::vector<RegionIn>& regs);
    int launch(Slot& s, bool copy_in);
    cudaError_t launch_f64(Slot& s);
    int finish(Slot& s, const std::vector<RegionIn>& regs, bool copy_out);
    uint64_t epilogue(const Slot& s, const std::vector<RegionIn>& regs, const float* f32, const double* f64, bool have64) const;
    void epilogue_from_device(const Slot& s, const std::vector<RegionIn>& regs, bool have64) const;
    int execute(Slot& s, const std::vector<RegionIn>& regs);
    void drain(Slot& s);                       // after an error: nothing of this slot may still be in flight when it is reused
    void add_span(Slot& s);                    // fold a finished batch's kernel span into stats.gpu_ms (union, not sum)
    void worker_main(uint32_t idx);
    void enqueue(std::function<void(Slot&)> fn, uint64_t work);
};

struct phmm_resident
{
    PhmmDevice* dev = nullptr;
    phmm_flat_batch fb{};                // a copy of the caller's batch descriptor (its arrays must outlive the resident batch)
    Slot slot;
    bool ran = false;                    // the device holds results of a phmm_resident_run
    DevBuf d_marg;                       // phmm_resident_marginalize: sites + maps in, marginal matrices out
    HostBuf h_marg;
    std::vector<RegionIn> regs;
    std::vector<phmm_hap> haps;
};

// The engine a caller holds: 1..N devices behind one submit / wait / run_flat front.  Batches go to the device with the least
// outstanding work (regions are independent, no collective; results land in the one caller's host buffers).
struct phmm_engine
{
    PhmmSettings cfg;
    std::vector<std::unique_ptr<PhmmDevice>> devs;
    std::atomic<uint32_t> rr{0};                 // tie-break of the dispatch

    PhmmDevice& pick_device();
    uint32_t total_workers() const { return cfg.n_workers * (uint32_t)std::max<size_t>(devs.size(), 1); }

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
        std::atomic<bool> done{false};   // set (release) after rc and err; a waiter polls it before it blocks on cv
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
    // busy.  A batch with a waiter goes out at once while fewer than pipeline_depth batches per device are in flight (the
    // pipeline is starving: plan / GPU / epilogue of different batches overlap); otherwise it keeps collecting the callers that
    // return from earlier batches and goes out when one of those completes, when it is big (batch_cells), or after linger_us.
    uint32_t in_flight = 0;            // sealed batches that have not completed
    uint32_t depth() const { return cfg.pipeline_depth * (uint32_t)std::max<size_t>(devs.size(), 1); }
    // Seals the open batch.  With `caller_runs` the batch may come back as a job for the calling thread (to be run after bm has
    // been released) instead of going to a worker.
    std::function<void()> seal_locked(bool caller_runs = false);
    bool caller_runs = true;             // PHMM_B200_CALLER_RUNS=0: every batch goes to a worker (the round-1 hand-over; A/B and tests)

    std::mutex sw_m;                   // phmm_sw_align: one call at a time
    float sw_last_ms = 0.f;
    uint64_t sw_last_cells = 0;
};

// rows per lane the throughput policy picks for a region of many reads of one length (phmm_plan.cpp)
int phmm_plan_rpt_for_uniform_reads(uint32_t read_len, bool refgaps);

// flat batches (phmm_front.cpp)
void flat_to_regions(const phmm_flat_batch* fb, phmm_flat_result* fr, uint32_t g0, uint32_t g1, std::vector<phmm_hap>& haps,
                     std::vector<RegionIn>& regs);
constexpr uint64_t kPairWeight = 16384;
constexpr size_t kHostOutBytes = 256 << 10;   // raw sums of a batch up to this size are written straight to mapped host memory
std::vector<uint64_t> flat_work_prefix(const phmm_flat_batch* fb);
std::vector<std::pair<uint32_t, uint32_t>> flat_partition(const std::vector<uint64_t>& pre, uint32_t n_regions, uint64_t batch_cells, uint32_t n_workers);
uint64_t region_cells(const RegionIn& r);

#endif
