// phmm_front.cpp — the engine a caller holds and its C ABI (include/phmm_b200.h).
//
// Mirrors what the reference's host wrapper does around its kernels (pairhmm/pairhmm_internal.cpp call_old_pairhmm :505-594 /
// call_new_pairhmm002 :596-753), re-cut for GPUs:
//   * regions from many callers are gathered into batches (the reference runs one region per call on a pool thread,
//     haplotypecaller_engine.cpp:119,250 — far too little work for a B200): tickets (phmm_submit / phmm_wait) with a
//     batch-while-busy sealing policy for its synchronous callers, phmm_run_flat for whole streams;
//   * ONE process drives 1..N GPUs, as the reference is one multi-threaded process (haplotypecaller.cpp:49,97): every
//     sealed batch / flat sub-batch goes to the device with the least outstanding work, the results land in the one caller's
//     host buffers.  Regions are independent: no collective, no peer access;
//   * per batch the host only plans (phmm_plan.cpp); quality capping, the PCR-indel model, the table look-ups, the forward
//     recurrence, the FP64 underflow recompute and (for big batches) the log10 / normalize / filter / transpose epilogue run
//     on the device (phmm_kernels.cu, phmm_device.cpp).
// There is no CPU implementation of the recurrence in this library and no fallback.
#include "phmm_host.h"

#include <nvtx3/nvToolsExt.h>

thread_local std::string g_phmm_err;

// NVTX ranges (header-only NVTX3: without a profiler attached a push / pop is a load and a branch)
PhmmRange::PhmmRange(const char* name) : on(true) { nvtxRangePushA(name); }
PhmmRange::~PhmmRange() { nvtxRangePop(); }

// ------------------------------------------------------------------------------------------------
// Dispatch: the device with the least outstanding work takes the next batch (ties go round)
// ------------------------------------------------------------------------------------------------

PhmmDevice& phmm_engine::pick_device()
{
    const size_t n = devs.size();
    if (n == 1) return *devs[0];
    const size_t start = rr.fetch_add(1, std::memory_order_relaxed) % n;
    size_t best = start;
    uint64_t best_w = devs[start]->outstanding.load(std::memory_order_relaxed);
    for (size_t k = 1; k < n; ++k) {
        const size_t i = (start + k) % n;
        const uint64_t w = devs[i]->outstanding.load(std::memory_order_relaxed);
        if (w < best_w) { best = i; best_w = w; }
    }
    return *devs[best];
}

// bm must be held
std::function<void()> phmm_engine::seal_locked(bool caller_runs)
{
    if (open.regs.empty()) return {};
    auto regs = std::make_shared<std::vector<RegionIn>>(std::move(open.regs));
    std::shared_ptr<BatchSync> sync = batch_state[open.id].sync;
    uint64_t work = open.cells;
    for (const RegionIn& r : *regs) work += (uint64_t)r.n_reads * r.n_haps * kPairWeight;
    open = OpenBatch{};
    ++in_flight;
    open_cv.notify_all();
    PhmmDevice& d = pick_device();
    auto body = [this, &d, regs, sync](Slot& s) {
        const int rc = d.execute(s, *regs);
        {
            std::lock_guard<std::mutex> lk(sync->m);
            sync->rc = rc;
            if (rc != PHMM_OK) sync->err = g_phmm_err;
            sync->done.store(true, std::memory_order_release);
        }
        sync->cv.notify_all();
        std::lock_guard<std::mutex> lk(bm);
        --in_flight;
        if (open.waiters > 0 && in_flight < depth()) seal_locked();   // a stage is free: send what has gathered meanwhile
    };
    if (caller_runs && regs->size() == 1) {
        // The pipeline has room and the thread that noticed is about to block on this very batch, its own region: it runs the
        // batch itself (plan, launches, completion poll, epilogue) on one of the device's caller slots.  A worker would do the
        // same work two thread wake-ups later.  (A batch that has gathered several callers' regions goes to a worker: with 32
        // and 64 callers that measured 4-7 % more regions per second, profiles/r02_callers_table.txt.)
        if (Slot* s = d.take_caller_slot()) {
            d.outstanding.fetch_add(work, std::memory_order_relaxed);
            return [&d, s, body, work] {
                int prev = -1;
                if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
                if (prev != d.ordinal) cudaSetDevice(d.ordinal);
                body(*s);
                d.outstanding.fetch_sub(work, std::memory_order_relaxed);
                d.return_caller_slot(s);
                if (prev >= 0 && prev != d.ordinal) cudaSetDevice(prev);      // the caller's own CUDA code keeps its device
            };
        }
    }
    d.enqueue(body, work);
    return {};
}

// ------------------------------------------------------------------------------------------------
// Flat batches
// ------------------------------------------------------------------------------------------------

// Regions [g0, g1) of a flat batch as RegionIn records.  The reads stay where they are (RegionIn.flat: the planner reads the
// caller's flat arrays directly — a view struct per read cost more than the planning it feeds); the haplotypes, few per region,
// get their views in `haps`.
void flat_to_regions(const phmm_flat_batch* fb, phmm_flat_result* fr, uint32_t g0, uint32_t g1, std::vector<phmm_hap>& haps,
                     std::vector<RegionIn>& regs)
{
    const uint64_t hbeg = fb->reg_hap_off[g0];
    const uint64_t nh = fb->reg_hap_off[g1] - hbeg;
    haps.resize(nh);
    for (uint64_t j = 0; j < nh; ++j) {
        haps[j].bases = fb->hap_bases + fb->hap_off[hbeg + j];
        haps[j].len = fb->hap_len[hbeg + j];
        haps[j].priority = fb->hap_priority ? fb->hap_priority[hbeg + j] : 0.0;
    }
    regs.resize(g1 - g0);
    for (uint32_t g = g0; g < g1; ++g) {
        RegionIn& r = regs[g - g0];
        const uint64_t r0 = fb->reg_read_off[g], h0 = fb->reg_hap_off[g];
        r.reads = nullptr;
        r.flat = fb;
        r.flat_r0 = r0;
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
std::vector<uint64_t> flat_work_prefix(const phmm_flat_batch* fb)
{
    std::vector<uint64_t> pre(fb->n_regions + 1, 0);
    auto work_of = [&](uint32_t g0, uint32_t g1) {                  // pre[g + 1] = work of region g, for g in [g0, g1)
        for (uint32_t g = g0; g < g1; ++g) {
            uint64_t rb = 0, hb = 0;
            for (uint64_t k = fb->reg_read_off[g]; k < fb->reg_read_off[g + 1]; ++k) rb += fb->read_len[k];
            for (uint64_t j = fb->reg_hap_off[g]; j < fb->reg_hap_off[g + 1]; ++j) hb += fb->hap_len[j];
            const uint64_t pairs = (fb->reg_read_off[g + 1] - fb->reg_read_off[g]) * (fb->reg_hap_off[g + 1] - fb->reg_hap_off[g]);
            pre[g + 1] = rb * hb + pairs * kPairWeight;
        }
    };
    // this runs on the caller's thread before any sub-batch can start: big streams (one engine over several GPUs takes
    // hundreds of thousands of regions per call) spread the pass over a few threads
    const uint32_t n = fb->n_regions;
    const bool big = n >= 131072 || (n > 0 && fb->reg_read_off[n] - fb->reg_read_off[0] >= 1000000);   // (~0.5 ms of adds per million reads)
    const uint32_t n_thr = big ? std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)) : 1u;
    if (n_thr > 1) {
        std::vector<std::thread> pool;
        for (uint32_t t = 0; t < n_thr; ++t)
            pool.emplace_back(work_of, (uint32_t)((uint64_t)n * t / n_thr), (uint32_t)((uint64_t)n * (t + 1) / n_thr));
        for (auto& t : pool) t.join();
    } else {
        work_of(0, n);
    }
    for (uint32_t g = 0; g < n; ++g) pre[g + 1] += pre[g];
    return pre;
}

// phmm_run_flat's cut of a stream of regions (pre = prefix sums of their work, flat_work_prefix) into sub-batches [first, end).
std::vector<std::pair<uint32_t, uint32_t>> flat_partition(const std::vector<uint64_t>& pre, uint32_t n_regions, uint64_t batch_cells,
                                                                 uint32_t n_workers)
{
    const uint64_t total = pre[n_regions];
    // sub-batch size: batch_cells at most, and small enough that every worker gets about two of them (host-bound streams of
    // small regions need the planning and the epilogues spread over all workers), but not below a quarter of batch_cells
    // (mixed-length batches much smaller than that run their per-bin kernels too small to fill the device)
    const uint64_t full = std::max<uint64_t>(1, std::min<uint64_t>(batch_cells, std::max<uint64_t>(total / (2ull * std::max(n_workers, 1u)), batch_cells >> 2)));
    std::vector<std::pair<uint32_t, uint32_t>> parts;
    uint32_t g = 0;
    // Ramp up: the workers plan the first sub-batches side by side, and the device has nothing to do until the first plan is done —
    // so the first ones are small (an eighth of the full size: planned in an eighth of the time), and there are several of each
    // size, so that the device still has work when the next size up arrives: 4 x full/8, 2 x full/4, 2 x full/2, then full.
    static const uint8_t kRamp[8] = {3, 3, 3, 3, 2, 2, 1, 1};
    uint32_t step = 0;
    while (g < n_regions) {
        const uint64_t left = total - pre[g];
        uint64_t want = step < 8 ? (full >> kRamp[step]) : full;
        ++step;
        // ramp down: never take more than half of what is left while more than full/4 remains
        if (left > (full >> 2)) want = std::min<uint64_t>(want, std::max<uint64_t>(left / 2, full >> 2));
        uint32_t end = g + 1;
        while (end < n_regions && pre[end] - pre[g] < want) ++end;
        parts.emplace_back(g, end);
        g = end;
    }
    return parts;
}

uint64_t region_cells(const RegionIn& r)
{
    uint64_t rb = 0, hb = 0;
    for (uint32_t k = 0; k < r.n_reads; ++k) rb += region_read_len(r, k);
    for (uint32_t j = 0; j < r.n_haps; ++j) hb += r.haps[j].len;
    return rb * hb;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------

static void add_stats(const PhmmDevice& d, phmm_stats* out)
{
    out->regions += d.stats.regions;
    out->pairs += d.stats.pairs;
    out->cells += d.stats.cells;
    out->fp64_pairs += d.stats.fp64_pairs;
    out->slow_pairs += d.stats.slow_pairs;
    out->h2d_bytes += d.stats.h2d;
    out->d2h_bytes += d.stats.d2h;
    out->batches += d.stats.batches;
    out->kernel_launches += d.stats.launches;
    out->device_epilogue_batches += d.stats.dev_epilogues;
    std::lock_guard<std::mutex> lk(const_cast<std::mutex&>(d.stats.m));
    out->gpu_ms += d.stats.gpu_ms;
    out->host_plan_ms += d.stats.plan_ms;
    out->host_epilogue_ms += d.stats.epilogue_ms;
}

extern "C" {

const char* phmm_last_error(void) { return g_phmm_err.c_str(); }
const char* phmm_version(void) { return "rovaca_b200 pairhmm 0.2 (sm_100a)"; }

int phmm_create_multi(const phmm_config* cfg, const int32_t* devices, uint32_t n_devices, phmm_engine** out)
{
    if (!out) return fail(PHMM_ERR_ARG, "phmm_create: out is null");
    *out = nullptr;
    if (n_devices && !devices) return fail(PHMM_ERR_ARG, "phmm_create_multi: device list is null");
    if (n_devices > 64) return fail(PHMM_ERR_ARG, "phmm_create_multi: more than 64 devices");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(PHMM_ERR_CUDA, std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") +
                                       " (this engine has no CPU fallback)");
    std::vector<int> ords;
    if (n_devices == 0) {
        for (int d = 0; d < ndev; ++d) ords.push_back(d);             // every visible device
    } else {
        for (uint32_t k = 0; k < n_devices; ++k) {
            int d = devices[k];
            if (d < 0) CU(cudaGetDevice(&d));                         // -1 = the calling thread's current device
            if (d >= ndev) return fail(PHMM_ERR_ARG, "phmm_create: device ordinal out of range");
            ords.push_back(d);                                        // an ordinal may be listed twice: two queues on one GPU
        }
    }
    int caller_dev = 0;
    cudaGetDevice(&caller_dev);
    std::unique_ptr<phmm_engine> e(new phmm_engine());
    PhmmSettings& s = e->cfg;
    // host workers per device: the box's cores are shared by the devices this process drives and — one rank per GPU being the
    // other deployment — by the ranks beside it
    const unsigned sharers = (unsigned)std::max<size_t>(std::max<size_t>(ords.size(), (size_t)ndev), 1);
    s.n_workers = std::max(2u, std::min(12u, std::thread::hardware_concurrency() / sharers));
    if (const char* env = std::getenv("PHMM_B200_WORKERS"))          // deployment override of the default only
        if (std::atoi(env) > 0) s.n_workers = std::min(32, std::atoi(env));
    if (cfg) {
        s.fma = cfg->fma_mode != PHMM_FMA_NONE;
        if (cfg->n_workers) s.n_workers = std::min<uint32_t>(cfg->n_workers, 32);
        if (cfg->batch_cells) s.batch_cells = cfg->batch_cells;
        if (cfg->linger_us) s.linger_us = cfg->linger_us;
        if (cfg->cta_warps) s.cta_warps = std::min<uint32_t>(cfg->cta_warps, PHMM_WARPS);
        if (cfg->pipeline_depth) s.pipeline_depth = std::min<uint32_t>(cfg->pipeline_depth, 32);
        if (cfg->lanes_per_read == 8 || cfg->lanes_per_read == 16 || cfg->lanes_per_read == 32) s.lanes_pref = cfg->lanes_per_read;
        s.fp64_always = cfg->fp64_mode == PHMM_FP64_ALWAYS;
    }
    if (const char* env = std::getenv("PHMM_B200_CALLER_RUNS")) e->caller_runs = std::atoi(env) != 0;
    std::unique_ptr<PhmmTables> t(new PhmmTables());
    phmm_build_tables(t.get());
    s.log10_mode = phmm_log10_pick_mode();
    {
        const uint32_t want = cfg ? cfg->epilogue : PHMM_EPILOGUE_AUTO;
        const bool pinned_libm = std::getenv("PHMM_B200_LIBM_LOG10F") != nullptr;
        const bool ok = want != PHMM_EPILOGUE_HOST && !pinned_libm && phmm_log10_restated_matches_libm();
        if (want == PHMM_EPILOGUE_DEVICE && !ok)
            return fail(PHMM_ERR_ARG, "PHMM_EPILOGUE_DEVICE: the restated log10f does not give this host's libm bits (or PHMM_B200_LIBM_LOG10F is set)");
        s.device_epilogue = ok;
        s.device_epilogue_min_pairs = want == PHMM_EPILOGUE_DEVICE ? 0 : 32768;
    }
    s.default_gapm = t->gapm[10];
    s.default_ii = t->ph2pr[10];
    s.default_gcp_d = t->ph2pr_d[10];
    s.default_gapm_d = 1.0 - t->ph2pr_d[10];                  // what RowParams<double> computes on the device: one IEEE subtraction
    int rc = PHMM_OK;
    for (size_t k = 0; k < ords.size() && rc == PHMM_OK; ++k) {
        e->devs.emplace_back(new PhmmDevice());
        rc = e->devs.back()->init(&e->cfg, ords[k], (uint32_t)k, *t);
    }
    cudaSetDevice(caller_dev);
    if (rc != PHMM_OK) {
        const std::string keep = g_phmm_err;
        for (auto& d : e->devs) d->shutdown();
        g_phmm_err = keep;
        return rc;
    }
    *out = e.release();
    return PHMM_OK;
}

int phmm_create(const phmm_config* cfg, phmm_engine** out)
{
    const int32_t dev = cfg ? cfg->device : -1;
    return phmm_create_multi(cfg, &dev, 1, out);
}

uint32_t phmm_device_count(const phmm_engine* e) { return e ? (uint32_t)e->devs.size() : 0; }

int phmm_device_ordinal(const phmm_engine* e, uint32_t index)
{
    if (!e || index >= e->devs.size()) return -1;
    return e->devs[index]->ordinal;
}

void phmm_destroy(phmm_engine* e)
{
    if (!e) return;
    {
        std::lock_guard<std::mutex> lk(e->bm);
        e->seal_locked();
    }
    for (auto& d : e->devs) d->shutdown();
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
    r.flat = nullptr;
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
    if (e->open.cells >= e->cfg.batch_cells) e->seal_locked();
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
    std::function<void()> mine;
    {
        std::unique_lock<std::mutex> lk(e->bm);
        auto it = e->batch_state.find(ticket);
        if (it == e->batch_state.end()) return fail(PHMM_ERR_STATE, "phmm_wait: unknown ticket");
        sync = it->second.sync;
        if (e->open.id == ticket && !e->open.regs.empty()) {
            if (e->in_flight < e->depth()) {
                mine = e->seal_locked(e->caller_runs);              // the pipeline is starving: go now, on this thread if it may
            } else {
                // busy: keep collecting returning callers; a completing batch, the size limit or the deadline seals it
                ++e->open.waiters;
                const auto deadline = e->open.first + std::chrono::microseconds(e->cfg.linger_us);
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
    if (mine) mine();
    if (!sync->done.load(std::memory_order_acquire)) {
        // (polling here for a while before blocking was measured: no gain for 1-8 callers, -12 % at 32 on 16 cores)
        std::unique_lock<std::mutex> lk(sync->m);
        sync->cv.wait(lk, [&] { return sync->done.load(std::memory_order_acquire); });
    }
    const int rc = sync->rc;
    if (rc != PHMM_OK) g_phmm_err = sync->err;
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
    // again at the end: the GPU starts after one small plan and only a small epilogue is left when it finishes.  Every
    // sub-batch goes to the device with the least outstanding work at that moment.
    const std::vector<uint64_t> pre = flat_work_prefix(fb);
    const std::vector<std::pair<uint32_t, uint32_t>> parts = flat_partition(pre, fb->n_regions, e->cfg.batch_cells, e->total_workers());
    std::mutex m;
    std::condition_variable cv;
    size_t remaining = parts.size();
    int rc_all = PHMM_OK;
    std::string err_all;
    for (const auto& part : parts) {
        PhmmDevice& d = e->pick_device();
        d.enqueue([&, part, dp = &d](Slot& s) {
            std::vector<phmm_hap> haps;
            std::vector<RegionIn> regs;
            flat_to_regions(fb, fr, part.first, part.second, haps, regs);
            const int rc = dp->execute(s, regs);
            std::lock_guard<std::mutex> lk(m);
            if (rc != PHMM_OK && rc_all == PHMM_OK) {
                rc_all = rc;
                err_all = g_phmm_err;
            }
            if (--remaining == 0) cv.notify_all();
        }, pre[part.second] - pre[part.first]);
    }
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return remaining == 0; });
    }
    if (rc_all != PHMM_OK) g_phmm_err = err_all;
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

int phmm_resident_upload_on(phmm_engine* e, uint32_t device_index, const phmm_flat_batch* fb, phmm_resident** out)
{
    if (!e || !fb || !out) return fail(PHMM_ERR_ARG, "phmm_resident_upload: null argument");
    if (device_index >= e->devs.size()) return fail(PHMM_ERR_ARG, "phmm_resident_upload: device index out of range");
    PhmmDevice& d = *e->devs[device_index];
    CU(cudaSetDevice(d.ordinal));
    std::unique_ptr<phmm_resident> r(new phmm_resident());
    r->dev = &d;
    int rc = r->slot.init();
    if (rc != PHMM_OK) return rc;
    r->fb = *fb;
    flat_to_regions(&r->fb, nullptr, 0, fb->n_regions, r->haps, r->regs);
    r->slot.force_dev_epi = true;
    if ((rc = d.plan(r->slot, r->regs)) != PHMM_OK) return rc;
    CU(cudaMemcpyAsync(r->slot.d_in.p, r->slot.h_in.p, r->slot.staged_bytes, cudaMemcpyHostToDevice, r->slot.stream));
    for (const Slot::DirectCopy& c : r->slot.direct)
        CU(cudaMemcpyAsync(static_cast<uint8_t*>(r->slot.d_in.p) + c.dst, c.src, c.len, cudaMemcpyHostToDevice, r->slot.stream));
    CU(cudaStreamSynchronize(r->slot.stream));
    d.stats.h2d += r->slot.in_bytes;
    *out = r.release();
    return PHMM_OK;
}

int phmm_resident_upload(phmm_engine* e, const phmm_flat_batch* fb, phmm_resident** out) { return phmm_resident_upload_on(e, 0, fb, out); }

int phmm_resident_run(phmm_engine* e, phmm_resident* r, float* gpu_ms, float* fwd_ms)
{
    if (!e || !r) return fail(PHMM_ERR_ARG, "phmm_resident_run: null argument");
    CU(cudaSetDevice(r->dev->ordinal));
    int rc = r->dev->launch(r->slot, false);
    if (rc != PHMM_OK) {
        r->dev->drain(r->slot);
        return rc;
    }
    CU(cudaStreamSynchronize(r->slot.stream));
    CU(cudaEventElapsedTime(&r->slot.gpu_ms, r->slot.ev[0], r->slot.ev[3]));
    CU(cudaEventElapsedTime(&r->slot.fwd_ms, r->slot.ev[1], r->slot.ev[2]));
    if (gpu_ms) *gpu_ms = r->slot.gpu_ms;
    if (fwd_ms) *fwd_ms = r->slot.fwd_ms;
    r->ran = true;
    return PHMM_OK;
}

// AlleleLikelihoods::marginalize for many variant sites on the matrices a resident batch left on the device.
int phmm_resident_marginalize(phmm_engine* e, phmm_resident* r, const phmm_marginal_site* sites, uint32_t n_sites, const uint64_t* out_off,
                              double* out)
{
    if (!e || !r || (n_sites && (!sites || !out_off || !out))) return fail(PHMM_ERR_ARG, "phmm_resident_marginalize: null argument");
    if (n_sites == 0) return PHMM_OK;
    Slot& s = r->slot;
    if (!r->ran) return fail(PHMM_ERR_STATE, "phmm_resident_marginalize: the batch has not been run (phmm_resident_run)");
    if (!s.dev_epi)
        return fail(PHMM_ERR_STATE, "phmm_resident_marginalize: this engine keeps the epilogue on the host (PHMM_EPILOGUE_HOST, or a libm whose "
                                    "log10f the device cannot reproduce): download the matrices and use phmm_marginalize");
    CU(cudaSetDevice(r->dev->ordinal));
    // sites + maps in one staging block, outputs behind them
    size_t n_map = 0, n_out = 0;
    for (uint32_t k = 0; k < n_sites; ++k) {
        if (sites[k].region >= s.regions.size()) return fail(PHMM_ERR_ARG, "phmm_resident_marginalize: region index out of range");
        const RegionPlan& rp = s.regions[sites[k].region];
        if (rp.n_haps && !sites[k].old_to_new) return fail(PHMM_ERR_ARG, "phmm_resident_marginalize: old_to_new is null");
        for (uint32_t a = 0; a < rp.n_haps; ++a)
            if (sites[k].old_to_new[a] >= 0 && (uint32_t)sites[k].old_to_new[a] >= sites[k].n_new)
                return fail(PHMM_ERR_ARG, "phmm_resident_marginalize: old_to_new names an allele past n_new");
        n_map += rp.n_haps;
        n_out = std::max<size_t>(n_out, out_off[k] + (size_t)sites[k].n_new * rp.n_reads);
    }
    const size_t o_sites = 0, o_maps = align_up(n_sites * sizeof(PhmmMargSite), 256), o_out = align_up(o_maps + n_map * sizeof(int32_t), 256);
    const size_t total = o_out + std::max<size_t>(n_out, 1) * sizeof(double);
    int rc;
    if ((rc = r->h_marg.ensure(total)) != PHMM_OK) return rc;
    if ((rc = r->d_marg.ensure(total)) != PHMM_OK) return rc;
    uint8_t* h = static_cast<uint8_t*>(r->h_marg.p);
    PhmmMargSite* hs = reinterpret_cast<PhmmMargSite*>(h + o_sites);
    int32_t* hm = reinterpret_cast<int32_t*>(h + o_maps);
    size_t at = 0;
    for (uint32_t k = 0; k < n_sites; ++k) {
        const RegionPlan& rp = s.regions[sites[k].region];
        hs[k] = PhmmMargSite{sites[k].region, sites[k].n_new, (uint32_t)at, 0u, out_off[k]};
        if (rp.n_haps) std::memcpy(hm + at, sites[k].old_to_new, rp.n_haps * sizeof(int32_t));
        at += rp.n_haps;
    }
    uint8_t* d = static_cast<uint8_t*>(r->d_marg.p);
    CU(cudaMemcpyAsync(d, h, o_out, cudaMemcpyHostToDevice, s.stream));
    CU(phmm_launch_marginalize(s.epi, reinterpret_cast<const PhmmMargSite*>(d + o_sites), reinterpret_cast<const int32_t*>(d + o_maps), n_sites,
                               reinterpret_cast<double*>(d + o_out), s.stream));
    CU(cudaMemcpyAsync(h + o_out, d + o_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    // only what the sites own is handed back (gaps between the sites' blocks, if the caller left any, are not touched)
    const double* ho = reinterpret_cast<const double*>(h + o_out);
    for (uint32_t k = 0; k < n_sites; ++k) {
        const size_t n = (size_t)sites[k].n_new * s.regions[sites[k].region].n_reads;
        std::memcpy(out + out_off[k], ho + out_off[k], n * sizeof(double));
    }
    r->dev->stats.launches += 1;
    r->dev->stats.h2d += o_out;
    r->dev->stats.d2h += n_out * sizeof(double);
    return PHMM_OK;
}

int phmm_resident_download(phmm_engine* e, phmm_resident* r, phmm_flat_result* fr)
{
    if (!e || !r || !fr) return fail(PHMM_ERR_ARG, "phmm_resident_download: null argument");
    CU(cudaSetDevice(r->dev->ordinal));
    fr->n_fp64 = 0;
    for (size_t g = 0; g < r->regs.size(); ++g) {
        RegionIn& x = r->regs[g];
        const uint64_t r0 = x.flat_r0;
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
    return r->dev->finish(r->slot, r->regs, true);
}

void phmm_resident_free(phmm_engine* e, phmm_resident* r)
{
    if (!r) return;
    (void)e;
    if (r->dev) cudaSetDevice(r->dev->ordinal);
    r->d_marg.release();
    r->h_marg.release();
    r->slot.destroy();
    cudaStreamSynchronize(phmm_alloc_stream());
    delete r;
}

uint32_t phmm_resident_launches(const phmm_resident* r) { return r ? r->slot.n_launches : 0; }

int phmm_get_stats(phmm_engine* e, phmm_stats* out)
{
    if (!e || !out) return fail(PHMM_ERR_ARG, "phmm_get_stats: null argument");
    std::memset(out, 0, sizeof(*out));
    for (const auto& d : e->devs) add_stats(*d, out);
    return PHMM_OK;
}

int phmm_get_device_stats(phmm_engine* e, uint32_t device_index, phmm_stats* out)
{
    if (!e || !out) return fail(PHMM_ERR_ARG, "phmm_get_device_stats: null argument");
    if (device_index >= e->devs.size()) return fail(PHMM_ERR_ARG, "phmm_get_device_stats: device index out of range");
    std::memset(out, 0, sizeof(*out));
    add_stats(*e->devs[device_index], out);
    return PHMM_OK;
}

int phmm_kernel_info(uint32_t read_len, int fma_mode, int* regs, int* local_bytes, int* lanes, int* rows_per_lane)
{
    // what the planner picks for a big region of reads of this one length (the reference's own gap qualities, throughput policy)
    const bool fma = fma_mode != PHMM_FMA_NONE;
    if (read_len == 0 || read_len > phmm_max_read_len(fma))
        return fail(PHMM_ERR_ARG, "phmm_kernel_info: the wavefront kernel does not take reads of this length (per-pair kernel)");
    const int rpt = phmm_plan_rpt_for_uniform_reads(read_len, fma);
    PhmmKernelInfo ki{};
    CU(phmm_fwd_kernel_info(rpt, fma, fma, &ki));
    if (regs) *regs = ki.regs;
    if (local_bytes) *local_bytes = ki.local_bytes;
    if (lanes) *lanes = (int)phmm_read_lanes(read_len, rpt);
    if (rows_per_lane) *rows_per_lane = rpt;
    return PHMM_OK;
}

// Measured FP32 peak for the roofline denominator: runs the issue-rate micro-benchmark (phmm_kernels.cu) on every SM of the
// engine's first device and returns the sustained rate in FMA lane-operations per second (one FFMA2 counts 2 per lane).  Runs
// on a stream, events and buffer of its own: never on a worker's.
int phmm_ubench(phmm_engine* e, int mode, uint32_t iters, float* ms_out, double* fma_per_s)
{
    if (!e) return fail(PHMM_ERR_ARG, "phmm_ubench: null engine");
    PhmmDevice& d = *e->devs[0];
    std::lock_guard<std::mutex> lk(d.ub_m);
    CU(cudaSetDevice(d.ordinal));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, d.ordinal));
    const uint32_t blocks = (uint32_t)prop.multiProcessorCount * 8;
    int rc = d.ub_buf.ensure(256);
    if (rc != PHMM_OK) return rc;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(d.ub_ev[0], d.ub_stream));
        CU(phmm_launch_ubench(mode, blocks, iters, static_cast<float*>(d.ub_buf.p), d.ub_stream));
        CU(cudaEventRecord(d.ub_ev[1], d.ub_stream));
        CU(cudaStreamSynchronize(d.ub_stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, d.ub_ev[0], d.ub_ev[1]));
        if (rep > 0 && ms < best) best = ms;
    }
    if (ms_out) *ms_out = best;
    if (fma_per_s) *fma_per_s = (double)blocks * 256.0 * (double)iters * 32.0 * 2.0 / ((double)best * 1e-3);
    return PHMM_OK;
}

}  // extern "C"
