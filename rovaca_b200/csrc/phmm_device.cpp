// phmm_device.cpp — one GPU of the engine: the device side of a batch (H2D, prep / forward / per-pair / epilogue launches,
// D2H), the host side after it (the reference's log10 / normalize / filter / transpose, pairhmm_internal.cpp:578-584, :53-63,
// :161-193, when the device epilogue is not used), and the worker threads that drive batches through both.
#include "phmm_host.h"

// PHMM_B200_PHASES=1: where a batch's wall time goes (plan / CUDA calls of launch() / waiting for the device / epilogue and
// the rest of finish()), summed over the process and printed by the last device's shutdown — tools/caller_bench's latency table.
static const bool g_phases = std::getenv("PHMM_B200_PHASES") != nullptr;
static std::atomic<uint64_t> g_phase_ns[4], g_phase_batches;
thread_local uint64_t tl_wait_ns = 0;

// One allocation stream per device for the whole process (engines come and go; the stream is a few hundred bytes).
cudaStream_t phmm_alloc_stream()
{
    static std::mutex m;
    static std::vector<cudaStream_t> streams;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return nullptr;
    std::lock_guard<std::mutex> lk(m);
    if (streams.size() <= (size_t)dev) streams.resize((size_t)dev + 1, nullptr);
    if (!streams[dev] && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        streams[dev] = nullptr;                    // the legacy default stream still gives stream-ordered allocations
    }
    return streams[dev];
}

// ------------------------------------------------------------------------------------------------
// Device side of one batch
// ------------------------------------------------------------------------------------------------

// The FP64 recompute of the pairs the forward kernels queued (pairhmm_internal.cpp:578-581).  The second half of the pair
// buffer is the fast kernel's hand-back list; its counter sits behind the first one.
cudaError_t PhmmDevice::launch_f64(Slot& s)
{
    PhmmPair* redo = static_cast<PhmmPair*>(s.d_pairs.p) + std::max<uint64_t>(s.n_pairs, 1);
    uint32_t* redo_count = s.dev.f64_count + 1;
    cudaError_t e = cudaMemsetAsync(redo_count, 0, 4, s.stream);
    if (e != cudaSuccess) return e;
    const bool general_only = std::getenv("PHMM_B200_FP64_GENERAL") != nullptr;              // tests: the general kernel alone (read per launch)
    const char* trip = std::getenv("PHMM_B200_FP64_TRIP_HI");                                  // tests: a higher floor forces hand-backs
    const uint32_t trip_hi = trip ? (uint32_t)std::strtoul(trip, nullptr, 0) : PHMM_PAIR64_TRIP_HI;
    return phmm_launch_pairs_f64(s.dev, s.dev.f64_pairs, s.dev.f64_count, static_cast<double*>(s.d_scr64.p), kPairThreads, s.max_cols,
                                 general_only ? nullptr : redo, redo_count, cfg->default_gapm_d, cfg->default_gcp_d, trip_hi, s.stream);
}

int PhmmDevice::launch(Slot& s, bool copy_in)
{
    if (copy_in && s.in_bytes) {
        CU(cudaMemcpyAsync(s.d_in.p, s.h_in.p, s.staged_bytes, cudaMemcpyHostToDevice, s.stream));
        for (const Slot::DirectCopy& c : s.direct)
            CU(cudaMemcpyAsync(static_cast<uint8_t*>(s.d_in.p) + c.dst, c.src, c.len, cudaMemcpyHostToDevice, s.stream));
        stats.h2d += s.in_bytes;
    }
    // (the prep kernel empties the list of underflowing pairs: one call less per batch)
    // PHMM_B200_POISON=1 (the GPU tests set it): every raw sum starts as NaN, so a pair no kernel wrote is loud instead of
    // keeping a stale value from an earlier batch
    static const bool poison = std::getenv("PHMM_B200_POISON") != nullptr;
    if (poison && s.n_pairs) CU(cudaMemsetAsync(s.dev.out_f32, 0xFF, s.n_pairs * sizeof(float), s.stream));
    CU(cudaEventRecord(s.ev[0], s.stream));
    CU(phmm_launch_prep(s.dev, s.stream));
    if (!s.lite) CU(cudaEventRecord(s.ev[1], s.stream));
    // One launch per read-length bin.  A batch of mixed lengths has many of them, none big enough to fill the GPU to the
    // end of its last wave, so they go out on up to four streams (forked after the prep kernel, joined before the FP64 pass).
    const size_t n_side = std::min<size_t>(Slot::kAux, s.launches.size() > 1 ? s.launches.size() - 1 : 0);
    for (size_t a = 0; a < n_side; ++a) CU(cudaStreamWaitEvent(s.aux[a], s.ev[1], 0));
    for (size_t i = 0; i < s.launches.size(); ++i) {
        const FwdLaunch& fl = s.launches[i];
        const size_t lane = i % (n_side + 1);
        cudaStream_t st = lane == 0 ? s.stream : s.aux[lane - 1];
        const cudaError_t le = phmm_launch_fwd(s.dev, fl.rpt, fl.refgaps, cfg->fma, fl.fixed_lanes, fl.duo, fl.cta_base, fl.n_ctas, fl.cta_warps, fl.smem, st);
        if (le != cudaSuccess)
            return fail(PHMM_ERR_CUDA, std::string("forward kernel launch (") + std::to_string(fl.rpt) + " rows per lane, " +
                                           std::to_string(fl.n_ctas) + " CTAs of " + std::to_string(fl.cta_warps) + " warps, " +
                                           std::to_string(fl.smem) + " B shared): " + cudaGetErrorString(le));
    }
    for (size_t a = 0; a < n_side; ++a) {
        CU(cudaEventRecord(s.aux_done[a], s.aux[a]));
        CU(cudaStreamWaitEvent(s.stream, s.aux_done[a], 0));
    }
    if (!s.lite) CU(cudaEventRecord(s.ev[2], s.stream));
    if (s.n_slow)
        CU(phmm_launch_pairs_f32(s.dev, cfg->fma, s.dev.slow_pairs, s.n_slow, static_cast<float*>(s.d_scr32.p), kPairThreads, s.max_cols,
                                 s.stream));
    s.f64_done = s.want_f64 && s.n_pairs;
    if (s.f64_done)
        CU(launch_f64(s));
    if (s.dev_epi && s.n_pairs) CU(phmm_launch_epilogue(s.epi, s.stream));
    CU(cudaEventRecord(s.ev[3], s.stream));
    stats.launches += s.n_launches;
    return PHMM_OK;
}

// D2H + the reference's host epilogue.
int PhmmDevice::finish(Slot& s, const std::vector<RegionIn>& regs, bool copy_out)
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
    const bool need_f32 = (!dev || asks_raw) && !s.host_out;       // host_out: the kernels wrote the raw sums into h_f32 themselves
    if (copy_out && s.n_pairs && !s.host_out) {
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
        const auto t_wait = std::chrono::steady_clock::now();
        constexpr int poll_us = 400;                // (no polling, or blocking syncs, measured the same or worse with 16-128 callers)
        if (s.n_cells < 1500000000ull) {            // < ~0.5 ms of device work; big batches sleep instead of burning a core
            const auto t_poll = std::chrono::steady_clock::now();
            while ((q = cudaStreamQuery(s.stream)) == cudaErrorNotReady) {
                if (std::chrono::steady_clock::now() - t_poll > std::chrono::microseconds(poll_us)) break;
                _mm_pause();
            }
            if (q != cudaSuccess && q != cudaErrorNotReady) return fail(PHMM_ERR_CUDA, std::string("cudaStreamQuery: ") + cudaGetErrorString(q));
        }
        if (q == cudaErrorNotReady) CU(cudaStreamSynchronize(s.stream));
        if (g_phases) tl_wait_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_wait).count();
    }
    // The FP64 sums are read only by a caller that asked for them (raw_f64) or under PHMM_FP64_ALWAYS; see `have64` below.
    bool have64 = s.f64_done;
    if (s.n_pairs && !s.host_out) {
        n64 = *static_cast<uint32_t*>(s.h_cnt.p);
        if (n64 > 0 && !have64) {
            if (asks_f64) {                                     // a resident batch learns at download time who wants raw sums
                CU(launch_f64(s));
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
    if (s.lite)
        s.fwd_ms = s.gpu_ms;                                   // (prep included: a few microseconds)
    else
        CU(cudaEventElapsedTime(&s.fwd_ms, s.ev[1], s.ev[2]));

    const auto t_epi = std::chrono::steady_clock::now();
    if (dev) {
        epilogue_from_device(s, regs, have64);
        stats.dev_epilogues += 1;
    }
    else {
        const uint64_t under = epilogue(s, regs, static_cast<const float*>(s.h_f32.p), static_cast<const double*>(s.h_f64.p), have64);
        if (s.host_out) n64 = (uint32_t)under;                 // the device's count stayed on the device
    }
    stats.regions += regs.size();
    stats.pairs += s.n_pairs;
    stats.cells += s.n_cells;
    stats.fp64_pairs += n64;
    stats.slow_pairs += s.n_slow;
    stats.batches += 1;
    add_span(s);
    {
        std::lock_guard<std::mutex> lk2(stats.m);
        stats.epilogue_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_epi).count();
    }
    return PHMM_OK;
}

// The epilogue ran on the device (phmm_epilogue_kernel): hand every region its matrix, surviving reads and best haplotypes.
void PhmmDevice::epilogue_from_device(const Slot& s, const std::vector<RegionIn>& regs, bool have64) const
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
uint64_t PhmmDevice::epilogue(const Slot& s, const std::vector<RegionIn>& regs, const float* f32, const double* f64, bool have64) const
{
    uint64_t under = 0;                            // pairs below MIN_ACCEPTED in this batch
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
        const uint32_t reg64 = (uint32_t)phmm_log10_block(fr, lk.data(), np, cfg->log10_mode);
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
                                                      r.out.kept_idx, cfg->log10_mode);
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
        under += reg64;
    }
    _mm_setcsr(old_csr);
    return under;
}

// After an error part-way through a batch: copies or kernels of this slot may still be reading its pinned arena and its
// device buffers; nothing may be in flight when the worker reuses the slot for the next plan().
void PhmmDevice::drain(Slot& s)
{
    if (s.h_in.host_only) return;
    if (s.stream) cudaStreamSynchronize(s.stream);
    for (auto& a : s.aux)
        if (a) cudaStreamSynchronize(a);
    cudaGetLastError();                                    // the error has been reported through fail(); do not leave it sticky
}

// Batches of different workers overlap on the device, so the device time of N batches is the union of their kernel spans,
// not the sum (phmm_stats.gpu_ms).  `epoch` marks the device time up to which spans have been counted; a finished batch adds
// what it has beyond that and hands over its end event as the new mark (the slot takes the old one for its next batch), so
// every difference taken is a short one and float milliseconds stay exact however long the engine lives.
void PhmmDevice::add_span(Slot& s)
{
    std::lock_guard<std::mutex> lk(span_m);
    float t0 = 0.f, t1 = 0.f;
    if (cudaEventElapsedTime(&t0, epoch, s.ev[0]) != cudaSuccess || cudaEventElapsedTime(&t1, epoch, s.ev[3]) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (t1 <= 0.f) return;                                 // wholly inside what has been counted
    std::swap(epoch, s.ev[3]);
    std::lock_guard<std::mutex> lk2(stats.m);
    stats.gpu_ms += (double)t1 - (double)std::max(t0, 0.f);
}

int PhmmDevice::execute(Slot& s, const std::vector<RegionIn>& regs)
{
    int rc;
    const auto t0 = std::chrono::steady_clock::now();
    busy_batches.fetch_add(1, std::memory_order_relaxed);
    s.small_ok = true;
    {
        PhmmRange r("phmm plan");
        rc = plan(s, regs);
    }
    if (rc != PHMM_OK) {
        busy_batches.fetch_sub(1, std::memory_order_relaxed);
        return rc;
    }
    const auto t1 = std::chrono::steady_clock::now();
    {
        PhmmRange r("phmm launch");
        rc = launch(s, true);
    }
    const auto t2 = std::chrono::steady_clock::now();
    if (rc == PHMM_OK) {
        PhmmRange r("phmm finish");
        rc = finish(s, regs, true);
    }
    if (g_phases) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ns = [](auto a, auto b) { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
        g_phase_ns[0] += ns(t0, t1);
        g_phase_ns[1] += ns(t1, t2);
        g_phase_ns[2] += tl_wait_ns;
        g_phase_ns[3] += ns(t2, t3) - tl_wait_ns;
        g_phase_batches += 1;
    }
    if (rc != PHMM_OK) {
        const std::string keep = g_phmm_err;
        drain(s);
        g_phmm_err = keep;
    }
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

void PhmmDevice::worker_main(uint32_t idx)
{
    cudaSetDevice(ordinal);
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
        // Arenas grow to the biggest size any slot of the device has needed (pinning host pages takes milliseconds): a worker with
        // nothing queued does that now, not under the first big batch it is handed.
        bool idle;
        {
            std::lock_guard<std::mutex> lk(qm);
            idle = jobs.empty() && !stopping;
        }
        if (idle) catch_up(s);
    }
}

void PhmmDevice::catch_up(Slot& s)
{
    s.h_in.catch_up(high_water[0]);
    s.d_in.catch_up(high_water[0]);
    s.d_words.catch_up(high_water[1]);
    s.d_f32.catch_up(high_water[2]);
    s.h_f32.catch_up(high_water[2]);
    s.d_f64.catch_up(high_water[3]);
    s.h_f64.catch_up(high_water[3]);
    s.d_pairs.catch_up(high_water[4]);
    s.d_scr64.catch_up(high_water[5]);
    if (s.d_epi.cap) {                                 // (a slot that has never run a device epilogue keeps none)
        s.d_epi.catch_up(high_water[7]);
        s.h_epi.catch_up(high_water[7]);
    }
}

Slot* PhmmDevice::take_caller_slot()
{
    for (uint32_t i = 0; i < kCallerSlots; ++i) {
        if (!caller_slots[i]) continue;
        bool idle = false;
        if (caller_slot_busy[i].compare_exchange_strong(idle, true, std::memory_order_acquire)) return caller_slots[i].get();
    }
    return nullptr;
}

void PhmmDevice::return_caller_slot(Slot* s)
{
    for (uint32_t i = 0; i < kCallerSlots; ++i)
        if (caller_slots[i].get() == s) caller_slot_busy[i].store(false, std::memory_order_release);
}

// `work` (weighted cells, flat_work_prefix's unit) counts as outstanding on this device from now until the job has run: the
// front sends every new batch to the device with the least of it.
void PhmmDevice::enqueue(std::function<void(Slot&)> fn, uint64_t work)
{
    outstanding.fetch_add(work, std::memory_order_relaxed);
    {
        std::lock_guard<std::mutex> lk(qm);
        jobs.push_back([this, work, fn = std::move(fn)](Slot& s) {
            fn(s);
            outstanding.fetch_sub(work, std::memory_order_relaxed);
        });
    }
    qcv.notify_one();
}

int PhmmDevice::init(const PhmmSettings* settings, int cuda_ordinal, uint32_t idx, const PhmmTables& tables)
{
    cfg = settings;
    ordinal = cuda_ordinal;
    index = idx;
    CU(cudaSetDevice(ordinal));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ordinal));
    if (prop.major != 10)
        return fail(PHMM_ERR_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                       "; this library carries sm_100a code only");
    warp_slots = (uint32_t)prop.multiProcessorCount * 8;
    smem_optin = prop.sharedMemPerBlockOptin;
    smem_sm = prop.sharedMemPerMultiprocessor;
    int rc = d_tables.ensure(sizeof(PhmmTables));
    if (rc != PHMM_OK) return rc;
    CU(cudaMemcpy(d_tables.p, &tables, sizeof(PhmmTables), cudaMemcpyHostToDevice));
    CU(cudaEventCreate(&epoch));
    CU(cudaEventRecord(epoch, nullptr));
    CU(cudaEventSynchronize(epoch));
    CU(cudaStreamCreateWithFlags(&ub_stream, cudaStreamNonBlocking));
    for (auto& e : ub_ev) CU(cudaEventCreate(&e));
    for (uint32_t i = 0; i < cfg->n_workers; ++i) {
        slots.emplace_back(new Slot());
        if ((rc = slots.back()->init()) != PHMM_OK) return rc;
    }
    for (uint32_t i = 0; i < std::min<uint32_t>(kCallerSlots, std::max<uint32_t>(cfg->pipeline_depth, 1)); ++i) {
        caller_slots[i].reset(new Slot());
        if ((rc = caller_slots[i]->init()) != PHMM_OK) return rc;
    }
    for (uint32_t i = 0; i < cfg->n_workers; ++i) workers.emplace_back([this, i] { worker_main(i); });
    return PHMM_OK;
}

void PhmmDevice::shutdown()
{
    {
        std::lock_guard<std::mutex> lk(qm);
        stopping = true;
    }
    qcv.notify_all();
    for (auto& t : workers) t.join();
    workers.clear();
    cudaSetDevice(ordinal);
    for (auto& s : slots) s->destroy();
    slots.clear();
    for (auto& s : caller_slots)
        if (s) {
            s->destroy();
            s.reset();
        }
    for (auto& ss : sw.slot) {
        for (auto* b : {&ss.h_in, &ss.h_out}) b->release();
        for (auto* b : {&ss.d_in, &ss.d_bt, &ss.d_aux, &ss.d_cig, &ss.d_out}) b->release();
        for (auto& ev : ss.ev)
            if (ev) cudaEventDestroy(ev);
        if (ss.stream) cudaStreamDestroy(ss.stream);
        ss.stream = nullptr;
    }
    if (sw.base) cudaEventDestroy(sw.base);
    sw.base = nullptr;
    ub_buf.release();
    for (auto& e : ub_ev)
        if (e) cudaEventDestroy(e);
    if (ub_stream) cudaStreamDestroy(ub_stream);
    if (epoch) cudaEventDestroy(epoch);
    d_tables.release();
    cudaStreamSynchronize(phmm_alloc_stream());    // the arenas go back to the device now, not at some later synchronisation
    if (g_phases && g_phase_batches.load()) {
        const double n = (double)g_phase_batches.exchange(0);
        std::fprintf(stderr, "{\"phases_us_per_batch\": {\"plan\": %.1f, \"launch_calls\": %.1f, \"wait_for_device\": %.1f, \"finish_and_epilogue\": %.1f}, \"batches\": %.0f}\n",
                     g_phase_ns[0].exchange(0) / n / 1e3, g_phase_ns[1].exchange(0) / n / 1e3, g_phase_ns[2].exchange(0) / n / 1e3,
                     g_phase_ns[3].exchange(0) / n / 1e3, n);
    }
}
