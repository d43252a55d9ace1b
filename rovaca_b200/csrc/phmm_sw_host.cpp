// phmm_sw_host.cpp — host side of the batched Smith-Waterman (include/phmm_b200.h; kernels in sw_kernels.cu): the step right
// after the path, read -> best-haplotype realignment.
#include "phmm_host.h"

extern "C" {

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
    constexpr size_t kBtBudget = (size_t)6 << 30;                        // back-track words one sub-batch may hold on the device
    constexpr uint32_t kSubBatch = 32768;                                // alignments per sub-batch: several per big call, so that packing,
                                                                         // kernels and the hand-back of results overlap
    const bool wide_only = std::getenv("PHMM_B200_SW_WIDE") != nullptr;   // tests: the 32-bit fill kernel for everything
    std::lock_guard<std::mutex> lk(e->sw_m);
    // sub-batches go round the devices' slots (two per device): the host packs one while the others run
    struct Lane { PhmmDevice* d; SwSlot* ss; };
    std::vector<Lane> ring;
    for (int k = 0; k < 2; ++k)
        for (auto& d : e->devs) ring.push_back(Lane{d.get(), &d->sw.slot[k]});
    if (n_pairs <= kSubBatch) ring.resize(1);
    for (auto& d : e->devs) {
        SwContext& c = d->sw;
        CU(cudaSetDevice(d->ordinal));
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
    }

    // results of a finished sub-batch -> the caller's buffers
    auto finish = [&](Lane& ln) -> int {
        SwSlot& ss = *ln.ss;
        SwContext& c = ln.d->sw;
        if (!ss.busy) return PHMM_OK;
        ss.busy = false;
        CU(cudaSetDevice(ln.d->ordinal));
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
        ln.d->stats.launches += 2;
        ln.d->stats.h2d += ss.in_bytes;
        ln.d->stats.d2h += (size_t)n * (sizeof(PhmmSwOut) + (size_t)ss.cigar_cap * 4);
        return PHMM_OK;
    };

    uint32_t at = 0, turn = 0;
    int rc = PHMM_OK;
    while (at < n_pairs && rc == PHMM_OK) {
        Lane& ln = ring[turn++ % ring.size()];
        SwSlot& ss = *ln.ss;
        if ((rc = finish(ln)) != PHMM_OK) break;                          // this slot's previous sub-batch: its kernels ran while the later ones were packed
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
            {
                const int big = std::max(std::max(std::abs(params->match), std::abs(params->mismatch)), 1);
                const long bound = (long)std::min(p.ref_len, p.read_len) * big + std::abs(params->gap_open) + std::abs(params->gap_extend);
                t.narrow = (!wide_only && cpt <= 32 && params->match > params->mismatch && bound <= kPhmmSwNarrowBound) ? 1u : 0u;
            }
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
            uint32_t count[2 * kPhmmSwNumCpt + 1] = {};               // groups: (cpt, 32-bit) then (cpt, 16-bit) per cpt
            auto group_of = [&](const PhmmSwTask& t) { int q = 0; while (kPhmmSwCpt[q] != (int)t.cpt) ++q; return 2 * q + (int)t.narrow; };
            for (uint32_t k = 0; k < n; ++k) ++count[group_of(tasks[k]) + 1];
            for (int q = 0; q < 2 * kPhmmSwNumCpt; ++q) count[q + 1] += count[q];
            for (uint32_t k = 0; k < n; ++k) order[count[group_of(tasks[k])]++] = k;
        }
        size_t off = 0;
        auto place = [&](size_t nbytes) { size_t o = off; off = align_up(off + nbytes, 256); return o; };
        const size_t o_tasks = place(n * sizeof(PhmmSwTask)), o_order = place(n * sizeof(uint32_t)), o_bytes = place(bytes + 16);
        if (cudaSetDevice(ln.d->ordinal) != cudaSuccess) { rc = fail(PHMM_ERR_CUDA, "cudaSetDevice failed"); break; }
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
            CU(cudaSetDevice(ln.d->ordinal));
            CU(cudaMemcpyAsync(ss.d_in.p, ss.h_in.p, off, cudaMemcpyHostToDevice, ss.stream));
            CU(cudaEventRecord(ss.ev[0], ss.stream));
            for (uint32_t g0 = 0; g0 < n;) {                             // one fill launch per (cpt, width) group
                uint32_t g1 = g0;
                const PhmmSwTask& lead = tasks[order[g0]];
                while (g1 < n && tasks[order[g1]].cpt == lead.cpt && tasks[order[g1]].narrow == lead.narrow) ++g1;
                if (lead.narrow) {
                    const uint32_t warps = std::min<uint32_t>((g1 - g0 + 1) / 2, ln.d->warp_slots * 4);
                    CU(phmm_launch_sw_fill16(b, (int)lead.cpt, g0, g1 - g0, warps, ss.stream));
                } else {
                    const uint32_t warps = std::min<uint32_t>(g1 - g0, ln.d->warp_slots * 4);
                    CU(phmm_launch_sw_fill(b, (int)lead.cpt, g0, g1 - g0, warps, ss.stream));
                }
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
        for (Lane& ln : ring) {
            if (ln.ss->busy) {
                cudaSetDevice(ln.d->ordinal);
                cudaStreamSynchronize(ln.ss->stream);
            }
            ln.ss->busy = false;
        }
        return rc;
    }
    // drain: oldest first
    for (size_t k = 0; k < ring.size(); ++k)
        if ((rc = finish(ring[(turn + k) % ring.size()])) != PHMM_OK) return rc;
    // the devices ran side by side: the call's kernel time is the longest device's union of spans
    e->sw_last_ms = 0.f;
    e->sw_last_cells = 0;
    for (auto& d : e->devs) {
        e->sw_last_ms = std::max(e->sw_last_ms, d->sw.last_ms);
        e->sw_last_cells += d->sw.last_cells;
    }
    return PHMM_OK;
}

int phmm_sw_last_timing(phmm_engine* e, float* gpu_ms, uint64_t* cells)
{
    if (!e) return fail(PHMM_ERR_ARG, "phmm_sw_last_timing: null engine");
    std::lock_guard<std::mutex> lk(e->sw_m);
    if (gpu_ms) *gpu_ms = e->sw_last_ms;
    if (cells) *cells = e->sw_last_cells;
    return PHMM_OK;
}

}  // extern "C"
