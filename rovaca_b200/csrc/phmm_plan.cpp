// phmm_plan.cpp — the planner: a list of regions -> read-length bins, haplotype chunks of two balanced column streams, CTAs,
// and the one pinned arena that carries all of it to the device.  Replaces the reference's 16-lane CPU
// grouping (generate_histogram / generate_new_pairhmm_idx, pairhmm_internal.cpp:231-391), on which results do not depend.
#include "phmm_host.h"

#include <array>

namespace
{

// What a padded row costs at `rpt` rows per lane: 1000 / (fraction of the FP32 roofline the forward kernel reaches on reads that
// fill their lanes exactly), measured on a B200 (profiles/r02_rpt_curve.txt: 64 reads x 8 haplotypes of 300, reads of
// 8 x rpt - 1 bases).  Few rows per lane pay a step's shuffles, loads and loop for little arithmetic; 12 rows is the sweet
// spot (165 registers: three CTAs, i.e. three warps per scheduler); from 13 up the register file holds two warps per scheduler.
// A region costs warps x steps x rpt x this.
constexpr uint32_t kRowCost[PHMM_MAX_RPT + 1] = {0,    4680, 2188, 1603, 1531, 1456, 1404, 1344, 1307, 1310,
                                                 1282, 1294, 1229, 1314, 1312, 1308, 1289, 1284, 1289, 1274};
inline uint32_t step_cost(int rpt) { return (uint32_t)rpt * kRowCost[rpt]; }
// Rows-per-lane counts the throughput policy chooses among: every distinct count in a batch is a launch of its own with its own
// tail, and neighbouring counts differ by a per cent or two — so a few well-separated ones.
constexpr int kRptChoices[] = {12, 16, 19};            // first choice first
constexpr int kRptChoicesGeneric[] = {12, 10, 8};
constexpr uint32_t kRptMargin = 6;

// Pack reads that need lanes_of[k] consecutive lanes each into warps of 32 lanes, biggest first into the first warp that still
// has room (by_lanes[n] = how many reads need n lanes).  Returns the warps needed; with `out`, also which reads go where:
// out_warp_count[w] = reads of warp w, order = read slots (indices into the region's reads) warp after warp.
struct LanePacker
{
    uint32_t count[33];
    uint32_t first[34];                        // final packing: where the read slots of every lane count start in `slots`
    std::vector<uint32_t> slots;               // read ids sorted by lane count (counting sort, input order kept)
    uint32_t pack(std::vector<uint32_t>* order, std::vector<uint8_t>* warp_count)
    {
        uint32_t left = 0, warps = 0;
        uint32_t cnt[33], cursor[33];
        int top = 0;
        for (int n = 1; n <= 32; ++n) {
            cnt[n] = count[n];
            cursor[n] = first[n];
            left += count[n];
            if (count[n]) top = n;
        }
        while (left) {
            uint32_t room = 32, in_warp = 0;
            while (top > 0 && cnt[top] == 0) --top;
            for (int n = top; n >= 1 && room; --n)
                while (cnt[n] && (uint32_t)n <= room) {
                    --cnt[n];
                    --left;
                    room -= (uint32_t)n;
                    ++in_warp;
                    if (order) order->push_back(slots[cursor[n]++]);
                }
            if (warp_count) warp_count->push_back((uint8_t)in_warp);
            ++warps;
        }
        return warps;
    }
};

}  // namespace

int phmm_plan_rpt_for_uniform_reads(uint32_t read_len, bool refgaps)
{
    const int max_rpt = phmm_max_rpt(refgaps), rpt_lo = (int)((read_len + 32) / 32);
    const int* choices = refgaps ? kRptChoices : kRptChoicesGeneric;
    int rpt = 0;
    uint64_t best = UINT64_MAX;
    for (int q = 2; q >= 0; --q) {
        const int c = std::max(rpt_lo, choices[q]);
        if (c > max_rpt || c == rpt) continue;
        const uint32_t per_warp = 32u / phmm_read_lanes(read_len, c);
        const uint64_t cost = (uint64_t)step_cost(c) * 1024 / std::max(per_warp, 1u);
        if (cost < best) { best = cost; rpt = c; }
    }
    return rpt ? rpt : max_rpt;
}

// ------------------------------------------------------------------------------------------------
// Planning + packing
// ------------------------------------------------------------------------------------------------

#ifndef PHMM_PLAN_TIMING      // section timers for tools/plan_timing.cpp; compiled out of the library
#define PT_NOW() 0
#define PT_ADD(i, a, b) (void)0
#endif

int PhmmDevice::plan(Slot& s, const std::vector<RegionIn>& regs)
{
    [[maybe_unused]] const auto pt0 = PT_NOW();
    uint64_t n_reads = 0, n_haps = 0, read_bytes = 0, hap_bytes = 0, n_pairs = 0, n_cells = 0;
    bool any_ins = false, any_del = false, any_gcp = false;
    s.want_f64 = cfg->fp64_always;
    s.want_best = s.want_raw = false;
    const RegionIn *flat_first = nullptr, *flat_last = nullptr;
    bool flat_contig = true;         // the batch's read bytes are one contiguous range of every blob of one flat batch
    for (const RegionIn& r : regs) {
        if (r.n_reads == 0 || r.n_haps == 0) {                   // pairhmm_internal.cpp:508-510
            flat_contig &= r.n_reads == 0;                       // reads that are skipped would sit inside the range
            continue;
        }
        s.want_f64 |= r.out.raw_f64 != nullptr;
        s.want_raw |= r.out.raw_f32 != nullptr || r.out.raw_f64 != nullptr;
        s.want_best |= r.out.best_hap != nullptr;
        uint64_t rb = 0, hb = 0;
        if (r.flat) {
            const uint32_t* len = r.flat->read_len + r.flat_r0;
            const uint64_t* off = r.flat->read_off + r.flat_r0;
            uint64_t gaps = 0;                                       // bytes between consecutive reads in the caller's blobs
            for (uint32_t k = 0; k < r.n_reads; ++k) {
                rb += len[k];
                gaps += (k + 1 < r.n_reads) ? off[k + 1] - (off[k] + len[k]) : 0;
            }
            any_ins |= r.flat->read_ins != nullptr;
            any_del |= r.flat->read_del != nullptr;
            any_gcp |= r.flat->read_gcp != nullptr;
            // one run per blob for the whole batch?  Regions of one flat batch, consecutive, their reads back to back
            if (flat_first == nullptr) {
                flat_first = &r;
            } else {
                flat_contig &= r.flat == flat_last->flat && r.flat_r0 == flat_last->flat_r0 + flat_last->n_reads &&
                               off[0] == flat_last->flat->read_off[r.flat_r0 - 1] + flat_last->flat->read_len[r.flat_r0 - 1];
            }
            flat_contig &= gaps == 0;
            flat_last = &r;
        } else {
            flat_contig = false;
            for (uint32_t k = 0; k < r.n_reads; ++k) {
                rb += r.reads[k].len;
                any_ins |= r.reads[k].ins != nullptr;
                any_del |= r.reads[k].del != nullptr;
                any_gcp |= r.reads[k].gcp != nullptr;
            }
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
    s.dev_epi = cfg->device_epilogue && (n_pairs >= cfg->device_epilogue_min_pairs || s.force_dev_epi);
    if (n_pairs >= (1ull << 31) || read_bytes >= (1ull << 31) || hap_bytes >= (1ull << 31))
        return fail(PHMM_ERR_ARG, "batch too large for 32-bit device offsets; lower batch_cells");
    // What the reference itself feeds its kernels — deletion qualities aliasing the insertion ones, one gap-continuation
    // quality per read (read_record.cpp:191-210, pairhmm_internal.cpp:425,559) — is served by the leaner kernel variant.
    const bool refgaps = cfg->fma && !any_del && !any_gcp;
    // A warp sweeps its chunk's columns one after the other (~0.2 us per column), so the chunk length is the latency of a
    // batch and chunks x reads its parallelism.  A big batch fills the GPU with long chunks (fewest seams and staging); a
    // small one (the synchronous per-region callers) is cut finer until it has about two warps per warp slot of the device.
    static const uint32_t env_cols = std::getenv("PHMM_B200_STREAM_COLS") ? (uint32_t)std::atoi(std::getenv("PHMM_B200_STREAM_COLS")) : 0u;   // experiments only
    const uint32_t full_cols = env_cols ? env_cols : kTargetStreamCols;
    uint32_t target_cols = full_cols;
    bool single_stream = false;    // one haplotype stream per chunk (latency policy below)
    int min_lanes = 8;             // lanes the longest read of a region should spread over at least (latency policy below)
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
        if (cfg->lanes_pref)
            min_lanes = (int)cfg->lanes_pref;
        else if (est_warps * 8 <= want)
            min_lanes = 32;
        else if (est_warps * 4 <= want)
            min_lanes = 16;
        // (est_warps counts 8-lane warps of full-length chunks: 32 lanes and a haplotype per chunk make 32 times as many)
        single_stream = min_lanes == 32 && !cfg->lanes_pref && est_warps * 64 <= want;
        est_warps *= (uint64_t)min_lanes / 8;
        // ... and a batch runs at 1 / (1 + slots / warps) of the device's rate while its last warps finish alone: below
        // kWavesWanted waves the chunks are cut shorter (a 300-column sweep costs a warp ~2 % in prologue, tools/chunk_curve.py)
        const uint64_t want_chunks = want * (kWavesWanted / 2);
        if (est_warps > 0 && est_warps < want_chunks)
            target_cols = (uint32_t)std::max<uint64_t>(kMinStreamCols, (uint64_t)full_cols * est_warps / want_chunks);
    }
    Slot::PlanScratch& ps = s.ps;
    // The read records go straight into the pinned arena (they are first in it): room for them and, as far as it can be told
    // now, for the rest — records per read and haplotype, the task list, the byte blobs — so that the arena is not moved later.
    {
        const size_t blobs = (size_t)(2 + (any_ins ? 1 : 0) + (any_del ? 1 : 0) + (any_gcp ? 1 : 0)) * (read_bytes + 272) + hap_bytes + 272;
        const size_t guess = (size_t)n_reads * (sizeof(PhmmRead) + 16) + (size_t)n_haps * 96 + blobs + ((size_t)1 << 16);
        const int rc0 = s.h_in.ensure(guess, &high_water[0]);
        if (rc0 != PHMM_OK) return rc0;
    }
    struct ReadRecords                               // the arena's head, written in place
    {
        HostBuf* buf;
        uint32_t n = 0;
        size_t size() const { return n; }
        void push_back(const PhmmRead& d) { static_cast<PhmmRead*>(buf->p)[n++] = d; }
    } reads{&s.h_in};
    std::vector<PhmmHap>& haps = ps.haps;
    std::vector<PhmmChunk>& chunks = ps.chunks;
    std::vector<PhmmPair>& slow = ps.slow;
    haps.clear();
    chunks.clear();
    slow.clear();
    ps.pri.clear();
    haps.reserve(n_haps);
    s.regions.clear();
    s.regions.reserve(regs.size());
    s.read_len.clear();
    s.read_len.reserve(n_reads);
    // per rows-per-lane count: (chunk, warps of reads) groups
    using Group = Slot::Group;
    const int max_rpt = phmm_max_rpt(refgaps);
    const uint32_t max_read = phmm_max_read_len(refgaps);
    std::vector<std::vector<Group>>& groups = ps.groups;
    std::vector<std::vector<uint32_t>>& group_reads = ps.group_reads;       // read ids, warp after warp
    std::vector<std::vector<uint8_t>>& group_warps = ps.group_warps;        // reads per warp
    constexpr int kFixedBin = PHMM_MAX_RPT + 1;                              // regions whose reads all take phmm_fixed_lanes() lanes
    constexpr int kDuoBin = PHMM_MAX_RPT + 2;                                // paired reads against a region's odd haplotype
    groups.resize(PHMM_MAX_RPT + 3);
    group_reads.resize(PHMM_MAX_RPT + 3);
    group_warps.resize(PHMM_MAX_RPT + 3);
    for (int b = 0; b <= PHMM_MAX_RPT + 2; ++b) {
        groups[b].clear();
        group_reads[b].clear();
        group_warps[b].clear();
    }
    std::vector<uint32_t>&order = ps.order, &region_chunks = ps.region_chunks;
    static thread_local LanePacker tl_packer;                                // per worker thread: no allocation once warm
    LanePacker& packer = tl_packer;
    // rows per lane picked for a region of n reads (capped at 255) of one length: [n][len], 0 = not seen yet; per worker
    // thread and per variant (the candidates differ)
    static thread_local std::vector<std::array<uint8_t, 32 * PHMM_MAX_RPT>> tl_uniform[2];
    std::vector<std::array<uint8_t, 32 * PHMM_MAX_RPT>>& uniform_rpt = tl_uniform[refgaps ? 1 : 0];
    if (uniform_rpt.empty()) uniform_rpt.assign(256, std::array<uint8_t, 32 * PHMM_MAX_RPT>{});
    // lanes a read of every length takes at every rows-per-lane count (a division per read and candidate otherwise)
    static const struct LanesLut
    {
        uint8_t v[PHMM_MAX_RPT + 1][32 * PHMM_MAX_RPT];
        LanesLut()
        {
            for (int c = 1; c <= PHMM_MAX_RPT; ++c)
                for (uint32_t len = 0; len < 32u * PHMM_MAX_RPT; ++len) v[c][len] = (uint8_t)std::min<uint32_t>(255u, phmm_read_lanes(len, c));
        }
        const uint8_t* operator[](int c) const { return v[c]; }
    } lanes_lut;

    static const int env_rpt = std::getenv("PHMM_B200_RPT") ? std::atoi(std::getenv("PHMM_B200_RPT")) : 0;   // experiments only
    // Paired reads (phmm_fwd_kernel<.., DUO>): a region with an odd number of haplotypes would leave the second stream of one
    // chunk a haplotype short — half of every lane idle for that long.  Its shortest haplotype is taken out of the two-stream
    // chunks and swept by the region's reads two at a time instead, when enough of them pair up (equal lane counts).
    static const bool env_no_duo = std::getenv("PHMM_B200_NO_DUO") != nullptr;                              // A/B and tests
    const bool duo_ok = !env_no_duo && refgaps && cfg->fma && !single_stream && min_lanes == 8 && env_rpt == 0;
    static thread_local LanePacker tl_duo_packer;
    LanePacker& duo_packer = tl_duo_packer;
    std::vector<uint32_t>&duo_a = ps.duo_a, &duo_b = ps.duo_b, &duo_order = ps.duo_order;

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
        // haplotypes -> chunks of two balanced streams.  The region's in-envelope haplotypes are dealt, longest first onto the
        // lightest, into 2 x n_chunks streams of about target_cols columns; streams of similar length are then paired into
        // chunks (a chunk sweeps max(x, y) steps: what one stream has beyond the other is idle lanes).  The haplotype records
        // are written in chunk order, so that a chunk owns a contiguous range of them; out_col keeps the caller's column.
        order.clear();
        uint64_t total = 0;
        for (uint32_t j = 0; j < r.n_haps; ++j) {
            max_cols = std::max(max_cols, r.haps[j].len);
            if (r.haps[j].len > 0 && r.haps[j].len <= kMaxStreamCols) {
                order.push_back(j);
                total += r.haps[j].len;
            }
        }
        const bool all_haps_in_stream = order.size() == r.n_haps;
        uint32_t duo_hap = UINT32_MAX, duo_chunk = UINT32_MAX;
        // Reads of mixed lengths mostly end up at 12 rows per lane in CTAs of one or two warps, where a full-length stream (a CTA
        // stages its chunk's streams whatever its size) keeps the SM's shared memory from holding the twelve warps the registers
        // allow; their chunks are cut at two thirds (measured: cfg5 streams +1.2 %, equal-length shapes want the full length).
        [[maybe_unused]] const auto ptp0 = PT_NOW();
        // one look at the read lengths for everything below that asks about them: which reads the wavefront kernel can take
        // (lens[k] != 0), their range, and how many of them take how many lanes at the paired-reads kernel's rows per lane
        std::vector<uint16_t>& lens = ps.fast_len;
        lens.resize(r.n_reads);
        uint32_t len_lo = UINT32_MAX, len_hi = 0, cls12[33] = {0};
        {
            const uint8_t* lut = lanes_lut[PHMM_DUO_RPT];
            for (uint32_t k = 0; k < r.n_reads; ++k) {
                const uint32_t len = region_read_len(r, k);
                const bool ok = len - 1u < max_read;
                lens[k] = ok ? (uint16_t)len : (uint16_t)0;
                if (ok) {
                    len_lo = std::min(len_lo, len);
                    len_hi = std::max(len_hi, len);
                    ++cls12[std::min<uint32_t>(lut[len], 32u)];
                }
            }
        }
        const uint32_t region_cols = (!env_cols && target_cols > kMixedStreamCols && len_lo < len_hi) ? kMixedStreamCols : target_cols;
        PT_ADD(5, ptp0, PT_NOW());
        [[maybe_unused]] const auto pth0 = PT_NOW();
        region_chunks.clear();
        std::vector<uint32_t>& rec_of = ps.rec_of;                     // caller's haplotype index -> its record in haps[]
        rec_of.assign(r.n_haps, 0u);
        {
            // byte offset of every haplotype in the blob (the bytes are gathered in the caller's order)
            std::vector<uint32_t>& boff = ps.hap_boff;
            boff.resize(r.n_haps);
            for (uint32_t j = 0; j < r.n_haps; ++j) {
                boff[j] = (uint32_t)hap_off;
                hap_off += r.haps[j].len;
            }
            auto emit = [&](uint32_t j, uint32_t pos, uint32_t ord_stream) {
                PhmmHap h{};
                h.byte_off = boff[j];
                h.len = r.haps[j].len;
                h.out_col = j;
                h.pos = pos;
                h.ord_stream = ord_stream;
                rec_of[j] = (uint32_t)haps.size();
                haps.push_back(h);
            };
            if (!order.empty()) {
                for (size_t a = 1; a < order.size(); ++a) {                // stable insertion sort, longest first (a handful of entries)
                    const uint32_t v = order[a];
                    size_t at2 = a;
                    while (at2 > 0 && r.haps[order[at2 - 1]].len < r.haps[v].len) {
                        order[at2] = order[at2 - 1];
                        --at2;
                    }
                    order[at2] = v;
                }
                if (duo_ok && (order.size() & 1)) {
                    // do enough reads pair up?  lanes at PHMM_DUO_RPT rows, read by read and pair by pair
                    uint32_t solo = 0, paired = 0;
                    for (uint32_t n = 1; n <= 32; ++n) {
                        solo += n * cls12[n];
                        paired += n * ((cls12[n] + 1) / 2);
                    }
                    if (len_hi <= 32u * PHMM_DUO_RPT - 1u && solo > 0 && paired * 10 <= solo * 7) {
                        duo_hap = order.back();                          // the shortest
                        order.pop_back();
                        total -= r.haps[duo_hap].len;
                    }
                }
                // never more chunks than pairs of haplotypes: a chunk with a single haplotype leaves its second stream empty —
                // unless the batch is so small that most of the device would idle anyway (the one-region calls of a lone
                // synchronous caller): then a haplotype per chunk halves the sweep, i.e. the call's latency
                if (!order.empty()) {                                         // (the odd haplotype may have been the only one)
                    const uint32_t per_chunk = single_stream ? 1u : 2u;
                    uint32_t n_chunks = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1, order.size() / per_chunk),
                                                                     std::max<uint64_t>(1, (total + per_chunk * (uint64_t)region_cols - 1) / (per_chunk * (uint64_t)region_cols)));
                    std::vector<uint32_t>&ssum = ps.stream_sum, &sown = ps.stream_of, &sidx = ps.stream_idx;
                    for (;;) {
                        const uint32_t S = per_chunk * n_chunks;
                        ssum.assign(S, 0u);
                        sown.resize(order.size());
                        bool fits = true;
                        for (size_t a = 0; a < order.size(); ++a) {
                            uint32_t best = 0;
                            for (uint32_t q = 1; q < S; ++q)
                                if (ssum[q] < ssum[best]) best = q;
                            sown[a] = best;
                            ssum[best] += r.haps[order[a]].len;
                            fits &= ssum[best] <= kMaxStreamCols;              // what one CTA can stage in shared memory
                        }
                        if (fits || per_chunk * n_chunks >= order.size()) break;
                        ++n_chunks;                                            // several very long haplotypes: spread them further
                    }
                    const uint32_t S = per_chunk * n_chunks;
                    sidx.resize(S);
                    for (uint32_t q = 0; q < S; ++q) sidx[q] = q;
                    for (uint32_t a = 1; a < S; ++a) {                         // streams by length, longest first
                        const uint32_t v = sidx[a];
                        uint32_t at2 = a;
                        while (at2 > 0 && ssum[sidx[at2 - 1]] < ssum[v]) {
                            sidx[at2] = sidx[at2 - 1];
                            --at2;
                        }
                        sidx[at2] = v;
                    }
                    for (uint32_t c = 0; c < n_chunks; ++c) {
                        const uint32_t sx = sidx[per_chunk * c], sy = single_stream ? UINT32_MAX : sidx[2 * c + 1];
                        if (ssum[sx] == 0) break;                              // fewer haplotypes than streams
                        PhmmChunk ch{};
                        ch.hap_begin = (uint32_t)haps.size();
                        uint32_t sum[2] = {0, 0}, cnt[2] = {0, 0};
                        for (size_t a = 0; a < order.size(); ++a) {
                            if (sown[a] != sx && sown[a] != sy) continue;
                            const uint32_t e = sown[a] == sx ? 0u : 1u;
                            emit(order[a], 32 + sum[e], (cnt[e] << 1) | e);
                            sum[e] += r.haps[order[a]].len;
                            ++cnt[e];
                        }
                        ch.n_haps = cnt[0] + cnt[1];
                        ch.stream_len = 32 + std::max(sum[0], sum[1]) + 1;
                        ch.n_xy = cnt[0] | (cnt[1] << 16);
                        region_chunks.push_back((uint32_t)chunks.size());
                        chunks.push_back(ch);
                    }
                }
                if (duo_hap != UINT32_MAX) {
                    PhmmChunk ch{};
                    ch.hap_begin = (uint32_t)haps.size();
                    emit(duo_hap, 32u, 0u);
                    ch.n_haps = 1;
                    ch.stream_len = 32 + r.haps[duo_hap].len + 1;
                    ch.n_xy = 1u;
                    duo_chunk = (uint32_t)chunks.size();
                    chunks.push_back(ch);
                }
            }
            // empty or over-long haplotypes: records of their own behind the chunks', served by the per-pair kernel (pos == 0)
            if (!all_haps_in_stream)
                for (uint32_t j = 0; j < r.n_haps; ++j)
                    if (!(r.haps[j].len > 0 && r.haps[j].len <= kMaxStreamCols)) emit(j, 0u, 0u);
        }
        [[maybe_unused]] const auto ptr0 = PT_NOW();
        PT_ADD(8, pth0, ptr0);
        // reads: one 32-byte record each, written straight into the pinned arena.  Everything but the length, the mapping quality
        // and the running offsets is the same for all reads of a region (of a flat batch: also which optional qualities exist).
        uint32_t longest_read = 0, shortest_read = UINT32_MAX, fast_reads = 0;
        {
            PhmmRead d{};
            d.n_haps = r.n_haps;
            d.hap_begin = rp.hap_begin;
            d.thr = (int16_t)std::max(-32768, std::min(32767, r.thr));
            // apply_pcr_error_model's switch takes HOSTILE / AGGRESSIVE / CONSERVATIVE and leaves the read alone for anything
            // else (pairhmm_internal.cpp:121-159, default: break)
            const uint32_t pcr = (r.pcr >= PHMM_PCR_HOSTILE && r.pcr <= PHMM_PCR_CONSERVATIVE) ? (uint32_t)r.pcr : (uint32_t)PHMM_PCR_NONE;
            const phmm_flat_batch* fb = r.flat;
            const uint32_t* flat_len = fb ? fb->read_len + r.flat_r0 : nullptr;
            const uint8_t* flat_mapq = fb ? fb->read_mapq + r.flat_r0 : nullptr;
            const uint32_t flat_fl = fb ? ((fb->read_ins ? PHMM_RD_HAS_INS : 0u) | (fb->read_del ? PHMM_RD_HAS_DEL : 0u) | (fb->read_gcp ? PHMM_RD_HAS_GCP : 0u)) : 0u;
            const bool has_chunk = !region_chunks.empty() || duo_chunk != UINT32_MAX;
            const size_t len_at = s.read_len.size();
            s.read_len.resize(len_at + r.n_reads);
            uint32_t* len_out = s.read_len.data() + len_at;
            uint64_t oo = out_off;
            for (uint32_t k = 0; k < r.n_reads; ++k, oo += r.n_haps) {
                uint32_t len, fl;
                if (fb) {
                    len = flat_len[k];
                    d.mapq = flat_mapq[k];
                    fl = flat_fl;
                } else {
                    const phmm_read& in = r.reads[k];
                    len = in.len;
                    d.mapq = in.mapq;
                    fl = (in.ins ? PHMM_RD_HAS_INS : 0u) | (in.del ? PHMM_RD_HAS_DEL : 0u) | (in.gcp ? PHMM_RD_HAS_GCP : 0u);
                }
                d.raw_off = (uint32_t)raw_off;
                d.word_off = (uint32_t)raw_off;
                d.len = len;
                d.out_off = (uint32_t)oo;
                d.model_flags = (uint8_t)(pcr | (fl << 4));
                raw_off += len;
                const uint32_t rid = (uint32_t)reads.size();
                reads.push_back(d);
                len_out[k] = len;
                const bool fast = lens[k] != 0 && has_chunk;                  // the wavefront kernel takes it
                if (fast) {
                    ++fast_reads;
                    longest_read = std::max(longest_read, len);
                    shortest_read = std::min(shortest_read, len);
                }
                if (!fast || !all_haps_in_stream)
                    for (uint32_t j = 0; j < r.n_haps; ++j) {
                        const bool hap_ok = r.haps[j].len > 0 && r.haps[j].len <= kMaxStreamCols;
                        if (!fast || !hap_ok) {
                            PhmmPair p{};
                            p.read = rid;
                            p.hap = rec_of[j];
                            p.out_idx = d.out_off + j;
                            slow.push_back(p);
                        }
                    }
            }
        }
        PT_ADD(9, ptr0, PT_NOW());
        [[maybe_unused]] const auto ptd0 = PT_NOW();
        if (fast_reads && duo_chunk != UINT32_MAX) {
            // the odd haplotype's sweep: reads of equal lane count two by two (neighbours in the caller's order, which is by
            // length), pairs packed into warps like single reads
            const uint8_t* lut = lanes_lut[PHMM_DUO_RPT];
            const uint32_t* cls = cls12;
            uint32_t at[34];
            at[1] = 0;
            for (int n = 1; n <= 32; ++n) at[n + 1] = at[n] + (cls[n] + 1) / 2;
            duo_a.assign(at[33], UINT32_MAX);
            duo_b.assign(at[33], UINT32_MAX);
            uint32_t seen[33] = {0};
            for (uint32_t k = 0; k < r.n_reads; ++k) {
                if (!lens[k]) continue;
                const uint32_t n = lut[lens[k]], q = seen[n]++;
                ((q & 1) ? duo_b : duo_a)[at[n] + q / 2] = rp.read_begin + k;
            }
            duo_packer.slots.resize(at[33]);
            for (uint32_t q = 0; q < at[33]; ++q) duo_packer.slots[q] = q;
            duo_packer.count[0] = 0;
            for (int n = 1; n <= 32; ++n) {
                duo_packer.count[n] = (cls[n] + 1) / 2;
                duo_packer.first[n] = at[n];
            }
            duo_order.clear();
            const uint32_t first_read = (uint32_t)group_reads[kDuoBin].size(), first_warp = (uint32_t)group_warps[kDuoBin].size();
            const uint32_t n_warps = duo_packer.pack(&duo_order, &group_warps[kDuoBin]);
            for (uint32_t q : duo_order) {
                group_reads[kDuoBin].push_back(duo_a[q]);
                group_reads[kDuoBin].push_back(duo_b[q]);
            }
            groups[kDuoBin].push_back(Group{duo_chunk, chunks[duo_chunk].stream_len, first_read, first_warp, n_warps});
        }
        PT_ADD(6, ptd0, PT_NOW());
        [[maybe_unused]] const auto ptm0 = PT_NOW();
        if (fast_reads && !region_chunks.empty()) {
            // Rows per lane for this region.  A read of len bases takes ceil((len + 1) / rpt) consecutive lanes; the region's
            // reads are packed into warps lane by lane.  Throughput: the rpt with the least warps x step cost (many rows per
            // lane amortise a step's shuffles and loads, but pad every read up to a whole lane).  Latency (a small batch on an
            // idle device): the longest read spread over about min_lanes lanes, so that a step is short.
            int rpt = 0;
            const int rpt_lo = (int)((longest_read + 32) / 32);                 // the longest read must fit 32 lanes
            // (lens[]: the lengths of the reads the wavefront kernel takes, 0 = not one of them)
            auto exact_cost = [&](int c) {                                      // warps after packing x step cost
                for (int n = 0; n <= 32; ++n) packer.count[n] = 0;
                const uint8_t* lut = lanes_lut[c];
                for (uint32_t k = 0; k < r.n_reads; ++k) ++packer.count[lut[lens[k]]];   // lut[0] = 1 lane: subtract the non-reads below
                packer.count[1] -= r.n_reads - fast_reads;
                return packer.pack(nullptr, nullptr) * step_cost(c);
            };
            if (env_rpt > 0) {
                rpt = std::max(rpt_lo, std::min(max_rpt, env_rpt));
            } else if (min_lanes > 8) {
                rpt = std::max(rpt_lo, std::min(max_rpt, (int)((longest_read + (uint32_t)min_lanes) / (uint32_t)min_lanes)));
            } else if (shortest_read == longest_read && (rpt = uniform_rpt[std::min<uint32_t>(fast_reads, 255u)][longest_read]) != 0) {
                // a region of equal-length reads (the usual short-read case): remembered from the last such region
            } else {
                // the few candidate counts, each packed for real: least warps x step cost (ties: more rows per lane)
                // The first choice is the most efficient instantiation; another count has to beat it by kRptMargin per cent to
                // be worth a launch of its own (reads that pack badly at the first choice do: 150-bp reads take 13 lanes at
                // 12 rows, two per warp, but 8 lanes at 19 rows, four per warp).
                uint64_t best = UINT64_MAX;
                const int* choices = refgaps ? kRptChoices : kRptChoicesGeneric;
                const bool uniform = shortest_read == longest_read;
                for (int q = 0; q < 3; ++q) {
                    const int c = std::max(rpt_lo, choices[q]);
                    if (c > max_rpt || c == rpt) continue;
                    uint64_t cost;
                    if (uniform) {
                        cost = exact_cost(c);                                   // packed for real, and remembered below
                    } else {
                        // mixed lengths: all lanes over what a warp holds on average — 32 less half a read's lanes minus one,
                        // the room the next read does not fit into (an estimate: packing every candidate for every region
                        // of a stream of small regions costs more host time than it saves device time)
                        const uint8_t* lut = lanes_lut[c];
                        uint32_t lanes = 0;
                        for (uint32_t k = 0; k < r.n_reads; ++k) lanes += lut[lens[k]];
                        lanes -= r.n_reads - fast_reads;
                        const uint32_t room64 = 64u * 32u - 32u * (lanes - fast_reads) / fast_reads;      // capacity per warp x 64
                        cost = (uint64_t)((64u * lanes + room64 - 1) / room64) * step_cost(c);
                    }
                    cost *= (q == 0 ? 100u - kRptMargin : 100u);
                    if (cost < best) { best = cost; rpt = c; }
                }
                if (rpt == 0) rpt = max_rpt;
                if (uniform) uniform_rpt[std::min<uint32_t>(fast_reads, 255u)][longest_read] = (uint8_t)rpt;
            }
            // the packing itself: read ids by lane count (a counting sort that keeps the caller's order), then warp after warp
            {
                const uint8_t* lut = lanes_lut[rpt];
                for (int n = 0; n <= 32; ++n) packer.count[n] = 0;
                for (uint32_t k = 0; k < r.n_reads; ++k) ++packer.count[lens[k] ? lut[lens[k]] : 0];
                uint32_t at = 0;
                for (int n = 1; n <= 32; ++n) { packer.first[n] = at; at += packer.count[n]; }
                packer.slots.resize(at);
                uint32_t cur[33];
                for (int n = 1; n <= 32; ++n) cur[n] = packer.first[n];
                for (uint32_t k = 0; k < r.n_reads; ++k)
                    if (lens[k]) packer.slots[cur[lut[lens[k]]]++] = rp.read_begin + k;
            }
            // a region whose reads all take the fixed lane count of this rows-per-lane count goes to the fixed-lanes kernel
            const uint32_t fixed = cfg->fma ? phmm_fixed_lanes(rpt, refgaps) : 0u;
            const int bin = (fixed && packer.count[fixed] == fast_reads) ? kFixedBin : rpt;
            const uint32_t first_read = (uint32_t)group_reads[bin].size(), first_warp = (uint32_t)group_warps[bin].size();
            const uint32_t n_warps = packer.pack(&group_reads[bin], &group_warps[bin]);
            // every chunk of the region sweeps the same warps: one copy of the lists, one group per chunk
            for (uint32_t ci : region_chunks) groups[bin].push_back(Group{ci, chunks[ci].stream_len, first_read, first_warp, n_warps});
        }
        PT_ADD(10, ptm0, PT_NOW());
        out_off += (uint64_t)r.n_reads * r.n_haps;
    }

    [[maybe_unused]] const auto pt2 = PT_NOW();
    PT_ADD(1, pt1, pt2);
    // CTAs per rows-per-lane count, longest streams first
    std::vector<PhmmCta>& ctas = ps.ctas;
    std::vector<uint32_t>& task_reads = ps.task_reads;
    ctas.clear();
    task_reads.clear();
    s.launches.clear();
    for (int bin = 1; bin <= kDuoBin; ++bin) {
        const bool duo = bin == kDuoBin;
        const int b = duo ? PHMM_DUO_RPT : bin == kFixedBin ? PHMM_MAX_RPT : bin;   // rows per lane of this launch
        if (bin < kFixedBin && bin > max_rpt) continue;
        if (groups[bin].empty()) continue;
        {
            // stable, longest stream first: a counting sort on the stream length (at most kMaxStreamCols + 33 values)
            uint32_t longest = 0;
            for (const Group& g : groups[bin]) longest = std::max(longest, g.stream_len);
            ps.bucket.assign((size_t)longest + 2, 0u);
            for (const Group& g : groups[bin]) ++ps.bucket[longest - g.stream_len + 1];
            for (uint32_t k = 1; k <= longest + 1; ++k) ps.bucket[k] += ps.bucket[k - 1];
            ps.sorted.resize(groups[bin].size());
            for (const Group& g : groups[bin]) ps.sorted[ps.bucket[longest - g.stream_len]++] = g;
            groups[bin].swap(ps.sorted);
        }
        FwdLaunch fl{};
        fl.rpt = b;
        fl.fixed_lanes = bin == kFixedBin ? phmm_fixed_lanes(b, refgaps) : 0u;
        fl.duo = duo;
        fl.refgaps = refgaps;
        fl.cta_base = (uint32_t)ctas.size();
        size_t smem = 0;
        // Warps per CTA of this launch: bigger CTAs stage each chunk's streams fewer times and run a little faster, but a
        // region's warps then come in bigger lumps; take the biggest size that wastes at most 8 % more warp slots than
        // single-warp CTAs.  (A rule that maximised resident warps per SM from the shared memory and the registers a CTA size
        // needs picked smaller CTAs for short chunks and measured 5 % slower on post-trim streams, equal elsewhere:
        // profiles/r02_experiments.txt.)
        uint32_t cta_warps = cfg->cta_warps;
        uint32_t longest_stream = 0, most_haps = 0;
        for (const Group& g : groups[bin]) {
            longest_stream = std::max(longest_stream, chunks[g.chunk].stream_len);
            most_haps = std::max(most_haps, chunks[g.chunk].n_haps);
        }
        if (cta_warps == 0) {
            uint64_t slots[3] = {0, 0, 0};
            for (const Group& g : groups[bin])
                for (int k = 0; k < 3; ++k) {
                    const uint32_t per = 1u << k;
                    slots[k] += (uint64_t)((g.n_warps + per - 1) / per) * per;
                }
            cta_warps = 1;
            for (int k = 1; k < 3; ++k)
                if (slots[k] * 100 <= slots[0] * 108) cta_warps = 1u << k;
        }
        // ... that still fits the device's shared memory with this launch's longest chunk (streams of very long haplotypes)
        for (;;) {
            smem = phmm_fwd_smem_bytes(longest_stream, most_haps, b, cta_warps, duo);
            if (smem <= smem_optin || cta_warps == 1) break;
            cta_warps >>= 1;
        }
        const std::vector<uint32_t>& gr = group_reads[bin];
        const std::vector<uint8_t>& gw = group_warps[bin];
        for (const Group& g : groups[bin]) {
            uint32_t at_read = g.first_read;
            for (uint32_t w0 = 0; w0 < g.n_warps; w0 += cta_warps) {
                PhmmCta c{};
                c.chunk = g.chunk;
                c.read_begin = (uint32_t)task_reads.size();
                const uint32_t w1 = std::min(g.n_warps, w0 + cta_warps);
                for (uint32_t w = w0; w < w1; ++w) {
                    const uint32_t n = gw[g.first_warp + w];
                    c.warp_reads |= n << (8 * (w - w0));
                    c.n_reads += n;
                }
                const uint32_t entries = c.n_reads * (duo ? 2u : 1u);     // paired reads: two task entries per read slot
                task_reads.insert(task_reads.end(), gr.begin() + at_read, gr.begin() + at_read + entries);
                at_read += entries;
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
        return (uint64_t)x.n_ctas * x.cta_warps * (uint32_t)x.rpt > (uint64_t)y.n_ctas * y.cta_warps * (uint32_t)y.rpt;
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
    if ((rc = s.h_in.ensure_keep(off, reads.size() * sizeof(PhmmRead), &high_water[0])) != PHMM_OK) return rc;
    if ((rc = s.d_in.ensure(off, &high_water[0])) != PHMM_OK) return rc;
    uint8_t* h = static_cast<uint8_t*>(s.h_in.p);
    static_assert(sizeof(PhmmRead) == 32, "the arena's head is an array of PhmmRead");
    if (o_reads != 0) return fail(PHMM_ERR_STATE, "planner: the read records must be first in the arena");
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
        const bool one_run = flat_contig && flat_first != nullptr;
        if (one_run) {                               // a flat batch with its reads back to back: one run per blob, no per-read walk
            const phmm_flat_batch* fb = flat_first->flat;
            const uint64_t o = fb->read_off[flat_first->flat_r0];
            rb_run.add(fb->read_bases + o, read_bytes, o_rb);
            rq_run.add(fb->read_qual + o, read_bytes, o_rq);
            if (fb->read_ins) ri_run.add(fb->read_ins + o, read_bytes, o_ri);
            if (fb->read_del) rd_run.add(fb->read_del + o, read_bytes, o_rd);
            if (fb->read_gcp) rg_run.add(fb->read_gcp + o, read_bytes, o_rg);
        }
        for (const RegionIn& r : regs) {
            if (r.n_reads == 0 || r.n_haps == 0) continue;
            pinned &= r.pinned_src;
            for (uint32_t k = 0; k < r.n_reads && !one_run; ++k) {
                const ReadView in = region_read(r, k);
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
    if ((rc = s.d_pairs.ensure(2 * pairs_alloc * sizeof(PhmmPair), &high_water[4])) != PHMM_OK) return rc;   // + the FP64 hand-back list
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
    // A small batch of the ticket path (synchronous callers: one or a few regions) is all latency.  Its kernels write the raw
    // sums straight into the pinned result buffer (a few KB of posted PCIe writes) — no copy back, no counter to fetch (the host
    // epilogue sees which sums are under MIN_ACCEPTED anyway) — and a single forward launch is not timed on its own.
    s.lite = s.small_ok && s.launches.size() <= 1 && slow.empty() && !s.want_f64 && !s.dev_epi;
    s.host_out = false;
    if (s.small_ok && !s.dev_epi && !s.want_f64 && !s.h_in.host_only && n_pairs && n_pairs * sizeof(float) <= kHostOutBytes) {
        void* mapped = nullptr;
        if (cudaHostGetDevicePointer(&mapped, s.h_f32.p, 0) == cudaSuccess && mapped) {
            b.out_f32 = static_cast<float*>(mapped);
            s.host_out = true;
        } else {
            cudaGetLastError();
        }
    }
    b.out_f64 = static_cast<double*>(s.d_f64.p);
    b.f64_pairs = static_cast<PhmmPair*>(s.d_pairs.p);
    b.f64_count = static_cast<uint32_t*>(s.d_cnt.p);
    b.tables = static_cast<const PhmmTables*>(d_tables.p);
    b.n_reads = (uint32_t)reads.size();
    b.rg_gapm = cfg->default_gapm;      // gap_conts() is the constant '+' - 33 = 10 for every read, read_record.cpp:205-210
    b.rg_ii = cfg->default_ii;
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
