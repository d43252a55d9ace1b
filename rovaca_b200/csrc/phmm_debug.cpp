// phmm_debug.cpp — phmm_debug_* exports: what the tests use to look inside the host side without a GPU (planner audit, the
// host epilogue on given raw sums, the log10f restatement against libm, the flat partition) and one device probe.  Not part of
// include/phmm_b200.h.
#include "phmm_host.h"

extern "C" {

// For tests (no GPU needed): run the planner alone on a flat batch and audit what it produced.  Every (read, haplotype) pair
// must be computed exactly once — by exactly one forward-kernel warp slot whose chunk holds the haplotype, or by the per-pair
// kernel — every read must sit in a bin that holds it, every CTA must respect its launch's shape.  out[0..15]:
//   0 pairs  1 pairs covered by the wavefront kernel  2 per-pair-kernel pairs  3 pairs covered more than once  4 pairs missed
//   5 warps launched  6 lanes of those warps  7 lanes that hold a read  8 reads on too few lanes  9 launches  10 chunks
//   11 lane-steps (sum over warps of 32 x RPT x steps)  12 cells (sum R x H over wavefront pairs)  13 CTA shape violations
int phmm_debug_plan(const phmm_flat_batch* fb, int fma_mode, uint32_t cta_warps, uint32_t busy_batches, uint32_t lanes_per_read, uint64_t* out)
{
    if (!fb || !out) return fail(PHMM_ERR_ARG, "phmm_debug_plan: null argument");
    PhmmSettings cfg;
    cfg.fma = fma_mode != PHMM_FMA_NONE;
    cfg.cta_warps = std::min<uint32_t>(cta_warps, PHMM_WARPS);
    cfg.lanes_pref = lanes_per_read;
    PhmmDevice e;
    e.cfg = &cfg;
    e.busy_batches.store(busy_batches);
    Slot s;
    s.set_host_only();
    std::vector<phmm_hap> haps;
    std::vector<RegionIn> regs;
    flat_to_regions(fb, nullptr, 0, fb->n_regions, haps, regs);
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
        const uint32_t rpt = (uint32_t)fl.rpt;
        for (uint32_t c = fl.cta_base; c < fl.cta_base + fl.n_ctas; ++c) {
            const PhmmCta& cta = T[c];
            const PhmmChunk& ch = C[cta.chunk];
            const uint32_t ts = fl.duo ? 2u : 1u;                             // paired reads: two task entries per read slot
            if (cta.n_reads == 0 || cta.read_begin + ts * cta.n_reads > s.n_tasks) ++out[13];
            if (fl.duo && (rpt != PHMM_DUO_RPT || ch.n_haps != 1 || (ch.n_xy >> 16) != 0)) ++out[13];
            uint32_t at = cta.read_begin, in_warps = 0;
            for (uint32_t w = 0; w < PHMM_WARPS; ++w) {
                const uint32_t n = (cta.warp_reads >> (8 * w)) & 0xffu;
                in_warps += n;
                if (n == 0) continue;
                if (w >= fl.cta_warps) ++out[13];                         // a warp the launch does not have
                uint32_t lanes = 0, widest = 0;
                for (uint32_t k = 0; k < n; ++k) {
                    const PhmmRead& rd = R[tasks[at + ts * k]];
                    const uint32_t nl = phmm_read_lanes(rd.len, (int)rpt);
                    if (nl * rpt < rd.len + 1 || rd.len == 0) ++out[8];
                    if (fl.fixed_lanes && nl != fl.fixed_lanes) ++out[13];   // the fixed-lanes kernel takes reads of that lane count only
                    if (fl.duo && tasks[at + 2 * k + 1] != UINT32_MAX) {     // its partner must take the same lanes
                        const PhmmRead& rb = R[tasks[at + 2 * k + 1]];
                        if (rb.len == 0 || phmm_read_lanes(rb.len, (int)rpt) != nl) ++out[8];
                    }
                    lanes += nl;
                    widest = std::max(widest, nl);
                }
                if (lanes > 32) ++out[13];                                // the kernel gives every read consecutive lanes of ONE warp
                const uint64_t steps = ch.stream_len - 1 - (32 - widest);
                out[5] += 1;
                out[6] += 32;
                out[7] += lanes;
                out[11] += (uint64_t)32u * rpt * steps;
                for (uint32_t k = 0; k < ts * n; ++k) {
                    if (tasks[at + k] == UINT32_MAX) continue;               // a pair of one
                    const PhmmRead& rd = R[tasks[at + k]];
                    for (uint32_t j = 0; j < ch.n_haps; ++j) {
                        const PhmmHap& hp = H[ch.hap_begin + j];
                        if (hp.pos == 0) { ++out[13]; continue; }        // chunks own placed haplotypes only
                        const uint64_t idx = (uint64_t)rd.out_off + hp.out_col;
                        if (idx >= s.n_pairs || hp.out_col >= rd.n_haps) { ++out[13]; continue; }
                        if (seen[idx]++) ++out[3];
                        ++out[1];
                        out[12] += (uint64_t)rd.len * hp.len;
                    }
                }
                at += ts * n;
            }
            if (in_warps != cta.n_reads) ++out[13];
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
    if (in_use) *in_use = e->cfg.device_epilogue ? 1 : 0;
    CU(cudaSetDevice(e->devs[0]->ordinal));
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
    PhmmSettings cfg;
    cfg.log10_mode = mode < 0 ? phmm_log10_pick_mode() : mode;
    PhmmDevice e;
    e.cfg = &cfg;
    Slot s;
    s.set_host_only();
    std::vector<phmm_hap> haps;
    std::vector<RegionIn> regs;
    flat_to_regions(fb, fr, 0, fb->n_regions, haps, regs);
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
    PhmmSettings cfg;
    const std::vector<uint64_t> pre = flat_work_prefix(fb);
    const auto parts = flat_partition(pre, fb->n_regions, batch_cells ? batch_cells : cfg.batch_cells, n_workers ? n_workers : cfg.n_workers);
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
