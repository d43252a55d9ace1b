/*
 * ref_shim.cpp — C-ABI door into the UNMODIFIED reference PairHMM, for tests and the CPU baseline.
 * TEST INFRASTRUCTURE ONLY (see phmm_oracle.h).
 *
 * oracle/Makefile compiles this file together with the reference's own translation units, taken
 * in place from /root/reference (nothing is copied into the repo):
 *     src/haplotypecaller/pairhmm/pairhmm_internal.cpp           (call_old_pairhmm, call_new_pairhmm002)
 *     src/haplotypecaller/pairhmm/core/avx{2,512}_{s,d}.cpp      (per-pair kernels)
 *     src/haplotypecaller/pairhmm/rovaca/avx_512_float.cc        (16 reads x 1 haplotype kernel)
 *     src/haplotypecaller/genotype/read_record.cpp               (ReadRecord accessors)
 *     src/haplotypecaller/common/utils/rovaca_memory_pool.cpp    (pmr bump pool)
 * into oracle/_ref/libphmm_ref.so, with the reference's per-file ISA flags
 * (pairhmm/CMakeLists.txt:3-5, CMakeLists.txt:49).  boost::dynamic_bitset is only named in a
 * typedef on this path (genotype/forward.h:4,71) and is satisfied by oracle/stubs/.
 *
 * What this file adds is glue only: fake bam1_t records so that the real ReadRecord can be
 * driven, raw-memory Haplotype objects (only the inline get_display_string()/length() are used by
 * the path), the three htslib symbols read_record.cpp links against, and the function-pointer
 * binding pairhmm_engine.cpp:21-53 performs (that file needs the spdlog logger, so its ten lines
 * of binding logic are restated in bind_kernels()).
 */
#include "ref_glue.h"

/* ---- htslib symbols read_record.cpp references (sam.c / hts.c); not exercised by the path ---- */
extern "C" {
const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
hts_pos_t bam_cigar2qlen(int n_cigar, const uint32_t* cigar)
{
    hts_pos_t l = 0;
    for (int k = 0; k < n_cigar; ++k)
        if (bam_cigar_type(bam_cigar_op(cigar[k])) & 1) l += bam_cigar_oplen(cigar[k]);
    return l;
}
hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t* cigar)
{
    hts_pos_t l = 0;
    for (int k = 0; k < n_cigar; ++k)
        if (bam_cigar_type(bam_cigar_op(cigar[k])) & 2) l += bam_cigar_oplen(cigar[k]);
    return l;
}
}

/* ---- the two globals pairhmm_engine.cpp:8-9 defines and pairhmm_internal.cpp:15-16 imports ---- */
float (*s_compute_full_prob_float_old_pairhmm)(const TestCase& tc) = nullptr;
double (*s_compute_full_prob_double_old_pairhmm)(const TestCase& tc) = nullptr;

namespace
{

bool g_avx512 = false;
bool g_bound = false;

/* pairhmm_engine.cpp:21-53 */
void bind_kernels()
{
    if (g_bound) return;
    g_avx512 = is_avx512_supported();
    ConvertChar::init();
    if (g_avx512) {
        s_compute_full_prob_float_old_pairhmm = compute_fp_avx512_s;
        s_compute_full_prob_double_old_pairhmm = compute_fp_avx512_d;
    } else {
        s_compute_full_prob_float_old_pairhmm = compute_fp_avx2_s;
        s_compute_full_prob_double_old_pairhmm = compute_fp_avx2_d;
    }
    g_bound = true;
}

struct CsrGuard
{
    unsigned old;
    CsrGuard() : old(_mm_getcsr()) {}
    ~CsrGuard() { _mm_setcsr(old); }
};

using ref_glue::FakeBam;

struct Region
{
    std::vector<FakeBam> bams;
    std::vector<uint8_t> pool_buf;
};

}  // namespace

extern "C" {

int ref_has_avx512(void)
{
    bind_kernels();
    return g_avx512 ? 1 : 0;
}

/* isa: 0 = AVX2 TU, 1 = AVX-512 TU.  One pair through the reference's per-pair kernels
 * (core/avx2_s.cpp:234, core/avx512_s.cpp:259). */
float ref_pair_f32(int isa, const uint8_t* hap, uint32_t H, const uint8_t* rs, const uint8_t* q, const uint8_t* ins,
                   const uint8_t* del, const uint8_t* gcp, uint32_t R)
{
    bind_kernels();
    CsrGuard g;
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    ::TestCase tc{};
    tc.rslen = R; tc.haplen = H; tc.hap = hap; tc.rs = rs; tc.q = q; tc.i = ins; tc.d = del; tc.c = gcp;
    if (isa == 1) {
        if (!g_avx512) return NAN;
        return compute_fp_avx512_s(tc);
    }
    return compute_fp_avx2_s(tc);
}

/* The reference's FP64 kernels *as built* (defective) — reported, not trusted. */
double ref_pair_f64(int isa, const uint8_t* hap, uint32_t H, const uint8_t* rs, const uint8_t* q, const uint8_t* ins,
                    const uint8_t* del, const uint8_t* gcp, uint32_t R)
{
    bind_kernels();
    CsrGuard g;
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    ::TestCase tc{};
    tc.rslen = R; tc.haplen = H; tc.hap = hap; tc.rs = rs; tc.q = q; tc.i = ins; tc.d = del; tc.c = gcp;
    if (isa == 1) {
        if (!g_avx512) return NAN;
        return compute_fp_avx512_d(tc);
    }
    return compute_fp_avx2_d(tc);
}

/* n (<=16) reads x 1 haplotype through rovaca/avx_512_float.cc:29.  Plane layout and the per-base
 * table look-ups are those of prepare_data (pairhmm_internal.cpp:450-483): [j*16 + lane].
 * Unused lanes repeat read 0.  Returns 0, or -1 without AVX-512. */
int ref_1xn_f32(const uint8_t* hap, uint32_t H, uint32_t n, const uint8_t* const* rs, const uint8_t* const* q,
                const uint8_t* const* ins, const uint8_t* const* del, const uint8_t* const* gcp, const uint32_t* lens,
                float* out16)
{
    bind_kernels();
    if (!g_avx512 || n == 0 || n > 16) return -1;
    CsrGuard g;
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    static const uint32_t conv[5][2] = {{'A', 1}, {'C', 2}, {'T', 4}, {'G', 8}, {'N', 15}};
    auto oh = [&](uint8_t c) -> int32_t {
        for (auto& e : conv) if (e[0] == c) return (int32_t)e[1];
        return 0;
    };
    uint32_t maxl = 0, minl = UINT32_MAX;
    uint32_t la[16];
    for (uint32_t k = 0; k < 16; ++k) {
        la[k] = lens[k < n ? k : 0];
        maxl = std::max(maxl, la[k]);
        minl = std::min(minl, la[k]);
    }
    const size_t plane = (size_t)16 * maxl;
    float* f = (float*)aligned_alloc(64, sizeof(float) * plane * 8 + 64);
    int32_t* rd = (int32_t*)aligned_alloc(64, sizeof(int32_t) * plane + 64);
    int32_t* hp = (int32_t*)aligned_alloc(64, sizeof(int32_t) * (H + 16));
    uint32_t* rl = (uint32_t*)aligned_alloc(64, 64);
    uint32_t* hl = (uint32_t*)aligned_alloc(64, 64);
    std::memset(f, 0, sizeof(float) * plane * 8);
    std::memset(rd, 0, sizeof(int32_t) * plane);
    rovaca::TestCase tc{};
    tc.read_num = 16; tc.hap_num = 16;
    tc.mm = f; tc.mi = f + plane; tc.ii = f + 2 * plane; tc.md = f + 3 * plane; tc.dd = f + 4 * plane;
    tc.gapm = f + 5 * plane; tc.distm = f + 6 * plane; tc._1_distm = f + 7 * plane;
    tc.reads = rd; tc.haps = hp; tc.read_len_arr = rl; tc.hap_len_arr = hl;
    rovaca::Context<float> ctx{};
    for (uint32_t k = 0; k < 16; ++k) {
        const uint32_t s = k < n ? k : 0;
        rl[k] = la[k];
        for (uint32_t j = 0; j < la[k]; ++j) {
            const int _i = ins[s][j] & 127, _d = del[s][j] & 127, _c = gcp[s][j] & 127, _q = q[s][j] & 127;
            const size_t at = k + (size_t)j * 16;
            rd[at] = oh(rs[s][j]);
            tc.mm[at] = ctx.set_mm_prob(_i, _d);
            tc.gapm[at] = 1.0f - rovaca::Context<float>::ph2pr[_c];
            tc.mi[at] = rovaca::Context<float>::ph2pr[_i];
            tc.ii[at] = rovaca::Context<float>::ph2pr[_c];
            tc.md[at] = rovaca::Context<float>::ph2pr[_d];
            tc.dd[at] = rovaca::Context<float>::ph2pr[_c];
            const float dist = rovaca::Context<float>::ph2pr[_q];
            tc._1_distm[at] = 1.0f - dist;
            tc.distm[at] = dist / 3.0f;
        }
    }
    for (uint32_t c = 0; c < H; ++c) hp[c] = oh(hap[c]);
    tc.max_read_len = maxl; tc.min_read_len = minl;
    tc.min_hap_len = tc.max_hap_len = H;
    rovaca::compute_full_prob_avx512_float_1xn(tc, out16);
    free(f); free(rd); free(hp); free(rl); free(hl);
    return 0;
}

/* The reference's float tables, for table-level checks (core/context.h:144-146, 49-59). */
void ref_tables_f32(float* ph2pr128, float* mm_tri_8256)
{
    rovaca::Context<float> ctx{};
    for (int x = 0; x < 128; ++x) ph2pr128[x] = rovaca::Context<float>::ph2pr[x];
    for (int k = 0; k < 128 * 129 / 2; ++k) mm_tri_8256[k] = rovaca::Context<float>::matchToMatchProb[k];
}

int32_t ref_find_tandem_repeat_units(const uint8_t* bases, uint32_t n, uint32_t offset)
{
    std::vector<uint8_t> buf(sizeof(rovaca::Bases) + n + 1);
    auto* b = reinterpret_cast<rovaca::pBases>(buf.data());
    b->num = n;
    std::memcpy(b->data, bases, n);
    return rovaca::find_tandem_repeat_units(b, offset);
}

void ref_apply_pcr_error_model(const uint8_t* bases, uint8_t* gap_qual, uint32_t n, int model)
{
    std::vector<uint8_t> b1(sizeof(rovaca::Bases) + n + 1), b2(sizeof(rovaca::Bases) + n + 1);
    auto* b = reinterpret_cast<rovaca::pBases>(b1.data());
    auto* g = reinterpret_cast<rovaca::pBases>(b2.data());
    b->num = n; g->num = n;
    std::memcpy(b->data, bases, n);
    std::memcpy(g->data, gap_qual, n);
    rovaca::apply_pcr_error_model(b, g, static_cast<PcrIndelModel>(model));
    std::memcpy(gap_qual, g->data, n);
}

/*
 * One region through the reference's own call_pairhmm.
 *   path: 0 = call_old_pairhmm (pairhmm_internal.cpp:505), 1 = call_new_pairhmm002 (:596, needs
 *         AVX-512 and reads sorted by ascending length, :602), -1 = what init_pairhmm_ptr(false)
 *         would bind on this CPU (pairhmm_engine.cpp:37-42).
 * Flat inputs: read k has bases/qual at rbase+roff[k], length rlen[k]; haplotype j at hbase+hoff[j].
 * out: [n_haps][n_reads] doubles exactly as returned (NaN tail), kept_idx: original indices of the
 * reads left in `rs` after the call.  Returns 0, or <0 on a contract violation.
 */
int ref_region(int path, uint32_t n_reads, const uint8_t* rbase, const uint8_t* rqual, const uint64_t* roff,
               const uint32_t* rlen, const uint8_t* rmapq, uint32_t n_haps, const uint8_t* hbase, const uint64_t* hoff,
               const uint32_t* hlen, int32_t min_quality_threshold, int pcr_model, double* out, uint32_t* kept_idx,
               uint32_t* n_kept)
{
    bind_kernels();
    if (path < 0) path = g_avx512 ? 1 : 0;
    if (path == 1 && !g_avx512) return -2;
    CsrGuard g;

    static thread_local std::unique_ptr<uint8_t[]> buffer(new uint8_t[(size_t)64 << 20]);   /* not zeroed, reused per thread */
    rovaca::RovacaMemoryPool pool{buffer.get(), (size_t)64 << 20};
    std::vector<FakeBam> bams(n_reads);
    rovaca::ReadVector rs{&pool};
    std::unordered_map<rovaca::pReadRecord, uint32_t> index_of;
    for (uint32_t k = 0; k < n_reads; ++k) {
        bams[k].build(rbase + roff[k], rqual + roff[k], rlen[k], rmapq[k]);
        rovaca::pReadRecord r = rovaca::ReadRecord::create(&pool, nullptr, &bams[k].b);
        rs.push_back(r);
        index_of[r] = k;
        if (path == 1 && k > 0 && rlen[k] < rlen[k - 1]) return -3;
    }
    rovaca::HaplotypeVector hs{&pool};
    for (uint32_t j = 0; j < n_haps; ++j) {
        auto* b = static_cast<rovaca::pBases>(pool.allocate(sizeof(rovaca::Bases) + hlen[j] + 1, 8));
        b->num = hlen[j];
        std::memcpy(b->data, hbase + hoff[j], hlen[j]);
        b->data[hlen[j]] = 0;
        void* mem = pool.allocate(sizeof(rovaca::Haplotype), alignof(rovaca::Haplotype));
        std::memset(mem, 0, sizeof(rovaca::Haplotype));
        auto* h = static_cast<rovaca::Haplotype*>(mem);   /* never constructed: only inline accessors are used */
        h->_bases = b;
        h->_is_symbolic = 0;
        hs.push_back(h);
    }
    rovaca::DoubleVector2D m = (path == 1)
                                   ? rovaca::call_new_pairhmm002(hs, rs, min_quality_threshold, static_cast<PcrIndelModel>(pcr_model), &pool)
                                   : rovaca::call_old_pairhmm(hs, rs, min_quality_threshold, static_cast<PcrIndelModel>(pcr_model), &pool);
    if (m.size() != n_haps && !(m.empty() && (n_haps == 0 || n_reads == 0))) return -4;
    for (uint32_t j = 0; j < m.size(); ++j) {
        if (m[j].size() != n_reads) return -5;
        std::memcpy(out + (size_t)j * n_reads, m[j].data(), sizeof(double) * n_reads);
    }
    *n_kept = (uint32_t)rs.size();
    for (uint32_t c = 0; c < rs.size(); ++c) kept_idx[c] = index_of[rs[c]];
    return 0;
}

/*
 * CPU baseline: many regions (flat, same layout as ref_region with per-region offsets) through
 * the reference's call_pairhmm, whole regions per thread as the reference parallelises
 * (haplotypecaller_engine.cpp:119), one std::thread per requested core.  Returns wall seconds over
 * the whole pass (building the fake BAM records included; it is <1 % of the work), or <0 on error.
 */
double ref_bench_regions(int path, int n_threads, uint32_t n_regions, const uint64_t* reg_read_off, const uint64_t* reg_hap_off,
                         const uint8_t* rbase, const uint8_t* rqual, const uint64_t* roff, const uint32_t* rlen,
                         const uint8_t* rmapq, const uint8_t* hbase, const uint64_t* hoff, const uint32_t* hlen,
                         int32_t min_quality_threshold, int pcr_model, double* checksum)
{
    bind_kernels();
    if (path < 0) path = g_avx512 ? 1 : 0;
    if (path == 1 && !g_avx512) return -2.0;
    if (n_threads < 1) n_threads = 1;
    std::atomic<uint32_t> next{0};
    std::vector<double> sums(n_threads, 0.0);
    std::vector<double> busy(n_threads, 0.0);
    auto worker = [&](int t) {
        CsrGuard g;
        std::vector<uint8_t> buffer((size_t)100 << 20);   /* k_target_mem, haplotypecaller.cpp:25 */
        std::vector<FakeBam> bams;
        for (;;) {
            const uint32_t reg = next.fetch_add(1);
            if (reg >= n_regions) break;
            rovaca::RovacaMemoryPool pool{buffer.data(), buffer.size()};
            const uint64_t r0 = reg_read_off[reg], r1 = reg_read_off[reg + 1];
            const uint64_t h0 = reg_hap_off[reg], h1 = reg_hap_off[reg + 1];
            bams.resize(r1 - r0);
            rovaca::ReadVector rs{&pool};
            for (uint64_t k = r0; k < r1; ++k) {
                bams[k - r0].build(rbase + roff[k], rqual + roff[k], rlen[k], rmapq[k]);
                rs.push_back(rovaca::ReadRecord::create(&pool, nullptr, &bams[k - r0].b));
            }
            rovaca::HaplotypeVector hs{&pool};
            for (uint64_t j = h0; j < h1; ++j) {
                auto* b = static_cast<rovaca::pBases>(pool.allocate(sizeof(rovaca::Bases) + hlen[j] + 1, 8));
                b->num = hlen[j];
                std::memcpy(b->data, hbase + hoff[j], hlen[j]);
                b->data[hlen[j]] = 0;
                void* mem = pool.allocate(sizeof(rovaca::Haplotype), alignof(rovaca::Haplotype));
                std::memset(mem, 0, sizeof(rovaca::Haplotype));
                auto* h = static_cast<rovaca::Haplotype*>(mem);
                h->_bases = b;
                hs.push_back(h);
            }
            if (path == 1)   /* the caller's sort, haplotypecaller_engine.cpp:245-248 */
                std::sort(rs.begin(), rs.end(), [](rovaca::pReadRecord a, rovaca::pReadRecord b) { return a->seq_length() < b->seq_length(); });
            auto t0 = std::chrono::steady_clock::now();
            rovaca::DoubleVector2D m = (path == 1)
                                           ? rovaca::call_new_pairhmm002(hs, rs, min_quality_threshold, static_cast<PcrIndelModel>(pcr_model), &pool)
                                           : rovaca::call_old_pairhmm(hs, rs, min_quality_threshold, static_cast<PcrIndelModel>(pcr_model), &pool);
            auto t1 = std::chrono::steady_clock::now();
            busy[t] += std::chrono::duration<double>(t1 - t0).count();
            double s = 0.0;
            for (auto& row : m)
                for (size_t c = 0; c < rs.size(); ++c) s += row[c];
            sums[t] += s;
        }
    };
    auto w0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
    auto w1 = std::chrono::steady_clock::now();
    if (checksum) {
        double s = 0.0;
        for (double v : sums) s += v;
        *checksum = s;
    }
    return std::chrono::duration<double>(w1 - w0).count();
}

/*
 * Many regions (same flat layout) through the reference's call_pairhmm on n_threads threads, results kept: region g's
 * [n_haps][n_reads] matrix at out + lik_off[g], the original indices of the reads left in `rs` at kept_idx + reg_read_off[g],
 * their count in n_kept[g].  For sampled parity checks of big streams.  Reads of a region must come in ascending length
 * when the new path is taken (the caller's sort, haplotypecaller_engine.cpp:245-248, is the input's business here, so that
 * column order is defined).  Returns 0, or <0 on a contract violation.
 */
int ref_regions_flat(int path, int n_threads, uint32_t n_regions, const uint64_t* reg_read_off, const uint64_t* reg_hap_off,
                     const uint8_t* rbase, const uint8_t* rqual, const uint64_t* roff, const uint32_t* rlen, const uint8_t* rmapq,
                     const uint8_t* hbase, const uint64_t* hoff, const uint32_t* hlen, int32_t min_quality_threshold, int pcr_model,
                     const uint64_t* lik_off, double* out, uint32_t* kept_idx, uint32_t* n_kept)
{
    bind_kernels();
    if (path < 0) path = g_avx512 ? 1 : 0;
    if (path == 1 && !g_avx512) return -2;
    if (n_threads < 1) n_threads = 1;
    std::atomic<uint32_t> next{0};
    std::atomic<int> rc_all{0};
    auto worker = [&]() {
        for (;;) {
            const uint32_t reg = next.fetch_add(1);
            if (reg >= n_regions) break;
            const uint64_t r0 = reg_read_off[reg], h0 = reg_hap_off[reg];
            const uint32_t nr = (uint32_t)(reg_read_off[reg + 1] - r0), nh = (uint32_t)(reg_hap_off[reg + 1] - h0);
            if (nr == 0 || nh == 0) {
                n_kept[reg] = 0;
                continue;
            }
            /* ref_region takes offsets relative to the blobs it is handed: pass the whole blobs and this region's slices */
            const int rc = ref_region(path, nr, rbase, rqual, roff + r0, rlen + r0, rmapq + r0, nh, hbase, hoff + h0, hlen + h0,
                                      min_quality_threshold, pcr_model, out + lik_off[reg], kept_idx + r0, n_kept + reg);
            if (rc != 0) rc_all.store(rc);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
    for (auto& x : th) x.join();
    return rc_all.load();
}

}  // extern "C"
