// workgen.cpp — synthetic active-region streams for the benches and the full-size tests, as a small C library
// of its own (libphmm_workgen.so; no CUDA, not part of the engine).  The numpy generators in rovaca_b200/workloads.py build
// regions one Python object at a time, which is fine for a few thousand regions and hopeless for the >= 1 M regions of the
// stream configuration (BASELINE.json configs[4]).
//
// Every region is a pure function of (spec, region index): its shape comes from one counter-seeded random stream, its content
// from a second one.  So any process can generate any slice of a stream by itself — which is what lets N ranks shard ONE common
// stream without shipping it — and slices generated separately concatenate to exactly the stream generated in one piece.
//
// A region: one reference haplotype of H uniform ACGT, alternates carrying 1-3 SNPs and (with probability `indel`) one
// insertion or deletion of 1-8 bases; reads sampled from a random haplotype at a random offset with `sub_rate` substitutions,
// or (with probability `junk`, or when no haplotype is long enough) uniform random sequence; base qualities 80 % Q30-40,
// 15 % Q20-29, 5 % Q2-19; mapq 60 (5 %: 20-40); reads sorted by length as the caller does on AVX-512 hosts
// (haplotypecaller_engine.cpp:245-248).  Shapes as in SURVEY.md §8d: haplotypes per region and their lengths follow the
// apply/test/casedata/gatk_result.txt distribution (mean 8.8, 201-499), reads per region Poisson(41); a `post_trim` fraction of
// the regions has the short post-trim shapes (H 41-120, R <= H; genotype/utils/adapter_utils.cpp:92-220).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

extern "C" {

typedef struct wg_spec
{
    uint64_t seed;
    double hap_lambda;        // haplotypes per region = clamp(Poisson(hap_lambda) + 1, hap_min, hap_max); 0 = hap_min exactly
    uint32_t hap_min, hap_max;
    double read_lambda;       // reads per region = max(1, Poisson(read_lambda)); 0 = read_fixed exactly
    uint32_t read_fixed;
    uint32_t hap_lo, hap_hi;  // reference haplotype length, uniform
    uint32_t read_lo, read_hi;
    double post_trim;         // fraction of regions with post-trim shapes
    uint32_t pt_hap_lo, pt_hap_hi, pt_read_lo, pt_read_hi;   // ... H uniform in [pt_hap_lo, pt_hap_hi], R in [pt_read_lo, min(H, pt_read_hi)]
    double junk, indel, sub_rate;
} wg_spec;

}  // extern "C"

namespace
{

struct Rng   // splitmix64: a counter-seeded stream per (region, purpose)
{
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }          // [0, n)
    uint32_t range(uint32_t lo, uint32_t hi) { return lo + below(hi - lo + 1); }                      // [lo, hi]
    uint32_t poisson(double lambda)   // Knuth; lambda <= ~50 here
    {
        const double limit = std::exp(-lambda);
        double p = 1.0;
        uint32_t k = 0;
        do {
            ++k;
            p *= uni();
        } while (p > limit);
        return k - 1;
    }
};

inline uint64_t region_seed(const wg_spec& sp, uint64_t g, uint64_t purpose)
{
    Rng r(sp.seed * 0xD1342543DE82EF95ull + g * 0x2545F4914F6CDD1Dull + purpose * 0x9E3779B97F4A7C15ull + 1);
    r.next();
    return r.next();
}

struct HapPlan
{
    uint32_t len, n_snp, indel_len;
    bool has_indel, is_ins;
};

struct Shape
{
    uint32_t H, n_haps, n_reads;
    std::vector<HapPlan> haps;
    std::vector<uint32_t> read_len;   // ascending
};

void region_shape(const wg_spec& sp, uint64_t g, Shape& out, bool lengths)
{
    Rng r(region_seed(sp, g, 0));
    const bool pt = r.uni() < sp.post_trim;
    out.n_haps = sp.hap_lambda > 0 ? std::min(sp.hap_max, std::max(sp.hap_min, r.poisson(sp.hap_lambda) + 1)) : sp.hap_min;
    out.n_reads = sp.read_lambda > 0 ? std::max(1u, r.poisson(sp.read_lambda)) : sp.read_fixed;
    out.H = pt ? r.range(sp.pt_hap_lo, sp.pt_hap_hi) : r.range(sp.hap_lo, sp.hap_hi);
    if (!lengths) return;
    out.haps.resize(out.n_haps);
    out.haps[0] = HapPlan{out.H, 0, 0, false, false};
    for (uint32_t j = 1; j < out.n_haps; ++j) {
        HapPlan h{};
        h.n_snp = r.range(1, 3);
        h.has_indel = r.uni() < sp.indel && out.H > 20;
        h.is_ins = r.uni() < 0.5;
        h.indel_len = r.range(1, 8);
        h.len = out.H;
        if (h.has_indel) h.len = h.is_ins ? out.H + h.indel_len : out.H - std::min(h.indel_len, out.H - 2);
        out.haps[j] = h;
    }
    const uint32_t rlo = pt ? sp.pt_read_lo : sp.read_lo;
    const uint32_t rhi = pt ? std::max(rlo, std::min(out.H, sp.pt_read_hi)) : sp.read_hi;
    out.read_len.resize(out.n_reads);
    for (uint32_t k = 0; k < out.n_reads; ++k) out.read_len[k] = r.range(rlo, rhi);
    std::sort(out.read_len.begin(), out.read_len.end());
}

const uint8_t kAcgt[4] = {'A', 'C', 'G', 'T'};

inline uint8_t qual_of(Rng& r)
{
    const double u = r.uni();
    if (u < 0.80) return (uint8_t)r.range(30, 40);
    if (u < 0.95) return (uint8_t)r.range(20, 29);
    return (uint8_t)r.range(2, 19);
}

void region_fill(const wg_spec& sp, uint64_t g, const Shape& sh, uint8_t* hap_bases, uint8_t* read_bases, uint8_t* read_qual, uint8_t* read_mapq)
{
    Rng r(region_seed(sp, g, 1));
    std::vector<uint8_t*> hp(sh.n_haps);
    {
        uint8_t* at = hap_bases;
        for (uint32_t j = 0; j < sh.n_haps; ++j) {
            hp[j] = at;
            at += sh.haps[j].len;
        }
    }
    uint8_t* ref = hp[0];
    for (uint32_t c = 0; c < sh.H; c += 32) {                       // 2 bits per base, 32 bases per draw
        uint64_t w = r.next();
        const uint32_t n = std::min(32u, sh.H - c);
        for (uint32_t k = 0; k < n; ++k, w >>= 2) ref[c + k] = kAcgt[w & 3];
    }
    for (uint32_t j = 1; j < sh.n_haps; ++j) {
        const HapPlan& h = sh.haps[j];
        uint8_t* d = hp[j];
        if (!h.has_indel) {
            memcpy(d, ref, sh.H);
        } else if (h.is_ins) {
            const uint32_t p = r.range(1, sh.H - 1);
            memcpy(d, ref, p);
            for (uint32_t k = 0; k < h.indel_len; ++k) d[p + k] = kAcgt[r.below(4)];
            memcpy(d + p + h.indel_len, ref + p, sh.H - p);
        } else {
            const uint32_t cut = sh.H - h.len;
            const uint32_t p = r.range(1, sh.H - cut - 1);
            memcpy(d, ref, p);
            memcpy(d + p, ref + p + cut, sh.H - p - cut);
        }
        for (uint32_t s = 0; s < h.n_snp; ++s) {
            const uint32_t p = r.below(h.len);
            const uint8_t b = d[p];
            const uint32_t cur = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3;
            d[p] = kAcgt[(cur + 1 + r.below(3)) & 3];
        }
    }
    uint8_t* rb = read_bases;
    uint8_t* rq = read_qual;
    const uint32_t sub_thr = (uint32_t)(sp.sub_rate * 4294967296.0);
    for (uint32_t k = 0; k < sh.n_reads; ++k) {
        const uint32_t R = sh.read_len[k];
        const uint32_t j = r.below(sh.n_haps);
        const bool junk = r.uni() < sp.junk || sh.haps[j].len < R;
        if (junk) {
            for (uint32_t c = 0; c < R; c += 32) {
                uint64_t w = r.next();
                const uint32_t n = std::min(32u, R - c);
                for (uint32_t q = 0; q < n; ++q, w >>= 2) rb[c + q] = kAcgt[w & 3];
            }
        } else {
            const uint32_t o = r.below(sh.haps[j].len - R + 1);
            memcpy(rb, hp[j] + o, R);
            for (uint32_t c = 0; c < R; ++c)
                if ((uint32_t)(r.next() >> 32) < sub_thr) rb[c] = kAcgt[r.below(4)];
        }
        for (uint32_t c = 0; c < R; ++c) rq[c] = qual_of(r);
        read_mapq[k] = r.uni() < 0.05 ? (uint8_t)r.range(20, 40) : 60;
        rb += R;
        rq += R;
    }
}

template <typename F>
void parallel_regions(uint32_t n, uint32_t threads, F fn)
{
    threads = std::max(1u, std::min(threads, n / 64 + 1));
    if (threads == 1) {
        fn(0u, n);
        return;
    }
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < threads; ++t) {
        const uint32_t a = (uint32_t)((uint64_t)n * t / threads), b = (uint32_t)((uint64_t)n * (t + 1) / threads);
        pool.emplace_back([=] { fn(a, b); });
    }
    for (auto& t : pool) t.join();
}

}  // namespace

extern "C" {

// Step 1: reads and haplotypes of regions [first, first + n).
void wg_counts(const wg_spec* sp, uint64_t first, uint32_t n, uint32_t* n_reads, uint32_t* n_haps, uint32_t threads)
{
    parallel_regions(n, threads, [=](uint32_t a, uint32_t b) {
        Shape sh;
        for (uint32_t g = a; g < b; ++g) {
            region_shape(*sp, first + g, sh, false);
            n_reads[g] = sh.n_reads;
            n_haps[g] = sh.n_haps;
        }
    });
}

// Cells (sum of read lengths x sum of haplotype lengths) and pairs (reads x haplotypes) of every region of [first, first + n),
// without generating a byte:
// what a sharding plan needs of a stream.
void wg_region_cells(const wg_spec* sp, uint64_t first, uint32_t n, uint64_t* cells, uint64_t* pairs, uint32_t threads)
{
    parallel_regions(n, threads, [=](uint32_t a, uint32_t b) {
        Shape sh;
        for (uint32_t g = a; g < b; ++g) {
            region_shape(*sp, first + g, sh, true);
            uint64_t rb = 0, hb = 0;
            for (uint32_t v : sh.read_len) rb += v;
            for (const HapPlan& h : sh.haps) hb += h.len;
            cells[g] = rb * hb;
            if (pairs) pairs[g] = (uint64_t)sh.n_reads * sh.n_haps;
        }
    });
}

// Step 2: their lengths.  reg_read_off / reg_hap_off: [n + 1] running sums of step 1's counts.
void wg_lengths(const wg_spec* sp, uint64_t first, uint32_t n, const uint64_t* reg_read_off, const uint64_t* reg_hap_off,
                uint32_t* read_len, uint32_t* hap_len, uint32_t threads)
{
    parallel_regions(n, threads, [=](uint32_t a, uint32_t b) {
        Shape sh;
        for (uint32_t g = a; g < b; ++g) {
            region_shape(*sp, first + g, sh, true);
            for (uint32_t k = 0; k < sh.n_reads; ++k) read_len[reg_read_off[g] + k] = sh.read_len[k];
            for (uint32_t j = 0; j < sh.n_haps; ++j) hap_len[reg_hap_off[g] + j] = sh.haps[j].len;
        }
    });
}

// Step 3: the bytes.  read_off / hap_off: byte offset of every read / haplotype in its blob (running sums of step 2's lengths).
void wg_fill(const wg_spec* sp, uint64_t first, uint32_t n, const uint64_t* reg_read_off, const uint64_t* reg_hap_off,
             const uint64_t* read_off, const uint64_t* hap_off, uint8_t* read_bases, uint8_t* read_qual, uint8_t* read_mapq,
             uint8_t* hap_bases, uint32_t threads)
{
    parallel_regions(n, threads, [=](uint32_t a, uint32_t b) {
        Shape sh;
        for (uint32_t g = a; g < b; ++g) {
            region_shape(*sp, first + g, sh, true);
            const uint64_t r0 = reg_read_off[g], h0 = reg_hap_off[g];
            region_fill(*sp, first + g, sh, hap_bases + hap_off[h0], read_bases + read_off[r0], read_qual + read_off[r0], read_mapq + r0);
        }
    });
}

// 64-bit digest of every region's result for identity checks between runs, devices and ranks: FNV-1a over the region's
// [n_haps][n_reads] matrix bytes, its kept indices and its kept count.
void wg_digest_regions(uint32_t n, const uint64_t* reg_read_off, const uint64_t* reg_hap_off, const uint64_t* lik_off, const double* lik,
                       const uint32_t* kept_idx, const uint32_t* n_kept, uint64_t* digest, uint32_t threads)
{
    parallel_regions(n, threads, [=](uint32_t a, uint32_t b) {
        for (uint32_t g = a; g < b; ++g) {
            uint64_t h = 0xCBF29CE484222325ull;
            auto eat = [&h](const void* p, size_t bytes) {
                const uint8_t* c = static_cast<const uint8_t*>(p);
                size_t k = 0;
                for (; k + 8 <= bytes; k += 8) {
                    uint64_t w;
                    memcpy(&w, c + k, 8);
                    h = (h ^ w) * 0x100000001B3ull;
                }
                for (; k < bytes; ++k) h = (h ^ c[k]) * 0x100000001B3ull;
            };
            const uint64_t nr = reg_read_off[g + 1] - reg_read_off[g], nh = reg_hap_off[g + 1] - reg_hap_off[g];
            eat(lik + lik_off[g], (size_t)(nr * nh) * 8);
            eat(kept_idx + reg_read_off[g], (size_t)n_kept[g] * 4);
            h = (h ^ n_kept[g]) * 0x100000001B3ull;
            digest[g] = h;
        }
    });
}

}  // extern "C"
