/* al_shim.cpp — TEST INFRASTRUCTURE ONLY.  Glue around the reference's own AlleleLikelihoods class template
 * (/root/reference/src/haplotypecaller/genotype/allele_likelihoods.hpp, included in place, never copied): instantiates it
 * over a stand-in allele type, fills it with a likelihood matrix and asks it what the steps after PairHMM ask —
 * best_alleles_breaking_ties (:364-383 -> search_best_allele :589-657) and marginalize (:189-216 -> :702-732) — so that
 * oracle/phmm_oracle.c's restatements (orc_best_alleles, orc_marginalize) can be pinned against the reference itself.
 * Nothing of the reference's logic is restated here. */
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "allele_likelihoods.hpp"
#include "common/utils/rovaca_memory_pool.h"

namespace
{

struct ShimAllele                       /* all the template asks of an allele: is_reference() (index_of_reference, interface_allele_list.hpp:73-81) */
{
    bool ref;
    uint32_t index;
    bool is_reference() const { return ref; }
};

struct OneSample : rovaca::InterfaceSampleList
{
    std::string name = "s";
    size_t number_of_samples() const override { return 1; }
    const std::string& get_sample(size_t) const override { return name; }
    int32_t index_of_sample(const std::string& n) const override { return n == name ? 0 : -1; }
};

using Likelihoods = rovaca::AlleleLikelihoods<const uint32_t*, ShimAllele*>;

Likelihoods* make(rovaca::pMemoryPool pool, OneSample* sample, std::vector<ShimAllele>& alleles, std::vector<uint32_t>& evidence,
                  const double* values, uint32_t n_alleles, uint32_t stride, uint32_t n_evidence, int32_t reference_index)
{
    alleles.resize(n_alleles);
    std::pmr::vector<ShimAllele*> list{pool};
    for (uint32_t a = 0; a < n_alleles; ++a) {
        alleles[a] = ShimAllele{(int32_t)a == reference_index, a};
        list.push_back(&alleles[a]);
    }
    evidence.resize(n_evidence);
    Likelihoods::Int32ToTeVectorMap by_sample{pool};
    Likelihoods::TeVector ev{pool};
    for (uint32_t e = 0; e < n_evidence; ++e) {
        evidence[e] = e;
        ev.push_back(&evidence[e]);
    }
    by_sample.insert({0, ev});
    rovaca::DoubleVector3D lik{pool};
    rovaca::DoubleVector2D sample_values{pool};
    for (uint32_t a = 0; a < n_alleles; ++a) {
        rovaca::DoubleVector row{pool};
        row.assign(values + (size_t)a * stride, values + (size_t)a * stride + n_evidence);
        sample_values.push_back(std::move(row));
    }
    lik.push_back(std::move(sample_values));
    auto* allele_list = rovaca::IndexedAlleleList<ShimAllele*>::create(list, pool);
    void* mem = pool->allocate(sizeof(Likelihoods), alignof(Likelihoods));
    return new (mem) Likelihoods{pool, sample, allele_list, by_sample, std::move(lik)};
}

}  // namespace

extern "C" {

/* best_alleles_breaking_ties(0, priority) on values[n_alleles][stride] (first n_evidence columns), priorities[a] per allele
 * (the reference's default priority function casts the allele to its own Allele class, which the stand-in is not: always
 * pass priorities).  reference_index < 0: no reference allele. */
int ref_best_alleles(const double* values, uint32_t n_alleles, uint32_t stride, uint32_t n_evidence, const double* priorities,
                     int32_t reference_index, uint32_t* best, uint32_t* second, double* best_lik, double* second_lik, double* confidence)
{
    if (!values || !best || !second || !priorities || n_alleles == 0) return -1;
    std::vector<uint8_t> buffer((size_t)64 << 20);
    rovaca::RovacaMemoryPool pool{buffer.data(), buffer.size()};
    OneSample sample;
    std::vector<ShimAllele> alleles;
    std::vector<uint32_t> evidence;
    Likelihoods* lk = make(&pool, &sample, alleles, evidence, values, n_alleles, stride, n_evidence, reference_index);
    auto result = lk->best_alleles_breaking_ties(0, [&](void* a) -> double { return priorities[static_cast<ShimAllele*>(a)->index]; });
    if (result.size() != n_evidence) return -2;
    for (uint32_t e = 0; e < n_evidence; ++e) {
        auto* ba = result[e];
        best[e] = ba->best_allele ? ba->best_allele->index : UINT32_MAX;
        second[e] = ba->second_best_allele ? ba->second_best_allele->index : UINT32_MAX;
        if (best_lik) best_lik[e] = ba->likelihood;
        if (second_lik) second_lik[e] = ba->second_best_likelihood;
        if (confidence) confidence[e] = ba->confidence;
    }
    return 0;
}

/* marginalize(new_alleles, new_to_old_allele_map) built from old_to_new[a] (< 0: allele a maps to nothing); the new matrix
 * [n_new][n_evidence] is read back through sample_likelihoods(0). */
int ref_marginalize(const double* values, uint32_t n_old, uint32_t stride, uint32_t n_evidence, const int32_t* old_to_new,
                    uint32_t n_new, double* out)
{
    if (!values || !old_to_new || !out) return -1;
    std::vector<uint8_t> buffer((size_t)64 << 20);
    rovaca::RovacaMemoryPool pool{buffer.data(), buffer.size()};
    OneSample sample;
    std::vector<ShimAllele> alleles, fresh(n_new);
    std::vector<uint32_t> evidence;
    Likelihoods* lk = make(&pool, &sample, alleles, evidence, values, n_old, stride, n_evidence, 0);
    std::pmr::vector<ShimAllele*> new_alleles{&pool};
    std::pmr::map<ShimAllele*, std::pmr::vector<ShimAllele*>> new_to_old{&pool};
    for (uint32_t n = 0; n < n_new; ++n) {
        fresh[n] = ShimAllele{n == 0, n};
        new_alleles.push_back(&fresh[n]);
        new_to_old.insert({&fresh[n], std::pmr::vector<ShimAllele*>{&pool}});
    }
    for (uint32_t a = 0; a < n_old; ++a)
        if (old_to_new[a] >= 0) new_to_old.at(&fresh[old_to_new[a]]).push_back(&alleles[a]);
    auto* m = lk->marginalize<ShimAllele*>(new_alleles, new_to_old);
    auto& v = m->sample_likelihoods(0);
    if (v.size() != n_new) return -2;
    for (uint32_t n = 0; n < n_new; ++n) {
        if (v[n].size() != n_evidence) return -3;
        for (uint32_t e = 0; e < n_evidence; ++e) out[(size_t)n * n_evidence + e] = v[n][e];
    }
    return 0;
}

}  // extern "C"
