/*
 * ref_glue.h — shared by ref_shim.cpp and adapter_shim.cpp: the reference headers the shims need and the fake BAM
 * record that lets the real ReadRecord be driven.  TEST INFRASTRUCTURE ONLY (see phmm_oracle.h).
 */
#ifndef ORACLE_REF_GLUE_H_
#define ORACLE_REF_GLUE_H_

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory_resource>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <list>
#include <bitset>
#include <memory>
#include <xmmintrin.h>

#include "htslib/sam.h"

/* Reach the two private fields of Allele the path reads (genotype/allele.h:66-67). */
#define private public
#define protected public
#include "genotype/haplotype.h"
#undef private
#undef protected
#include "genotype/read_record.h"
#include "common/utils/rovaca_memory_pool.h"
#include "pairhmm/pairhmm_internal.h"
#include "pairhmm/core/avx.h"
#include "pairhmm/core/avx_impl.h"
#include "pairhmm/rovaca/avx_512_float.h"
#include "pairhmm/rovaca/context.h"


namespace ref_glue
{

inline uint8_t nt16_code(uint8_t c)
{
    switch (c) {
        case 'A': return 1;
        case 'C': return 2;
        case 'G': return 4;
        case 'T': return 8;
        case 'N': return 15;
        default: return 15;
    }
}

/* A BAM record carrying just what ReadRecord's accessors read (read_record.cpp:73-95,176-210). */
struct FakeBam
{
    bam1_t b;
    std::vector<uint8_t> data;
    void build(const uint8_t* bases, const uint8_t* qual, uint32_t len, uint8_t mapq)
    {
        std::memset(&b, 0, sizeof(b));
        const uint32_t l_qname = 4;
        data.assign(l_qname + (len + 1) / 2 + len, 0);
        data[0] = 'r';
        uint8_t* seq = data.data() + l_qname;
        for (uint32_t i = 0; i < len; ++i) seq[i >> 1] |= nt16_code(bases[i]) << ((~i & 1) << 2);
        std::memcpy(seq + (len + 1) / 2, qual, len);
        b.core.l_qname = l_qname;
        b.core.l_extranul = 2;
        b.core.n_cigar = 0;
        b.core.l_qseq = (int32_t)len;
        b.core.qual = mapq;
        b.data = data.data();
        b.l_data = (int)data.size();
        b.m_data = (uint32_t)data.size();
    }
};


/* Haplotype objects the path can read: only the inline accessors get_display_string()/length() are used
 * (genotype/allele.h:104,107), so the object is raw memory with the two fields set. */
inline rovaca::pHaplotype make_haplotype(std::pmr::memory_resource* pool, const uint8_t* bases, uint32_t len)
{
    auto* b = static_cast<rovaca::pBases>(pool->allocate(sizeof(rovaca::Bases) + len + 1, 8));
    b->num = len;
    std::memcpy(b->data, bases, len);
    b->data[len] = 0;
    void* mem = pool->allocate(sizeof(rovaca::Haplotype), alignof(rovaca::Haplotype));
    std::memset(mem, 0, sizeof(rovaca::Haplotype));
    auto* h = static_cast<rovaca::Haplotype*>(mem);
    h->_bases = b;
    h->_is_symbolic = 0;
    return h;
}

}  // namespace ref_glue
#endif
