/*
 * sw_shim.c — C door into the UNMODIFIED reference Smith-Waterman (common/base/lib/sw_avx/sw_avx.c, sw_avx2.c, compiled in
 * place by oracle/Makefile).  TEST INFRASTRUCTURE ONLY (see sw_oracle.h).  Glue only: fills lib_sw_avx the way
 * AlignmentUtils::create_read_aligned_to_ref does (genotype/utils/alignment_utils.cpp:574-579) and copies the result out.
 */
#include <string.h>

#include "smithwaterman_common.h"

/* isa: 0 = sw_avx2_run, 1 = sw_avx512_run, -1 = what sw_avx_init binds on this CPU.  is_assemble picks the parameter set
 * (sw_avx.c:13-14, 432-436).  Returns the CIGAR length, -1 when the reference gave up (sequence >= 2048), -2 on bad use. */
int ref_sw_align(int isa, int is_assemble, const uint8_t* seq1, int len1, const uint8_t* seq2, int len2, uint32_t* cigar, int cigar_cap,
                 int* alignment_offset, int* score, int* bound_isa)
{
    if (len1 < 0 || len2 < 0 || len1 > 2048 || len2 > 2048) return -2;
    p_lib_sw_avx sw = sw_avx_init((uint8_t)is_assemble);
    if (!sw) return -2;
    if (bound_isa) *bound_isa = sw->avx_function == sw_avx512_run ? 1 : 0;
    sw->len1 = (int16_t)len1;
    memcpy(sw->seq1, seq1, (size_t)len1);
    sw->len2 = (int16_t)len2;
    memcpy(sw->seq2, seq2, (size_t)len2);
    sw->alignment_offset = -7;
    sw->cigar_count = 0;
    if (isa == 1) sw_avx512_run(sw); else if (isa == 0) sw_avx2_run(sw); else sw->avx_function(sw);
    int n = (int)sw->cigar_count;
    *alignment_offset = sw->alignment_offset;
    *score = sw->score;
    if ((len1 > len2 ? len1 : len2) >= 2048) n = -1;
    else {
        if (n > cigar_cap) n = cigar_cap;
        memcpy(cigar, sw->cigar, (size_t)n * sizeof(uint32_t));
    }
    sw_avx_finit(sw);
    return n;
}

/* CPU baseline: n alignments (flat blobs, pair k = ref at rbytes + roff[k], read at qbytes + qoff[k]) through the reference's
 * sw->avx_function, one lib_sw_avx per thread as the reference keeps one per pool worker.  Returns wall seconds; checksum =
 * sum of scores + alignment offsets + CIGAR lengths. */
#include <pthread.h>
#include <time.h>

typedef struct {
    int tid, nthreads, is_assemble;
    uint32_t n;
    const uint8_t *rbytes, *qbytes;
    const uint64_t *roff, *qoff;
    const uint32_t *rlen, *qlen;
    long long sum;
} sw_bench_arg;

static void* sw_bench_worker(void* v)
{
    sw_bench_arg* a = (sw_bench_arg*)v;
    p_lib_sw_avx sw = sw_avx_init((uint8_t)a->is_assemble);
    long long s = 0;
    for (uint32_t k = (uint32_t)a->tid; k < a->n; k += (uint32_t)a->nthreads) {
        sw->len1 = (int16_t)a->rlen[k];
        memcpy(sw->seq1, a->rbytes + a->roff[k], a->rlen[k]);
        sw->len2 = (int16_t)a->qlen[k];
        memcpy(sw->seq2, a->qbytes + a->qoff[k], a->qlen[k]);
        sw->avx_function(sw);
        s += sw->score + sw->alignment_offset + (long long)sw->cigar_count;
    }
    a->sum = s;
    sw_avx_finit(sw);
    return 0;
}

double ref_sw_bench(int nthreads, int is_assemble, uint32_t n, const uint8_t* rbytes, const uint64_t* roff, const uint32_t* rlen,
                    const uint8_t* qbytes, const uint64_t* qoff, const uint32_t* qlen, long long* checksum)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    sw_bench_arg arg[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; ++t) {
        arg[t] = (sw_bench_arg){t, nthreads, is_assemble, n, rbytes, qbytes, roff, qoff, rlen, qlen, 0};
        pthread_create(&th[t], 0, sw_bench_worker, &arg[t]);
    }
    long long sum = 0;
    for (int t = 0; t < nthreads; ++t) {
        pthread_join(th[t], 0);
        sum += arg[t].sum;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (checksum) *checksum = sum;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
