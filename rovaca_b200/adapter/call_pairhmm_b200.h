// call_pairhmm_b200.h — the B200 engine behind the reference's own PairHMM seam, in the reference's own types.
//
// Replaces (paths relative to /root/reference/src/haplotypecaller/):
//   rovaca::init_pairhmm_ptr(bool)          pairhmm/pairhmm_engine.cpp:21-53   -> rovaca::init_pairhmm_ptr_b200(device)
//   what rovaca::call_pairhmm points at     pairhmm/pairhmm_engine.h:14-15     -> rovaca::call_pairhmm_b200
// Compiles inside the reference tree (it includes the reference's headers) and links against libphmm_b200.so through
// include/phmm_b200.h.  INTEGRATION.md shows the three-line change at haplotypecaller.cpp:49.
#ifndef ROVACA_CALL_PAIRHMM_B200_H_
#define ROVACA_CALL_PAIRHMM_B200_H_

#include <cstddef>

#include "common/enum.h"
#include "genotype/forward.h"

namespace rovaca
{

// Creates the engine on CUDA device `device` (-1 = current) and binds the global rovaca::call_pairhmm pointer to
// call_pairhmm_b200.  Throws std::runtime_error when no sm_100 device / library is usable: there is no CPU fallback.
// If the environment names devices (ROVACA_B200_DEVICES=all, or a comma list of CUDA ordinals) that list wins, so that a
// deployed binary can be pointed at every GPU of the box without a rebuild.
void init_pairhmm_ptr_b200(int device = -1);
// The same over several GPUs: the reference is ONE process whose pool workers all call through the one pointer
// (haplotypecaller.cpp:49,97; haplotypecaller_engine.cpp:250), so one engine drives all of them — every batch of regions goes
// to the device with the least outstanding work.  n_devices == 0: every visible device.
void init_pairhmm_ptr_b200_devices(const int* devices, int n_devices);
void shutdown_pairhmm_b200();

// Same contract as call_old_pairhmm / call_new_pairhmm002 (pairhmm/pairhmm_internal.cpp:505-594, 596-753): returns
// [haplotype][read] log10 likelihoods allocated from `pool`, first rs.size()-after-filtering columns valid, tail NaN;
// erases the poorly modelled reads from `rs`.  Re-entrant: callable from all of the reference's pool workers at once.
DoubleVector2D call_pairhmm_b200(const HaplotypeVector& hs, ReadVector& rs, int32_t min_quality_threshold, PcrIndelModel pcr_option,
                                 pMemoryPool pool);

// The asynchronous form of the same call, for the region loop of a pool task (haplotypecaller_engine.cpp:145-284 walks the
// regions of one RegionSource serially; one synchronous call per region leaves a GPU idle unless hundreds of tasks are in
// flight).  submit hands the region to the engine and returns at once; wait returns what call_pairhmm_b200 would have and
// erases the filtered reads from `rs`.  Between the two, `hs`, `rs`, the ReadRecords and `pool` must stay alive and the pool
// must not be reset or rewound (the result and the staging live in it): a caller that keeps W regions in flight needs W
// region pools (INTEGRATION.md shows the loop).
struct PairhmmCallB200;
PairhmmCallB200* call_pairhmm_b200_submit(const HaplotypeVector& hs, const ReadVector& rs, int32_t min_quality_threshold,
                                          PcrIndelModel pcr_option, pMemoryPool pool);
DoubleVector2D call_pairhmm_b200_wait(PairhmmCallB200* call, ReadVector& rs);

// A window of regions at once: submits all of them, then collects them in order.  regions[k].likelihoods receives what
// call_pairhmm(…) would have returned for region k (allocated from regions[k].pool), regions[k].rs loses its filtered reads.
struct PairhmmRegionB200
{
    const HaplotypeVector* hs;
    ReadVector* rs;
    pMemoryPool pool;
    DoubleVector2D likelihoods;
};
void call_pairhmm_b200_window(PairhmmRegionB200* regions, size_t n_regions, int32_t min_quality_threshold, PcrIndelModel pcr_option);

}  // namespace rovaca

#endif
