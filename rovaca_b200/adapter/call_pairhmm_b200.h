// call_pairhmm_b200.h — the B200 engine behind the reference's own PairHMM seam, in the reference's own types.
//
// Replaces (paths relative to /root/reference/src/haplotypecaller/):
//   rovaca::init_pairhmm_ptr(bool)          pairhmm/pairhmm_engine.cpp:21-53   -> rovaca::init_pairhmm_ptr_b200(device)
//   what rovaca::call_pairhmm points at     pairhmm/pairhmm_engine.h:14-15     -> rovaca::call_pairhmm_b200
// Compiles inside the reference tree (it includes the reference's headers) and links against libphmm_b200.so through
// include/phmm_b200.h.  INTEGRATION.md shows the three-line change at haplotypecaller.cpp:49.
#ifndef ROVACA_CALL_PAIRHMM_B200_H_
#define ROVACA_CALL_PAIRHMM_B200_H_

#include "common/enum.h"
#include "genotype/forward.h"

namespace rovaca
{

// Creates the engine on CUDA device `device` (-1 = current) and binds the global rovaca::call_pairhmm pointer to
// call_pairhmm_b200.  Throws std::runtime_error when no sm_100 device / library is usable: there is no CPU fallback.
void init_pairhmm_ptr_b200(int device = -1);
void shutdown_pairhmm_b200();

// Same contract as call_old_pairhmm / call_new_pairhmm002 (pairhmm/pairhmm_internal.cpp:505-594, 596-753): returns
// [haplotype][read] log10 likelihoods allocated from `pool`, first rs.size()-after-filtering columns valid, tail NaN;
// erases the poorly modelled reads from `rs`.  Re-entrant: callable from all of the reference's pool workers at once.
DoubleVector2D call_pairhmm_b200(const HaplotypeVector& hs, ReadVector& rs, int32_t min_quality_threshold, PcrIndelModel pcr_option,
                                 pMemoryPool pool);

}  // namespace rovaca

#endif
