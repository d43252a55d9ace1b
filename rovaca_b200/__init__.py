"""rovaca_b200 — a B200-native (sm_100a) PairHMM likelihood engine behind Rovaca's ``call_pairhmm`` contract.

The product is ``libphmm_b200.so`` (CUDA kernels + C++ host engine, C ABI in ``include/phmm_b200.h``);
this package is the thin Python door used by the tests and the bench.  There is no CPU fallback: if the
library is missing or no sm_100 GPU is present, construction fails loudly.
"""
from .pairhmm import (PairHmmEngine, PhmmError, PCR_NONE, PCR_HOSTILE, PCR_AGGRESSIVE, PCR_CONSERVATIVE,
                      FMA_AVX512, FMA_NONE, FP64_WHEN_OBSERVABLE, FP64_ALWAYS, EPILOGUE_AUTO, EPILOGUE_HOST, EPILOGUE_DEVICE,
                      SW_REALIGN, SW_ASSEMBLE, marginalize, library_path, build_library)
from .workloads import RegionBatch

__all__ = ["PairHmmEngine", "PhmmError", "RegionBatch", "PCR_NONE", "PCR_HOSTILE", "PCR_AGGRESSIVE", "PCR_CONSERVATIVE",
           "FMA_AVX512", "FMA_NONE", "FP64_WHEN_OBSERVABLE", "FP64_ALWAYS", "EPILOGUE_AUTO", "EPILOGUE_HOST", "EPILOGUE_DEVICE",
           "SW_REALIGN", "SW_ASSEMBLE", "marginalize", "library_path", "build_library"]
