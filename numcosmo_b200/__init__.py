"""numcosmo_b200 -- B200 (sm_100a) implementation of NumCosmo's APES density-estimation hot path.

Layout:
  csrc/        CUDA kernels + the C ABI (include/ncm_sd_gpu.h) -> lib/libncm_sd_gpu.so
  host/        C++ mirror of the reference's ncm_stats_dist_* / APES interface over the C ABI
  capi.py      ctypes binding of the C ABI
  stats_dist.py, apes.py   Python mirror of numcosmo_py's Ncm.StatsDist* / APES usage

There is no CPU fallback: importing is cheap, but creating a context without the built library
or without an sm_100 GPU raises.
"""
__version__ = "0.1.0"
