"""deeparc-sfm_b200 — B200-native bundle-adjustment engine for the hot path of pureexe/deeparc-sfm.

Layout
  csrc/        CUDA kernels (sm_100a) + the C ABI of include/deeparc_ba.h -> lib/libdeeparc_ba.so
  host/        C++ mirror of the reference's host surface (DeepArcManager, ParameterBlock, solve(),
               the sfm driver) that calls the engine through the C ABI
  capi.py      ctypes binding of the C ABI
  synthetic.py workload generators (the reference datasets are not available)

The directory name carries a hyphen (it mirrors the reference repository's name); import it as
``deeparc_sfm_b200`` (alias package at the repository root).
"""
from . import synthetic  # noqa: F401
from . import capi  # noqa: F401

__all__ = ["synthetic", "capi"]
