"""suriko-b200: B200-native bundle-adjustment engine, drop-in for suriko-engine's Kanatani BA path.

The compute path is the sm_100a shared library surikatoko_b200/_lib/libsrk_ba.so (C ABI: include/srk/ba_c_api.h).
This package is the thin Python binding used by the tests and the benchmark; the C++ drop-in adapter is
include/suriko_compat/bundle-adj-kanatani.h.  There is no CPU fallback.
"""
from .capi import (BAProblem, BAOptions, BAReport, Engine, SrkError, load_library, STOP_REASONS,  # noqa: F401
                   SOLVER_AUTO, SOLVER_DENSE_CHOLESKY, SOLVER_BLOCK_PCG)
from .ba import BundleAdjustmentKanatani, BundleAdjustmentKanataniTermCriteria  # noqa: F401
from .mvf import MultiViewIterativeFactorizer  # noqa: F401
