"""hgaprec_b200 -- B200-native CAVI engine for (Hierarchical) Poisson Factorization.

The product is the C-ABI shared library ``libhpf_b200.so`` (include/hpf_cuda.h,
sources under hgaprec_b200/csrc).  This package is a thin ctypes binding used by
the tests and bench.py; it contains no numerical code and has no CPU fallback.
"""
from .capi import (Engine, HpfError, HIER, BIAS, BINARY, JACOBI, LOGL, THETA, BETA, THETARATE, BETARATE,
                   THETABIAS, BETABIAS, LIB_PATH, load_library, build_library, comm_unique_id, partition_users)

__all__ = ["Engine", "HpfError", "HIER", "BIAS", "BINARY", "JACOBI", "LOGL", "THETA", "BETA", "THETARATE",
           "BETARATE", "THETABIAS", "BETABIAS", "LIB_PATH", "load_library", "build_library",
           "comm_unique_id", "partition_users"]
