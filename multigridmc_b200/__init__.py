"""multigridmc_b200 -- B200-native (sm_100a) MultigridMC sampling hot path.

The product is the CUDA library ``csrc/libmgmc_b200.so`` behind the C ABI of
``include/mgmc_b200.h``; this package is the thin ctypes loader used by the tests and ``bench.py``.
The reference-compatible C++ host layer lives in ``host/``.  There is no CPU fallback: importing
works without a GPU, every compute call needs one.
"""
from .capi import (Context, Desc, MgmcError, lib, build, host_stencil, strip_partition,  # noqa: F401
                   host_coefficients, periodic_kappa_sq, host_stencil3, make_desc)

__all__ = ["Context", "Desc", "MgmcError", "lib", "build", "host_stencil", "strip_partition",
           "host_coefficients", "periodic_kappa_sq", "host_stencil3", "make_desc"]
