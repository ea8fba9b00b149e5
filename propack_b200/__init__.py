"""propack_b200 -- B200-native (sm_100a) implementation of PROPACK's Lanczos-bidiagonalisation hot path.

The product is ``lib/libpropack_b200.so`` (hand-written CUDA kernels + host drivers behind the
reference's Fortran calling convention, see ``include/propack_b200.h``).  This package is the thin
host-side mirror of the interfaces a PROPACK user knows:

* :mod:`propack_b200.f77`   -- the Fortran-ABI entry points (``dlansvd_`` ...) through ctypes, host arrays;
* :func:`propack_b200.svdp` -- same signature as ``scipy.sparse.linalg._svdp._svdp`` (SciPy's PROPACK wrapper), and
  :func:`propack_b200.svds` -- ``scipy.sparse.linalg.svds(..., solver='propack')``; both also take torch / DLPack device arrays;
* :mod:`propack_b200.matio` / :mod:`propack_b200.hb` -- the matrix files of the reference's example programs.

There is no CPU fallback: importing works anywhere (so the symbol table can be checked), but every
compute call needs the CUDA library and a B200.
"""
from ._lib import lib, library_path, last_error, counters, reset_counters, phase_ms, set_profile  # noqa: F401
from .svdp import svdp, svds, Operator  # noqa: F401

__all__ = ["svdp", "svds", "Operator", "lib", "library_path", "last_error", "counters", "reset_counters", "phase_ms",
           "set_profile"]
