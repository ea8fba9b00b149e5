"""ctypes loader for libpropack_b200.so.  Fails loudly when the CUDA library has not been built."""
from __future__ import annotations

import ctypes as C
import glob
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

COUNTER_NAMES = ("nopx nreorth ndot nitref nrestart nbsvd nlandim nsing nsteps reorth_passes reorth_cols "
                 "reorth_elems reorth_vec_elems launches host_syncs reserved").split()
PHASE_NAMES = "aprod reorth level1 getu0 ritzvec restart host_bsvd".split()


def library_path() -> str:
    return os.path.join(_HERE, "lib", "libpropack_b200.so")


def find_lapack() -> str | None:
    env = os.environ.get("PROPACK_B200_LAPACK")
    if env:
        return env
    try:
        import scipy
        root = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
        hits = sorted(glob.glob(os.path.join(root, "libscipy_openblas*.so")))
        return hits[0] if hits else None
    except Exception:  # scipy absent: the library falls back to liblapack.so.3 / libopenblas.so.0
        return None


def lib():
    """The loaded C-ABI library (no device is touched until a compute entry point is called)."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C propack_b200/csrc`). propack_b200 has no CPU fallback.")
        L = C.CDLL(path)
        L.propack_b200_last_error.restype = C.c_char_p
        L.propack_b200_op_bytes.restype = C.c_double
        L.propack_b200_bench_reorth_d.restype = C.c_double
        L.propack_b200_bench_spmv.restype = C.c_double
        L.propack_b200_bench_gemm_d.restype = C.c_double
        lp = find_lapack()
        if lp:
            L.propack_b200_set_lapack(lp.encode())
        _LIB = L
    return _LIB


def last_error() -> str:
    return (lib().propack_b200_last_error() or b"").decode()


def check(code: int, what: str):
    if code < 0:
        raise RuntimeError(f"propack_b200: {what} failed (code {code}): {last_error()}")
    return code


def counters() -> dict:
    out = (C.c_longlong * 16)()
    lib().propack_b200_get_counters(out)
    return dict(zip(COUNTER_NAMES, list(out)))


def reset_counters():
    lib().propack_b200_reset_counters()


def set_profile(on: bool):
    lib().propack_b200_set_profile(int(bool(on)))


def phase_ms() -> dict:
    ms = (C.c_double * 8)()
    ln = (C.c_longlong * 8)()
    lib().propack_b200_get_phase_ms(ms, ln)
    return {n: {"ms": ms[i], "launches": ln[i]} for i, n in enumerate(PHASE_NAMES)}
