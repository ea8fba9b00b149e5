"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front end for ``oracle/libpropack_oracle.so`` (the CPU restatement of PROPACK's
Lanczos-bidiagonalization path, see ``propack_oracle.hpp``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; the
product package ``propack_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_SFX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d",
        np.dtype(np.complex64): "c", np.dtype(np.complex128): "z"}
_REAL = {"s": np.float32, "d": np.float64, "c": np.float32, "z": np.float64}


def find_lapack() -> str:
    """Path of the LP64 LAPACK the image ships inside scipy (``scipy_`` symbol prefix)."""
    env = os.environ.get("PROPACK_LAPACK_SO")
    if env:
        return env
    import scipy
    root = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    hits = sorted(glob.glob(os.path.join(root, "libscipy_openblas*.so")))
    if not hits:
        raise RuntimeError("no scipy_openblas found; set PROPACK_LAPACK_SO")
    return hits[0]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpropack_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "libpropack_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        if L.oracle_init_lapack(find_lapack().encode()) != 0:
            raise RuntimeError("oracle: could not bind LAPACK bdsqr/bdsdc")
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def csr_pair(A, dtype=None):
    """CSR of A and CSR of A^T (= CSC of A) with sorted int32 indices, as numpy arrays."""
    import scipy.sparse as sp
    A = sp.csr_array(A)
    if dtype is not None:
        A = A.astype(dtype)
    A.sort_indices()
    At = sp.csr_array(A.T)
    At.sort_indices()
    f = lambda M: (M.indptr.astype(np.int32), M.indices.astype(np.int32), np.ascontiguousarray(M.data))
    return f(A), f(At)


class Operator:
    """Either a sparse matrix (built-in OpenMP CSR aprod) or a dense ndarray (numpy callback)."""

    def __init__(self, A, dtype):
        import scipy.sparse as sp
        self.dtype = np.dtype(dtype)
        self.sfx = _SFX[self.dtype]
        self.shape = A.shape
        self.cb = None
        self.arrays = [None] * 6
        if sp.issparse(A):
            (rp, ci, va), (trp, tci, tva) = csr_pair(A, self.dtype)
            self.arrays = [rp, ci, va, trp, tci, tva]
        else:
            Ad = np.asarray(A, dtype=self.dtype)
            AH = np.ascontiguousarray(Ad.conj().T)
            AT = np.ascontiguousarray(Ad.T)
            m, n = Ad.shape
            ptr = C.POINTER(C.c_char)

            def cb(transa, m_, n_, x, y):
                t = chr(transa).lower()
                nx, ny = (n, m) if t == "n" else (m, n)
                xv = np.frombuffer((C.c_char * (nx * self.dtype.itemsize)).from_address(C.addressof(x.contents)), dtype=self.dtype)
                yv = np.frombuffer((C.c_char * (ny * self.dtype.itemsize)).from_address(C.addressof(y.contents)), dtype=self.dtype)
                if t == "n":
                    yv[:] = Ad @ xv
                elif t == "c":
                    yv[:] = AH @ xv
                else:
                    yv[:] = AT @ xv

            self._cbtype = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_int, ptr, ptr)
            self.cb = self._cbtype(cb)

    def args(self):
        return [_p(a) for a in self.arrays] + [self.cb if self.cb is not None else None]


def stats_reset():
    lib().oracle_stats_reset()


def stats():
    out = (C.c_longlong * 10)()
    lib().oracle_stats_get(out)
    names = "nopx nreorth ndot nitref nrestart nbsvd nlandim nsing nsteps reorth_cols".split()
    return dict(zip(names, list(out)))


def _rc(sfx, x):
    return (C.c_float if _REAL[sfx] is np.float32 else C.c_double)(x)


def lansvd(A, k, kmax, tol=0.0, u0=None, delta=None, eta=None, anorm=0.0, cgs=False, elr=True,
           jobu=True, jobv=True, dtype=np.float64):
    """xLANSVD (dlansvd.F:1-291).  Returns dict(U, sigma, bnd, V, info, k)."""
    op = Operator(A, dtype)
    sfx, R = op.sfx, _REAL[op.sfx]
    m, n = op.shape
    kmax = min(m + 1, n + 1, kmax)
    eps = np.finfo(R).eps
    doption = np.array([np.sqrt(eps) if delta is None else delta, eps ** 0.75 if eta is None else eta, anorm], dtype=R)
    ioption = np.array([int(cgs), int(elr)], dtype=np.int32)
    U = np.zeros((m, kmax + 1), dtype=op.dtype, order="F")
    V = np.zeros((n, kmax), dtype=op.dtype, order="F")
    if u0 is not None:
        U[:, 0] = u0
    sigma = np.zeros(k, dtype=R)
    bnd = np.zeros(k, dtype=R)
    kk = C.c_int(k)
    info = C.c_int(0)
    fn = getattr(lib(), f"oracle_lansvd_{sfx}")
    fn(C.c_int(int(jobu)), C.c_int(int(jobv)), C.c_int(m), C.c_int(n), C.byref(kk), C.c_int(kmax), *op.args(),
       _p(U), C.c_long(m), _p(sigma), _p(bnd), _p(V), C.c_long(n), _rc(sfx, tol), _p(doption), _p(ioption),
       C.byref(info))
    kc = kk.value
    return dict(U=U[:, :kc], sigma=sigma[:kc], bnd=bnd[:kc], V=V[:, :kc], info=info.value, k=kc,
                anorm=float(doption[2]))


def lansvd_irl(A, k, dim, p=None, which="L", maxiter=1000, tol=0.0, u0=None, delta=None, eta=None, anorm=0.0,
               cgs=False, elr=True, min_relgap=0.002, jobu=True, jobv=True, dtype=np.float64):
    """xLANSVD_IRL (dlansvd_irl.F:1-419).  ``dim`` = Krylov dimension, ``p`` = shifts per restart."""
    op = Operator(A, dtype)
    sfx, R = op.sfx, _REAL[op.sfx]
    m, n = op.shape
    dim = min(m + 1, n + 1, dim)
    if p is None:
        p = dim - k
    eps = np.finfo(R).eps
    doption = np.array([np.sqrt(eps) if delta is None else delta, eps ** 0.75 if eta is None else eta, anorm,
                        min_relgap], dtype=R)
    ioption = np.array([int(cgs), int(elr)], dtype=np.int32)
    U = np.zeros((m, dim + 1), dtype=op.dtype, order="F")
    V = np.zeros((n, dim), dtype=op.dtype, order="F")
    if u0 is not None:
        U[:, 0] = u0
    sigma = np.zeros(dim + 1, dtype=R)
    bnd = np.zeros(dim + 1, dtype=R)
    dd = C.c_int(dim)
    neig = C.c_int(k)
    info = C.c_int(0)
    fn = getattr(lib(), f"oracle_lansvd_irl_{sfx}")
    fn(C.c_int(ord(which[0].lower())), C.c_int(int(jobu)), C.c_int(int(jobv)), C.c_int(m), C.c_int(n), C.byref(dd),
       C.c_int(p), C.byref(neig), C.c_int(maxiter), *op.args(), _p(U), C.c_long(m), _p(sigma), _p(bnd), _p(V),
       C.c_long(n), _rc(sfx, tol), _p(doption), _p(ioption), C.byref(info))
    kc = neig.value
    return dict(U=U[:, :kc], sigma=sigma[:kc], bnd=bnd[:kc], V=V[:, :kc], info=info.value, k=kc,
                anorm=float(doption[2]))


def larnv(n, dtype=np.float64, iseed=(1, 3, 5, 7)):
    sfx = _SFX[np.dtype(dtype)]
    seed = np.array(iseed, dtype=np.int32)
    x = np.zeros(n, dtype=dtype)
    getattr(lib(), f"oracle_larnv_{sfx}")(_p(seed), C.c_long(n), _p(x))
    return x, seed


def reorth(V, vnew, normvnew, index, alpha, iflag, k=None):
    """dreorth (dreorth.F:5-101) on column-major V[:, :k]; returns (vnew', norm')."""
    V = np.asfortranarray(V)
    sfx = _SFX[V.dtype]
    R = _REAL[sfx]
    n = V.shape[0]
    k = V.shape[1] if k is None else k
    v = np.array(vnew, dtype=V.dtype, copy=True)
    idx = np.array(index, dtype=np.int32)
    nrm = (C.c_float if R is np.float32 else C.c_double)(normvnew)
    getattr(lib(), f"oracle_reorth_{sfx}")(C.c_long(n), C.c_int(k), _p(V), C.c_long(V.shape[0]), _p(v), C.byref(nrm),
                                            _p(idx), _rc(sfx, alpha), C.c_int(iflag))
    return v, nrm.value


def gemm_ovwr_left(transb, A, B, n, k):
    """dgemm_ovwr_left (dgemm_ovwr.F:56-87): returns A[:, :k] @ op(B) as the first n columns of a copy of A."""
    A = np.array(A, order="F", copy=True)
    sfx = _SFX[A.dtype]
    B = np.asfortranarray(B, dtype=_REAL[sfx])
    getattr(lib(), f"oracle_gemm_ovwr_left_{sfx}")(C.c_int(ord(transb)), C.c_long(A.shape[0]), C.c_int(n), C.c_int(k),
                                                    _p(A), C.c_long(A.shape[0]), _p(B), C.c_int(B.shape[0]))
    return A


def csr_aprod(transa, A, x, dtype=np.float64):
    op = Operator(A, dtype)
    m, n = op.shape
    x = np.ascontiguousarray(x, dtype=op.dtype)
    y = np.zeros(m if transa == "n" else n, dtype=op.dtype)
    getattr(lib(), f"oracle_csr_aprod_{op.sfx}")(C.c_int(ord(transa)), C.c_int(m), C.c_int(n), *op.args()[:6], _p(x), _p(y))
    return y


def nrm2(x):
    """pdnrm2 / pdznrm2 (dblasext.F:6-30, OpenMP branch)."""
    x = np.ascontiguousarray(x)
    sfx = _SFX[x.dtype]
    f = getattr(lib(), f"oracle_nrm2_{sfx}")
    f.restype = C.c_float if _REAL[sfx] is np.float32 else C.c_double
    return float(f(C.c_long(x.size), _p(x)))


def dotc(x, y):
    """pddot / pzdotc (dblasext.F:121-147): conj(x).y"""
    x = np.ascontiguousarray(x); y = np.ascontiguousarray(y, dtype=x.dtype)
    sfx = _SFX[x.dtype]
    out = np.zeros(1, dtype=x.dtype)
    getattr(lib(), f"oracle_dotc_{sfx}")(C.c_long(x.size), _p(x), _p(y), _p(out))
    return out[0]


def axpy(alpha, x, y):
    """pdaxpy / pzaxpy (dblasext.F:92-115): returns alpha*x + y"""
    x = np.ascontiguousarray(x); y = np.array(y, dtype=x.dtype, copy=True)
    sfx = _SFX[x.dtype]
    a = np.array([alpha], dtype=x.dtype)
    getattr(lib(), f"oracle_axpy_{sfx}")(C.c_long(x.size), _p(a), _p(x), _p(y))
    return y


def scal(alpha, x):
    """pdscal / pzdscal (dblasext.F:38-60)"""
    x = np.array(x, copy=True)
    sfx = _SFX[x.dtype]
    getattr(lib(), f"oracle_scal_{sfx}")(C.c_long(x.size), _rc(sfx, alpha), _p(x))
    return x


def ritzvec(which, U, V, D, E, k, jobu=True, jobv=True):
    """dritzvec (dritzvec.F:1-199).  U (m, dim+1), V (n, dim) Lanczos bases; returns (U[:, :k], V[:, :k], D_out)."""
    U = np.array(U, order="F", copy=True); V = np.array(V, order="F", copy=True)
    sfx = _SFX[U.dtype]
    R = _REAL[sfx]
    D = np.array(D, dtype=R, copy=True); E = np.array(E, dtype=R, copy=True)
    dim = D.size
    getattr(lib(), f"oracle_ritzvec_{sfx}")(C.c_int(ord(which)), C.c_int(int(jobu)), C.c_int(int(jobv)), C.c_int(U.shape[0]),
                                            C.c_int(V.shape[0]), C.c_int(k), C.c_int(dim), _p(D), _p(E), _p(U),
                                            C.c_long(U.shape[0]), _p(V), C.c_long(V.shape[0]))
    return U[:, :k], V[:, :k], D


def lanbpro(A, k0, k, U, V, B, rnorm, delta=None, eta=None, anorm=0.0, cgs=False, elr=True, dtype=np.float64):
    """xLANBPRO (dlanbpro.F:1-549) in place on U (m,k+1), V (n,k), B (k,2).  Returns (k, rnorm, ierr, anorm)."""
    op = Operator(A, dtype)
    sfx, R = op.sfx, _REAL[op.sfx]
    m, n = op.shape
    eps = np.finfo(R).eps
    doption = np.array([np.sqrt(eps) if delta is None else delta, eps ** 0.75 if eta is None else eta, anorm], dtype=R)
    ioption = np.array([int(cgs), int(elr)], dtype=np.int32)
    kk, ierr = C.c_int(k), C.c_int(0)
    rn = _rc(sfx, rnorm)
    getattr(lib(), f"oracle_lanbpro_{sfx}")(C.c_int(m), C.c_int(n), C.c_int(k0), C.byref(kk), *op.args(), _p(U), C.c_long(U.shape[0]),
                                            _p(V), C.c_long(V.shape[0]), _p(B), C.c_int(B.shape[0]), C.byref(rn), _p(doption),
                                            _p(ioption), C.byref(ierr))
    return kk.value, float(rn.value), ierr.value, float(doption[2])
