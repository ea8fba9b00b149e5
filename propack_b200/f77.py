"""Fortran-ABI entry points of libpropack_b200.so through ctypes (host numpy arrays).

Every function here calls the symbol a Fortran program linked against PROPACK would call
(``dlansvd_``, ``dlansvd_irl_``, ``dlanbpro_``, ``dreorth_``, ``dgetu0_``, ``dgemm_ovwr_left_`` ... and
the s/c/z variants) with the reference's argument lists (double/dlansvd.F:1-3 etc.): all scalars by
reference, hidden CHARACTER lengths appended.  Used by the parity tests and by :mod:`propack_b200.svdp`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib

PREFIX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
          np.dtype(np.complex128): "z"}
REAL = {"s": np.float32, "d": np.float64, "c": np.float32, "z": np.float64}
CPLX_GEMM = {"c": "csgemm_ovwr_left_", "z": "zdgemm_ovwr_left_"}


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i(v):
    return C.byref(C.c_int(int(v)))


def _r(pfx, v):
    return C.byref((C.c_float if REAL[pfx] is np.float32 else C.c_double)(v))


def _is_torch_tensor(A):
    mod = type(A).__module__
    return mod == "torch" or mod.startswith("torch.")


class Operator:
    """A linear operator the drivers can use.

    * scipy sparse matrix / dense ndarray -> registered on the device (built-in APROD, handle in IPARM(1));
    * torch tensor (dense or sparse CSR; a CUDA tensor never leaves the device) or any object with ``__dlpack__`` -> same;
    * anything with ``matvec``/``rmatvec`` (a LinearOperator) -> Python APROD callback with the reference's
      contract (dlansvd.F:20-33), host-staged.
    """

    def __init__(self, A, dtype=None):
        import scipy.sparse as sp
        L = lib()
        self.handle = 0
        self._cb = None
        if sp.issparse(A):
            A = sp.csr_array(A)
            dtype = np.dtype(dtype or A.dtype)
            if dtype not in PREFIX:
                dtype = np.dtype(np.complex128 if np.iscomplexobj(A) else np.float64)
            A = A.astype(dtype)
            A.sort_indices()
            self.dtype, self.pfx, self.shape = dtype, PREFIX[dtype], A.shape
            rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
            ci = np.ascontiguousarray(A.indices, dtype=np.int32)
            va = np.ascontiguousarray(A.data)
            self.handle = check(getattr(L, f"propack_b200_csr_create_{self.pfx}")(
                C.c_int(A.shape[0]), C.c_int(A.shape[1]), _p(rp), _p(ci), _p(va), C.c_int(0)), "csr_create")
            self.nnz = int(A.nnz)
        elif _is_torch_tensor(A):
            self._init_from_torch(A, dtype)
        elif isinstance(A, np.ndarray):
            dtype = np.dtype(dtype or A.dtype)
            if dtype not in PREFIX:
                dtype = np.dtype(np.complex128 if np.iscomplexobj(A) else np.float64)
            Af = np.asfortranarray(A, dtype=dtype)
            self.dtype, self.pfx, self.shape = dtype, PREFIX[dtype], Af.shape
            self.handle = check(getattr(L, f"propack_b200_dense_create_{self.pfx}")(
                C.c_int(Af.shape[0]), C.c_int(Af.shape[1]), _p(Af), C.c_long(Af.shape[0])), "dense_create")
        elif hasattr(A, "__dlpack__") and not hasattr(A, "matvec"):
            import torch
            self._init_from_torch(torch.from_dlpack(A), dtype)
        else:  # LinearOperator-like
            dtype = np.dtype(dtype or getattr(A, "dtype", np.float64))
            if dtype not in PREFIX:
                dtype = np.dtype(np.float64)
            self.dtype, self.pfx, self.shape = dtype, PREFIX[dtype], tuple(A.shape)
            m, n = self.shape
            isz = dtype.itemsize
            vp = C.c_void_p

            def cb(transa, m_, n_, x, y, parm, iparm, tlen):
                t = transa[0].decode().lower()
                nx, ny = (n, m) if t == "n" else (m, n)
                xv = np.frombuffer((C.c_char * (nx * isz)).from_address(x), dtype=dtype)
                yv = np.frombuffer((C.c_char * (ny * isz)).from_address(y), dtype=dtype)
                yv[:] = A.matvec(xv) if t == "n" else A.rmatvec(xv)

            self._cbtype = C.CFUNCTYPE(None, C.POINTER(C.c_char), C.POINTER(C.c_int), C.POINTER(C.c_int), vp, vp, vp, vp, C.c_size_t)
            self._cb = self._cbtype(cb)
        self.iparm = np.array([self.handle, 0], dtype=np.int32)
        self.parm = np.zeros(2, dtype=self.dtype)

    def _init_from_torch(self, A, dtype):
        """A torch tensor (or anything exposing ``__dlpack__`` was converted to one): a CUDA tensor stays on the device -- a dense one
        is laid out column-major with zeroed padding rows by a device-to-device copy and adopted without passing through the
        host; a sparse CSR one hands its device arrays to ``csr_create`` (cudaMemcpyDefault).  CPU tensors go through numpy."""
        import torch
        L = lib()
        tmap = {torch.float32: np.float32, torch.float64: np.float64, torch.complex64: np.complex64, torch.complex128: np.complex128}
        if A.layout == torch.sparse_csr:
            npdt = np.dtype(dtype or tmap.get(A.dtype, np.float64))
            tdt = {v: k for k, v in tmap.items()}[npdt.type]
            crow = A.crow_indices().to(torch.int32).contiguous()
            col = A.col_indices().to(torch.int32).contiguous()
            val = A.values().to(tdt).contiguous()
            self.dtype, self.pfx, self.shape = npdt, PREFIX[npdt], tuple(A.shape)
            self.handle = check(getattr(L, f"propack_b200_csr_create_{self.pfx}")(
                C.c_int(A.shape[0]), C.c_int(A.shape[1]), C.c_void_p(crow.data_ptr()), C.c_void_p(col.data_ptr()),
                C.c_void_p(val.data_ptr()), C.c_int(0)), "csr_create")
            if A.is_cuda:
                torch.cuda.synchronize()
            self.nnz = int(val.numel())
            return
        if not A.is_cuda:
            return self.__init__(A.detach().cpu().numpy(), dtype)
        npdt = np.dtype(dtype or tmap.get(A.dtype, np.float64))
        tdt = {v: k for k, v in tmap.items()}[npdt.type]
        m, n = A.shape
        ld = (m + 31) // 32 * 32
        buf = torch.zeros((n, ld), dtype=tdt, device=A.device)       # row j of buf = column j of A: column-major, lda = ld
        buf[:, :m] = A.detach().to(tdt).t()
        torch.cuda.synchronize(A.device)
        self._device_store = buf                                      # adopted, not copied: keep it alive
        self.dtype, self.pfx, self.shape = npdt, PREFIX[npdt], (m, n)
        self.handle = check(getattr(L, f"propack_b200_dense_adopt_device_{self.pfx}")(
            C.c_int(m), C.c_int(n), C.c_void_p(buf.data_ptr()), C.c_long(ld)), "dense_adopt_device")

    @property
    def aprod(self):
        if self._cb is not None:
            return self._cb
        return getattr(lib(), f"propack_b200_aprod_{self.pfx}_")

    def bytes_per_product(self, adjoint=False) -> float:
        return float(lib().propack_b200_op_bytes(C.c_int(self.handle), C.c_int(int(adjoint)))) if self.handle else 0.0

    def close(self):
        if self.handle:
            lib().propack_b200_op_destroy(C.c_int(self.handle))
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _options(pfx, delta, eta, anorm, min_relgap=None):
    R = REAL[pfx]
    eps = np.finfo(R).eps
    v = [np.sqrt(eps) if delta is None else delta, eps ** 0.75 if eta is None else eta, anorm]
    if min_relgap is not None:
        v.append(min_relgap)
    return np.array(v, dtype=R)


def _basis_buffers(op, m, n, ucols, vcols, k, U, V, u0):
    """The caller-owned U, V of the Fortran interface.  The reference needs (m, kmax+1) / (n, kmax); this library keeps the
    Krylov bases in HBM and only reads U(:,1) and writes the first k columns, so caller-provided (e.g. pinned) buffers
    with >= max(k,1) columns are accepted too."""
    if U is None:
        U = np.zeros((m, ucols), dtype=op.dtype, order="F")
    if V is None:
        V = np.zeros((n, vcols), dtype=op.dtype, order="F")
    for name, a, rows in (("U", U, m), ("V", V, n)):
        if a.dtype != op.dtype or not a.flags.f_contiguous or a.shape[0] != rows or a.shape[1] < max(k, 1):
            raise ValueError(f"{name} must be Fortran-ordered {op.dtype} with {rows} rows and >= {max(k, 1)} columns")
    if u0 is not None:
        U[:, 0] = u0
    return U, V


def lansvd(op: Operator, k, kmax, tol=0.0, u0=None, delta=None, eta=None, anorm=0.0, cgs=False, elr=True,
           jobu=True, jobv=True, U=None, V=None):
    """xLANSVD through the Fortran ABI (reference double/dlansvd.F:1-3).  Host arrays in and out."""
    pfx, R = op.pfx, REAL[op.pfx]
    m, n = op.shape
    kmax = min(m + 1, n + 1, kmax)
    U, V = _basis_buffers(op, m, n, kmax + 1, kmax, k, U, V, u0)
    sigma = np.zeros(max(k, 1), dtype=R)
    bnd = np.zeros(max(k, 1), dtype=R)
    doption = _options(pfx, delta, eta, anorm)
    ioption = np.array([int(cgs), int(elr)], dtype=np.int32)
    work = np.zeros(8, dtype=R)
    iwork = np.zeros(8, dtype=np.int32)
    kk, info = C.c_int(k), C.c_int(0)
    args = [b"y" if jobu else b"n", b"y" if jobv else b"n", _i(m), _i(n), C.byref(kk), _i(kmax), op.aprod, _p(U), _i(m),
            _p(sigma), _p(bnd), _p(V), _i(n), _r(pfx, tol), _p(work), _i(work.size)]
    if pfx in "cz":
        zwork = np.zeros(8, dtype=op.dtype)
        args += [_p(zwork), _i(zwork.size)]
    args += [_p(iwork), _i(iwork.size), _p(doption), _p(ioption), C.byref(info), _p(op.parm), _p(op.iparm),
             C.c_size_t(1), C.c_size_t(1)]
    getattr(lib(), f"{pfx}lansvd_")(*args)
    if info.value <= -99:
        check(info.value, f"{pfx}lansvd_")
    kc = kk.value
    return dict(U=U[:, :kc], sigma=sigma[:kc], bnd=bnd[:kc], V=V[:, :kc], info=info.value, k=kc, anorm=float(doption[2]))


def lansvd_irl(op: Operator, k, dim, p=None, which="L", maxiter=1000, tol=0.0, u0=None, delta=None, eta=None,
               anorm=0.0, cgs=False, elr=True, min_relgap=0.002, jobu=True, jobv=True, U=None, V=None):
    """xLANSVD_IRL through the Fortran ABI (reference double/dlansvd_irl.F:1-3)."""
    pfx, R = op.pfx, REAL[op.pfx]
    m, n = op.shape
    dim = min(m + 1, n + 1, dim)
    if p is None:
        p = dim - k
    U, V = _basis_buffers(op, m, n, dim + 1, dim, k, U, V, u0)
    sigma = np.zeros(dim + 1, dtype=R)
    bnd = np.zeros(dim + 1, dtype=R)
    doption = _options(pfx, delta, eta, anorm, min_relgap)
    ioption = np.array([int(cgs), int(elr)], dtype=np.int32)
    work = np.zeros(8, dtype=R)
    iwork = np.zeros(8, dtype=np.int32)
    dd, neig, info = C.c_int(dim), C.c_int(k), C.c_int(0)
    args = [which[:1].lower().encode(), b"y" if jobu else b"n", b"y" if jobv else b"n", _i(m), _i(n), C.byref(dd), _i(p),
            C.byref(neig), _i(maxiter), op.aprod, _p(U), _i(m), _p(sigma), _p(bnd), _p(V), _i(n), _r(pfx, tol), _p(work),
            _i(work.size)]
    if pfx in "cz":
        zwork = np.zeros(8, dtype=op.dtype)
        args += [_p(zwork), _i(zwork.size)]
    args += [_p(iwork), _i(iwork.size), _p(doption), _p(ioption), C.byref(info), _p(op.parm), _p(op.iparm),
             C.c_size_t(1), C.c_size_t(1), C.c_size_t(1)]
    getattr(lib(), f"{pfx}lansvd_irl_")(*args)
    if info.value <= -99:
        check(info.value, f"{pfx}lansvd_irl_")
    kc = neig.value
    return dict(U=U[:, :kc], sigma=sigma[:kc], bnd=bnd[:kc], V=V[:, :kc], info=info.value, k=kc, anorm=float(doption[2]))


def reorth(V, vnew, normvnew, index, alpha, iflag, k=None):
    """xREORTH (reference double/dreorth.F:5-6) on host arrays; returns (vnew', ||vnew'||)."""
    V = np.asfortranarray(V)
    pfx = PREFIX[V.dtype]
    R = REAL[pfx]
    n = V.shape[0]
    k = V.shape[1] if k is None else k
    v = np.array(vnew, dtype=V.dtype, copy=True)
    idx = np.array(list(index) + [0, 0], dtype=np.int32)
    nrm = (C.c_float if R is np.float32 else C.c_double)(normvnew)
    work = np.zeros(k + 1, dtype=V.dtype)
    getattr(lib(), f"{pfx}reorth_")(_i(n), _i(k), _p(V), _i(V.shape[0]), _p(v), C.byref(nrm), _p(idx), _r(pfx, alpha),
                                     _p(work), _i(iflag))
    if nrm.value < 0:
        check(-99, f"{pfx}reorth_")
    return v, nrm.value


def gemm_ovwr_left(transb, A, B, n, k, alpha=1.0):
    """xGEMM_OVWR_LEFT (reference double/dgemm_ovwr.F:56-57; complex: zgemm_ovwr.F:6): A <- alpha*A*op(B)."""
    A = np.array(A, order="F", copy=True)
    pfx = PREFIX[A.dtype]
    R = REAL[pfx]
    B = np.asfortranarray(B, dtype=R)
    m = A.shape[0]
    work = np.zeros(8, dtype=A.dtype)
    if pfx in "sd":
        getattr(lib(), f"{pfx}gemm_ovwr_left_")(transb.encode(), _i(m), _i(n), _i(k), _r(pfx, alpha), _p(A), _i(m),
                                                 _r(pfx, 0.0), _p(B), _i(B.shape[0]), _p(work), _i(work.size), C.c_size_t(1))
    else:
        getattr(lib(), CPLX_GEMM[pfx])(transb.encode(), _i(m), _i(n), _i(k), _p(A), _i(m), _p(B), _i(B.shape[0]), _p(work),
                                       _i(work.size), C.c_size_t(1))
    return A


def getu0(op: Operator, transa, j, ntry, basis, icgs=1):
    """xGETU0 (reference double/dgetu0.F:11-12).  Returns (u0, u0norm, anormest, ierr)."""
    pfx, R = op.pfx, REAL[op.pfx]
    m, n = op.shape
    rows = m if transa == "n" else n
    Ub = np.asfortranarray(basis, dtype=op.dtype) if j > 0 else np.zeros((rows, 1), dtype=op.dtype, order="F")
    u0 = np.zeros(rows, dtype=op.dtype)
    cr = C.c_float if R is np.float32 else C.c_double
    u0norm, anormest, ierr = cr(0), cr(0), C.c_int(0)
    work = np.zeros(8, dtype=op.dtype)
    getattr(lib(), f"{pfx}getu0_")(transa.encode(), _i(m), _i(n), _i(j), _i(ntry), _p(u0), C.byref(u0norm), _p(Ub), _i(rows),
                                    op.aprod, _p(op.parm), _p(op.iparm), C.byref(ierr), _i(icgs), C.byref(anormest), _p(work),
                                    C.c_size_t(1))
    if ierr.value <= -99:
        check(ierr.value, f"{pfx}getu0_")
    return u0, u0norm.value, anormest.value, ierr.value


def safescal(x, alpha):
    x = np.array(x, copy=True)
    pfx = PREFIX[x.dtype]
    getattr(lib(), f"{pfx}safescal_")(_i(x.size), _r(pfx, alpha), _p(x))
    return x


def aprod(op: Operator, transa, x):
    """Call the exported built-in APROD directly with host vectors (dlansvd.F:20-33 contract)."""
    m, n = op.shape
    x = np.ascontiguousarray(x, dtype=op.dtype)
    y = np.zeros(m if transa == "n" else n, dtype=op.dtype)
    getattr(lib(), f"propack_b200_aprod_{op.pfx}_")(transa.encode(), _i(m), _i(n), _p(x), _p(y), _p(op.parm), _p(op.iparm),
                                                     C.c_size_t(1))
    return y


def lanbpro(op: Operator, k0, k, U, V, B, rnorm, delta=None, eta=None, anorm=0.0, cgs=False, elr=True):
    """xLANBPRO (reference double/dlanbpro.F:1-2).  U (m,k+1), V (n,k), B (k,2) Fortran-ordered, updated in place."""
    pfx, R = op.pfx, REAL[op.pfx]
    m, n = op.shape
    doption = _options(pfx, delta, eta, anorm)
    ioption = np.array([int(cgs), int(elr)], dtype=np.int32)
    cr = C.c_float if R is np.float32 else C.c_double
    kk, rn, ierr = C.c_int(k), cr(rnorm), C.c_int(0)
    work = np.zeros(8, dtype=R)
    iwork = np.zeros(8, dtype=np.int32)
    args = [_i(m), _i(n), _i(k0), C.byref(kk), op.aprod, _p(U), _i(U.shape[0]), _p(V), _i(V.shape[0]), _p(B), _i(B.shape[0]),
            C.byref(rn), _p(doption), _p(ioption), _p(work)]
    if pfx in "cz":
        args.append(_p(np.zeros(8, dtype=op.dtype)))
    args += [_p(iwork), _p(op.parm), _p(op.iparm), C.byref(ierr)]
    getattr(lib(), f"{pfx}lanbpro_")(*args)
    if ierr.value <= -99:
        check(ierr.value, f"{pfx}lanbpro_")
    return kk.value, rn.value, ierr.value, float(doption[2])


def ritzvec(which, U, V, D, E, k, jobu=True, jobv=True):
    """xRITZVEC (reference double/dritzvec.F:1-2; complex zritzvec.F:1-2).  U (m, dim+1), V (n, dim) hold the Lanczos bases.
    Returns (U[:, :k], V[:, :k], D_out) with D_out = singular values of the bidiagonal (descending)."""
    U = np.array(U, order="F", copy=True)
    V = np.array(V, order="F", copy=True)
    pfx = PREFIX[U.dtype]
    R = REAL[pfx]
    D = np.array(D, dtype=R, copy=True)
    E = np.array(E, dtype=R, copy=True)
    S = np.zeros(max(k, 1), dtype=R)
    dim = D.size
    m, n = U.shape[0], V.shape[0]
    work = np.zeros(8, dtype=R)
    iwork = np.zeros(8, dtype=np.int32)
    args = [which[:1].encode(), b"y" if jobu else b"n", b"y" if jobv else b"n", _i(m), _i(n), _i(k), _i(dim), _p(D), _p(E), _p(S),
            _p(U), _i(m), _p(V), _i(n), _p(work), _i(work.size)]
    if pfx in "cz":
        zwork = np.zeros(8, dtype=U.dtype)
        args += [_p(zwork), _i(zwork.size)]
    args += [_p(iwork), C.c_size_t(1), C.c_size_t(1), C.c_size_t(1)]
    getattr(lib(), f"{pfx}ritzvec_")(*args)
    return U[:, :k], V[:, :k], D


def gemm_ovwr(transa, A, B, m, n, k, alpha=1.0, beta=0.0):
    """xGEMM_OVWR (reference double/dgemm_ovwr.F:5-6): returns alpha*op(A)*B + beta*B, written over B (m x n)."""
    B = np.array(B, order="F", copy=True)
    pfx = PREFIX[B.dtype]
    A = np.asfortranarray(A, dtype=B.dtype)
    work = np.zeros(max(m, 1), dtype=B.dtype)
    getattr(lib(), f"{pfx}gemm_ovwr_")(transa.encode(), _i(m), _i(n), _i(k), _r(pfx, alpha), _p(A), _i(A.shape[0]), _r(pfx, beta),
                                       _p(B), _i(B.shape[0]), _p(work), _i(work.size), C.c_size_t(1))
    return B[:m, :n]


_NRM2 = {"s": "psnrm2_", "d": "pdnrm2_", "c": "pscnrm2_", "z": "pdznrm2_"}
_DOT = {"s": "psdot_", "d": "pddot_", "c": "pcdotc_", "z": "pzdotc_"}
_AXPY = {"s": "psaxpy_", "d": "pdaxpy_", "c": "pcaxpy_", "z": "pzaxpy_"}
_SCAL = {"s": "psscal_", "d": "pdscal_", "c": "pcsscal_", "z": "pzdscal_"}
_ZERO = {"s": "pszero_", "d": "pdzero_", "c": "pczero_", "z": "pzzero_"}


class _C8(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


class _C16(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


def _strided(x, inc):
    """A buffer holding x at stride |inc| (BLAS layout for the given increment)."""
    x = np.ascontiguousarray(x)
    if inc == 1:
        return x.copy()
    buf = np.zeros(1 + (x.size - 1) * abs(inc) if x.size else 1, dtype=x.dtype)
    if inc > 0:
        buf[::inc][:x.size] = x
    else:
        buf[::-inc][:x.size] = x[::-1]
    return buf


def _unstrided(buf, n, inc):
    if inc == 1:
        return buf[:n].copy()
    v = buf[::abs(inc)][:n]
    return v.copy() if inc > 0 else v[::-1].copy()


def nrm2(x, incx=1):
    """pdnrm2 / psnrm2 / pdznrm2 / pscnrm2 (reference double/dblasext.F:6, complex16/zblasext.F:6)."""
    x = np.ascontiguousarray(x)
    pfx = PREFIX[x.dtype]
    f = getattr(lib(), _NRM2[pfx])
    f.restype = C.c_float if REAL[pfx] is np.float32 else C.c_double
    return float(f(_i(x.size), _p(_strided(x, incx)), _i(incx)))


def dotc(x, y, incx=1, incy=1):
    """pddot / psdot / pzdotc / pcdotc (reference double/dblasext.F:121, complex16/zblasext.F:167): conj(x).y"""
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y, dtype=x.dtype)
    pfx = PREFIX[x.dtype]
    f = getattr(lib(), _DOT[pfx])
    f.restype = {"s": C.c_float, "d": C.c_double, "c": _C8, "z": _C16}[pfx]
    r = f(_i(x.size), _p(_strided(x, incx)), _i(incx), _p(_strided(y, incy)), _i(incy))
    return complex(r.re, r.im) if pfx in "cz" else float(r)


def axpy(alpha, x, y, incx=1, incy=1):
    """pdaxpy / psaxpy / pzaxpy / pcaxpy (reference double/dblasext.F:92, complex16/zblasext.F:113): alpha*x + y"""
    x = np.ascontiguousarray(x)
    pfx = PREFIX[x.dtype]
    yb = _strided(np.ascontiguousarray(y, dtype=x.dtype), incy)
    a = np.array([alpha], dtype=x.dtype)
    getattr(lib(), _AXPY[pfx])(_i(x.size), _p(a), _p(_strided(x, incx)), _i(incx), _p(yb), _i(incy))
    return _unstrided(yb, x.size, incy)


def scal(alpha, x, incx=1):
    """pdscal / psscal / pzdscal / pcsscal (reference double/dblasext.F:38, complex16/zblasext.F:60): real alpha * x"""
    x = np.ascontiguousarray(x)
    pfx = PREFIX[x.dtype]
    xb = _strided(x, incx)
    getattr(lib(), _SCAL[pfx])(_i(x.size), _r(pfx, alpha), _p(xb), _i(incx))
    return _unstrided(xb, x.size, incx)


def zero(x, incx=1):
    """pdzero / pszero / pzzero / pczero (reference double/dblasext.F:202, complex16/zblasext.F:344)"""
    x = np.ascontiguousarray(x)
    pfx = PREFIX[x.dtype]
    xb = _strided(x, incx)
    getattr(lib(), _ZERO[pfx])(_i(x.size), _p(xb), _i(incx))
    return xb


def sell_arrays(op: Operator, adjoint=False, panel=0):
    """The sliced jagged-ELL copy of a registered CSR operator (one column panel of it): dict(joff, len8, ci, va, panels, long)."""
    info = (C.c_longlong * 4)()
    check(lib().propack_b200_csr_sell_info(C.c_int(op.handle), C.c_int(int(adjoint)), C.c_int(panel), info), "csr_sell_info")
    ns, stored = int(info[0]), int(info[1])
    rows = op.shape[1] if adjoint else op.shape[0]
    joff = np.zeros(ns + 1, dtype=np.int64)
    len8 = np.zeros(max(rows, 1), dtype=np.uint8)
    ci = np.zeros(max(stored, 1), dtype=np.int32)
    va = np.zeros(max(stored, 1), dtype=op.dtype)
    check(lib().propack_b200_csr_get_sell(C.c_int(op.handle), C.c_int(int(adjoint)), C.c_int(panel), _p(joff), _p(len8), _p(ci), _p(va)),
          "csr_get_sell")
    return dict(joff=joff, len8=len8[:rows], ci=ci[:stored], va=va[:stored], panels=int(info[2]), long=int(info[3]))
