"""Row-sharded multi-GPU front end: one process per GPU.

Rows of A and of the left Lanczos basis U are block-partitioned over the ranks, V-vectors likewise; every rank
holds the CSR of its row block of A (for ``A v``) and the CSR of the transpose of its *column* block (for
``A^H u``), so both products are "all-gather the input vector, then a purely local SpMV".  The host control flow of
xLANSVD / xLANSVD_IRL runs replicated on every rank (it only sees all-reduced scalars, so it is identical
everywhere); the O(k^2) bidiagonal SVD is done redundantly.

The integer work here (partition bounds, shard extraction) is bit-exact against ``scipy.sparse`` slicing and is what
the CPU (gloo, world_size 2) tests exercise; the device side is ``propack_b200_csr_create_sharded_*`` +
``propack_b200_comm_init`` in the C-ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib
from .f77 import PREFIX, REAL, _options, _p


# ----------------------------------------------------------------------------------------------------------
# partition (pure integer arithmetic; mirrors csrc/comm.hpp::shard_slice / shard_bounds)
# ----------------------------------------------------------------------------------------------------------
def slice_len(dim: int, world: int) -> int:
    """Common padded slice length: ceil(dim/world) rounded up to a multiple of 32 elements."""
    per = (dim + world - 1) // world
    return (per + 31) // 32 * 32


def shard_bounds(dim: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the indices owned by `rank` (trailing ranks may own fewer, or none)."""
    s = slice_len(dim, world)
    return min(s * rank, dim), min(s * (rank + 1), dim)


def shard_csr(A, world: int, rank: int):
    """(A[r0:r1, :] as CSR,  (A[:, c0:c1])^T as CSR) with sorted int32 indices -- the two local operands of a rank."""
    import scipy.sparse as sp
    A = sp.csr_array(A)
    m, n = A.shape
    r0, r1 = shard_bounds(m, world, rank)
    c0, c1 = shard_bounds(n, world, rank)
    rows = sp.csr_array(A[r0:r1, :])
    rows.sort_indices()
    colt = sp.csc_array(A[:, c0:c1])          # CSC of the column block == CSR of its transpose
    colt.sort_indices()
    colt_csr = sp.csr_array((colt.data, colt.indices, colt.indptr), shape=(c1 - c0, m))
    return rows, colt_csr


# ----------------------------------------------------------------------------------------------------------
# communicator bootstrap through torch.distributed (any backend: the 128-byte NCCL id is broadcast from rank 0)
# ----------------------------------------------------------------------------------------------------------
def init_comm():
    """Create the library's NCCL communicator over the ranks of the default torch.distributed group."""
    import torch
    import torch.distributed as dist
    L = lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    check(L.propack_b200_init(), "init")
    if world == 1:
        return rank, world
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        check(L.propack_b200_comm_unique_id(buf), "comm_unique_id")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    check(L.propack_b200_comm_init(C.c_int(rank), C.c_int(world), raw), "comm_init")
    return rank, world


def finalize_comm():
    lib().propack_b200_comm_finalize()


class ShardedOperator:
    """This rank's share of a sparse matrix, resident on its GPU (``propack_b200_csr_create_sharded_*``)."""

    def __init__(self, A, rank: int, world: int, dtype=None):
        import scipy.sparse as sp
        A = sp.csr_array(A)
        dtype = np.dtype(dtype or A.dtype)
        if dtype not in PREFIX:
            dtype = np.dtype(np.complex128 if np.iscomplexobj(A) else np.float64)
        A = A.astype(dtype)
        A.sort_indices()
        self.dtype, self.pfx, self.shape = dtype, PREFIX[dtype], A.shape
        self.rank, self.world = rank, world
        self.rows = shard_bounds(A.shape[0], world, rank)
        self.cols = shard_bounds(A.shape[1], world, rank)
        rows, colt = shard_csr(A, world, rank)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self._keep = (i32(rows.indptr), i32(rows.indices), np.ascontiguousarray(rows.data),
                      i32(colt.indptr), i32(colt.indices), np.ascontiguousarray(colt.data))
        a = self._keep
        self.nnz_local = (int(rows.nnz), int(colt.nnz))
        self.handle = check(getattr(lib(), f"propack_b200_csr_create_sharded_{self.pfx}")(
            C.c_int(A.shape[0]), C.c_int(A.shape[1]), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]), C.c_int(0)),
            "csr_create_sharded")
        self._keep = None

    def bytes_per_product(self, adjoint=False) -> float:
        return float(lib().propack_b200_op_bytes(C.c_int(self.handle), C.c_int(int(adjoint))))

    def close(self):
        if self.handle:
            lib().propack_b200_op_destroy(C.c_int(self.handle))
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedDenseOperator:
    """This rank's row block of a dense matrix, resident on its GPU (``propack_b200_dense_create_sharded_*``).

    ``A`` is the full (m x n) array (every rank uploads only its own rows ``A[r0:r1]``).  ``synthetic=(seed, table)``
    evaluates the rows of the BASELINE config-3 matrix on the device instead (``A`` is then just the global shape)."""

    def __init__(self, A, rank: int, world: int, dtype=None, synthetic=None):
        L = lib()
        self.rank, self.world = rank, world
        if synthetic is not None:
            m, n = A
            seed, table = synthetic
            self.dtype, self.pfx, self.shape = np.dtype(np.float64), "d", (int(m), int(n))
            T = np.ascontiguousarray(table, dtype=np.float64)
            L.propack_b200_dense_create_synthetic_sharded_d.argtypes = [C.c_int, C.c_int, C.c_ulonglong, C.c_void_p]
            self.handle = check(L.propack_b200_dense_create_synthetic_sharded_d(int(m), int(n), int(seed), T.ctypes.data_as(C.c_void_p)),
                                "dense_create_synthetic_sharded")
        else:
            A = np.asarray(A)
            dtype = np.dtype(dtype or A.dtype)
            if dtype not in PREFIX:
                dtype = np.dtype(np.complex128 if np.iscomplexobj(A) else np.float64)
            self.dtype, self.pfx = dtype, PREFIX[dtype]
            r0, r1 = shard_bounds(A.shape[0], world, rank)
            self.shape = (int(A.shape[0]), int(A.shape[1]))
            blk = np.asfortranarray(A[r0:r1].astype(dtype, copy=False))
            if blk.shape[0] == 0:
                blk = np.zeros((1, A.shape[1]), dtype=dtype, order="F")
            fn = getattr(L, f"propack_b200_dense_create_sharded_{self.pfx}")
            fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_long]
            self.handle = check(fn(self.shape[0], self.shape[1], _p(blk), blk.shape[0]), "dense_create_sharded")
        self.rows = shard_bounds(self.shape[0], world, rank)
        self.cols = shard_bounds(self.shape[1], world, rank)

    def bytes_per_product(self, adjoint=False) -> float:
        return float(lib().propack_b200_op_bytes(C.c_int(self.handle), C.c_int(int(adjoint))))

    def close(self):
        if self.handle:
            lib().propack_b200_op_destroy(C.c_int(self.handle))
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Solver:
    """A solver session over a (sharded or single-GPU) operator: bases stay in HBM, results come back as this
    rank's row slices."""

    def __init__(self, op, ucols: int, vcols: int):
        self.op = op
        L = lib()
        self.id = check(L.propack_b200_solver_create(C.c_int(op.handle), C.c_int(ucols), C.c_int(vcols)), "solver_create")
        ml, nl, ldu, ldv = C.c_int(0), C.c_int(0), C.c_long(0), C.c_long(0)
        check(L.propack_b200_solver_local_rows(C.c_int(self.id), C.byref(ml), C.byref(nl), C.byref(ldu), C.byref(ldv)), "local_rows")
        self.m_local, self.n_local = ml.value, nl.value

    def set_start(self, u0_global=None):
        if u0_global is None:
            check(lib().propack_b200_solver_set_start(C.c_int(self.id), None), "set_start")
            return
        r0, r1 = getattr(self.op, "rows", (0, self.op.shape[0]))
        loc = np.ascontiguousarray(np.asarray(u0_global, dtype=self.op.dtype)[r0:r1])
        if loc.size == 0:
            loc = np.zeros(1, dtype=self.op.dtype)
        check(lib().propack_b200_solver_set_start(C.c_int(self.id), _p(loc)), "set_start")

    def lansvd(self, k, kmax, tol=0.0, delta=None, eta=None, anorm=0.0, cgs=True, elr=True, jobu=True, jobv=True, U_out=None,
               V_out=None):
        pfx = self.op.pfx
        R = REAL[pfx]
        sigma = np.zeros(max(k, 1), dtype=R); bnd = np.zeros(max(k, 1), dtype=R)
        dopt = _options(pfx, delta, eta, anorm)
        iopt = np.array([int(bool(cgs)), int(bool(elr))], dtype=np.int32)
        kk, info = C.c_int(k), C.c_int(0)
        check(lib().propack_b200_solver_lansvd(C.c_int(self.id), C.c_int(int(jobu)), C.c_int(int(jobv)), C.byref(kk), C.c_int(kmax),
                                               _p(sigma), _p(bnd), C.c_double(tol), _p(dopt), _p(iopt), C.byref(info)), "solver_lansvd")
        return self._result(kk.value, info.value, sigma, bnd, jobu, jobv, U_out, V_out)

    def lansvd_irl(self, which, dim, p, neig, maxiter, tol=0.0, delta=None, eta=None, anorm=0.0, min_relgap=0.002, cgs=True,
                   elr=True, jobu=True, jobv=True, U_out=None, V_out=None):
        pfx = self.op.pfx
        R = REAL[pfx]
        sigma = np.zeros(max(neig, 1), dtype=R); bnd = np.zeros(max(neig, 1), dtype=R)
        dopt = _options(pfx, delta, eta, anorm, min_relgap)
        iopt = np.array([int(bool(cgs)), int(bool(elr))], dtype=np.int32)
        d, ne, info = C.c_int(dim), C.c_int(neig), C.c_int(0)
        smallest = int(str(which).lower().startswith("s"))
        check(lib().propack_b200_solver_lansvd_irl(C.c_int(self.id), C.c_int(smallest), C.c_int(int(jobu)), C.c_int(int(jobv)),
                                                   C.byref(d), C.c_int(p), C.byref(ne), C.c_int(maxiter), _p(sigma), _p(bnd),
                                                   C.c_double(tol), _p(dopt), _p(iopt), C.byref(info)), "solver_lansvd_irl")
        return self._result(ne.value, info.value, sigma, bnd, jobu, jobv, U_out, V_out)

    def _result(self, k, info, sigma, bnd, jobu, jobv, U_out=None, V_out=None):
        """This rank's row slices of the k Ritz vectors.  `U_out` / `V_out`: caller-owned result buffers (e.g. pinned host memory,
        which the device-to-host copy reaches at PCIe speed) -- Fortran-ordered, the operator's dtype, at least
        max(m_local, 1) / max(n_local, 1) rows and k columns; otherwise fresh (pageable) arrays are returned."""
        out = {"k": k, "info": info, "sigma": sigma[:k].copy(), "bnd": bnd[:k].copy(), "U": None, "V": None}
        dt = self.op.dtype

        def buffer(given, rows, name):
            rows = max(rows, 1)
            if given is None:
                return np.zeros((rows, k), dtype=dt, order="F")
            if (given.dtype != dt or given.ndim != 2 or not given.flags.f_contiguous or given.shape[0] < rows or given.shape[1] < k
                    or not given.flags.writeable):
                raise ValueError(f"{name} must be a writeable Fortran-ordered {dt} array with >= {rows} rows and >= {k} columns")
            return given

        if jobu and k > 0:
            U = buffer(U_out, self.m_local, "U_out")
            check(lib().propack_b200_solver_get_u(C.c_int(self.id), C.c_int(k), _p(U), C.c_long(U.shape[0])), "get_u")
            out["U"] = U[:self.m_local, :k]
        if jobv and k > 0:
            V = buffer(V_out, self.n_local, "V_out")
            check(lib().propack_b200_solver_get_v(C.c_int(self.id), C.c_int(k), _p(V), C.c_long(V.shape[0])), "get_v")
            out["V"] = V[:self.n_local, :k]
        return out

    def close(self):
        if self.id:
            lib().propack_b200_solver_destroy(C.c_int(self.id))
            self.id = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_rows(local: np.ndarray, dim: int) -> np.ndarray:
    """Assemble the full (dim x k) array from every rank's row slice (torch.distributed all_gather_object)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    full = np.concatenate([p for p in parts if p is not None and p.size], axis=0)
    assert full.shape[0] == dim, (full.shape, dim)
    return full
