"""CPU: host-side logic of the row-sharded multi-GPU path (propack_b200/dist.py) -- partition arithmetic, shard
extraction (bit-exact vs scipy slicing) and, under torch.distributed `gloo` with world_size 2, that "all-gather the
input slices, then a purely local product" reproduces A x and A^H u bit for bit."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from propack_b200 import dist as pdist  # noqa: E402


@pytest.mark.parametrize("dim", [1, 31, 32, 33, 712, 1850, 4096, 1_000_000, 10_000_019])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_bounds_partition(dim, world):
    s = pdist.slice_len(dim, world)
    assert s % 32 == 0 and s * world >= dim
    prev = 0
    for r in range(world):
        lo, hi = pdist.shard_bounds(dim, world, r)
        assert lo == prev and lo <= hi <= dim and hi - lo <= s
        assert lo == min(r * s, dim)            # slices are contiguous at multiples of s: global index == gathered index
        prev = hi
    assert prev == dim


def test_shard_bounds_match_c_library():
    from propack_b200 import _lib
    L = _lib.lib()
    L.propack_b200_shard_slice.restype = C.c_long
    L.propack_b200_shard_slice.argtypes = [C.c_long, C.c_int]
    L.propack_b200_shard_bounds.argtypes = [C.c_long, C.c_int, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    for dim in (1, 40, 712, 1850, 999_999, 10_000_000):
        for world in (1, 2, 4, 8):
            assert L.propack_b200_shard_slice(dim, world) == pdist.slice_len(dim, world)
            for r in range(world):
                lo, hi = C.c_long(0), C.c_long(0)
                L.propack_b200_shard_bounds(dim, world, r, C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == pdist.shard_bounds(dim, world, r)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_shard_csr_bit_exact(dtype, world):
    rng = np.random.default_rng(3)
    A = sp.random_array((301, 77), density=0.05, format="csr", rng=rng, data_sampler=rng.standard_normal).astype(dtype)
    if dtype is np.complex128:
        A = sp.csr_array(A + 1j * sp.random_array(A.shape, density=0.05, format="csr", rng=rng, data_sampler=rng.standard_normal))
    A.sort_indices()
    nnz_r = nnz_c = 0
    for r in range(world):
        rows, colt = pdist.shard_csr(A, world, r)
        r0, r1 = pdist.shard_bounds(A.shape[0], world, r)
        c0, c1 = pdist.shard_bounds(A.shape[1], world, r)
        assert rows.shape == (r1 - r0, A.shape[1]) and colt.shape == (c1 - c0, A.shape[0])
        assert rows.has_sorted_indices and colt.has_sorted_indices
        assert np.array_equal(rows.toarray(), A.toarray()[r0:r1])                 # values and positions bit-exact
        assert np.array_equal(colt.toarray(), A.toarray()[:, c0:c1].T)
        # CSR of the row block is literally the corresponding slice of A's arrays
        p0, p1 = A.indptr[r0], A.indptr[r1]
        assert np.array_equal(rows.indices, A.indices[p0:p1]) and np.array_equal(rows.data, A.data[p0:p1])
        assert np.array_equal(rows.indptr, A.indptr[r0:r1 + 1] - p0)
        nnz_r += rows.nnz; nnz_c += colt.nnz
    assert nnz_r == A.nnz and nnz_c == A.nnz


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        m, n = 1850, 712   # illc1850's shape: slices 928 / 384, the last rank's are shorter
        A = sp.random_array((m, n), density=0.01, format="csr", rng=rng, data_sampler=rng.standard_normal)
        A.sort_indices()
        x = rng.standard_normal(n); u = rng.standard_normal(m)
        rows, colt = pdist.shard_csr(A, world, rank)
        sm, sn = pdist.slice_len(m, world), pdist.slice_len(n, world)
        r0, r1 = pdist.shard_bounds(m, world, rank); c0, c1 = pdist.shard_bounds(n, world, rank)

        def allgather(local, s):   # what ncclAllGather does with equal padded slices
            pad = np.zeros(s); pad[:local.size] = local
            parts = [torch.zeros(s, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(pad))
            return torch.cat(parts).numpy()

        xf = allgather(x[c0:c1], sn)                    # gathered index == global index
        assert np.array_equal(xf[:n], x)
        y_loc = rows @ xf[:n]                           # local SpMV over global column ids
        uf = allgather(u[r0:r1], sm)
        assert np.array_equal(uf[:m], u)
        v_loc = colt @ uf[:m]
        y = pdist.gather_rows(y_loc, m); v = pdist.gather_rows(v_loc, n)
        ok = np.array_equal(y, A @ x) and np.array_equal(v, A.T.tocsr() @ u)
        # cross-rank reduction of norm partials, as the library does (sum of squares all-reduced, then sqrt)
        part = torch.tensor([float(np.dot(y_loc, y_loc))], dtype=torch.float64)
        dist.all_reduce(part)
        ok = ok and abs(float(part.sqrt()) - np.linalg.norm(A @ x)) < 1e-12 * np.linalg.norm(A @ x)
        # dense row-sharded operator (ShardedDenseOperator): A x from the gathered n-vector and the local row block;
        # A^T u as local GEMV^T + all-reduce of the n coefficients, every rank keeping its slice
        D = rng.standard_normal((m, 24)); xd = rng.standard_normal(24)
        dn = pdist.slice_len(24, world); d0, d1 = pdist.shard_bounds(24, world, rank)
        xdf = allgather(xd[d0:d1], dn)
        yd = pdist.gather_rows(D[r0:r1] @ xdf[:24], m)
        ok = ok and np.allclose(yd, D @ xd, rtol=0, atol=1e-12)
        coef = torch.from_numpy(D[r0:r1].T @ u[r0:r1])
        dist.all_reduce(coef)
        vd = pdist.gather_rows(coef.numpy()[d0:d1], 24)
        ok = ok and np.allclose(vd, D.T @ u, rtol=0, atol=1e-10)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharded_products_match_global():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_solver_result_buffers(monkeypatch):
    """dist.Solver._result: fresh arrays by default; caller-owned (e.g. pinned) Fortran-ordered buffers are filled in place, with their
    own leading dimension, and unsuitable ones are refused (no device needed: the two copy-back entry points are faked)."""
    calls = []

    class FakeLib:
        def _fill(self, which, sid, ncols, ptr, ld):
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(ld.value * ncols.value,))
            rows = 5 if which == "u" else 3
            for j in range(ncols.value):
                a[j * ld.value: j * ld.value + rows] = (100 if which == "u" else 200) + 10 * j + np.arange(rows)
            calls.append((which, ncols.value, ld.value))
            return 0

        def propack_b200_solver_get_u(self, sid, ncols, ptr, ld):
            return self._fill("u", sid, ncols, ptr, ld)

        def propack_b200_solver_get_v(self, sid, ncols, ptr, ld):
            return self._fill("v", sid, ncols, ptr, ld)

    monkeypatch.setattr(pdist, "lib", lambda: FakeLib())
    sv = pdist.Solver.__new__(pdist.Solver)
    sv.id, sv.m_local, sv.n_local = 0, 5, 3           # id 0: close() / __del__ do nothing
    sv.op = type("Op", (), {"dtype": np.dtype(np.float64)})()
    sigma = np.array([3.0, 2.0, 1.0]); bnd = np.zeros(3)
    r = sv._result(2, 0, sigma, bnd, True, True)
    assert r["U"].shape == (5, 2) and r["V"].shape == (3, 2) and r["U"][4, 1] == 114.0 and r["V"][2, 1] == 212.0
    assert calls == [("u", 2, 5), ("v", 2, 3)]
    # caller-owned buffers: more rows and columns than needed, Fortran order (a transposed C array, like a pinned torch tensor)
    Ub = np.zeros((4, 8)).T; Vb = np.zeros((2, 6)).T
    assert Ub.flags.f_contiguous and Ub.shape == (8, 4)
    calls.clear()
    r = sv._result(2, 0, sigma, bnd, True, True, U_out=Ub, V_out=Vb)
    assert calls == [("u", 2, 8), ("v", 2, 6)]
    assert r["U"].shape == (5, 2) and np.shares_memory(r["U"], Ub) and r["U"][0, 0] == 100.0 and r["U"][4, 1] == 114.0
    assert r["V"].shape == (3, 2) and np.shares_memory(r["V"], Vb) and r["V"][2, 1] == 212.0
    assert np.array_equal(r["sigma"], [3.0, 2.0])
    for bad in (np.zeros((8, 4)), np.zeros((4, 8), dtype=np.float32).T, np.zeros((1, 8)).T, np.zeros((4, 4)).T):
        with pytest.raises(ValueError):
            sv._result(2, 0, sigma, bnd, True, False, U_out=bad)
