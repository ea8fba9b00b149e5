"""CPU: the fast route for the per-iteration Ritz values / error bounds of the non-restarted driver
(host::ritz_leading: dqds + inverse iteration on the Golub-Kahan tridiagonal) against the reference route
(dbdqr + dbdsqr, dlansvd.F:193-209) on real Lanczos bidiagonals produced by the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def lanczos_bidiagonal(oracle):
    """alpha, beta of 420 steps of DLANBPRO (oracle) on a 20000 x 20000 random sparse matrix; also rnorm = beta_j."""
    rng = np.random.default_rng(0)
    m = n = 20000
    A = sp.random_array((m, n), density=5e-4, format="csr", rng=rng, data_sampler=rng.standard_normal)
    A.sort_indices()
    steps = 420
    L = oracle.lib()
    op = oracle.Operator(A, np.float64)
    U = np.zeros((m, steps + 1), order="F"); V = np.zeros((n, steps), order="F")
    U[:, 0] = np.random.default_rng(1).uniform(size=m)
    B = np.zeros((steps, 2), order="F")
    eps = np.finfo(np.float64).eps
    dopt = np.array([np.sqrt(eps), eps ** 0.75, 0.0]); iopt = np.array([1, 1], dtype=np.int32)
    kk, rn, ierr = C.c_int(steps), C.c_double(float(np.linalg.norm(U[:, 0]))), C.c_int(0)
    L.oracle_lanbpro_d(C.c_int(m), C.c_int(n), C.c_int(0), C.byref(kk), *op.args(), _p(U), C.c_long(m), _p(V), C.c_long(n), _p(B),
                       C.c_int(steps), C.byref(rn), _p(dopt), _p(iopt), C.byref(ierr))
    assert kk.value == steps
    return B[:, 0].copy(), B[:, 1].copy()


def _bounds(j, alpha, beta, K, method):
    from propack_b200 import _lib
    L = _lib.lib()
    th, last = np.zeros(K), np.zeros(K)
    rc = L.propack_b200_host_ritz_bounds_d(C.c_int(j), _p(np.ascontiguousarray(alpha[:j])), _p(np.ascontiguousarray(beta[:j])), C.c_int(K),
                                           C.c_int(method), _p(th), _p(last))
    return rc, th, last


@pytest.mark.parametrize("j", [128, 151, 226, 326, 420])
def test_fast_route_matches_reference_route(lanczos_bidiagonal, j):
    alpha, beta = lanczos_bidiagonal
    K = 54
    rc0, th0, b0 = _bounds(j, alpha, beta, K, 0)
    rc1, th1, b1 = _bounds(j, alpha, beta, K, 1)
    assert rc0 == 0 and rc1 == 0
    assert np.max(np.abs(th1 - th0) / th0) < 1e-13                       # Ritz values
    # last components of the left singular vectors: both routes are normwise accurate (errors ~ eps / gap); what the
    # driver compares them with is tol * theta / rnorm >= 16 eps, so agreement to 1e-12 absolute decides identically
    assert np.max(np.abs(b1 - b0)) < 1e-12
    big = b0 > 1e-9
    assert np.max(np.abs(b1[big] - b0[big]) / b0[big]) < 1e-6
    # same convergence decisions as the driver takes (dlansvd.F:207-236) for a range of tolerances
    rnorm = beta[j - 1]
    for tol in (1e-6, 1e-10, 1e-13):
        conv = lambda th, b: int(np.argmin(np.append(np.abs(rnorm * b) <= tol * th, False)))
        assert conv(th0, b0) == conv(th1, b1)


def test_fast_route_declines_on_reducible_and_repeated_input():
    from propack_b200 import _lib
    j, K = 140, 20
    rng = np.random.default_rng(0)
    alpha, beta = rng.uniform(1, 2, j), rng.uniform(1, 2, j)
    beta[70] = 0.0                                   # B splits: the Golub-Kahan matrix is reducible
    assert _bounds(j, alpha, beta, K, 1)[0] == 1
    alpha[:] = 1.0; beta[:] = 1e-16                  # negligible coupling
    assert _bounds(j, alpha, beta, K, 1)[0] == 1
    assert _bounds(j, alpha, beta, K, 0)[0] == 0     # the reference route always answers


def _ritz_w(dim, alpha, beta, k, method):
    from propack_b200 import _lib
    L = _lib.lib()
    WU = np.zeros((dim + 1, k), order="F"); WV = np.zeros((dim, k), order="F")
    rc = L.propack_b200_host_ritz_vectors_d(C.c_int(dim), _p(np.ascontiguousarray(alpha[:dim])), _p(np.ascontiguousarray(beta[:dim])),
                                            C.c_int(k), C.c_int(method), _p(WU), _p(WV))
    return rc, WU, WV


@pytest.mark.parametrize("dim,k", [(128, 20), (226, 50), (420, 50), (420, 100)])
def test_fast_ritz_vector_matrices_are_singular_vectors_of_B(lanczos_bidiagonal, dim, k):
    """The fast route's WU, WV are the k leading singular vector pairs of the (dim+1) x dim lower bidiagonal: orthonormal,
    B v = sigma u, B^T u = sigma v to working precision, and equal to the reference route's (dbdqr + dbdsdc) up to sign."""
    alpha, beta = lanczos_bidiagonal
    B = np.zeros((dim + 1, dim))
    B[np.arange(dim), np.arange(dim)] = alpha[:dim]
    B[np.arange(1, dim + 1), np.arange(dim)] = beta[:dim]
    s = np.linalg.svd(B, compute_uv=False)[:k]
    rc0, U0, V0 = _ritz_w(dim, alpha, beta, k, 0)
    rc1, U1, V1 = _ritz_w(dim, alpha, beta, k, 1)
    assert rc0 == 0 and rc1 == 0
    for U, V in ((U0, V0), (U1, V1)):
        assert np.max(np.abs(U.T @ U - np.eye(k))) < 5e-13 and np.max(np.abs(V.T @ V - np.eye(k))) < 5e-13
        assert np.max(np.abs(B @ V - U * s)) < 1e-12 * s[0]
        assert np.max(np.abs(B.T @ U - V * s)) < 1e-12 * s[0]
    # same vectors up to a common sign per pair (the leading values of this spectrum are distinct)
    su = np.sign(np.sum(U0 * U1, axis=0))
    assert np.all(su != 0)
    assert np.max(np.abs(U1 * su - U0)) < 1e-9 and np.max(np.abs(V1 * su - V0)) < 1e-9


def test_restart_sweeps_row_parallel_is_bit_identical():
    """The implicit-restart rotation accumulation (dlansvd_irl.F:350-363): the recorded-rotation, row-parallel route of the
    driver equals the sequential dbsvdstep accumulation bit for bit, for 1 and several host threads, and P^T B Q is the
    updated bidiagonal."""
    import ctypes as C
    from propack_b200 import _lib
    L = _lib.lib()
    fn = L.propack_b200_host_restart_sweeps_d
    fn.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int]
    rng = np.random.default_rng(7)
    for dim, k in ((12, 5), (60, 20), (300, 100)):
        a0 = rng.uniform(0.5, 2.0, dim + 1); b0 = rng.uniform(0.1, 1.0, dim + 1)
        B = np.zeros((dim + 1, dim)); B[np.arange(dim), np.arange(dim)] = a0[:dim]; B[np.arange(1, dim + 1), np.arange(dim)] = b0[:dim]
        shift = np.sort(np.linalg.svd(B, compute_uv=False))[:dim - k][::-1].copy()      # p exact shifts, largest first like the driver
        shift = np.concatenate([shift, np.zeros(k + 1)])
        out = {}
        for nt in (0, 1, 5):
            a, b = a0.copy(), b0.copy()
            P = np.zeros((dim + 1, dim + 1), order="F"); Q = np.zeros((dim, dim), order="F")
            assert fn(dim, k, shift.ctypes.data, a.ctypes.data, b.ctypes.data, P.ctypes.data, Q.ctypes.data, nt) == 0
            out[nt] = (a, b, P, Q)
        for nt in (1, 5):
            for x, y in zip(out[0], out[nt]):
                assert np.array_equal(x, y)
        a, b, P, Q = out[0]
        assert np.max(np.abs(P.T @ P - np.eye(dim + 1))) < 1e-13 and np.max(np.abs(Q.T @ Q - np.eye(dim))) < 1e-13
        Bp = P.T @ B @ Q
        assert np.max(np.abs(np.diag(Bp)[:k] - a[:k])) < 1e-12 * np.abs(a0).max() * dim
        assert np.max(np.abs(np.diag(Bp, -1)[:k - 1] - b[:k - 1])) < 1e-12 * np.abs(a0).max() * dim
