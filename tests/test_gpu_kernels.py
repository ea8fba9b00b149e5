"""GPU: kernel-level parity of the CUDA path against the CPU oracle, called through the Fortran C-ABI
(dreorth_, dgemm_ovwr_left_, dgetu0_, dsafescal_, propack_b200_aprod_*_ ...) on the same seeded inputs.

Bars (BASELINE.json): integer work bit-exact; floating point within 1e-10 (double / complex16) or
1e-4 (single / complex8) relative.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import DTYPES, TOL, rand_sparse, rand_vec

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


# ---------------------------------------------------------------------------------------------------------
# APROD: CSR SpMV, both directions, ragged rows (empty rows, warp rows, CTA rows)
# ---------------------------------------------------------------------------------------------------------
def ragged_matrix(rng, m, n, dtype):
    """Rows of length 0, 1..40, a few of ~300 (warp-per-row bin) and two > 2048 (CTA-per-row bin)."""
    lens = rng.integers(0, 41, size=m)
    lens[rng.integers(0, m, size=m // 10)] = 0
    lens[:6] = [300, 257, 513, 2500, 3000, 0]
    lens = np.minimum(lens, n)
    rows = np.repeat(np.arange(m), lens)
    cols = np.concatenate([np.sort(rng.choice(n, size=l, replace=False)) for l in lens]) if lens.sum() else np.zeros(0, int)
    data = rand_vec(rng, rows.size, dtype)
    A = sp.csr_array((data, (rows, cols)), shape=(m, n))
    A.sort_indices()
    return A


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(700, 3100), (4000, 3500)])
def test_aprod_matches_oracle(oracle, dtype, shape):
    from propack_b200 import f77
    rng = np.random.default_rng(0)
    A = ragged_matrix(rng, *shape, dtype)
    op = f77.Operator(A)
    cplx = np.iscomplexobj(np.zeros(1, dtype=dtype))
    for transa in ("n", "c" if cplx else "t"):
        x = rand_vec(rng, A.shape[1] if transa == "n" else A.shape[0], dtype)
        got = f77.aprod(op, transa, x)
        want = oracle.csr_aprod(transa, A, x, dtype=dtype)
        assert rel(got, want) < TOL[dtype] * 1e-2
    op.close()


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_csr_transpose_is_bit_exact(dtype):
    """Integer work: the device-side CSR of A^T equals scipy's sorted tocsc() arrays exactly."""
    from propack_b200 import f77, _lib
    rng = np.random.default_rng(1)
    A = ragged_matrix(rng, 900, 1300, dtype)
    op = f77.Operator(A)
    At = sp.csr_array(A.T)
    At.sort_indices()
    trp = np.zeros(A.shape[1] + 1, dtype=np.int32)
    tci = np.zeros(A.nnz, dtype=np.int32)
    tva = np.zeros(A.nnz, dtype=dtype)
    rc = _lib.lib().propack_b200_csr_get_transpose(C.c_int(op.handle), trp.ctypes.data_as(C.c_void_p), tci.ctypes.data_as(C.c_void_p),
                                                    tva.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert np.array_equal(trp, At.indptr.astype(np.int32))
    assert np.array_equal(tci, At.indices.astype(np.int32))
    assert np.array_equal(tva, At.data)
    op.close()


def test_aprod_empty_matrix_rows_and_single_entry():
    from propack_b200 import f77
    A = sp.csr_array(([2.5], ([3], [1])), shape=(5, 4))
    op = f77.Operator(A)
    assert np.array_equal(f77.aprod(op, "n", np.arange(1.0, 5.0)), [0, 0, 0, 5.0, 0])
    assert np.array_equal(f77.aprod(op, "t", np.arange(1.0, 6.0)), [0, 10.0, 0, 0])
    op.close()


# ---------------------------------------------------------------------------------------------------------
# dreorth: iterated Gram-Schmidt as GEMV pairs
# ---------------------------------------------------------------------------------------------------------
def semi_orthonormal_basis(rng, n, k, dtype):
    Q, _ = np.linalg.qr(rand_vec(rng, n * k, dtype).reshape(n, k).astype(np.complex128 if np.iscomplexobj(np.zeros(1, dtype)) else np.float64))
    return np.asfortranarray(Q.astype(dtype))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,k,index", [
    (1000, 7, [1, 7, 8]),                    # single interval, odd sizes (remainder columns, masked tail pack)
    (5003, 40, [1, 40, 41]),                 # full interval
    (5003, 40, [3, 9, 15, 15, 20, 38, 41]),  # three intervals incl. a single column
    (70001, 300, [1, 300, 301]),             # > GT_CHUNK columns: two column slices
    (333, 12, [13]),                         # empty list: only the norm is recomputed
])
@pytest.mark.parametrize("iflag", [1, 0])
def test_reorth_matches_oracle(oracle, dtype, n, k, index, iflag):
    from propack_b200 import f77
    rng = np.random.default_rng(k)
    V = semi_orthonormal_basis(rng, n, k, dtype)
    v = rand_vec(rng, n, dtype)
    v = (v + V @ rand_vec(rng, k, dtype) * 3).astype(dtype)   # large components inside span(V)
    nrm0 = float(np.linalg.norm(v))
    got, gn = f77.reorth(V, v, nrm0, index, 0.717, iflag)
    want, wn = oracle.reorth(V, v, nrm0, index, 0.717, iflag)
    tol = TOL[dtype]
    assert abs(gn - wn) <= tol * wn
    assert rel(got, want) < tol
    # the defining property: orthogonal to the selected columns to working precision
    sel = np.concatenate([np.arange(index[i] - 1, index[i + 1]) for i in range(0, len(index) - 1, 2)]) if len(index) > 1 else np.zeros(0, int)
    if sel.size:
        eps = np.finfo(dtype).eps
        assert np.max(np.abs(V[:, sel].conj().T @ got)) < 50 * eps * max(gn, 1.0) * np.sqrt(n)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_reorth_vector_in_span_is_zeroed(oracle, dtype):
    """dreorth.F:85-98: a vector that keeps failing the DGKS test for NTRY passes is zeroed.  With V = unit
    vectors the projection is exact, so every pass leaves exactly 0 (0 > 0.717*0 is false) on both sides."""
    from propack_b200 import f77
    rng = np.random.default_rng(5)
    V = np.asfortranarray(np.eye(2000, 10, dtype=dtype))
    v = (V @ rand_vec(rng, 10, dtype)).astype(dtype)
    got, gn = f77.reorth(V, v, float(np.linalg.norm(v)), [1, 10, 11], 0.717, 1)
    want, wn = oracle.reorth(V, v, float(np.linalg.norm(v)), [1, 10, 11], 0.717, 1)
    assert gn == 0.0 and wn == 0.0
    assert not got.any() and not want.any()


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_reorth_vector_in_span_leaves_rounding_noise(oracle, dtype):
    """A vector in span(V) of a generic basis: pass 1 removes it down to rounding noise, pass 2 finds that noise
    already orthogonal and accepts it (dreorth.F:94) -- the result is O(eps*||v||) on both sides, not zeroed."""
    from propack_b200 import f77
    rng = np.random.default_rng(5)
    V = semi_orthonormal_basis(rng, 2000, 10, dtype)
    v = (V @ rand_vec(rng, 10, dtype)).astype(dtype)
    nv = float(np.linalg.norm(v))
    got, gn = f77.reorth(V, v, nv, [1, 10, 11], 0.717, 1)
    want, wn = oracle.reorth(V, v, nv, [1, 10, 11], 0.717, 1)
    assert gn <= 1e-13 * nv and wn <= 1e-13 * nv
    assert abs(np.linalg.norm(got) - gn) <= 1e-3 * max(gn, 1e-300)


# ---------------------------------------------------------------------------------------------------------
# dgemm_ovwr_left: tall in-place GEMM on the FP64 tensor pipe
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,k,n,transb", [
    (1000, 21, 10, "t"),     # Ritz vectors (dritzvec.F:160): op(B) = B^T
    (1000, 21, 11, "n"),     # IRL restart (dlansvd_irl.F:387)
    (4099, 201, 101, "t"),   # BASELINE config 5 shape, odd row count
    (3001, 150, 150, "n"),   # N > 128: slab path through the scratch panel
    (17, 5, 3, "n"),
])
def test_gemm_ovwr_left_matches_oracle(oracle, dtype, m, k, n, transb):
    from propack_b200 import f77
    rng = np.random.default_rng(m)
    A = np.asfortranarray(rand_vec(rng, m * k, dtype).reshape(m, k))
    R = np.float32 if dtype in (np.float32, np.complex64) else np.float64
    B = rng.standard_normal((n, k) if transb == "t" else (k, n)).astype(R)
    got = f77.gemm_ovwr_left(transb, A, B, n, k)
    want = oracle.gemm_ovwr_left(transb, A, B, n, k)
    assert rel(got[:, :n], want[:, :n]) < TOL[dtype] * 1e-1
    ref = A.astype(np.complex128) @ (B.T if transb == "t" else B).astype(np.float64)
    assert rel(got[:, :n], ref) < TOL[dtype] * 1e-1
    if n < k:
        assert np.array_equal(got[:, n:], A[:, n:])  # columns beyond n are untouched


# ---------------------------------------------------------------------------------------------------------
# dgetu0 / dsafescal
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("transa", ["n", "t"])
def test_getu0_start_vector(oracle, dtype, transa):
    """u0 = op(A) r with r the LAPACK xLARNV(2,(1,3,5,7)) stream (bit-exact on the device), j = 0 and j = 5."""
    from propack_b200 import f77
    import ctypes
    rng = np.random.default_rng(2)
    A = rand_sparse(rng, 400, 300, 0.03, dtype)
    cplx = np.iscomplexobj(np.zeros(1, dtype=dtype))
    t = transa if (transa == "n" or not cplx) else "c"
    op = f77.Operator(A)
    r, _ = oracle.larnv(A.shape[1] if t == "n" else A.shape[0], dtype=dtype)
    want0 = oracle.csr_aprod(t, A, r, dtype=dtype)
    u0, u0norm, anormest, ierr = f77.getu0(op, t, 0, 1, None)
    assert ierr == 0
    assert rel(u0, want0) < TOL[dtype] * 1e-2
    assert abs(u0norm - np.linalg.norm(want0)) < TOL[dtype] * u0norm
    assert abs(anormest - np.linalg.norm(want0) / np.linalg.norm(r)) < TOL[dtype] * anormest
    rows = A.shape[0] if t == "n" else A.shape[1]
    Ub = semi_orthonormal_basis(rng, rows, 5, dtype)
    u5, n5, _, ierr = f77.getu0(op, t, 5, 3, Ub, icgs=1)
    assert ierr == 0 and n5 > 0
    assert np.max(np.abs(Ub.conj().T @ u5)) < 100 * np.finfo(dtype).eps * np.linalg.norm(want0)
    assert abs(np.linalg.norm(u5) - n5) < TOL[dtype] * n5
    op.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_safescal(oracle, dtype):
    from propack_b200 import f77
    rng = np.random.default_rng(0)
    x = rand_vec(rng, 1001, dtype)
    assert rel(f77.safescal(x, 3.7), x / dtype(3.7)) < 4 * np.finfo(dtype).eps
    if dtype in (np.float64, np.complex128):   # |alpha| < sfmin: dlascl branch (dsafescal.F:49-53)
        tiny = 1e-310
        xs = x * 1e-300
        y = f77.safescal(xs, tiny)
        want = xs.real / tiny + 1j * (xs.imag / tiny) if np.iscomplexobj(xs) else xs / tiny  # numpy's complex '/' overflows here
        assert rel(y, want) < 1e-12
