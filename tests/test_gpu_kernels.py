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


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,k,index", [
    (300_001, 37, [1, 37, 38]),               # long vectors: the TMA-staged GEMV^T (cp.async.bulk ring), odd row count
    (524_288, 300, [2, 150, 160, 299, 301]),  # > 256 columns: two column slices per row range; two intervals
])
def test_reorth_long_vectors_tma_path(oracle, dtype, n, k, index):
    """dreorth_ on vectors long enough for the TMA-staged kernel (L >= 262144), against the oracle."""
    from propack_b200 import f77
    rng = np.random.default_rng(k)
    V = np.asfortranarray(rand_vec(rng, n * k, dtype).reshape(n, k) / np.sqrt(n)).astype(dtype)
    v = rand_vec(rng, n, dtype)
    nrm0 = float(np.linalg.norm(v))
    got, gn = f77.reorth(V, v, nrm0, index, 0.717, 1)
    want, wn = oracle.reorth(V, v, nrm0, index, 0.717, 1)
    tol = TOL[dtype]
    assert abs(gn - wn) <= tol * wn
    assert rel(got, want) < tol


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


# ---------------------------------------------------------------------------------------------------------
# sliced jagged-ELL construction and column-panel split (integer work: bit-exact), SpMV on awkward shapes
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("shape", [(700, 3100), (2500, 900), (1024, 64)])
def test_sell_build_is_bit_exact(dtype, shape):
    """The device-built jagged-ELL copies of A and A^T equal the numpy restatement (tests/sell_ref.py) array for array."""
    from propack_b200 import f77
    from sell_ref import sell_ref
    rng = np.random.default_rng(2)
    A = ragged_matrix(rng, *shape, dtype)
    op = f77.Operator(A)
    At = sp.csr_array(A.T)
    At.sort_indices()
    for adjoint, M in ((False, A), (True, At)):
        got = f77.sell_arrays(op, adjoint)
        assert got["panels"] == 1
        want = sell_ref(M, long_thr=got["long"])
        for key in ("joff", "len8", "ci", "va"):
            assert np.array_equal(got[key], want[key]), key
    op.close()


def test_column_panels_are_bit_exact_and_products_agree(oracle):
    """Column blocking (operands whose gathered vector outgrows L2; forced here with PROPACK_B200_SPMV_COLBLOCKS): every panel's
    jagged-ELL copy equals the numpy restatement, and the panel-by-panel product equals the one-panel product's oracle value."""
    import subprocess, sys, os, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent("""
        import sys, numpy as np, scipy.sparse as sp
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from propack_b200 import f77
        from sell_ref import sell_ref, column_panels
        from oracle import oracle_py as O
        from test_gpu_kernels import ragged_matrix, rel
        from conftest import rand_vec
        rng = np.random.default_rng(6)
        for dtype in (np.float64, np.complex128):
            A = ragged_matrix(rng, 3000, 2600, dtype)
            At = sp.csr_array(A.T); At.sort_indices()
            op = f77.Operator(A)
            for adjoint, M in ((False, A), (True, At)):
                want = column_panels(M, 4)
                for g in range(4):
                    got = f77.sell_arrays(op, adjoint, panel=g)
                    assert got["panels"] == 4
                    ref = sell_ref(want[g], long_thr=got["long"])
                    for key in ("joff", "len8", "ci", "va"):
                        assert np.array_equal(got[key], ref[key]), (adjoint, g, key)
            cplx = np.iscomplexobj(A.data)
            for transa in ("n", "c" if cplx else "t"):
                x = rand_vec(rng, A.shape[1] if transa == "n" else A.shape[0], dtype)
                assert rel(f77.aprod(op, transa, x), O.csr_aprod(transa, A, x, dtype=dtype)) < 1e-12
            op.close()
        print("PANELS_OK")
    """ % (root, os.path.join(root, "tests")))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, PROPACK_B200_SPMV_COLBLOCKS="4"))
    sys.stdout.write(p.stdout[-2000:]); sys.stderr.write(p.stderr[-3000:])
    assert p.returncode == 0 and "PANELS_OK" in p.stdout


@pytest.mark.parametrize("dtype", DTYPES)
def test_aprod_uniform_rows_and_tiny_shapes(oracle, dtype):
    """Exactly-10-per-row matrices (the config 5 pattern, no padding at all), a 1 x 1 matrix and a single column."""
    from propack_b200 import f77
    rng = np.random.default_rng(3)
    m, n = 5000, 4100
    cols = np.sort(rng.integers(0, n, size=(m, 10)), axis=1)
    A = sp.csr_array((rand_vec(rng, m * 10, dtype), cols.ravel(), np.arange(0, m * 10 + 1, 10)), shape=(m, n))
    A.sum_duplicates(); A.sort_indices()
    cplx = np.iscomplexobj(np.zeros(1, dtype=dtype))
    cases = [A, sp.csr_array(np.array([[3.0]], dtype=dtype)), sp.csr_array(rand_vec(rng, 77, dtype).reshape(77, 1))]
    for M in cases:
        op = f77.Operator(M)
        for transa in ("n", "c" if cplx else "t"):
            x = rand_vec(rng, M.shape[1] if transa == "n" else M.shape[0], dtype)
            got = f77.aprod(op, transa, x)
            want = oracle.csr_aprod(transa, M, x, dtype=dtype)
            assert rel(got, want) < TOL[dtype] * 1e-2
        op.close()


def test_csr_create_rejects_malformed_input():
    """Row pointers that decrease / overshoot, unsorted rows and out-of-range columns are refused (no device faults)."""
    from propack_b200 import _lib
    L = _lib.lib()
    f = L.propack_b200_csr_create_d
    va = np.ones(4)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    bad_rp = np.array([0, 3, 2, 4], dtype=np.int32); ci = np.array([0, 1, 2, 0], dtype=np.int32)
    assert f(3, 3, p(bad_rp), p(ci), p(va), 0) < 0 and "non-decreasing" in _lib.last_error()
    rp = np.array([0, 2, 3, 4], dtype=np.int32); unsorted = np.array([1, 0, 2, 0], dtype=np.int32)
    assert f(3, 3, p(rp), p(unsorted), p(va), 0) < 0 and "sorted" in _lib.last_error()
    oob = np.array([0, 3, 2, 0], dtype=np.int32)
    assert f(3, 3, p(rp), p(oob), p(va), 0) < 0 and "out of range" in _lib.last_error()


# ---------------------------------------------------------------------------------------------------------
# blasext level-1 (pdnrm2 / pddot / pdaxpy / pdscal / pdzero and s/c/z variants), through the exported symbols
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 31, 4097, 300_001])
def test_blasext_level1_matches_oracle(oracle, dtype, n):
    from propack_b200 import f77
    rng = np.random.default_rng(n)
    x, y = rand_vec(rng, n, dtype), rand_vec(rng, n, dtype)
    # the oracle accumulates in the working precision like the reference BLAS (rounding error ~ sqrt(n) eps in single),
    # the device kernels in double: the bar against the oracle carries that term, the bar against a float64 sum does not
    tol = max(TOL[dtype] * 1e-2, 4.0 * np.sqrt(n) * np.finfo(dtype).eps)
    assert abs(f77.nrm2(x) - oracle.nrm2(x)) <= tol * oracle.nrm2(x)
    exact = np.linalg.norm(x.astype(np.complex128 if np.iscomplexobj(x) else np.float64))
    assert abs(f77.nrm2(x) - exact) <= 4 * np.finfo(dtype).eps * exact
    d_got, d_want = f77.dotc(x, y), oracle.dotc(x, y)
    assert abs(d_got - d_want) <= tol * np.linalg.norm(x) * np.linalg.norm(y)
    alpha = dtype(0.75) if not np.iscomplexobj(x) else dtype(0.75 - 0.5j)
    assert rel(f77.axpy(alpha, x, y), oracle.axpy(alpha, x, y)) < tol
    assert rel(f77.scal(-1.5, x), oracle.scal(-1.5, x)) < tol
    assert not np.any(f77.zero(x))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_blasext_level1_increments(dtype):
    """Non-unit and negative increments follow the BLAS convention (dblasext.F:154-196 forwards them to the BLAS)."""
    from propack_b200 import f77
    rng = np.random.default_rng(5)
    n = 1000
    x, y = rand_vec(rng, n, dtype), rand_vec(rng, n, dtype)
    assert abs(f77.nrm2(x, incx=3) - np.linalg.norm(x)) < 1e-12 * np.linalg.norm(x)
    assert abs(f77.dotc(x, y, incx=2, incy=-1) - np.vdot(x, y)) < 1e-10
    assert rel(f77.axpy(2.0, x, y, incx=-2, incy=3), 2.0 * x + y) < 1e-13
    assert rel(f77.scal(0.5, x, incx=2), 0.5 * x) < 1e-15


# ---------------------------------------------------------------------------------------------------------
# xRITZVEC and xGEMM_OVWR entry points
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("which", ["L", "S"])
def test_ritzvec_matches_oracle(oracle, dtype, which):
    """dritzvec_ on a genuine Lanczos factorisation (bases and bidiagonal from dlanbpro_): the same Ritz vectors as the oracle's
    dritzvec, and D returns the singular values of B."""
    from propack_b200 import f77
    rng = np.random.default_rng(11)
    A = rand_sparse(rng, 1200, 700, 0.02, dtype)
    op = f77.Operator(A)
    m, n = A.shape
    dim, k = 40, 6
    R = np.float32 if dtype in (np.float32, np.complex64) else np.float64
    U = np.zeros((m, dim + 1), dtype=dtype, order="F"); V = np.zeros((n, dim), dtype=dtype, order="F")
    B = np.zeros((dim, 2), dtype=R, order="F")
    u0 = rng.uniform(size=m).astype(dtype)
    U[:, 0] = u0
    kk, rnorm, ierr, _ = f77.lanbpro(op, 0, dim, U, V, B, float(np.linalg.norm(u0)), cgs=True)
    assert kk == dim
    gu, gv, gd = f77.ritzvec(which, U, V, B[:, 0], B[:, 1], k)
    wu, wv, wd = oracle.ritzvec(which, U, V, B[:, 0], B[:, 1], k)
    tol = TOL[dtype]
    assert np.max(np.abs(gd - wd)) < tol * wd[0]
    for i in range(k):
        for g, w in ((gu, wu), (gv, wv)):
            ph = np.vdot(w[:, i], g[:, i])
            ph = ph / abs(ph)
            assert np.linalg.norm(g[:, i] - ph * w[:, i]) < (1e-3 if R is np.float32 else 1e-8)
    op.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("transa", ["n", "t"])
def test_gemm_ovwr_matches_definition(dtype, transa):
    """dgemm_ovwr_: B <- alpha*op(A)*B + beta*B (dgemm_ovwr.F:5-53) for the (dim) x (dim+1) shapes dritzvec uses and a ragged one."""
    from propack_b200 import f77
    rng = np.random.default_rng(9)
    for (m, n, k) in [(40, 41, 40), (17, 9, 23), (1, 1, 1)]:
        A = rng.standard_normal((m, k) if transa == "n" else (k, m)).astype(dtype)
        B = rng.standard_normal((max(m, k), n)).astype(dtype)
        opA = A if transa == "n" else A.T
        for alpha, beta in ((1.0, 0.0), (-0.5, 2.0)):
            want = alpha * (opA.astype(np.float64) @ B[:k].astype(np.float64)) + beta * B[:m].astype(np.float64)
            got = f77.gemm_ovwr(transa, A, B, m, n, k, alpha, beta)
            assert rel(got, want) < TOL[dtype] * 1e-1
