"""GPU: driver-level parity -- xLANSVD / xLANSVD_IRL / xLANBPRO through the Fortran C-ABI against the
oracle, the committed golden vectors (dense LAPACK SVD + SciPy's PROPACK translation of the same inputs)
and the residual / orthogonality properties the north star names."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import DTYPES, GOLDEN, TOL, rand_sparse

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))))


def subspace_dist(X, Y):
    """max principal-angle sine between equal-dimension subspaces with orthonormal columns (sign/phase free)."""
    s = np.linalg.svd(X.conj().T @ Y, compute_uv=False)
    return float(np.sqrt(max(0.0, 1.0 - min(s) ** 2)))


def check_triplets(A, r, tol, dtype, ref=None):
    """Residuals, orthogonality.  With partial reorthogonalisation the basis is only semi-orthogonal (sqrt(eps)), so at
    tol ~ 1e-12 the true residual of the *reference algorithm* sits above tol*sigma_1; when the oracle's triplets are
    given, the bar is "no worse than 5x the oracle's own residual"."""
    U, S, V = r["U"], r["sigma"], r["V"]
    k = S.size
    eps = np.finfo(dtype).eps
    AH = A.conj().T
    bar = max(tol, 1e3 * eps) * S[0] * 10
    if ref is not None:
        bar = max(bar, 5 * np.max(np.linalg.norm(A @ ref["V"] - ref["U"] * ref["sigma"], axis=0)))
    assert np.max(np.linalg.norm(A @ V - U * S, axis=0)) < bar
    assert np.max(np.linalg.norm(AH @ U - V * S, axis=0)) < max(np.sqrt(tol), 1e3 * eps) * S[0] * 10
    assert np.max(np.abs(U.conj().T @ U - np.eye(k))) < 200 * np.sqrt(eps)
    assert np.max(np.abs(V.conj().T @ V - np.eye(k))) < 200 * np.sqrt(eps)


@pytest.mark.parametrize("cgs", [0, 1])
def test_illc1850_lansvd_k10(oracle, examples, cgs):
    """BASELINE config 1: illc1850, k=10, DLANSVD non-restarted."""
    from propack_b200 import f77
    import propack_b200
    g, A = examples["g"], examples["illc1850"]
    op = f77.Operator(A)
    propack_b200.reset_counters()
    got = f77.lansvd(op, 10, 100, tol=1e-12, u0=g["illc1850_u0"], cgs=bool(cgs))
    ctr = propack_b200.counters()
    oracle.stats_reset()
    ref = oracle.lansvd(A, 10, 100, tol=1e-12, u0=g["illc1850_u0"], cgs=bool(cgs))
    assert got["info"] == 0 and got["k"] == 10
    assert relerr(got["sigma"], ref["sigma"]) < 1e-10                     # vs oracle
    assert relerr(got["sigma"], g["illc1850_svd"][:10]) < 1e-10           # vs dense LAPACK
    assert relerr(got["sigma"], g[f"illc1850_scipy_lansvd_k10_cgs{cgs}_sigma"]) < 1e-10   # vs SciPy's PROPACK
    check_triplets(A, got, 1e-12, np.float64)
    for i in range(10):   # singular vectors up to sign
        assert min(np.linalg.norm(got["U"][:, i] - ref["U"][:, i]), np.linalg.norm(got["U"][:, i] + ref["U"][:, i])) < 1e-7
        assert min(np.linalg.norm(got["V"][:, i] - ref["V"][:, i]), np.linalg.norm(got["V"][:, i] + ref["V"][:, i])) < 1e-7
    # same Krylov trajectory as the reference algorithm: identical step and matvec counts
    st = oracle.stats()
    assert ctr["nsteps"] == st["nsteps"] and ctr["nopx"] == st["nopx"] and ctr["nbsvd"] == st["nbsvd"]
    assert ctr["launches"] > 0
    op.close()


def test_illc1850_irl_k10(oracle, examples):
    from propack_b200 import f77
    g, A = examples["g"], examples["illc1850"]
    op = f77.Operator(A)
    got = f77.lansvd_irl(op, 10, 50, p=40, tol=1e-12, u0=g["illc1850_u0"])
    ref = oracle.lansvd_irl(A, 10, 50, p=40, tol=1e-12, u0=g["illc1850_u0"])
    assert got["info"] == 0 and got["k"] == 10
    assert relerr(got["sigma"], ref["sigma"]) < 1e-10
    assert relerr(got["sigma"], g["illc1850_svd"][:10]) < 1e-10
    assert relerr(got["sigma"], g["illc1850_scipy_irl_k10_dim50_sigma"]) < 1e-10
    check_triplets(A, got, 1e-12, np.float64)
    assert subspace_dist(got["U"], ref["U"]) < 1e-5   # Ritz vectors of a semiorthogonal (sqrt(eps)) basis
    op.close()


def test_illc1850_lansvd_k200(examples):
    """The reference's own example run (README:133-138): k=200 of 712."""
    from propack_b200 import f77
    g, A = examples["g"], examples["illc1850"]
    op = f77.Operator(A)
    got = f77.lansvd(op, 200, 712, tol=0.0, u0=g["illc1850_u0"], cgs=True)
    assert got["info"] == 0 and got["k"] == 200
    assert relerr(got["sigma"], g["illc1850_svd"][:200]) < 1e-10
    assert relerr(got["sigma"], g["illc1850_scipy_lansvd_k200_sigma"]) < 1e-10
    U, S, V = got["U"], got["sigma"], got["V"]
    assert np.linalg.norm(U.T @ U - np.eye(200)) < 1e-6
    assert np.linalg.norm(V.T @ V - np.eye(200)) < 1e-6
    Ad = A.toarray()
    u, s, vt = np.linalg.svd(Ad, full_matrices=False)
    assert np.linalg.norm((U * S) @ V.T - (u[:, :200] * s[:200]) @ vt[:200]) < 1e-8   # SciPy test_examples check
    op.close()


def test_mhd1280b_zlansvd(oracle, examples):
    """Complex example matrix of the reference (README:89-118), ZLANSVD ('c' products, zreorth)."""
    from propack_b200 import f77
    g, A = examples["g"], examples["mhd1280b"]
    op = f77.Operator(A)
    got = f77.lansvd(op, 10, 200, tol=1e-12, u0=g["mhd1280b_u0"], cgs=True)
    ref = oracle.lansvd(A, 10, 200, tol=1e-12, u0=g["mhd1280b_u0"], cgs=True, dtype=np.complex128)
    assert got["info"] == 0 and got["k"] == 10
    assert relerr(got["sigma"], ref["sigma"]) < 1e-10
    assert relerr(got["sigma"], g["mhd1280b_svd"][:10]) < 1e-10
    assert relerr(got["sigma"], g["mhd1280b_scipy_lansvd_k10_sigma"]) < 1e-10
    check_triplets(A, got, 1e-12, np.complex128, ref)
    op.close()


def test_single_precision_examples(examples):
    from propack_b200 import f77
    g = examples["g"]
    op = f77.Operator(examples["illc1850"], dtype=np.float32)
    got = f77.lansvd(op, 10, 100, tol=1e-5, u0=g["illc1850_u0"].astype(np.float32))
    assert got["k"] == 10 and relerr(got["sigma"], g["illc1850_svd"][:10]) < 1e-4
    assert relerr(got["sigma"], g["illc1850_scipy_slansvd_k10_sigma"]) < 1e-4
    op.close()
    op = f77.Operator(examples["mhd1280b"], dtype=np.complex64)
    got = f77.lansvd(op, 10, 200, tol=1e-5, u0=g["mhd1280b_u0"].astype(np.complex64))
    assert got["k"] == 10 and relerr(got["sigma"], g["mhd1280b_svd"][:10]) < 1e-4
    assert relerr(got["sigma"], g["mhd1280b_scipy_clansvd_k10_sigma"]) < 1e-4
    op.close()


@pytest.mark.parametrize("name", ["s", "d", "c", "z"])
@pytest.mark.parametrize("irl", [False, True])
@pytest.mark.parametrize("fmt", ["dense", "csr"])
def test_scipy_test_svdp_cases(name, irl, fmt):
    """SciPy test_propack.py::test_svdp through the scipy-compatible svdp() (k=3, 10x20)."""
    from propack_b200 import svdp
    t = np.load(os.path.join(GOLDEN, "small_dense.npz"))
    A, want = t[f"A_{name}"], t[f"svd_{name}"][:3]
    Ain = sp.csr_array(A) if fmt == "csr" else A
    u, s, vh, _ = svdp(Ain, 3, which="LM", irl_mode=irl, kmax=11 if irl else None, full_output=True, rng=np.random.default_rng(0))
    tol = TOL[A.dtype.type]
    assert relerr(s, want) < tol
    assert np.allclose(np.abs(u.conj().T @ u), np.eye(3), atol=100 * np.sqrt(np.finfo(A.dtype).eps))
    assert np.linalg.norm(A @ vh.conj().T - u * s) < max(tol, 1e-5 if name in "sc" else 1e-9) * s[0] * 10


def test_smallest_triplets_irl(oracle):
    """which='S' (dlansvd_irl.F:262-275, 318-332) on a well-conditioned matrix."""
    from propack_b200 import f77
    rng = np.random.default_rng(4)
    Q1, _ = np.linalg.qr(rng.standard_normal((120, 60)))
    Q2, _ = np.linalg.qr(rng.standard_normal((60, 60)))
    s = np.linspace(1.0, 4.0, 60)
    A = (Q1 * s) @ Q2.T
    op = f77.Operator(A)
    u0 = rng.uniform(size=120)
    got = f77.lansvd_irl(op, 3, 30, p=10, which="S", tol=1e-10, u0=u0, maxiter=300)
    ref = oracle.lansvd_irl(A, 3, 30, p=10, which="S", tol=1e-10, u0=u0, maxiter=300)
    assert got["k"] == ref["k"] == 3
    assert relerr(np.sort(got["sigma"]), np.sort(s)[:3]) < 1e-8
    assert relerr(np.sort(got["sigma"]), np.sort(ref["sigma"])) < 1e-8
    op.close()


@pytest.mark.parametrize("dim,p", [(30, 20), (24, 18), (40, 10)])
def test_smallest_triplets_irl_many_shifts(dim, p):
    """which='S' with p > dim/2 -- the case the reference documents as broken (Changelog:56, dlansvd_irl.F:318-332: only dim-p
    shifts are filled for p sweeps) -- and p < dim/2, against dense LAPACK SVD (the oracle reproduces the reference's defect)."""
    from propack_b200 import f77
    rng = np.random.default_rng(4)
    Q1, _ = np.linalg.qr(rng.standard_normal((120, 60)))
    Q2, _ = np.linalg.qr(rng.standard_normal((60, 60)))
    s = np.linspace(1.0, 4.0, 60)
    A = (Q1 * s) @ Q2.T
    op = f77.Operator(A)
    got = f77.lansvd_irl(op, 3, dim, p=p, which="S", tol=1e-10, u0=rng.uniform(size=120), maxiter=2000)
    op.close()
    assert got["info"] == 0 and got["k"] == 3
    assert relerr(np.sort(got["sigma"]), np.sort(s)[:3]) < 1e-8
    U, S, V = got["U"], got["sigma"], got["V"]
    assert np.max(np.linalg.norm(A @ V - U * S, axis=0)) < 1e-7


def test_thin_hilbert_and_fat_random(oracle):
    """SciPy test_thin_hilbert (200x4, k=4: j == min(m,n), dbdqr ignorelast) and test_fat_random (3x100, k=3)."""
    from propack_b200 import f77
    i, j = np.meshgrid(np.arange(200), np.arange(4), indexing="ij")
    A = 1.0 / (i + j + 1)
    op = f77.Operator(A)
    got = f77.lansvd(op, 4, 5, u0=np.random.default_rng(0).uniform(size=200))
    assert got["k"] == 4 and relerr(got["sigma"], np.linalg.svd(A, compute_uv=False)) < 1e-8
    op.close()
    B = np.random.default_rng(0).standard_normal((3, 100))
    op = f77.Operator(B)
    got = f77.lansvd(op, 3, 4, u0=np.random.default_rng(1).uniform(size=3))
    assert got["k"] == 3 and relerr(got["sigma"], np.linalg.svd(B, compute_uv=False)) < 1e-10
    op.close()


def test_zero_start_vector_uses_lapack_rng(oracle, examples):
    """U(:,1) = 0 => dgetu0 start vector from dlarnv(2,(1,3,5,7)) (dlansvd.F:165-169): same run as the oracle."""
    from propack_b200 import f77
    A = examples["illc1850"]
    op = f77.Operator(A)
    got = f77.lansvd(op, 6, 80, tol=1e-12, u0=None, cgs=True)
    ref = oracle.lansvd(A, 6, 80, tol=1e-12, u0=None, cgs=True)
    assert got["k"] == ref["k"] == 6
    assert relerr(got["sigma"], ref["sigma"]) < 1e-10
    op.close()


def test_generic_callback_aprod(oracle, examples):
    """A user APROD (LinearOperator-like, host callback with the dlansvd.F:20-33 contract) gives the same answer
    as the device-resident operator."""
    from propack_b200 import f77
    from scipy.sparse.linalg import aslinearoperator
    g, A = examples["g"], examples["illc1850"]
    lop = aslinearoperator(A)
    op = f77.Operator(lop)
    got = f77.lansvd(op, 5, 60, tol=1e-12, u0=g["illc1850_u0"], cgs=True)
    assert got["k"] == 5 and relerr(got["sigma"], g["illc1850_svd"][:5]) < 1e-10


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_lanbpro_factorisation_and_extension(oracle, dtype):
    """dlanbpro: A V_k = U_{k+1} B_k, then extend k0 -> k (dlanbpro.F:15-16,231-275)."""
    from propack_b200 import f77
    rng = np.random.default_rng(7)
    A = rand_sparse(rng, 900, 500, 0.02, dtype)
    op = f77.Operator(A)
    m, n = A.shape
    k = 30
    U = np.zeros((m, k + 1), dtype=dtype, order="F")
    V = np.zeros((n, k), dtype=dtype, order="F")
    B = np.zeros((k, 2), dtype=np.float64, order="F")
    u0 = rng.uniform(size=m).astype(dtype)
    U[:, 0] = u0
    k1, rnorm, ierr, anorm = f77.lanbpro(op, 0, 12, U, V, B, float(np.linalg.norm(u0)))
    assert k1 == 12 and ierr >= 0
    k2, rnorm, ierr, anorm = f77.lanbpro(op, 12, k, U, V, B, rnorm, anorm=anorm)
    assert k2 == k
    Bm = np.zeros((k + 1, k))
    Bm[np.arange(k), np.arange(k)] = B[:, 0]
    Bm[np.arange(1, k + 1), np.arange(k)] = B[:, 1]
    assert np.linalg.norm(A @ V - U @ Bm) < 1e-12 * np.linalg.norm(Bm)
    assert np.max(np.abs(U.conj().T @ U - np.eye(k + 1))) < 1e-7
    assert np.max(np.abs(V.conj().T @ V - np.eye(k))) < 1e-7
    assert abs(rnorm - B[k - 1, 1]) == 0
    # the same bidiagonal as the oracle's dlanbpro on the same inputs (alpha = B(:,1), beta = B(:,2)), both legs
    Uo = np.zeros_like(U); Vo = np.zeros_like(V); Bo = np.zeros_like(B)
    Uo[:, 0] = u0
    ko, rn_o, ierr_o, an_o = oracle.lanbpro(A, 0, 12, Uo, Vo, Bo, float(np.linalg.norm(u0)), dtype=dtype)
    ko, rn_o, ierr_o, an_o = oracle.lanbpro(A, 12, k, Uo, Vo, Bo, rn_o, anorm=an_o, dtype=dtype)
    assert ko == k
    assert np.max(np.abs(B - Bo) / np.abs(Bo)) < 1e-10
    assert abs(rn_o - rnorm) < 1e-10 * rnorm and abs(an_o - anorm) < 1e-10 * anorm
    op.close()


def test_complex_irl_with_restarts(oracle):
    """Complex IRL after >= 1 restart agrees with dense SVD (the reference applies P^T/Q^T here: SURVEY 2.3)."""
    from propack_b200 import f77
    import propack_b200
    rng = np.random.default_rng(3)
    A = (rng.standard_normal((300, 200)) + 1j * rng.standard_normal((300, 200))) @ np.diag(np.linspace(1, 5, 200) ** 2)
    op = f77.Operator(A)
    propack_b200.reset_counters()
    got = f77.lansvd_irl(op, 5, 12, p=6, u0=rng.uniform(size=300) + 0j, tol=1e-12)
    assert propack_b200.counters()["nrestart"] > 0
    assert got["k"] == 5
    assert relerr(got["sigma"], np.linalg.svd(A, compute_uv=False)[:5]) < 1e-10
    op.close()


def test_medium_synthetic_matches_oracle(oracle):
    """BASELINE config 2 recipe at 1/10 scale (100k x 100k, ~10 nnz/row, k=20): sigma vs the oracle, residuals."""
    from propack_b200 import f77
    rng = np.random.default_rng(0)
    A = sp.random_array((100_000, 100_000), density=1e-4, format="csr", rng=rng, data_sampler=rng.standard_normal)
    A.sort_indices()
    u0 = np.random.default_rng(1).uniform(size=A.shape[0])
    op = f77.Operator(A)
    got = f77.lansvd(op, 20, 400, tol=1e-10, u0=u0, cgs=True)
    ref = oracle.lansvd(A, 20, 400, tol=1e-10, u0=u0, cgs=True)
    assert got["info"] == 0 and got["k"] == ref["k"] == 20
    assert relerr(got["sigma"], ref["sigma"]) < 1e-10
    check_triplets(A, got, 1e-10, np.float64)
    assert subspace_dist(got["V"], ref["V"]) < 1e-5
    op.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_powerlaw_rows_lansvd(oracle, dtype):
    """BASELINE config 4 in miniature: power-law row lengths (rows of 600..2400 non-zeros next to empty and short rows), so
    the SpMV's long-row kernel, the run packing around long rows and the fused norm all take part in a full solve."""
    from propack_b200 import f77
    rng = np.random.default_rng(4)
    m, n = 3000, 2600
    lens = np.minimum(rng.zipf(2.0, size=m) * 3, 2400)
    lens[:4] = [2400, 700, 0, 129]
    lens[m // 2] = 600
    rows = np.repeat(np.arange(m), lens)
    cols = rng.integers(0, n, size=rows.size)
    vals = rng.standard_normal(rows.size)
    if np.issubdtype(dtype, np.complexfloating):
        vals = vals + 1j * rng.standard_normal(rows.size)
    A = sp.csr_array(sp.coo_array((vals, (rows, cols)), shape=(m, n))).astype(dtype)
    A.sum_duplicates(); A.sort_indices()
    u0 = rng.uniform(size=m).astype(dtype)
    op = f77.Operator(A)
    k = 6
    got = f77.lansvd(op, k, 150, tol=1e-10 if dtype != np.float32 else 1e-5, u0=u0, cgs=True)
    ref = oracle.lansvd(A, k, 150, tol=1e-10 if dtype != np.float32 else 1e-5, u0=u0, cgs=True, dtype=dtype)
    sd = np.linalg.svd(A.toarray().astype(np.complex128 if np.iscomplexobj(A.data) else np.float64), compute_uv=False)[:k]
    tol = 1e-10 if dtype != np.float32 else 1e-4
    assert got["info"] == 0 and got["k"] == k
    assert relerr(got["sigma"], sd) < tol
    assert relerr(got["sigma"], ref["sigma"]) < tol
    res = np.max(np.linalg.norm(A @ got["V"] - got["U"] * got["sigma"], axis=0))
    assert res < (1e-8 if dtype != np.float32 else 1e-2) * got["sigma"][0]
    op.close()


@pytest.mark.gpu
def test_dense_synthetic_generator_bit_exact_and_irl(oracle):
    """BASELINE config 3 in miniature: the on-device dense generator equals its numpy replica bit for bit (unit-vector
    products return exact columns / rows), and DLANSVD_IRL on it matches dense LAPACK SVD and the oracle."""
    from propack_b200 import f77, synth
    m, n, seed = 3001, 200, 7
    T = synth.planted_table(synth.planted_coefficients(m, n))
    from oracle import synth_ref
    A = synth_ref.dense_planted(m, n, seed, T)
    op = synth.device_dense_planted(m, n, seed, T)
    for j in (0, 1, 63, 64, 199):
        e = np.zeros(n); e[j] = 1.0
        assert np.array_equal(f77.aprod(op, "n", e), A[:, j])
    for i in (0, 255, 256, 3000):
        e = np.zeros(m); e[i] = 1.0
        assert np.array_equal(f77.aprod(op, "t", e), A[i, :])
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n)
    assert rel_vec(f77.aprod(op, "n", x), A @ x) < 1e-13
    u0 = rng.uniform(size=m)
    k = 10
    got = f77.lansvd_irl(op, k, 40, p=20, maxiter=100, tol=1e-10, u0=u0, cgs=True)
    ref = oracle.lansvd_irl(A, k, 40, p=20, maxiter=100, tol=1e-10, u0=u0, cgs=True)
    sd = np.linalg.svd(A, compute_uv=False)[:k]
    assert got["info"] == 0 and got["k"] == k
    assert relerr(got["sigma"], sd) < 1e-10 and relerr(got["sigma"], ref["sigma"]) < 1e-10
    assert np.max(np.linalg.norm(A @ got["V"] - got["U"] * got["sigma"], axis=0)) < 1e-8 * got["sigma"][0]
    op.close()


def rel_vec(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


@pytest.mark.gpu
def test_c2_full_size_parity_and_properties(oracle):
    """BASELINE configs[1] at its full size (1M x 1M, 1e7 non-zeros, k=50, kmax=600): sigma against the CPU oracle on the same
    inputs to 1e-10 relative, and the size-independent properties -- residuals, orthogonality, ordering."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from propack_b200 import f77
    A, u0, k, kmax, tol = bench.make_matrix("c2")
    op = f77.Operator(A)
    got = f77.lansvd(op, k, kmax, tol=tol, u0=u0, cgs=True)
    op.close()
    assert got["info"] == 0 and got["k"] == k
    S, U, V = got["sigma"], got["U"], got["V"]
    assert np.all(np.diff(S) <= 0) and S[0] < 10 and S[-1] > 7          # sqrt(10)*(1+1) ~ 6.3 bulk edge + finite-size tail
    res = np.linalg.norm(A @ V - U * S, axis=0)
    resT = np.linalg.norm(A.T.tocsr() @ U - V * S, axis=0)
    assert res.max() < 1e-8 * S[0] and resT.max() < 1e-6 * S[0]
    assert np.max(np.abs(U.T @ U - np.eye(k))) < 1e-6 and np.max(np.abs(V.T @ V - np.eye(k))) < 1e-6
    ref = oracle.lansvd(A, k, kmax, tol=tol, u0=u0, cgs=True, jobu=False, jobv=False)
    assert ref["k"] == k and relerr(S, ref["sigma"]) < 1e-10


@pytest.mark.gpu
def test_example_driver_on_illc1850_rra():
    """BASELINE configs[0]: the reference's example program (example.F on illc1850.rra, k = 10, DLANSVD) end to end:
    Harwell-Boeing file in, singular values compared with the stored reference values as `compare` does."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "examples", "example.py"), os.path.join(GOLDEN, "illc1850.rra"), "--k", "10",
                        "--kmax", "100", "--compare", os.path.join(GOLDEN, "Sigma_illc1850.ascii")], capture_output=True, text=True, timeout=600)
    sys.stdout.write(p.stdout[-3000:]); sys.stderr.write(p.stderr[-2000:])
    assert p.returncode == 0 and "max relative error of sigma" in p.stdout


@pytest.mark.gpu
def test_example_programs_all_formats(tmp_path):
    """The rest of the reference's Examples layer (README:89-121): example_irl on illc1850.rra, the complex example on
    mhd1280b.cua, and example.py on coordinate / binary-dense / diagonal files, each checked against stored singular values."""
    import importlib.util
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_example_data", os.path.join(root, "examples", "make_example_data.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    data = mod.main(os.path.join(tmp_path, "data"))
    ex, ex_irl = os.path.join(root, "examples", "example.py"), os.path.join(root, "examples", "example_irl.py")
    sig_real = os.path.join(GOLDEN, "Sigma_illc1850.ascii")
    runs = [
        [ex_irl, os.path.join(GOLDEN, "illc1850.rra"), "--k", "10", "--kmax", "50", "--p", "40", "--compare", sig_real],
        [ex, os.path.join(data, "mhd1280b.cua"), "--k", "10", "--kmax", "200", "--compare", os.path.join(data, "Sigma_mhd1280b.ascii")],
        [ex_irl, os.path.join(data, "mhd1280b.cua"), "--k", "6", "--kmax", "40", "--p", "20", "--compare", os.path.join(data, "Sigma_mhd1280b.ascii")],
        [ex, os.path.join(data, "illc1850.coord"), "--k", "10", "--kmax", "100", "--compare", sig_real],
        [ex, os.path.join(data, "illc1850.cbin"), "--k", "10", "--kmax", "100", "--precision", "single", "--tol", "1e-6", "--compare", sig_real],
        [ex, os.path.join(data, "illc1850.bin"), "--k", "10", "--kmax", "100", "--compare", sig_real],
        [ex, os.path.join(data, "band4000.diag"), "--k", "5", "--kmax", "400"],
    ]
    for cmd in runs:
        p = subprocess.run([sys.executable] + cmd, capture_output=True, text=True, timeout=600)
        sys.stdout.write(p.stdout[-1500:]); sys.stderr.write(p.stderr[-1500:])
        assert p.returncode == 0, cmd
        if "--compare" in cmd:
            assert "max relative error of sigma" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_svdp_accepts_torch_device_arrays(dtype):
    """svdp / svds on arrays that already live on the GPU (torch CUDA tensors, dense and sparse CSR, and DLPack exporters): the
    operator is built by device-to-device copies, and the answer matches dense LAPACK SVD."""
    import torch
    from propack_b200 import svdp, svds
    rng = np.random.default_rng(2)
    A = rng.standard_normal((301, 77)) @ np.diag(np.linspace(1, 3, 77))
    if np.issubdtype(dtype, np.complexfloating):
        A = A + 1j * rng.standard_normal(A.shape)
    A = A.astype(dtype)
    want = np.linalg.svd(A.astype(np.complex128 if np.iscomplexobj(A) else np.float64), compute_uv=False)
    tol = TOL[dtype]
    At = torch.from_numpy(A).cuda()
    u, s, vh, _ = svdp(At, 5, irl_mode=False, kmax=60, rng=np.random.default_rng(0))
    assert relerr(s, want[:5]) < tol
    assert np.linalg.norm(A @ vh.conj().T - u * s) < (1e-4 if dtype == np.float32 else 1e-9) * s[0] * 10
    Asp = sp.csr_array(np.where(np.abs(A) > 1.0, A, 0))
    wsp = np.linalg.svd(Asp.toarray().astype(np.complex128 if np.iscomplexobj(A) else np.float64), compute_uv=False)
    Tsp = torch.sparse_csr_tensor(torch.from_numpy(Asp.indptr.astype(np.int64)), torch.from_numpy(Asp.indices.astype(np.int64)),
                                  torch.from_numpy(Asp.data), size=Asp.shape).cuda()
    u, s, vh, _ = svdp(Tsp, 4, irl_mode=True, kmax=40, rng=np.random.default_rng(0))
    assert relerr(s, wsp[:4]) < tol

    class Exporter:                      # anything with __dlpack__ (cupy / jax / ... arrays)
        def __init__(self, t): self.t = t
        def __dlpack__(self, **kw): return self.t.__dlpack__(**kw)
        def __dlpack_device__(self): return self.t.__dlpack_device__()
    uu, ss, vv = svds(Exporter(At), k=3, maxiter=40, rng=np.random.default_rng(0))
    assert relerr(ss[::-1], want[:3]) < tol and np.all(np.diff(ss) >= 0)      # scipy's svds returns ascending order
    assert svds(At, k=2, return_singular_vectors=False, maxiter=30, rng=np.random.default_rng(0)).shape == (2,)
