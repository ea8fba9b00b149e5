"""CPU: pin the oracle (oracle/propack_oracle.hpp) against the committed golden vectors.

Sources (tools/make_golden.py): dense LAPACK SVD of the two PROPACK example matrices, SciPy's C
translation of PROPACK run on fixed start vectors, LAPACK xLARNV known answers.  Tolerances are the
reference's own: "of the order 1e-15" for double / complex16 and "1e-6" for single / complex8
(reference README:152-157), relaxed to the BASELINE parity bars 1e-10 / 1e-4.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, TOL


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))))


@pytest.mark.parametrize("name", ["d", "s", "z", "c"])
def test_larnv_known_answer(oracle, name):
    """dlarnv/zlarnv idist=2, iseed=(1,3,5,7): bit-exact stream and final seed (dgetu0.F:41-44,69)."""
    kat = np.load(os.path.join(GOLDEN, "larnv_kat.npz"))
    want = kat[f"{name}larnv_x"]
    x, seed = oracle.larnv(want.size, dtype=want.dtype)
    assert np.array_equal(x, want)
    assert np.array_equal(seed, kat[f"{name}larnv_seed_after"])


def test_larnv_survey_vector(oracle):
    x, seed = oracle.larnv(5)
    assert x.tolist() == [0.3957424639187579, 0.0008649603975001696, -0.9227205789982591, -0.9165671495278005,
                          0.1175963848841306]
    assert seed.tolist() == [2288, 3429, 3993, 2627]


@pytest.mark.parametrize("cgs", [0, 1])
def test_illc1850_lansvd_k10(oracle, examples, cgs):
    g, A = examples["g"], examples["illc1850"]
    r = oracle.lansvd(A, 10, 100, tol=1e-12, u0=g["illc1850_u0"], cgs=bool(cgs))
    assert r["info"] == 0 and r["k"] == 10
    assert relerr(r["sigma"], g["illc1850_svd"][:10]) < 1e-13           # dense LAPACK
    assert relerr(r["sigma"], g[f"illc1850_scipy_lansvd_k10_cgs{cgs}_sigma"]) < 1e-13  # SciPy's PROPACK
    U, S, V = r["U"], r["sigma"], r["V"]
    assert np.max(np.linalg.norm(A @ V - U * S, axis=0)) < 1e-11
    # the left residual carries the (gap-refined) Lanczos bound: |sigma error| ~ residual^2/gap (dbsvd.F:207-230)
    assert np.max(np.linalg.norm(A.T @ U - V * S, axis=0)) < 1e-7
    assert np.max(np.abs(U.T @ U - np.eye(10))) < 1e-10


def test_illc1850_irl_k10(oracle, examples):
    g, A = examples["g"], examples["illc1850"]
    r = oracle.lansvd_irl(A, 10, 50, p=40, tol=1e-12, u0=g["illc1850_u0"])
    assert r["info"] == 0 and r["k"] == 10
    assert relerr(r["sigma"], g["illc1850_svd"][:10]) < 1e-13
    assert relerr(r["sigma"], g["illc1850_scipy_irl_k10_dim50_sigma"]) < 1e-13
    assert oracle.stats()["nrestart"] >= 0


def test_illc1850_lansvd_k200(oracle, examples):
    """The reference's own example size: k=200 (README:133-138), sigma error 'of the order 1e-15'."""
    g, A = examples["g"], examples["illc1850"]
    r = oracle.lansvd(A, 200, 712, tol=0.0, u0=g["illc1850_u0"])
    assert r["info"] == 0 and r["k"] == 200
    assert relerr(r["sigma"], g["illc1850_svd"][:200]) < 1e-12
    assert relerr(r["sigma"], g["illc1850_scipy_lansvd_k200_sigma"]) < 1e-12


def test_mhd1280b_zlansvd(oracle, examples):
    g, A = examples["g"], examples["mhd1280b"]
    r = oracle.lansvd(A, 10, 200, tol=1e-12, u0=g["mhd1280b_u0"], dtype=np.complex128)
    assert r["info"] == 0 and r["k"] == 10
    assert relerr(r["sigma"], g["mhd1280b_svd"][:10]) < 1e-12
    assert relerr(r["sigma"], g["mhd1280b_scipy_lansvd_k10_sigma"]) < 1e-12
    U, S, V = r["U"], r["sigma"], r["V"]
    # semiorthogonal basis (delta = sqrt(eps)): residuals sit at ~delta*eps^(1/2)*sigma_1, far below tol*sigma
    assert np.max(np.linalg.norm(A @ V - U * S, axis=0)) < 1e-9 * S[0]
    assert np.max(np.linalg.norm(A.conj().T @ U - V * S, axis=0)) < 1e-6


def test_single_precision_examples(oracle, examples):
    g = examples["g"]
    r = oracle.lansvd(examples["illc1850"], 10, 100, tol=1e-5, u0=g["illc1850_u0"], dtype=np.float32)
    assert r["k"] == 10 and relerr(r["sigma"], g["illc1850_svd"][:10]) < 1e-4
    assert relerr(r["sigma"], g["illc1850_scipy_slansvd_k10_sigma"]) < 1e-4
    r = oracle.lansvd(examples["mhd1280b"], 10, 200, tol=1e-5, u0=g["mhd1280b_u0"], dtype=np.complex64)
    assert r["k"] == 10 and relerr(r["sigma"], g["mhd1280b_svd"][:10]) < 1e-4
    assert relerr(r["sigma"], g["mhd1280b_scipy_clansvd_k10_sigma"]) < 1e-4


@pytest.mark.parametrize("name", ["s", "d", "c", "z"])
@pytest.mark.parametrize("irl", [False, True])
def test_scipy_test_svdp_matrices(oracle, name, irl):
    """The 10x20 cases of SciPy's test_propack.py::test_svdp (k=3), sigma vs dense SVD."""
    t = np.load(os.path.join(GOLDEN, "small_dense.npz"))
    A, want = t[f"A_{name}"], t[f"svd_{name}"][:3]
    u0 = np.random.default_rng(1).uniform(size=10).astype(A.dtype)
    if irl:
        r = oracle.lansvd_irl(A, 3, 11, p=8, u0=u0, dtype=A.dtype)
    else:
        r = oracle.lansvd(A, 3, 11, u0=u0, dtype=A.dtype)
    assert r["k"] == 3
    assert relerr(r["sigma"], want) < TOL[A.dtype.type]


def test_complex_irl_with_restarts_is_correct(oracle):
    """the reference (and SciPy's port) mis-apply P/Q in the complex restart; the oracle
    implements the intended product, so complex IRL must agree with dense SVD even after restarts."""
    rng = np.random.default_rng(3)
    A = (rng.standard_normal((300, 200)) + 1j * rng.standard_normal((300, 200)))
    A = A @ np.diag(np.linspace(1, 5, 200) ** 2)
    u0 = rng.uniform(size=300) + 0j
    r = oracle.lansvd_irl(A, 5, 12, p=6, u0=u0, dtype=np.complex128, tol=1e-12)
    assert oracle.stats()["nrestart"] > 0
    assert r["k"] == 5
    assert relerr(r["sigma"], np.linalg.svd(A, compute_uv=False)[:5]) < 1e-10


def test_thin_hilbert_invariant_subspace(oracle):
    """SciPy test_thin_hilbert: 200x4 Hilbert, k=4 = full rank (j == min(m,n) path, dbdqr ignorelast)."""
    i, j = np.meshgrid(np.arange(200), np.arange(4), indexing="ij")
    A = 1.0 / (i + j + 1)
    r = oracle.lansvd(A, 4, 5, u0=np.random.default_rng(0).uniform(size=200))
    assert r["k"] == 4
    assert relerr(r["sigma"], np.linalg.svd(A, compute_uv=False)) < 1e-8
