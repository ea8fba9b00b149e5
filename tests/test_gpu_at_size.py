"""GPU: parity at (or near) the sizes BASELINE.json states, against the CPU oracle on the same seeded inputs.

configs[1] (C2) at full size lives in test_gpu_drivers.py; here: configs[3] (C4, complex16 power-law CSR 2M x 2M, ZLANSVD) at
full size, the configs[4] pattern (C5: exactly 10 non-zeros per row, k=100, DLANSVD_IRL dim=300 p=200 -- many restarts, the
101-column restart GEMM) at 1M rows, and a 100k-row replica of configs[2] (C3: dense x 4096 columns, DLANSVD_IRL dim=200).
Bars (BASELINE.json north_star): sigma within 1e-10 relative of the reference algorithm; residuals below tol.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))))


def test_c4_full_size_complex_powerlaw_parity(oracle):
    """BASELINE configs[3]: complex16 power-law CSR 2M x 2M (~2e7 nnz, rows up to ~15k long), k=64, kmax=700, ZLANSVD."""
    import bench
    from propack_b200 import f77
    A, u0, k, kmax, tol = bench.make_matrix("c4")
    assert A.shape == (2_000_000, 2_000_000) and A.dtype == np.complex128
    op = f77.Operator(A)
    got = f77.lansvd(op, k, kmax, tol=tol, u0=u0, cgs=True)
    op.close()
    assert got["info"] == 0 and got["k"] == k
    S, U, V = got["sigma"], got["U"], got["V"]
    assert np.all(np.diff(S) <= 0)
    res = np.linalg.norm(A @ V - U * S, axis=0)
    assert res.max() < 1e-8 * S[0]
    assert np.max(np.abs(U.conj().T @ U - np.eye(k))) < 1e-6 and np.max(np.abs(V.conj().T @ V - np.eye(k))) < 1e-6
    ref = oracle.lansvd(A, k, kmax, tol=tol, u0=u0, cgs=True, jobu=False, jobv=False, dtype=np.complex128)
    assert ref["k"] == k and relerr(S, ref["sigma"]) < 1e-10


def test_c5_pattern_1m_rows_irl_parity(oracle):
    """BASELINE configs[4] pattern at 1M x 1M: exactly 10 columns per row, k=100, DLANSVD_IRL dim=300 p=200 (the driver and shapes
    of the 10M-row target: restarts, the 101-column restart GEMM, Ritz vectors from dim=300)."""
    import bench
    import propack_b200
    from propack_b200 import f77
    A, u0, k, dim, tol = bench.make_matrix("c5-small")
    assert A.shape == (1_000_000, 1_000_000) and int(np.diff(A.indptr).max()) == 10
    p = bench.IRL_P["c5-small"]
    op = f77.Operator(A)
    propack_b200.reset_counters()
    got = f77.lansvd_irl(op, k, dim, p=p, maxiter=bench.IRL_MAXITER, tol=tol, u0=u0, cgs=True)
    ctr = propack_b200.counters()
    op.close()
    assert got["info"] == 0 and got["k"] == k and ctr["nrestart"] >= 2
    S, U, V = got["sigma"], got["U"], got["V"]
    res = np.linalg.norm(A @ V - U * S, axis=0)
    assert res.max() < 1e-8 * S[0]
    assert np.max(np.abs(U.T @ U - np.eye(k))) < 1e-6 and np.max(np.abs(V.T @ V - np.eye(k))) < 1e-6
    # the oracle's result for exactly this case is a committed fixture (tools/make_golden_atsize.py; it also holds SciPy's
    # _svdp result on the same inputs); without the file the oracle is run here (~1 min of CPU)
    gpath = os.path.join(ROOT, "tests", "golden", "c5_small_irl.npz")
    if os.path.exists(gpath):
        g = np.load(gpath)
        assert int(g["nnz"]) == A.nnz and int(g["checksum_indices"]) == int(A.indices.astype(np.int64).sum())
        ref = {"k": int(g["sigma_oracle"].size), "sigma": g["sigma_oracle"]}
        st = {"nsteps": int(g["nsteps"]), "nrestart": int(g["nrestart"])}
        if "sigma_scipy_svdp" in g:
            assert relerr(S, g["sigma_scipy_svdp"]) < 1e-10          # SciPy's PROPACK translation, same parameters
    else:
        oracle.stats_reset()
        ref = oracle.lansvd_irl(A, k, dim, p=p, which="L", maxiter=bench.IRL_MAXITER, tol=tol, u0=u0, cgs=True, jobu=False, jobv=False)
        st = oracle.stats()
    assert ref["k"] == k and relerr(S, ref["sigma"]) < 1e-10
    assert ctr["nrestart"] == st["nrestart"] and ctr["nsteps"] == st["nsteps"]      # the same restart trajectory


def test_c3_replica_100k_rows_dense_irl_parity(oracle):
    """BASELINE configs[2] replica: the first 100k rows of the 2M x 4096 synthetic dense operator (bit-identical on both sides),
    k=100, DLANSVD_IRL dim=200 p=100."""
    from oracle import synth_ref
    from propack_b200 import f77, synth
    import bench
    m_full, n, rows = 2_000_000, 4096, 100_000
    T = synth.planted_table(synth.planted_coefficients(m_full, n))
    A = synth_ref.dense_planted(m_full, n, bench.DENSE_SEED, T, rows=np.arange(rows))
    op = synth.device_dense_planted(rows, n, bench.DENSE_SEED, T)
    e = np.zeros(n); e[17] = 1.0
    assert np.array_equal(f77.aprod(op, "n", e), A[:, 17])          # the two operators are the same matrix
    u0 = np.random.default_rng(1).uniform(size=rows)
    k, dim, p = 100, 200, 100
    got = f77.lansvd_irl(op, k, dim, p=p, maxiter=50, tol=1e-10, u0=u0, cgs=True)
    op.close()
    assert got["info"] == 0 and got["k"] == k
    S, U, V = got["sigma"], got["U"], got["V"]
    assert np.max(np.linalg.norm(A @ V - U * S, axis=0)) < 1e-8 * S[0]
    ref = oracle.lansvd_irl(A, k, dim, p=p, which="L", maxiter=50, tol=1e-10, u0=u0, cgs=True, jobu=False, jobv=False)
    assert ref["k"] == k and relerr(S, ref["sigma"]) < 1e-10
