"""Checkpoint / resume of the Lanczos factorisation (propack_b200/checkpoint.py): file round trip on the CPU, resume on the GPU.
(Named to be collected last: it builds on entry points the earlier files test.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from propack_b200 import checkpoint  # noqa: E402


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_save_load_round_trip(tmp_path, dtype):
    rng = np.random.default_rng(0)
    m, n, k = 40, 30, 7
    U = np.asfortranarray(rng.standard_normal((m, 12)).astype(dtype))
    V = np.asfortranarray(rng.standard_normal((n, 11)).astype(dtype))
    B = np.asfortranarray(rng.uniform(1, 2, size=(11, 2)))
    p = str(tmp_path / "chk.npz")
    checkpoint.save(p, U, V, B, k, rnorm=1.25, anorm=3.5, meta={"workload": "unit"})
    st = checkpoint.load(p, kmax=20)
    assert st["k"] == k and st["rnorm"] == 1.25 and st["anorm"] == 3.5 and "unit" in st["meta"]
    assert st["U"].shape == (m, 21) and st["V"].shape == (n, 20) and st["B"].shape == (20, 2)
    assert st["U"].flags.f_contiguous and st["V"].flags.f_contiguous and st["B"].flags.f_contiguous
    assert st["U"].dtype == U.dtype and np.array_equal(st["U"][:, :k + 1], U[:, :k + 1]) and not np.any(st["U"][:, k + 1:])
    assert np.array_equal(st["V"][:, :k], V[:, :k]) and not np.any(st["V"][:, k:])
    assert np.array_equal(st["B"][:k], B[:k]) and not np.any(st["B"][k:])
    with pytest.raises(ValueError):
        checkpoint.load(p, kmax=k - 1)
    with pytest.raises(ValueError):
        checkpoint.save(p, U, V, B, 12, rnorm=1.0)          # U has only 12 columns: k+1 = 13 needed


@pytest.mark.gpu
def test_resume_continues_the_factorisation(tmp_path, oracle):
    """Stop after 12 steps, write the checkpoint, reload it and extend to 30 steps: the result is the bidiagonal the oracle's
    dlanbpro produces with the same split (dlanbpro.F:231-275), and A V = U B holds over all 30 steps."""
    import scipy.sparse as sp
    from propack_b200 import f77
    rng = np.random.default_rng(7)
    A = sp.random_array((900, 500), density=0.02, format="csr", rng=rng, data_sampler=rng.standard_normal)
    A.sort_indices()
    m, n = A.shape
    k1, k = 12, 30
    op = f77.Operator(A)
    U = np.zeros((m, k1 + 1), order="F"); V = np.zeros((n, k1), order="F"); B = np.zeros((k1, 2), order="F")
    u0 = rng.uniform(size=m)
    U[:, 0] = u0
    kd, rnorm, ierr, anorm = f77.lanbpro(op, 0, k1, U, V, B, float(np.linalg.norm(u0)))
    assert kd == k1 and ierr >= 0
    p = str(tmp_path / "lanczos.npz")
    checkpoint.save(p, U, V, B, kd, rnorm, anorm)
    st = checkpoint.load(p, kmax=k)
    kd2, rnorm2, ierr2, anorm2 = checkpoint.resume(op, st, k)
    op.close()
    assert kd2 == k and ierr2 >= 0 and st["k"] == k
    Bm = np.zeros((k + 1, k))
    Bm[np.arange(k), np.arange(k)] = st["B"][:, 0]
    Bm[np.arange(1, k + 1), np.arange(k)] = st["B"][:, 1]
    assert np.linalg.norm(A @ st["V"] - st["U"] @ Bm) < 1e-12 * np.linalg.norm(Bm)
    Uo = np.zeros((m, k + 1), order="F"); Vo = np.zeros((n, k), order="F"); Bo = np.zeros((k, 2), order="F")
    Uo[:, 0] = u0
    ko, rn_o, ierr_o, an_o = oracle.lanbpro(A, 0, k1, Uo, Vo, Bo, float(np.linalg.norm(u0)), dtype=np.float64)
    ko, rn_o, ierr_o, an_o = oracle.lanbpro(A, k1, k, Uo, Vo, Bo, rn_o, anorm=an_o, dtype=np.float64)
    assert ko == k and np.max(np.abs(st["B"] - Bo) / np.abs(Bo)) < 1e-10
