"""CPU: the C-ABI library loads and exports every symbol include/propack_b200.h declares; host-only
entry points (the O(k^2) bidiagonal algebra the reference keeps on the CPU) match the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "propack_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b([a-z][a-z0-9_]*_)\s*\(", text))            # Fortran-ABI names end in '_'
    names |= set(re.findall(r"\b(propack_b200_[a-z0-9_]+)\s*\(", text))
    names |= {"timing_"}
    return sorted(n for n in names if not n.startswith("pb200_aprod"))


def test_library_exports_every_declared_symbol():
    from propack_b200 import _lib
    L = _lib.lib()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in include/propack_b200.h but not exported: {missing}"
    assert len(declared_symbols()) >= 80


def test_reference_entry_points_present():
    """Symbols a program linked against the reference library resolves."""
    from propack_b200 import _lib
    L = _lib.lib()
    for p in "sdcz":
        for name in ("lansvd_", "lansvd_irl_", "lanbpro_", "reorth_", "getu0_", "safescal_"):
            assert hasattr(L, p + name)
    for name in ("dgemm_ovwr_left_", "sgemm_ovwr_left_", "zdgemm_ovwr_left_", "csgemm_ovwr_left_", "dbdqr_", "dbsvdstep_",
                 "drefinebounds_", "clearstat_", "printstat_", "timing_"):
        assert hasattr(L, name)


def test_timing_common_layout():
    """COMMON /timing/ is 27 four-byte words in the order of double/stat.h:12-15."""
    from propack_b200 import _lib
    L = _lib.lib()
    L.clearstat_()
    blk = (C.c_int * 27).in_dll(L, "timing_")
    assert list(blk) == [0] * 27


@pytest.mark.skipif(__import__("conftest")._have_gpu(), reason="checks the no-device failure mode")
def test_compute_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the driver reports info <= -100 and an error string."""
    from propack_b200 import f77, _lib
    import scipy.sparse as sp
    A = sp.random_array((50, 40), density=0.2, format="csr", rng=np.random.default_rng(0))
    with pytest.raises(RuntimeError, match="propack_b200"):
        f77.Operator(A)
    assert "CUDA" in _lib.last_error() or "device" in _lib.last_error()


# ---- host algebra of the product vs the oracle ------------------------------------------------------

def _cd(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("seed", range(6))
def test_compute_int_and_set_mu_match_oracle(oracle, seed):
    from propack_b200 import _lib
    L, O = _lib.lib(), oracle.lib()
    rng = np.random.default_rng(seed)
    j = int(rng.integers(1, 60))
    mu = 10.0 ** rng.uniform(-13, -6, size=j) * rng.choice([-1, 1], size=j)
    delta, eta = 1.5e-8, 1.8e-12
    a = np.zeros(2 * j + 8, dtype=np.int32)
    b = np.zeros(2 * j + 8, dtype=np.int32)
    L.dcompute_int_(_cd(mu), C.byref(C.c_int(j)), C.byref(C.c_double(delta)), C.byref(C.c_double(eta)), _cd(a))
    O.oracle_compute_int_d(_cd(mu), C.c_int(j), C.c_double(delta), C.c_double(eta), _cd(b))
    n = int(np.argmax(b > j)) + 1
    assert np.array_equal(a[:n], b[:n])
    m1, m2 = mu.copy(), mu.copy()
    L.dset_mu_(C.byref(C.c_int(j)), _cd(m1), _cd(a), C.byref(C.c_double(1e-16)))
    O.oracle_set_mu_d(C.c_int(j), _cd(m2), _cd(b), C.c_double(1e-16))
    assert np.array_equal(m1, m2)


def test_compute_int_hand_case(oracle):
    """|mu| = [small, BIG, mid, small, mid, BIG] -> runs of |mu|>=eta containing a |mu|>delta."""
    from propack_b200 import _lib
    L = _lib.lib()
    delta, eta = 1e-3, 1e-6
    mu = np.array([1e-9, 1e-2, 1e-5, 1e-9, 1e-5, 1e-2])
    idx = np.zeros(16, dtype=np.int32)
    L.dcompute_int_(_cd(mu), C.byref(C.c_int(6)), C.byref(C.c_double(delta)), C.byref(C.c_double(eta)), _cd(idx))
    assert idx[:5].tolist() == [2, 3, 5, 6, 7]


@pytest.mark.parametrize("j", [1, 2, 3, 17])
def test_update_mu_nu_match_oracle(oracle, j):
    from propack_b200 import _lib
    L, O = _lib.lib(), oracle.lib()
    rng = np.random.default_rng(j)
    alpha, beta = rng.uniform(0.5, 2, size=j + 2), rng.uniform(0.5, 2, size=j + 2)
    mu, nu = rng.standard_normal(j + 2) * 1e-10, rng.standard_normal(j + 2) * 1e-10
    for fn in ("update_mu", "update_nu"):
        m1, n1, m2, n2 = mu.copy(), nu.copy(), mu.copy(), nu.copy()
        x1, x2 = C.c_double(0), C.c_double(0)
        args = lambda x, m_, n_: (C.byref(x), _cd(m_), _cd(n_))
        if fn == "update_mu":
            L.dupdate_mu_(C.byref(x1), _cd(m1), _cd(n1), C.byref(C.c_int(j)), _cd(alpha), _cd(beta), C.byref(C.c_double(2.0)), C.byref(C.c_double(1e-14)))
            O.oracle_update_mu_d(C.byref(x2), _cd(m2), _cd(n2), C.c_int(j), _cd(alpha), _cd(beta), C.c_double(2.0), C.c_double(1e-14))
        else:
            L.dupdate_nu_(C.byref(x1), _cd(m1), _cd(n1), C.byref(C.c_int(j)), _cd(alpha), _cd(beta), C.byref(C.c_double(2.0)), C.byref(C.c_double(1e-14)))
            O.oracle_update_nu_d(C.byref(x2), _cd(m2), _cd(n2), C.c_int(j), _cd(alpha), _cd(beta), C.c_double(2.0), C.c_double(1e-14))
        # floating point: the oracle is built with FMA contraction (-march=x86-64-v3), the product is not
        assert np.isclose(x1.value, x2.value, rtol=1e-12, atol=0)
        assert np.allclose(m1, m2, rtol=1e-12, atol=0) and np.allclose(n1, n2, rtol=1e-12, atol=0)


@pytest.mark.parametrize("n,ignorelast", [(1, 0), (2, 0), (7, 0), (7, 1), (30, 0)])
def test_bdqr_matches_oracle_and_numpy(oracle, n, ignorelast):
    from propack_b200 import _lib
    L, O = _lib.lib(), oracle.lib()
    rng = np.random.default_rng(n)
    d, e = rng.uniform(0.5, 2, size=n), rng.uniform(0.5, 2, size=n)
    B = np.zeros((n + 1, n))
    B[np.arange(n), np.arange(n)] = d
    B[np.arange(1, n + 1), np.arange(n)] = e
    out = []
    for lib, is_f77 in ((L, True), (O, False)):
        dd, ee = d.copy(), e.copy()
        Qt = np.zeros((n + 1, n + 1), order="F")
        c1, c2 = C.c_double(0), C.c_double(0)
        if is_f77:
            lib.dbdqr_(C.byref(C.c_int(ignorelast)), b"Y", C.byref(C.c_int(n)), _cd(dd), _cd(ee), C.byref(c1), C.byref(c2), _cd(Qt),
                       C.byref(C.c_int(n + 1)), C.c_size_t(1))
        else:
            lib.oracle_bdqr_d(C.c_int(ignorelast), C.c_int(1), C.c_int(n), _cd(dd), _cd(ee), C.byref(c1), C.byref(c2), _cd(Qt), C.c_int(n + 1))
        out.append((dd, ee, Qt, c1.value, c2.value))
    for x, y in zip(out[0], out[1]):
        assert np.allclose(np.asarray(x), np.asarray(y), rtol=1e-12, atol=1e-15)
    if not ignorelast:
        dd, ee, Qt, c1, c2 = out[0]
        R = np.diag(dd) + np.diag(ee[:n - 1], 1)
        assert np.allclose(Qt @ B, np.vstack([R, np.zeros((1, n))]), atol=1e-13)
        assert abs(Qt[n - 1, n] - c1) < 1e-14 and abs(Qt[n, n] - c2) < 1e-14


@pytest.mark.parametrize("k", [2, 5, 9])
def test_bsvdstep_matches_oracle_and_preserves_singular_values(oracle, k):
    from propack_b200 import _lib
    L, O = _lib.lib(), oracle.lib()
    n = 9
    rng = np.random.default_rng(k)
    d, e = rng.uniform(0.5, 2, size=n), rng.uniform(0.5, 2, size=n)
    res = []
    for lib, is_f77 in ((L, True), (O, False)):
        dd, ee = d.copy(), e.copy()
        P, Q = np.eye(n + 1, order="F"), np.eye(n, order="F")
        if is_f77:
            lib.dbsvdstep_(b"y", b"y", C.byref(C.c_int(n + 1)), C.byref(C.c_int(n)), C.byref(C.c_int(k)), C.byref(C.c_double(0.7)),
                           _cd(dd), _cd(ee), _cd(P), C.byref(C.c_int(n + 1)), _cd(Q), C.byref(C.c_int(n)), C.c_size_t(1), C.c_size_t(1))
        else:
            lib.oracle_bsvdstep_d(C.c_int(1), C.c_int(1), C.c_int(n + 1), C.c_int(n), C.c_int(k), C.c_double(0.7), _cd(dd), _cd(ee),
                                  _cd(P), C.c_int(n + 1), _cd(Q), C.c_int(n))
        res.append((dd, ee, P, Q))
    for x, y in zip(res[0], res[1]):
        assert np.allclose(x, y, rtol=1e-12, atol=1e-15)
    dd, ee, P, Q = res[0]
    B0 = np.zeros((n + 1, n)); B0[np.arange(n), np.arange(n)] = d; B0[np.arange(1, n + 1), np.arange(n)] = e
    B1 = np.zeros((n + 1, n)); B1[np.arange(n), np.arange(n)] = dd; B1[np.arange(1, n + 1), np.arange(n)] = ee
    if k == n:  # a full-length sweep is an orthogonal equivalence: B+ = P^T B Q  (dlansvd_irl.F:346-348)
        assert np.allclose(P.T @ B0 @ Q, B1, atol=1e-13)
        assert np.allclose(np.linalg.svd(B0, compute_uv=False), np.linalg.svd(B1, compute_uv=False), atol=1e-13)
    assert np.allclose(P.T @ P, np.eye(n + 1), atol=1e-14)


def test_refinebounds_matches_oracle(oracle):
    from propack_b200 import _lib
    L, O = _lib.lib(), oracle.lib()
    rng = np.random.default_rng(0)
    for n, k in ((20, 12), (12, 12)):
        theta = np.sort(rng.uniform(1, 3, size=k))[::-1].copy()
        theta[3] = theta[2] * (1 - 1e-14)  # a cluster
        bound = 10.0 ** rng.uniform(-12, -3, size=k)
        b1, b2 = bound.copy(), bound.copy()
        L.drefinebounds_(C.byref(C.c_int(n)), C.byref(C.c_int(k)), _cd(theta), _cd(b1), C.byref(C.c_double(1e-13)), C.byref(C.c_double(1e-12)))
        O.oracle_refinebounds_d(C.c_int(n), C.c_int(k), _cd(theta), _cd(b2), C.c_double(1e-13), C.c_double(1e-12))
        assert np.allclose(b1, b2, rtol=1e-12, atol=0)


def test_fortran_interface_module_names_exist_in_the_library():
    """include/propack_b200.f90 (never compiled here) must at least bind to symbols the library really exports."""
    import re
    from propack_b200 import _lib
    L = _lib.lib()
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "propack_b200.f90")).read()
    names = re.findall(r"bind\(C, name='([A-Za-z0-9_]+)'\)", src)
    assert len(names) >= 8
    for n in names:
        assert hasattr(L, n), n
    for p in "sdcz":
        assert hasattr(L, f"propack_b200_aprod_{p}_")
