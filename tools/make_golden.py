#!/usr/bin/env python
"""Generate tests/golden/*.npz -- the fixtures that pin the oracle (and through it the CUDA path).

Run in the BUILD container only (it imports SciPy's PROPACK translation and SciPy's copy of the two
PROPACK example matrices); the GPU box just reads the committed .npz files.

Sources of truth recorded here:
  * illc1850 (real 1850x712) and mhd1280b (complex 1280x1280): the PROPACK example matrices
    (reference README:89-118), taken from scipy/sparse/linalg/tests/propack_test_data.npz;
  * dense LAPACK SVD of both (numpy.linalg.svd) -- ground truth for sigma;
  * scipy.sparse.linalg._svdp._svdp (SciPy's C translation of this PROPACK; same xLANSVD / xLANSVD_IRL
    argument lists) on fixed start vectors -- the "reference run here" the oracle is checked against;
  * LAPACK xLARNV(idist=2, iseed=(1,3,5,7)) known answers from the image's scipy_openblas, all four types
    (pins dgetu0's start vector bit-exactly, dgetu0.F:41-44,69);
  * the seeded 10x20 matrices of SciPy's test_propack.py::test_svdp.
"""
import ctypes as C
import glob
import os

import numpy as np
import scipy
import scipy.sparse as sp
from scipy.sparse.linalg._svdp import _svdp

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def coo_dict(A, prefix):
    A = sp.coo_array(A)
    return {f"{prefix}_row": A.row.astype(np.int32), f"{prefix}_col": A.col.astype(np.int32), f"{prefix}_data": A.data,
            f"{prefix}_shape": np.array(A.shape, dtype=np.int64)}


def run_svdp(A, k, irl, kmax, u0, tol, **kw):
    u, s, vt, b = _svdp(A, k, irl_mode=irl, kmax=kmax, v0=u0, tol=tol, full_output=True, rng=np.random.default_rng(0), **kw)
    return s, b


def main():
    d = np.load(os.path.join(os.path.dirname(scipy.__file__), "sparse/linalg/tests/propack_test_data.npz"), allow_pickle=True)
    A_real = sp.csr_array(d["A_real"].item())
    A_cplx = sp.csr_array(d["A_complex"].item())
    g = {}
    g.update(coo_dict(A_real, "illc1850"))
    g.update(coo_dict(A_cplx, "mhd1280b"))
    g["illc1850_svd"] = np.linalg.svd(A_real.toarray(), compute_uv=False)
    g["mhd1280b_svd"] = np.linalg.svd(A_cplx.toarray(), compute_uv=False)
    rng = np.random.default_rng(0)
    u0r = rng.uniform(size=A_real.shape[0])
    u0c = rng.uniform(size=A_cplx.shape[0]) + 1j * rng.uniform(size=A_cplx.shape[0])
    g["illc1850_u0"] = u0r
    g["mhd1280b_u0"] = u0c
    # SciPy's PROPACK translation, BASELINE config 1 (k=10, kmax=100, tol=1e-12), both drivers, both GS flavours
    for cgs in (0, 1):
        s, b = run_svdp(A_real, 10, False, 100, u0r, 1e-12, cgs=bool(cgs))
        g[f"illc1850_scipy_lansvd_k10_cgs{cgs}_sigma"], g[f"illc1850_scipy_lansvd_k10_cgs{cgs}_bnd"] = s, b
    s, b = run_svdp(A_real, 10, True, 50, u0r, 1e-12, shifts=40)
    g["illc1850_scipy_irl_k10_dim50_sigma"], g["illc1850_scipy_irl_k10_dim50_bnd"] = s, b
    s, b = run_svdp(A_real, 200, False, 712, u0r, 0.0)
    g["illc1850_scipy_lansvd_k200_sigma"] = s
    s, b = run_svdp(A_cplx, 10, False, 200, u0c, 1e-12)
    g["mhd1280b_scipy_lansvd_k10_sigma"] = s
    s, b = run_svdp(A_cplx.astype(np.complex64), 10, False, 200, u0c.astype(np.complex64), 1e-5)
    g["mhd1280b_scipy_clansvd_k10_sigma"] = s
    s, b = run_svdp(A_real.astype(np.float32), 10, False, 100, u0r.astype(np.float32), 1e-5)
    g["illc1850_scipy_slansvd_k10_sigma"] = s
    np.savez_compressed(os.path.join(OUT, "propack_examples.npz"), **g)

    # LAPACK xLARNV known answers
    lp = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so")))[0]
    L = C.CDLL(lp)
    k = {}
    for name, dt, mult in (("d", np.float64, 1), ("s", np.float32, 1), ("z", np.complex128, 1), ("c", np.complex64, 1)):
        n = 300  # spans several 64/128-value DLARUV blocks
        x = np.zeros(n, dtype=dt)
        seed = np.array([1, 3, 5, 7], dtype=np.int32)
        getattr(L, f"scipy_{name}larnv_")(C.byref(C.c_int(2)), seed.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)),
                                          x.ctypes.data_as(C.c_void_p))
        k[f"{name}larnv_x"] = x
        k[f"{name}larnv_seed_after"] = seed
    np.savez_compressed(os.path.join(OUT, "larnv_kat.npz"), **k)

    # SciPy test_propack.py::test_svdp inputs (10x20, 20 % zeroed, default_rng(0)) and dense sigma
    t = {}
    for name, dt in (("s", np.float32), ("d", np.float64), ("c", np.complex64), ("z", np.complex128)):
        rng = np.random.default_rng(0)
        A = rng.random((10, 20)).astype(dt)
        if np.iscomplexobj(A):
            A = (A + 1j * rng.random((10, 20))).astype(dt)
        A[A.real > 0.8] = 0  # sparsify like the SciPy test does ("20 % zeroed")
        t[f"A_{name}"] = A
        t[f"svd_{name}"] = np.linalg.svd(A.astype(np.complex128 if np.iscomplexobj(A) else np.float64), compute_uv=False)
    np.savez_compressed(os.path.join(OUT, "small_dense.npz"), **t)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
