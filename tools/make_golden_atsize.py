#!/usr/bin/env python
"""Generate tests/golden/c5_small_irl.npz: reference results for the BASELINE configs[4] pattern at 1M x 1M (bench workload
'c5-small': exactly 10 uniform columns per row, k=100, DLANSVD_IRL dim=300 p=200, tol=1e-10, CGS, start vector
default_rng(1).uniform) from
  * the CPU oracle (oracle/, C++ restatement of the reference): sigma, Lanczos steps, restarts;
  * SciPy's PROPACK translation scipy.sparse.linalg._svdp (irl_mode=True, same parameters): sigma  [--scipy, ~4 min].
The multi-GPU parity check (tests/dist_check.py) and the at-size GPU test compare against this file, so that the minutes of
CPU time are spent here, once, and not on a GPU lease.  Run in the build container; the .npz is committed.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

A, u0, k, dim, tol = bench.make_matrix("c5-small")
p = bench.IRL_P["c5-small"]
O.stats_reset()
t0 = time.perf_counter()
ref = O.lansvd_irl(A, k, dim, p=p, which="L", maxiter=bench.IRL_MAXITER, tol=tol, u0=u0, cgs=True, jobu=False, jobv=False)
st = O.stats()
print(f"oracle: {time.perf_counter() - t0:.1f} s, k={ref['k']} info={ref['info']} steps={st['nsteps']} restarts={st['nrestart']}")
out = {"sigma_oracle": ref["sigma"], "nsteps": np.int64(st["nsteps"]), "nrestart": np.int64(st["nrestart"]),
       "shape": np.array(A.shape, dtype=np.int64), "nnz": np.int64(A.nnz), "k": np.int64(k), "dim": np.int64(dim), "p": np.int64(p),
       "checksum_indices": np.int64(int(A.indices.astype(np.int64).sum())), "checksum_data": np.float64(A.data.sum())}
if "--scipy" in sys.argv:
    import scipy.sparse.linalg as spla
    from scipy.sparse.linalg._svdp import _svdp
    AT = A.T.tocsr()
    lop = spla.LinearOperator(A.shape, matvec=lambda x: A @ x, rmatvec=lambda x: AT @ x, dtype=A.dtype)
    t0 = time.perf_counter()
    u, sg, vh, bnd = _svdp(lop, k, which="LM", irl_mode=True, kmax=dim, v0=u0, tol=tol, cgs=True, shifts=p, maxiter=bench.IRL_MAXITER,
                           full_output=True, rng=np.random.default_rng(0))
    print(f"scipy _svdp: {time.perf_counter() - t0:.1f} s, max rel diff vs oracle {np.max(np.abs(np.sort(sg)[::-1] - ref['sigma']) / ref['sigma']):.2e}")
    out["sigma_scipy_svdp"] = np.sort(sg)[::-1]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c5_small_irl.npz"), **out)
print("written tests/golden/c5_small_irl.npz")
