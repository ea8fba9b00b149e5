#!/usr/bin/env python
"""The reference's example driver (README:89-118, ``double/Examples/example.F`` -- absent from the checkout) on this
library: read a Harwell-Boeing matrix, compute the k largest singular triplets with xLANSVD (or xLANSVD_IRL with
--irl), print them like the Fortran program does, optionally compare against stored singular values.

    python examples/example.py tests/golden/illc1850.rra --k 10 --compare tests/golden/Sigma_illc1850.ascii

BASELINE configs[0] is exactly this run (illc1850, k = 10, DLANSVD non-restarted).  Needs a B200.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from propack_b200 import f77, hb  # noqa: E402
import propack_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("matrix")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--kmax", type=int, default=None)
    ap.add_argument("--tol", type=float, default=1e-12)
    ap.add_argument("--irl", action="store_true", help="implicitly restarted driver (example_irl.F)")
    ap.add_argument("--p", type=int, default=None, help="shifts per restart (IRL)")
    ap.add_argument("--compare", default=None, help="file with reference singular values, one per line")
    args = ap.parse_args()

    A = hb.read_hb(args.matrix).tocsr()
    m, n = A.shape
    kmax = args.kmax or min(m, n, max(10 * args.k, 100))
    print(f" Matrix {os.path.basename(args.matrix)}: {m} x {n}, {A.nnz} non-zeros, {A.dtype}")
    op = f77.Operator(A)
    propack_b200.reset_counters()
    t0 = time.perf_counter()
    if args.irl:
        r = f77.lansvd_irl(op, args.k, kmax, p=args.p, tol=args.tol, cgs=True)
    else:
        r = f77.lansvd(op, args.k, kmax, tol=args.tol, cgs=True)    # all-zero U(:,1): the library draws the start vector
    dt = time.perf_counter() - t0
    op.close()
    print(f" info = {r['info']}, converged triplets = {r['k']}, time = {dt:.4f} s")
    print("    i        sigma(i)                  bnd(i)            ||A v - sigma u||")
    res = np.linalg.norm(A @ r["V"] - r["U"] * r["sigma"], axis=0)
    for i in range(r["k"]):
        print(f" {i + 1:4d}  {r['sigma'][i]:24.16e}  {r['bnd'][i]:12.4e}  {res[i]:12.4e}")
    c = propack_b200.counters()
    print(f" matrix-vector products = {c['nopx']}, reorthogonalisations = {c['nreorth']}, bidiagonal SVDs = {c['nbsvd']}, "
          f"Lanczos steps = {c['nsteps']}")
    rc = 0 if r["info"] == 0 and r["k"] == args.k else 1
    if args.compare:
        err = hb.compare(r["sigma"], hb.read_sigma_ascii(args.compare))
        print(f" max relative error of sigma vs {os.path.basename(args.compare)} = {err:.3e}")
        rc = rc or (0 if err < 1e-10 else 2)
    return rc


if __name__ == "__main__":
    sys.exit(main())
