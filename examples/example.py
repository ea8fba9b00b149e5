#!/usr/bin/env python
"""The reference's example driver (README:89-118, ``<precision>/Examples/example.F`` -- absent from the checkout) on this
library: read a matrix (Harwell-Boeing, coordinate, diagonal or dense; ASCII or binary: propack_b200/matio.py), compute
the k largest singular triplets with xLANSVD (or xLANSVD_IRL with --irl, see example_irl.py), print them like the Fortran
program does, optionally compare against stored singular values as the reference's ``compare`` program does.

    python examples/example.py tests/golden/illc1850.rra --k 10 --compare tests/golden/Sigma_illc1850.ascii
    python examples/make_example_data.py && python examples/example.py examples/data/mhd1280b.cua --k 10      # complex16
    python examples/example.py examples/data/illc1850.coord --format coord --precision single

BASELINE configs[0] is exactly the first run (illc1850, k = 10, DLANSVD non-restarted).  Needs a B200.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from propack_b200 import f77, hb, matio  # noqa: E402
import propack_b200  # noqa: E402


def main(argv=None, irl_default=False):
    ap = argparse.ArgumentParser()
    ap.add_argument("matrix")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--kmax", type=int, default=None)
    ap.add_argument("--tol", type=float, default=1e-12)
    ap.add_argument("--irl", action="store_true", default=irl_default, help="implicitly restarted driver (example_irl.F)")
    ap.add_argument("--p", type=int, default=None, help="shifts per restart (IRL)")
    ap.add_argument("--compare", default=None, help="file with reference singular values, one per line")
    ap.add_argument("--format", default=None, choices=matio.FORMATS, help="matrix file format (default: by extension, else Harwell-Boeing)")
    ap.add_argument("--complex", action="store_true", help="the coordinate / diagonal / dense ASCII file holds (re, im) pairs")
    ap.add_argument("--precision", default="double", choices=["single", "double"],
                    help="single: s/c drivers, double: d/z drivers (the four precision directories of the reference)")
    ap.add_argument("--which", default="L", choices=["L", "S"], help="largest / smallest triplets (S needs --irl)")
    args = ap.parse_args(argv)

    A = matio.read_matrix(args.matrix, args.format, complex_values=args.complex)
    cplx = np.iscomplexobj(A if isinstance(A, np.ndarray) else A.data)
    dtype = {("single", False): np.float32, ("single", True): np.complex64, ("double", False): np.float64,
             ("double", True): np.complex128}[(args.precision, bool(cplx))]
    A = A.astype(dtype)
    m, n = A.shape
    kmax = args.kmax or min(m, n, max(10 * args.k, 100))
    nnz = int(np.count_nonzero(A)) if isinstance(A, np.ndarray) else A.nnz
    print(f" Matrix {os.path.basename(args.matrix)}: {m} x {n}, {nnz} non-zeros, {A.dtype}")
    op = f77.Operator(A)
    propack_b200.reset_counters()
    t0 = time.perf_counter()
    if args.irl:
        r = f77.lansvd_irl(op, args.k, kmax, p=args.p, which=args.which, tol=args.tol, cgs=True)
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
        rc = rc or (0 if err < (1e-10 if args.precision == "double" else 1e-4) else 2)
    return rc


if __name__ == "__main__":
    sys.exit(main())
