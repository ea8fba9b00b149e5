"""``svdp`` -- the caller-facing mirror of SciPy's PROPACK wrapper.

Same signature, defaults, validation and return convention as
``scipy.sparse.linalg._svdp._svdp`` (the downstream binding of this PROPACK path), so
SciPy's own ``test_propack.py`` cases can be pointed at this function unchanged.  All arithmetic runs in
``libpropack_b200.so`` through the Fortran-ABI drivers (:mod:`propack_b200.f77`).
"""
from __future__ import annotations

import numpy as np

from . import f77
from .f77 import Operator


class LinAlgError(np.linalg.LinAlgError):
    pass


def svdp(A, k, which="LM", irl_mode=True, kmax=None, compute_u=True, compute_v=True, v0=None, full_output=False, tol=0,
         delta=None, eta=None, anorm=0, cgs=False, elr=True, min_relgap=0.002, shifts=None, maxiter=None, rng=None):
    """Singular value decomposition of a linear operator by Lanczos bidiagonalisation (PROPACK).

    Parameters, defaults and errors follow ``scipy.sparse.linalg._svdp._svdp``.  ``A`` may be a scipy sparse
    matrix or dense ndarray (device-resident operator), an :class:`Operator`, or a LinearOperator with
    ``matvec``/``rmatvec`` (host callback).  Returns ``(u, sigma, vh, bnd)``.
    """
    which = which.upper()
    if which not in {"LM", "SM"}:
        raise ValueError("`which` must be either 'LM' or 'SM'")
    if not irl_mode and which == "SM":
        raise ValueError("`which`='SM' requires irl_mode=True")
    rng = np.random.default_rng(rng)

    op = A if isinstance(A, Operator) else Operator(A)
    owns = not isinstance(A, Operator)
    try:
        m, n = op.shape
        typ = op.dtype
        if (k < 1) or (k > min(m, n)):
            raise ValueError("k must be positive and not greater than m or n")
        if kmax is None:
            kmax = 10 * k
        if maxiter is None:
            maxiter = 1000
        kmax = min(m + 1, n + 1, kmax)
        if kmax < k:
            raise ValueError(f"kmax must be greater than or equal to k, but kmax ({kmax}) < k ({k})")
        if v0 is None:
            u0 = rng.uniform(size=m).astype(typ)
            if np.iscomplexobj(u0):
                u0 = u0 + 1j * rng.uniform(size=m)
        else:
            u0 = np.asarray(v0)
            if u0.shape != (m,):
                raise ValueError(f"v0 must be of length {m}")
        R = f77.REAL[op.pfx]
        if delta is None:
            delta = np.sqrt(np.finfo(R).eps)
        if eta is None:
            eta = np.finfo(R).eps ** 0.75
        if irl_mode:
            if shifts is None:
                shifts = kmax - k
            if k > min(kmax - shifts, m, n):
                raise ValueError("shifts must satisfy k <= min(kmax-shifts, m, n)!")
            elif shifts < 0:
                raise ValueError("shifts must be >= 0!")
            r = f77.lansvd_irl(op, k, kmax, p=shifts, which="S" if which == "SM" else "L", maxiter=maxiter, tol=tol, u0=u0,
                               delta=delta, eta=eta, anorm=anorm, cgs=cgs, elr=elr, min_relgap=min_relgap,
                               jobu=compute_u, jobv=compute_v)
            if r["info"] == 0 and r["k"] < k:  # the Fortran driver is silent on maxiter exhaustion (dlansvd_irl.F:206,417)
                r["info"] = -1
        else:
            r = f77.lansvd(op, k, kmax, tol=tol, u0=u0, delta=delta, eta=eta, anorm=anorm, cgs=cgs, elr=elr,
                           jobu=compute_u, jobv=compute_v)
        if r["info"] > 0:
            raise LinAlgError(f"An invariant subspace of dimension {r['info']} was found.")
        elif r["info"] < 0:
            raise LinAlgError(f"k={k} singular triplets did not converge within kmax={kmax} iterations")
        return r["U"][:, :k], r["sigma"], r["V"][:, :k].conj().T, r["bnd"]
    finally:
        if owns:
            op.close()


def svds(A, k=6, ncv=None, tol=0, which="LM", v0=None, maxiter=None, return_singular_vectors=True, solver="propack", rng=None,
         options=None):
    """``scipy.sparse.linalg.svds(..., solver='propack')`` on this library: same arguments, defaults and return convention
    (singular values in ASCENDING order, ``u, s, vh``; ``return_singular_vectors`` in {True, False, "u", "vh"}).  SciPy's svds maps
    ``maxiter`` to PROPACK's ``kmax`` and always uses the implicitly restarted driver; so does this."""
    if solver != "propack":
        raise ValueError("propack_b200.svds only implements solver='propack'")
    if which not in {"LM", "SM"}:
        raise ValueError("`which` must be either 'LM' or 'SM'.")
    if return_singular_vectors not in {True, False, "u", "vh"}:
        raise ValueError("`return_singular_vectors` must be in {True, False, 'u', 'vh'}.")
    jobu = return_singular_vectors in {True, "u"}
    jobv = return_singular_vectors in {True, "vh"}
    u, s, vh, _ = svdp(A, k=k, tol=tol, which=which, maxiter=None, compute_u=jobu, compute_v=jobv, irl_mode=True, kmax=maxiter, v0=v0,
                       rng=rng)
    u, s, vh = u[:, ::-1], s[::-1], vh[::-1]
    if not return_singular_vectors:
        return s
    return (u if jobu else None), s, (vh if jobv else None)
