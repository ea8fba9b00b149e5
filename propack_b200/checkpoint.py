"""Checkpoint / resume of a Lanczos bidiagonalisation ("checkpoint dump of (U, V, B)").

The reference's own mechanism for continuing a factorisation is ``xLANBPRO`` with ``k0 > 0`` (double/dlanbpro.F:15-16, 231-275):
given ``U(:,1:k0+1)``, ``V(:,1:k0)``, ``B(1:k0,:)`` and ``rnorm = beta_{k0+1}`` it extends the factorisation to ``k`` steps.  This
module stores exactly that state in one ``.npz`` file and resumes from it through the same entry point
(``propack_b200.f77.lanbpro`` -> ``{s,d,c,z}lanbpro_``), so a long run can be split across processes or survive a restart of
the job: ``A V_k = U_{k+1} B_k`` holds after the resume as if the run had never been interrupted.
"""
from __future__ import annotations

import numpy as np

_FORMAT = 1


def save(path, U, V, B, k, rnorm, anorm=0.0, meta=None):
    """Write the state after ``k`` Lanczos steps: ``U[:, :k+1]``, ``V[:, :k]``, ``B[:k]`` (alpha, beta), ``rnorm``, ``anorm``."""
    U, V, B = np.asarray(U), np.asarray(V), np.asarray(B)
    k = int(k)
    if k < 1 or U.shape[1] < k + 1 or V.shape[1] < k or B.shape[0] < k or B.shape[1] != 2:
        raise ValueError("checkpoint.save: need U (m, >=k+1), V (n, >=k), B (>=k, 2) and k >= 1")
    np.savez(path, format=np.int64(_FORMAT), k=np.int64(k), rnorm=np.float64(rnorm), anorm=np.float64(anorm),
             U=np.ascontiguousarray(U[:, :k + 1]), V=np.ascontiguousarray(V[:, :k]), B=np.ascontiguousarray(B[:k]),
             meta=np.array(repr(meta) if meta is not None else ""))


def load(path, kmax=None):
    """Read a checkpoint.  Returns a dict with Fortran-ordered ``U (m, kmax+1)``, ``V (n, kmax)``, ``B (kmax, 2)`` whose leading
    ``k+1`` / ``k`` / ``k`` columns / rows hold the stored state (``kmax`` defaults to ``k``: room to extend is the caller's choice),
    plus ``k``, ``rnorm``, ``anorm``."""
    with np.load(path, allow_pickle=False) as z:
        if int(z["format"]) != _FORMAT:
            raise ValueError(f"checkpoint.load: unknown format {int(z['format'])}")
        k = int(z["k"])
        kmax = k if kmax is None else int(kmax)
        if kmax < k:
            raise ValueError("checkpoint.load: kmax is smaller than the stored number of steps")
        Us, Vs, Bs = z["U"], z["V"], z["B"]
        U = np.zeros((Us.shape[0], kmax + 1), dtype=Us.dtype, order="F")
        V = np.zeros((Vs.shape[0], kmax), dtype=Vs.dtype, order="F")
        B = np.zeros((kmax, 2), dtype=Bs.dtype, order="F")
        U[:, :k + 1] = Us; V[:, :k] = Vs; B[:k] = Bs
        return {"U": U, "V": V, "B": B, "k": k, "rnorm": float(z["rnorm"]), "anorm": float(z["anorm"]), "meta": str(z["meta"])}


def resume(op, state, k, **lanbpro_options):
    """Extend the stored factorisation to ``k`` steps on the device (``xLANBPRO`` with ``k0 = state['k']``).  ``state`` comes from
    :func:`load` with ``kmax >= k``; its arrays are updated in place.  Returns ``(k_done, rnorm, ierr, anorm)`` like ``f77.lanbpro``."""
    from . import f77
    if state["U"].shape[1] < k + 1 or state["V"].shape[1] < k or state["B"].shape[0] < k:
        raise ValueError("checkpoint.resume: load the checkpoint with kmax >= k")
    out = f77.lanbpro(op, state["k"], k, state["U"], state["V"], state["B"], state["rnorm"], anorm=state["anorm"], **lanbpro_options)
    state["k"], state["rnorm"], state["anorm"] = out[0], out[1], out[3]
    return out
