"""Synthetic workload definition of BASELINE config 3 (dense tall-skinny, generated on the device).

The 2M x 4096 matrix (65.5 GB) never exists on the host: ``csrc/dense_gen.cu`` evaluates a counter-based formula on the
GPU.  This module only builds the small parameter table of that formula and registers the operator; the bit-identical
numpy replica used by the parity tests and by the bench's CPU arm lives with the other test infrastructure in
``oracle/synth_ref.py``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def planted_coefficients(m: int, n: int, top: float = 100.0, halving: float = 20.0) -> np.ndarray:
    """c_r, r = 0..127: planted singular values ~ c_r * sqrt(m n) = top * 2^(-r/halving) * (noise edge)."""
    edge = (np.sqrt(m) + np.sqrt(n)) / np.sqrt(3.0)          # largest singular value of the uniform(-1,1) bulk
    r = np.arange(128, dtype=np.float64)
    return top * edge * 2.0 ** (-r / halving) / np.sqrt(float(m) * float(n))


def planted_table(c: np.ndarray) -> np.ndarray:
    """T[g][b] = sum_{t=0..7} (+1 | -1 by bit t of b) * c[8g+t], summed in the order t = 0..7 (16 x 256 doubles)."""
    T = np.zeros((16, 256), dtype=np.float64)
    b = np.arange(256)
    for g in range(16):
        acc = np.zeros(256)
        for t in range(8):
            acc = acc + np.where((b >> t) & 1, -c[8 * g + t], c[8 * g + t])
        T[g] = acc
    return T


def device_dense_planted(m: int, n: int, seed: int, table: np.ndarray):
    """Register the synthetic dense operator on the GPU; returns a ``propack_b200.f77.Operator`` (real*8)."""
    from . import f77
    from ._lib import check, lib
    L = lib()
    T = np.ascontiguousarray(table, dtype=np.float64)
    L.propack_b200_dense_create_synthetic_d.argtypes = [C.c_int, C.c_int, C.c_ulonglong, C.c_void_p]
    h = check(L.propack_b200_dense_create_synthetic_d(m, n, seed, T.ctypes.data_as(C.c_void_p)), "dense_create_synthetic")
    op = f77.Operator.__new__(f77.Operator)
    op.handle, op._cb, op.dtype, op.pfx, op.shape = h, None, np.dtype(np.float64), "d", (m, n)
    op.iparm = np.array([h, 0], dtype=np.int32)
    op.parm = np.zeros(2)
    return op
