"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/propack_oracle.hpp).

Bit-identical numpy replica of the device generator ``propack_b200/csrc/dense_gen.cu`` (BASELINE config 3): the parity
tests build a small instance on both sides and require equality; the bench's CPU arm builds a row-scaled replica.
"""
import numpy as np

from propack_b200.synth import splitmix64


def dense_planted(m: int, n: int, seed: int, table: np.ndarray, rows=None) -> np.ndarray:
    """Rows `rows` (default all) of the m x n synthetic matrix, Fortran order; bit-identical to the device generator."""
    seed = np.uint64(seed)
    i = np.arange(m, dtype=np.uint64) if rows is None else np.asarray(rows, dtype=np.uint64)
    j = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x0 = splitmix64(seed ^ (np.uint64(0x1000000000000000) + np.uint64(2) * i))
        x1 = splitmix64(seed ^ (np.uint64(0x1000000000000000) + np.uint64(2) * i + np.uint64(1)))
        y0 = splitmix64(seed ^ (np.uint64(0x2000000000000000) + np.uint64(2) * j))
        y1 = splitmix64(seed ^ (np.uint64(0x2000000000000000) + np.uint64(2) * j + np.uint64(1)))
        hi = splitmix64(seed ^ i)
    A = np.empty((i.size, n), dtype=np.float64, order="F")
    T = np.ascontiguousarray(table, dtype=np.float64)
    for jj in range(n):
        with np.errstate(over="ignore"):
            h = splitmix64(hi ^ (np.uint64(0x3000000000000000) + j[jj]))
        a = ((h >> np.uint64(40)).astype(np.int64).astype(np.float64) + (-8388607.5)) * (1.0 / 8388608.0)
        w = x0 ^ y0[jj]
        for g in range(8):
            a = a + T[g][((w >> np.uint64(8 * g)) & np.uint64(255)).astype(np.intp)]
        w = x1 ^ y1[jj]
        for g in range(8):
            a = a + T[8 + g][((w >> np.uint64(8 * g)) & np.uint64(255)).astype(np.intp)]
        A[:, jj] = a
    return A


