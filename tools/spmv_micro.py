#!/usr/bin/env python
"""SpMV micro-benchmark on the bench workloads: the registered operator's fused product (y = A x + coef*prev, ||y||) for
both directions, L2 flushed between launches, CUDA events.  PROPACK_B200_SPMV=csr selects the round-1 CSR warp-group
kernel, the default is the SELL-32-sigma kernel (compare by running the script twice)."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from propack_b200 import _lib, f77  # noqa: E402

L = _lib.lib()
_lib.check(L.propack_b200_init(), "init")
out = {"kernel": os.environ.get("PROPACK_B200_SPMV", "sell")}
def banded(m, per=10, half=512):
    """10 columns per row inside a +-`half` band around the diagonal: the same bytes per product as config 5, but the x
    gathers of a 32-row slice fall into a few KB -- shows what the kernel does when the gather is not the bound."""
    import numpy as np
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    off = rng.integers(-half, half + 1, size=(m, per), dtype=np.int64)
    cols = np.clip(np.arange(m, dtype=np.int64)[:, None] + off, 0, m - 1).astype(np.int32)
    cols.sort(axis=1)
    A = sp.csr_array((rng.standard_normal(size=(m, per)).ravel(), cols.ravel(), np.arange(0, m * per + 1, per, dtype=np.int64)), shape=(m, m))
    A.sum_duplicates(); A.sort_indices()
    return A


for wl in sys.argv[1:] or ["c2"]:
    if wl.startswith("banded"):
        A = banded(int(wl.split(":")[1]) if ":" in wl else 10_000_000)
    else:
        A, u0, k, kmax, tol = bench.make_matrix(wl)
    t0 = time.perf_counter()
    op = f77.Operator(A)
    out[f"{wl}_create_s"] = time.perf_counter() - t0
    w = A.dtype.itemsize
    for adj in (0, 1):
        for flush in (1, 0):
            t = L.propack_b200_bench_spmv(C.c_int(op.handle), C.c_int(adj), C.c_int(20), C.c_int(flush))
            nbytes = op.bytes_per_product(bool(adj)) + w * (A.shape[1] if adj else A.shape[0])
            out[f"{wl}_{'t' if adj else 'n'}{'' if flush else '_noflush'}"] = {"us": 1e3 * t, "gbs": nbytes / t / 1e6}
    op.close()
    del A
print(json.dumps(out))
