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
for wl in sys.argv[1:] or ["c2"]:
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
