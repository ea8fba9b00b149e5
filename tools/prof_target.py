#!/usr/bin/env python
"""Short target for `ncu --set full`: a few launches of each hot kernel on the C2 operands."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from propack_b200 import _lib, f77  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
A, u0, k, kmax, tol = bench.make_matrix(wl)
L = _lib.lib()
_lib.check(L.propack_b200_init(), "init")
op = bench.make_operator(A)
L.propack_b200_bench_reorth_d.argtypes = [C.c_long, C.c_int, C.c_int, C.c_int]
L.propack_b200_bench_gemm_d.argtypes = [C.c_long, C.c_int, C.c_int, C.c_int]
m = A.shape[0]
out = {}
REPS = int(os.environ.get("PROF_REPS", "3"))   # ncu captures: PROF_REPS=1 keeps the launch count small
for persist in ((1, 0) if os.environ.get("PROF_L2_AB") else (1,)):
    L.propack_b200_set_option(b"l2_persist", C.c_int(persist))
    for adj in (0, 1):
        for flush in ((1, 0) if os.environ.get("PROF_L2_AB") else (1,)):
            t = L.propack_b200_bench_spmv(C.c_int(op.handle), C.c_int(adj), C.c_int(REPS), C.c_int(flush))
            tag = f"spmv_{adj}" + ("" if persist else "_nopersist") + ("" if flush else "_noflush")
            out[tag + "_ms"] = t
            out[tag + "_gbs"] = (op.bytes_per_product(bool(adj)) + 8.0 * m) / t / 1e6
L.propack_b200_set_option(b"l2_persist", C.c_int(1))
for l in [int(x) for x in os.environ.get("PROF_LS", "16,64,256,300,537").split(",")]:
    t = L.propack_b200_bench_reorth_d(m, l, REPS, 1)
    out[f"reorth_l{l}_ms"] = t
    out[f"reorth_l{l}_gbs"] = 8.0 * m * (2 * l + 3) / t / 1e6
for (N, K) in ((50, 538), (101, 301)):
    t = L.propack_b200_bench_gemm_d(m, N, K, REPS)
    out[f"gemm_N{N}_K{K}_ms"] = t
    out[f"gemm_N{N}_K{K}_tflops"] = 2.0 * m * N * K / t / 1e9
    out[f"gemm_N{N}_K{K}_gbs"] = 8.0 * m * (N + K) / t / 1e6
print(out)
