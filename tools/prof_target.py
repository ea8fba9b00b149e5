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
NO_SPMV = bool(os.environ.get("PROF_NO_SPMV"))   # GEMV / GEMM launches only: no need to build the workload matrix
L = _lib.lib()
_lib.check(L.propack_b200_init(), "init")
if not NO_SPMV:
    A, u0, k, kmax, tol = bench.make_matrix(wl)
    op = bench.make_operator(A)
L.propack_b200_bench_reorth_d.argtypes = [C.c_long, C.c_int, C.c_int, C.c_int]
L.propack_b200_bench_gemm_d.argtypes = [C.c_long, C.c_int, C.c_int, C.c_int]
m = bench.WORKLOADS[wl][0]
out = {}
REPS = int(os.environ.get("PROF_REPS", "3"))   # ncu captures: PROF_REPS=1 keeps the launch count small
for adj in (() if NO_SPMV else (0, 1)):
    t = L.propack_b200_bench_spmv(C.c_int(op.handle), C.c_int(adj), C.c_int(REPS), C.c_int(1))
    out[f"spmv_{adj}_ms"] = t
    out[f"spmv_{adj}_gbs"] = (op.bytes_per_product(bool(adj)) + 8.0 * m) / t / 1e6
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
