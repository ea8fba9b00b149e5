#!/usr/bin/env python
"""Row-sharded solve of a bench workload under several settings in ONE torchrun launch (the matrix is generated once):
    torchrun --nproc-per-node N tools/sharded_sweep.py c5 "PHASES=4,PUSH=32" "PHASES=2,PUSH=32" ...
Each setting: operator + solver created afresh, 1 warm-up solve, 2 timed solves (CUDA events, max over ranks), then one
profiled solve for the per-phase split.  Rank 0 prints one JSON line per setting."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import propack_b200  # noqa: E402
from propack_b200 import _lib, dist as pdist  # noqa: E402

wl = sys.argv[1]
settings = sys.argv[2:] or ["PHASES=4,PUSH=32"]
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = pdist.init_comm()
L = _lib.lib()
stream = torch.cuda.current_stream()
_lib.check(L.propack_b200_set_stream(C.c_void_p(stream.cuda_stream)), "set_stream")
A, u0, k, kmax, tol = bench.make_matrix(wl)
m, n = A.shape
lanmax = min(m + 1, n + 1, kmax)
for st in settings:
    kv = dict(x.split("=") for x in st.split(","))
    os.environ["PROPACK_B200_SPMV_PHASES"] = kv.get("PHASES", "4")
    os.environ["PROPACK_B200_PUSH_CTAS"] = kv.get("PUSH", "64")
    os.environ["PROPACK_B200_SELL_VARIANT"] = kv.get("VARIANT", "1")  # SELL kernel shape (sell.cu): 1 = 8 chains x 32 warps/SM, 2 = 4 chains x 64 warps/SM, ...
    os.environ["PROPACK_B200_PUSH"] = kv.get("MODE", "sm")            # sm (push kernel, default) | ce (copy engines)
    op = pdist.ShardedOperator(A, rank, world)
    sv = pdist.Solver(op, lanmax + 1, lanmax)

    def solve():
        sv.set_start(u0)
        propack_b200.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record(stream)
        sigma, kc, info = bench.session_solve(L, sv.id, wl, k, kmax, tol)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), propack_b200.counters(), sigma, kc, info
    solve()
    ts = []
    for _ in range(2):
        ms, ctr, sigma, kc, info = solve()
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(float(t.item()))
    propack_b200.set_profile(True)
    pms, _, _, _, _ = solve()
    ph = {kname: round(v["ms"], 1) for kname, v in propack_b200.phase_ms().items()}
    propack_b200.set_profile(False)
    # the local panel launches alone (no all-gather, whatever the gather buffer holds) and the un-staged product (NCCL all-gather +
    # panels), both directions, CUDA events, L2 flushed: what the transport has to hide
    iso = {}
    for skip in (1, 0):
        L.propack_b200_set_option(b"bench_skip_gather", C.c_int(skip))
        for adj in (0, 1):
            dist.barrier()
            t_ms = L.propack_b200_bench_spmv(C.c_int(op.handle), C.c_int(adj), C.c_int(10), C.c_int(1))
            iso[("local_panels" if skip else "nccl_gather+panels") + ("_t" if adj else "_n") + "_us"] = round(1e3 * t_ms, 1)
    L.propack_b200_set_option(b"bench_skip_gather", C.c_int(0))
    if rank == 0:
        print(json.dumps({"workload": wl, "spmv_isolated": iso, "world": world, "setting": st, "ms": ts, "steps": ctr["nsteps"], "converged": kc, "info": info,
                          "sigma_1": float(sigma[0]) if kc else None, "profiled_ms": round(pms, 1), "phases_ms": ph}), flush=True)
    sv.close(); op.close()
pdist.finalize_comm()
dist.destroy_process_group()
