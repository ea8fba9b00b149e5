#!/usr/bin/env python
"""bench.py -- headline measurement of the PROPACK Lanczos-bidiagonalisation hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference|scipy] [--workload c5|c2|c3|c4|...]

One "step" = one complete solve (time to k triplets) of the workload; the metric is whole-job Lanczos steps per second
(BASELINE.json metric "time to k triplets + Lanczos steps/s"), with the time to k triplets reported as ms_per_step.

Default workload at every N (BASELINE.json configs[4], "C5", the north-star target): synthetic CSR 10M x 10M, exactly 10
uniform columns per row (1e8 nnz), N(0,1) values, default_rng(0); top-100 triplets by DLANSVD_IRL (dim=300, p=200),
tol=1e-10, double, CGS reorthogonalisation, start vector default_rng(1).uniform.  It fits one GPU (1.2 GB matrix x 2
directions x 2 formats + 2 x 24 GB bases); N > 1 row-shards the same problem (strong scaling).  Inputs are far larger
than the 126 MB L2, so no explicit L2 flush is needed between steps.  `--workload c2` is BASELINE configs[1].

Keys (see the task contract): value / ms_per_step are device-timed (CUDA events on the library stream) with the matrix
and start vector already resident in HBM; e2e goes through the Fortran-ABI `dlansvd[_irl]_` with HOST buffers (matrix
upload + start vector up, U/V/sigma back) inside the timed region; roofline describes the kernel with the largest share
of the solve, measured live with CUDA events in a profiled solve of the same workload; cpu_baseline is the CPU oracle (a
port of the reference, OpenMP) on a bounded sample.  `--impl reference` times the oracle's FULL solve of the same
workload (same driver, k, tol, start vector: time to k triplets, like for like) on all host cores, plus SciPy's PROPACK
translation (`scipy.sparse.linalg._svdp`) on a bounded 1/10-scale replica as an independent cross-check (at --gpus 1 only: once
per scaling sweep is enough).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (m, n, density, k, kmax, tol)           DLANSVD (non-restarted), BASELINE configs[1] and reduced copies
    "c2": (1_000_000, 1_000_000, 1e-5, 50, 600, 1e-10),
    "c2-small": (100_000, 100_000, 1e-4, 50, 600, 1e-10),
    "c2-tiny": (20_000, 20_000, 5e-4, 10, 200, 1e-10),
    # BASELINE configs[4] ("C5"): 10M x 10M, 10 distinct uniform columns per row (1e8 nnz), k=100, DLANSVD_IRL dim=300 p=200
    "c5": (10_000_000, 10_000_000, None, 100, 300, 1e-10),
    "c5-small": (1_000_000, 1_000_000, None, 100, 300, 1e-10),
}
WORKLOADS["c4"] = (2_000_000, 2_000_000, "powerlaw", 64, 700, 1e-10)        # BASELINE configs[3]: complex16, ZLANSVD
WORKLOADS["c4-small"] = (200_000, 200_000, "powerlaw", 32, 400, 1e-10)
COMPLEX = {"c4", "c4-small"}
# BASELINE configs[2] ("C3"): dense tall-skinny 2M x 4096 (65.5 GB, generated on the device), k=100, DLANSVD_IRL dim=200 p=100
WORKLOADS["c3"] = (2_000_000, 4096, "dense", 100, 200, 1e-10)
WORKLOADS["c3-small"] = (200_000, 1024, "dense", 50, 100, 1e-10)
DENSE = {"c3", "c3-small"}
DENSE_SEED = 0
DENSE_CPU_ROWS = {"c3": 100_000, "c3-small": 20_000}   # rows of the CPU arm's replica (cost per step is linear in rows)
IRL_P = {"c5": 200, "c5-small": 200, "c3": 100, "c3-small": 50}   # workloads solved with DLANSVD_IRL: shifts per restart (kmax column = dim)
IRL_MAXITER = 50
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at working size, from the committed
# `ncu --set full` captures (profiles/r02_ncu_extract_c5.txt); filled in by hand from those files, None = not captured
# (SpMV: per product = the 2 column-panel launches of spmv_sell_kernel on config 5, mean of A x and A^T u; reorth: one GEMV
# pair gemv_t_tma_kernel + gemv_n_kernel at L = 1e7, l = 300 -- 48.27 GB against 48.24 GB algorithmic)
NCU_TRAFFIC_BYTES_PER_LAUNCH = {("c5", "spmv"): 1.676e9, ("c5", "reorth"): 48.27e9, ("c2", "spmv"): 140.03e6 + 5.3e6, ("c2", "reorth"): None}
NCU_TRAFFIC_NOTE = {("c5", "spmv"): "dram__bytes_read+write of the 2 panel launches of one product (profiles/r02_ncu_extract_c5.txt [0]+[1] / [6]+[7]); "
                                    "algorithmic 1.48 GB + the second panel's read-modify-write of y (0.16 GB)",
                    ("c5", "reorth"): "one GEMV pair at L=1e7, l=300 (profiles/r02_ncu_extract_gemv_gemm.txt [4]+[5]): 48.27 GB measured vs 48.24 GB algorithmic"}
CPU_SAMPLE_STEPS = 150  # Lanczos steps of the same problem the in-line cpu_baseline runs (bounded sample, ~10-30 s)
REF_BUDGET_S = 500.0    # --impl reference: stop adding full CPU solves once the projected run time passes this
SCIPY_START_BEFORE_S = 400.0   # ... and only start the SciPy cross-check (~4 min on config 5's 1/10 replica) this early in the run


class DenseSpec:
    """The synthetic dense operator of config 3: never materialised on the host at full size (propack_b200/synth.py)."""

    def __init__(self, m, n, table):
        self.shape, self.table, self.dtype, self.nnz = (m, n), table, np.dtype(np.float64), m * n

    def replica(self, rows):
        from oracle import synth_ref
        return synth_ref.dense_planted(self.shape[0], self.shape[1], DENSE_SEED, self.table, rows=np.arange(rows))


def make_operator(A):
    """Device-resident operator for a workload matrix."""
    from propack_b200 import f77, synth
    if isinstance(A, DenseSpec):
        return synth.device_dense_planted(A.shape[0], A.shape[1], DENSE_SEED, A.table)
    return f77.Operator(A)


def make_matrix(name):
    import scipy.sparse as sp
    m, n, dens, k, kmax, tol = WORKLOADS[name]
    rng = np.random.default_rng(0)
    if dens == "dense":
        from propack_b200 import synth
        A = DenseSpec(m, n, synth.planted_table(synth.planted_coefficients(m, n)))
        return A, np.random.default_rng(1).uniform(size=m), k, kmax, tol
    if dens == "powerlaw":   # row lengths min(1 + floor(Zipf(2)), 10000), renormalised to ~10 per row, complex values
        lens = np.minimum(rng.zipf(2.0, size=m), 10_000).astype(np.float64)
        lens = np.maximum(1, np.rint(lens * (10.0 * m / lens.sum()))).astype(np.int64)
        lens = np.minimum(lens, n)
        rows = np.repeat(np.arange(m, dtype=np.int32), lens)
        cols = rng.integers(0, n, size=rows.size, dtype=np.int32)
        vals = rng.standard_normal(rows.size) + 1j * rng.standard_normal(rows.size)
        A = sp.csr_array(sp.coo_array((vals, (rows, cols)), shape=(m, n)))   # duplicates inside a row are summed
        A.sum_duplicates()
    elif dens is None:   # exactly 10 uniform columns per row (the rare duplicates inside a row are summed)
        per = 10
        cols = rng.integers(0, n, size=(m, per), dtype=np.int32)
        cols.sort(axis=1)
        vals = rng.standard_normal(size=(m, per))
        A = sp.csr_array((vals.ravel(), cols.ravel(), np.arange(0, m * per + 1, per, dtype=np.int64)), shape=(m, n))
        A.sum_duplicates()
    else:
        A = sp.random_array((m, n), density=dens, format="csr", rng=rng, data_sampler=rng.standard_normal)
    A.sort_indices()
    u0 = np.random.default_rng(1).uniform(size=m).astype(A.dtype)
    return A, u0, k, kmax, tol


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu = gpu_index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def wl_dtype(name):
    return ("c128", 16) if name in COMPLEX else ("f64", 8)


def reorth_bytes(ctr, w=8):
    """one Gram-Schmidt pass of a length-L vector against l columns moves w*L*(2l+3) bytes."""
    return w * (2 * ctr["reorth_elems"] + 3 * ctr["reorth_vec_elems"])


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (a port of the reference; the Fortran cannot be built here)
# ------------------------------------------------------------------------------------------------------
def cpu_sample(A, u0, steps):
    """Time `steps` Lanczos steps (DLANBPRO, k0=0) of the same problem on the host cores. Returns (steps/s, seconds)."""
    from oracle import oracle_py as O
    L = O.lib()
    dt_ = np.dtype(A.dtype)
    op = O.Operator(A, dt_)
    m, n = A.shape
    U = np.zeros((m, steps + 1), dtype=dt_, order="F"); V = np.zeros((n, steps), dtype=dt_, order="F")
    U[:, 0] = u0
    B = np.zeros((steps, 2), order="F")
    eps = np.finfo(np.float64).eps
    doption = np.array([np.sqrt(eps), eps ** 0.75, 0.0])
    ioption = np.array([1, 1], dtype=np.int32)
    kk, rn, ierr = C.c_int(steps), C.c_double(float(np.linalg.norm(u0))), C.c_int(0)
    O.stats_reset()
    t0 = time.perf_counter()
    getattr(L, "oracle_lanbpro_z" if dt_.kind == "c" else "oracle_lanbpro_d")(C.c_int(m), C.c_int(n), C.c_int(0), C.byref(kk), *op.args(), U.ctypes.data_as(C.c_void_p), C.c_long(m),
                       V.ctypes.data_as(C.c_void_p), C.c_long(n), B.ctypes.data_as(C.c_void_p), C.c_int(steps), C.byref(rn),
                       doption.ctypes.data_as(C.c_void_p), ioption.ctypes.data_as(C.c_void_p), C.byref(ierr))
    dt = time.perf_counter() - t0
    return O.stats()["nsteps"] / dt, dt, int(L.oracle_num_threads())


def cpu_arm(name, A, u0, steps):
    """CPU arm on the workload (sparse: the same matrix; dense config 3: a row-scaled replica, rate scaled back).
    Returns (steps/s on the full problem, seconds, cores, note)."""
    if isinstance(A, DenseSpec):
        rows = min(DENSE_CPU_ROWS[name], A.shape[0])
        Ar = A.replica(rows)
        v, dt, cores = cpu_sample(Ar, u0[:rows], steps)
        scale = rows / A.shape[0]
        return v * scale, dt, cores, (f"dense replica with the first {rows} of {A.shape[0]} rows (the full matrix is 65.5 GB and is only ever "
                                      f"generated on the device); measured {v:.1f} steps/s on the replica, scaled by {scale:.4f} because every "
                                      f"per-step cost is linear in the row count")
    v, dt, cores = cpu_sample(A, u0, steps)
    return v, dt, cores, ""


def cpu_full_solve(name, A, u0, k, kmax, tol):
    """The oracle's complete driver run of the workload (time to k triplets incl. Ritz vectors) on all host cores.
    Returns (seconds, steps, converged, info, cores, sigma, note)."""
    from oracle import oracle_py as O
    L = O.lib()
    note = ""
    if isinstance(A, DenseSpec):
        rows = min(DENSE_CPU_ROWS[name], A.shape[0])
        Am = A.replica(rows)
        u0 = u0[:rows]
        note = (f"dense replica with the first {rows} of {A.shape[0]} rows (the full matrix is 65.5 GB and is only ever generated on the "
                f"device); steps/s scaled by {rows / A.shape[0]:.4f} because every per-step cost is linear in the row count")
    else:
        Am = A
    O.stats_reset()
    t0 = time.perf_counter()
    if name in IRL_P:
        r = O.lansvd_irl(Am, k, kmax, p=IRL_P[name], which="L", maxiter=IRL_MAXITER, tol=tol, u0=u0, cgs=True, dtype=np.dtype(Am.dtype))
    else:
        r = O.lansvd(Am, k, kmax, tol=tol, u0=u0, cgs=True, dtype=np.dtype(Am.dtype))
    dt = time.perf_counter() - t0
    scale = (Am.shape[0] / A.shape[0]) if isinstance(A, DenseSpec) else 1.0
    return dt / scale, O.stats()["nsteps"], r["k"], r["info"], int(L.oracle_num_threads()), r["sigma"], note


def scipy_svdp_sample(name, budget_rows=1_000_000):
    """SciPy's own PROPACK translation (scipy.sparse.linalg._svdp, the importable build of the reference: SURVEY 8c/8d item 2)
    on a 1/10-scale replica of the sparse workloads, same driver parameters.  Returns a dict for the JSON line, or None."""
    small = {"c5": "c5-small", "c2": "c2-small", "c4": "c4-small"}.get(name, name if name.endswith("-small") or name.endswith("-tiny") else None)
    if small is None or small in DENSE:
        return None
    try:
        from scipy.sparse.linalg._svdp import _svdp
    except Exception as e:   # scipy without the PROPACK extension
        return {"unavailable": repr(e)}
    A, u0, k, kmax, tol = make_matrix(small)
    if A.shape[0] > budget_rows:
        return None
    AT = A.T.tocsr()
    calls = [0]
    import scipy.sparse.linalg as spla
    lop = spla.LinearOperator(A.shape, matvec=lambda x: (calls.__setitem__(0, calls[0] + 1), A @ x)[1],
                              rmatvec=lambda x: (calls.__setitem__(0, calls[0] + 1), AT.conj() @ x)[1], dtype=A.dtype)
    irl = small in IRL_P
    t0 = time.perf_counter()
    try:
        u, sg, vh, bnd = _svdp(lop, k, which="LM", irl_mode=irl, kmax=kmax, v0=u0, tol=tol, cgs=True, shifts=IRL_P.get(small),
                               maxiter=IRL_MAXITER if irl else None, full_output=True, rng=np.random.default_rng(0))
    except Exception as e:
        return {"workload": small, "error": repr(e)[:200]}
    dt = time.perf_counter() - t0
    return {"workload": f"{small}: {A.shape[0]}x{A.shape[1]}, nnz={A.nnz}, k={k}, {'IRL dim=%d p=%d' % (kmax, IRL_P[small]) if irl else 'kmax=%d' % kmax}",
            "time_to_k_triplets_s": dt, "matvecs": calls[0], "steps_per_s": calls[0] / 2.0 / dt, "sigma_1": float(sg.max()),
            "impl": "scipy.sparse.linalg._svdp (SciPy's C translation of PROPACK; scipy.sparse matvec callbacks)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm uses all the host cores it can (set before libgomp loads)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    A, u0, k, kmax, tol = make_matrix(args.workload)
    # Each step is a FULL solve (time to k triplets).  The run is bounded: after the first solve, further warm-up / timed solves
    # are only done while the projected total stays under REF_BUDGET_S -- on config 5 one solve is minutes of CPU time.
    secs, steps_per_solve, runs = [], 0, 0
    t_start = time.perf_counter()
    want = args.warmup + args.steps
    res = None
    while runs < want:
        res = cpu_full_solve(args.workload, A, u0, k, kmax, tol)
        runs += 1
        secs.append(res[0]); steps_per_solve = res[1]
        elapsed = time.perf_counter() - t_start
        if elapsed + res[0] > REF_BUDGET_S:
            break
    dt, nsteps, kc, info, cores, sigma, note = res
    timed = secs[min(args.warmup, len(secs) - 1):]          # drop warm-up solves when there was time for them
    mean_s = float(np.mean(timed))
    value = steps_per_solve / mean_s
    sample = (f"oracle {'DLANSVD_IRL' if args.workload in IRL_P else ('ZLANSVD' if args.workload in COMPLEX else 'DLANSVD')} (C++/OpenMP port of "
              f"the reference; Fortran not buildable here), FULL solve of the {args.workload} problem to k={k} converged triplets incl. Ritz "
              f"vectors ({nsteps} Lanczos steps, converged={kc}, info={info}), OpenMP CSR/CSC APROD; {len(secs)} solve(s) run, "
              f"{len(timed)} timed (requested warmup={args.warmup} steps={args.steps}, bounded by {REF_BUDGET_S:.0f} s)" + ("; " + note if note else ""))
    line = {
        "impl": "reference", "metric": "lanczos_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * mean_s, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": wl_dtype(args.workload)[0], "data": "synthetic",
        "config": config_dict(args.workload, A, k, kmax, tol),
        "time_to_k_triplets_s": mean_s, "lanczos_steps_per_solve": int(steps_per_solve), "converged": int(kc), "info": int(info),
        "sigma_1": float(sigma[0]) if kc else None, "sigma_k": float(sigma[kc - 1]) if kc else None,
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_scipy and args.gpus == 1 and time.perf_counter() - t_start < SCIPY_START_BEFORE_S:   # (once per scaling sweep is enough)
        sc = scipy_svdp_sample(args.workload)
        if sc:
            line["scipy_svdp"] = sc
    print(json.dumps(line))


def run_scipy(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    print(json.dumps({"impl": "scipy", "scipy_svdp": scipy_svdp_sample(args.workload, budget_rows=10**9)}))


def config_dict(name, A, k, kmax, tol):
    if name in DENSE:
        return {"workload": f"BASELINE configs[2] '{name}': synthetic dense {A.shape[0]}x{A.shape[1]} f64 generated on the device "
                            f"(uniform(-1,1) bulk + planted rank-128 part, propack_b200/synth.py), k={k}, DLANSVD_IRL double dim={kmax} "
                            f"p={IRL_P[name]}, tol={tol:g}, CGS, ELR",
                "driver": "dlansvd_irl", "k": k, "kmax": kmax, "tol": tol, "nnz": int(A.nnz), "rows": int(A.shape[0]),
                "cols": int(A.shape[1]), "l2_policy": "inputs larger than L2 (no flush needed)"}
    irl = name in IRL_P
    cz = name in COMPLEX
    which = "configs[4]" if irl else ("configs[3]" if cz else "configs[1]")
    drv = f"DLANSVD_IRL double dim={kmax} p={IRL_P[name]}" if irl else (f"kmax={kmax}, ZLANSVD complex16" if cz else f"kmax={kmax}, DLANSVD double")
    kind = "power-law-row complex CSR" if cz else "random CSR"
    return {"workload": f"BASELINE {which} '{name}': synthetic {kind} {A.shape[0]}x{A.shape[1]}, nnz={A.nnz} "
                        f"(~{A.nnz / A.shape[0]:.1f}/row, longest row {int(np.diff(A.indptr).max())}), k={k}, {drv}, tol={tol:g}, CGS, ELR",
            "driver": ("zlansvd" if cz else "dlansvd_irl" if irl else "dlansvd"), "k": k, "kmax": kmax, "tol": tol, "nnz": int(A.nnz),
            "rows": int(A.shape[0]), "cols": int(A.shape[1]), "l2_policy": "inputs larger than L2 (no flush needed)"}


def session_solve(L, solver, name, k, kmax, tol, jobu=1, jobv=1):
    """One driver call on a solver session (DLANSVD, or DLANSVD_IRL for the restarted workloads). -> (sigma, k, info)"""
    from propack_b200 import _lib
    eps = np.finfo(np.float64).eps
    sigma = np.zeros(k); bnd = np.zeros(k)
    iopt = np.array([1, 1], dtype=np.int32)
    info = C.c_int(0)
    if name in IRL_P:
        dopt = np.array([np.sqrt(eps), eps ** 0.75, 0.0, 0.002])
        dim, neig = C.c_int(kmax), C.c_int(k)
        _lib.check(L.propack_b200_solver_lansvd_irl(C.c_int(solver), C.c_int(0), C.c_int(jobu), C.c_int(jobv), C.byref(dim),
                                                    C.c_int(IRL_P[name]), C.byref(neig), C.c_int(IRL_MAXITER),
                                                    sigma.ctypes.data_as(C.c_void_p), bnd.ctypes.data_as(C.c_void_p), C.c_double(tol),
                                                    dopt.ctypes.data_as(C.c_void_p), iopt.ctypes.data_as(C.c_void_p), C.byref(info)),
                   "lansvd_irl")
        return sigma[:neig.value], neig.value, info.value
    dopt = np.array([np.sqrt(eps), eps ** 0.75, 0.0])
    kk = C.c_int(k)
    _lib.check(L.propack_b200_solver_lansvd(C.c_int(solver), C.c_int(jobu), C.c_int(jobv), C.byref(kk), C.c_int(kmax),
                                            sigma.ctypes.data_as(C.c_void_p), bnd.ctypes.data_as(C.c_void_p), C.c_double(tol),
                                            dopt.ctypes.data_as(C.c_void_p), iopt.ctypes.data_as(C.c_void_p), C.byref(info)), "lansvd")
    return sigma[:kk.value], kk.value, info.value


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import propack_b200
    from propack_b200 import _lib, f77
    L = _lib.lib()
    _lib.check(L.propack_b200_init(), "init")
    stream = torch.cuda.current_stream()
    _lib.check(L.propack_b200_set_stream(C.c_void_p(stream.cuda_stream)), "set_stream")

    A, u0, k, kmax, tol = make_matrix(args.workload)
    m, n = A.shape
    op = make_operator(A)                      # matrix resident in HBM from here on
    lanmax = min(m + 1, n + 1, kmax)
    solver = _lib.check(L.propack_b200_solver_create(C.c_int(op.handle), C.c_int(lanmax + 1), C.c_int(lanmax)), "solver_create")
    eps = np.finfo(np.float64).eps

    def solve_resident():
        _lib.check(L.propack_b200_solver_set_start(C.c_int(solver), u0.ctypes.data_as(C.c_void_p)), "set_start")
        propack_b200.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        sigma, kc, info = session_solve(L, solver, args.workload, k, kmax, tol)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), propack_b200.counters(), sigma, kc, info

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        solve_resident()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    times, steps_total, launches, last = [], 0, 0, None
    for _ in range(args.steps):
        ms, ctr, sigma, kc, info = solve_resident()
        times.append(ms); steps_total += ctr["nsteps"]; launches += ctr["launches"]; last = (ctr, sigma, kc, info)
    barrier()
    clk = clocks.stop()
    total_ms = float(np.sum(times))
    if world > 1:
        t = torch.tensor([total_ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); total_ms = float(t.item())
        s = torch.tensor([float(steps_total)], device="cuda"); dist.all_reduce(s); steps_total = int(s.item())
    value = steps_total / (total_ms * 1e-3)

    # ---- roofline: reorthogonalisation GEMV pair, CUDA-event phase timers in one profiled solve ------------
    propack_b200.set_profile(True)
    _, pctr, _, _, _ = solve_resident()
    ph = propack_b200.phase_ms()
    propack_b200.set_profile(False)
    peak, peak_src = peaks()
    dts, w = wl_dtype(args.workload)
    rb = reorth_bytes(pctr, w)
    reorth_ms = ph["reorth"]["ms"]
    achieved = rb / (reorth_ms * 1e-3) / 1e9 if reorth_ms > 0 else 0.0
    spmv_bytes = (pctr["nopx"] / 2.0) * (op.bytes_per_product(False) + op.bytes_per_product(True) + float(w) * (m + n))
    spmv_gbs = spmv_bytes / (ph["aprod"]["ms"] * 1e-3) / 1e9 if ph["aprod"]["ms"] > 0 else 0.0
    # the roofline object describes the phase with the larger share of the solve (SpMV / dense GEMV as APROD, or the
    # reorthogonalisation GEMV pair); both are HBM streams, timed with CUDA events on the library stream
    tot_ms = sum(v["ms"] for v in ph.values())
    aprod_ms = ph["aprod"]["ms"]
    traffic_known = NCU_TRAFFIC_BYTES_PER_LAUNCH
    if aprod_ms >= reorth_ms:
        kname = ("dense APROD = gemv_n_kernel / gemv_t_kernel + gemv_t_finalize over A itself" if isinstance(A, DenseSpec) else
                 "spmv_sell_kernel (sliced jagged-ELL gather SpMV, fused axpy + norm; one launch per column panel of the gathered "
                 "vector; + spmv_long_kernel on power-law rows)")
        roofline = {"bound": "hbm", "kernel": kname, "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
                    "peak_source": peak_src, "frac_of_nominal_8000": spmv_gbs / 8000.0,
                    "traffic": traffic_known.get((args.workload, "spmv")), "traffic_note": NCU_TRAFFIC_NOTE.get((args.workload, "spmv")),
                    "algorithmic_bytes": spmv_bytes,
                    "algorithmic_bytes_per_launch": spmv_bytes / max(pctr["nopx"], 1), "kernel_ms_in_solve": aprod_ms,
                    "share_of_solve": aprod_ms / tot_ms,
                    "note": ("random-column gathers bound this kernel by the L1TEX wavefront rate, not HBM: the measured gather floor is "
                             "~51% of the HBM peak at 10 nnz/row (profiles/r01_spmv_lab.md)") if not isinstance(A, DenseSpec) else "",
                    "how": "CUDA-event phase timers on the library stream in one extra profiled solve of the same workload"}
    else:
        roofline = {"bound": "hbm", "kernel": "reorthogonalisation GEMV pair (gemv_t_tma_kernel [TMA-staged; gemv_t_kernel for short vectors] + gemv_t_finalize + gemv_n_kernel)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "frac_of_nominal_8000": achieved / 8000.0, "traffic": traffic_known.get((args.workload, "reorth")),
                    "traffic_note": NCU_TRAFFIC_NOTE.get((args.workload, "reorth")),
                    "algorithmic_bytes": rb, "kernel_ms_in_solve": reorth_ms, "share_of_solve": reorth_ms / tot_ms,
                    "how": "CUDA-event phase timers on the library stream in one extra profiled solve of the same workload"}
    # isolated kernels (device-resident synthetic operands, L2 flushed between launches)
    L.propack_b200_bench_reorth_d.argtypes = [C.c_long, C.c_int, C.c_int, C.c_int]
    iso = {}
    for l in (64, 256):   # (f64 micro-benchmark of the GEMV pair; the complex kernels stream the same bytes per element pair)
        t_ms = L.propack_b200_bench_reorth_d(m, l, 5, 1)
        iso[f"reorth_f64_L{m}_l{l}_gbs"] = 8.0 * m * (2 * l + 3) / (t_ms * 1e-3) / 1e9
    for adj in (0, 1):
        t_ms = L.propack_b200_bench_spmv(C.c_int(op.handle), C.c_int(adj), C.c_int(10), C.c_int(1))
        iso[f"spmv_{'t' if adj else 'n'}_gbs"] = (op.bytes_per_product(bool(adj)) + float(w) * (n if adj else m)) / (t_ms * 1e-3) / 1e9

    # ---- e2e: Fortran-ABI dlansvd_ with host buffers (matrix upload, start vector up, U/V/sigma down) --------
    dense = isinstance(A, DenseSpec)
    pfx = "z" if args.workload in COMPLEX else "d"
    tdt = torch.complex128 if pfx == "z" else torch.float64
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    if dense:   # the 65.5 GB operator has no host copy: e2e = generate it on the device + solve + results to the host
        rp = ci = va = np.zeros(0)
        u0p = pin(u0)
    else:
        rp = np.ascontiguousarray(A.indptr, dtype=np.int32); ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        va = np.ascontiguousarray(A.data)
        rp, ci, va, u0p = pin(rp), pin(ci), pin(va), pin(u0)
    # caller-owned result buffers of the Fortran interface, in pinned memory (allocated once, outside the timed region)
    Upin = torch.empty((k + 1, m), dtype=tdt).pin_memory().numpy().T
    Vpin = torch.empty((k + 1, n), dtype=tdt).pin_memory().numpy().T

    e2e_create, e2e_call = [], []

    def solve_e2e():
        t0 = time.perf_counter()
        if dense:
            op2 = make_operator(A)
            torch.cuda.synchronize()
        else:
            op2 = f77.Operator.__new__(f77.Operator)
            h = _lib.check(getattr(L, f"propack_b200_csr_create_{pfx}")(C.c_int(m), C.c_int(n), rp.ctypes.data_as(C.c_void_p),
                                                                        ci.ctypes.data_as(C.c_void_p), va.ctypes.data_as(C.c_void_p), C.c_int(0)),
                           "csr_create")
            op2.handle, op2._cb, op2.dtype, op2.pfx, op2.shape = h, None, np.dtype(A.dtype), pfx, (m, n)
            op2.iparm = np.array([h, 0], dtype=np.int32); op2.parm = np.zeros(2, dtype=A.dtype)
        e2e_create.append(time.perf_counter() - t0)
        t1 = time.perf_counter()
        if args.workload in IRL_P:
            r = f77.lansvd_irl(op2, k, kmax, p=IRL_P[args.workload], maxiter=IRL_MAXITER, tol=tol, u0=u0p, cgs=True, U=Upin, V=Vpin)
        else:
            r = f77.lansvd(op2, k, kmax, tol=tol, u0=u0p, cgs=True, U=Upin, V=Vpin)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_call.append(time.perf_counter() - t1)
        op2.close()
        return dt, r

    _lib.check(L.propack_b200_solver_destroy(C.c_int(solver)), "solver_destroy")  # free the resident bases first
    if dense:
        op.close()   # one 65.5 GB operator at a time
    e2e_t, e2e_steps = [], 0
    for i in range(1 + max(1, min(args.steps, 3))):
        propack_b200.reset_counters()
        dt, r = solve_e2e()
        if i >= 1:
            e2e_t.append(dt); e2e_steps += propack_b200.counters()["nsteps"]
    e2e_val = e2e_steps / float(np.sum(e2e_t))
    if world > 1:
        t = torch.tensor([float(np.sum(e2e_t))], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([float(e2e_steps)], device="cuda"); dist.all_reduce(s)
        e2e_val = float(s.item()) / float(t.item())
    h2d = rp.nbytes + ci.nbytes + va.nbytes + u0.nbytes + (A.table.nbytes if dense else 0)
    d2h = (m + n) * k * w + 2 * k * 8

    # ---- cpu baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores, note = cpu_arm(args.workload, A, u0, CPU_SAMPLE_STEPS)
        cpu = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
               "sample": f"oracle DLANBPRO (C++/OpenMP port; the Fortran reference cannot be compiled in this image), first "
                         f"{CPU_SAMPLE_STEPS} Lanczos steps of the same problem ({dt:.1f} s) -- a bounded sample: early steps reorthogonalise "
                         f"against fewer columns than the run average, so this overstates the CPU rate; the like-for-like number is the full "
                         f"CPU solve timed by `bench.py --impl reference`" + ("; " + note if note else "")}
    if rank == 0:
        ctr, sigma, kc, info = last
        line = {
            "metric": "lanczos_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": dts, "data": "synthetic",
            "config": config_dict(args.workload, A, k, kmax, tol), "parallelism": "single GPU",
            "time_to_k_triplets_s": total_ms / args.steps / 1e3, "lanczos_steps_per_solve": ctr["nsteps"],
            "converged": kc, "info": info, "sigma_1": float(sigma[0]) if kc else None, "sigma_k": float(sigma[-1]) if kc else None,
            "gpu_launches": int(launches), "host_syncs_per_solve": ctr["host_syncs"],
            "e2e": {"value": e2e_val, "unit": "steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "time_to_k_triplets_s": float(np.mean(e2e_t)),
                    "operator_create_s": float(np.mean(e2e_create[1:])) if len(e2e_create) > 1 else None,
                    "driver_call_s": float(np.mean(e2e_call[1:])) if len(e2e_call) > 1 else None,
                    "path": "propack_b200_csr_create_d + dlansvd[_irl]_ (Fortran ABI; host CSR, start vector and the caller's U,V result buffers in pinned memory; U,V,sigma copied back)"},
            "roofline": roofline,
            "reorth": {"achieved_gbs_in_solve": achieved, "frac": achieved / peak, "reorth_ms_in_solve": reorth_ms,
                       "algorithmic_bytes": rb},
            "spmv": {"achieved_gbs_in_solve": spmv_gbs, "frac": spmv_gbs / peak, "aprod_ms_in_solve": ph["aprod"]["ms"]},
            "isolated_kernels_gbs": iso,
            "phases_ms": {kname: v["ms"] for kname, v in ph.items()},
            "clocks": clk,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_ours_sharded(args):
    """N > 1: the same workload, row-sharded over the N GPUs of the node (strong scaling)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL / torchrun banners must not land on stdout (rank 0 prints exactly one JSON line): fd 1 -> fd 2 until the end
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import propack_b200
    from propack_b200 import _lib, dist as pdist
    L = _lib.lib()
    pdist.init_comm()
    stream = torch.cuda.current_stream()
    _lib.check(L.propack_b200_set_stream(C.c_void_p(stream.cuda_stream)), "set_stream")

    A, u0, k, kmax, tol = make_matrix(args.workload)     # every rank builds the same seeded matrix, keeps its shard
    m, n = A.shape
    lanmax = min(m + 1, n + 1, kmax)
    dense = isinstance(A, DenseSpec)
    make_sharded = ((lambda: pdist.ShardedDenseOperator((m, n), rank, world, synthetic=(DENSE_SEED, A.table))) if dense else
                    (lambda: pdist.ShardedOperator(A, rank, world)))
    op = make_sharded()
    sv = pdist.Solver(op, lanmax + 1, lanmax)

    def barrier():
        dist.barrier(); torch.cuda.synchronize()

    # device-resident arm: Ritz vectors are formed on the device (jobu=jobv='y' inside the library) but not copied out
    def solve_resident():
        sv.set_start(u0)
        propack_b200.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        sigma, kc, info = session_solve(L, sv.id, args.workload, k, kmax, tol)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), propack_b200.counters(), sigma, kc, info

    for _ in range(args.warmup):
        solve_resident()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    times, steps, launches, last = [], 0, 0, None
    for _ in range(args.steps):
        ms, ctr, sigma, kc, info = solve_resident()
        times.append(ms); steps += ctr["nsteps"]; launches += ctr["launches"]; last = (ctr, sigma, kc, info)
    barrier()
    clk = clocks.stop()
    t = torch.tensor([float(np.sum(times))], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = steps / (total_ms * 1e-3)          # every rank executes the same Lanczos steps: count them once
    # per-phase CUDA-event timers of one extra solve (rank 0's view; adds an event sync per phase, so it is slower than the
    # timed solves; "aprod" includes the wait for the other ranks' slices, "level1" the fused normalise + all-gather push)
    propack_b200.set_profile(True)
    pms, _roofline_ctr, _, _, _ = solve_resident()
    ph = {kname: v["ms"] for kname, v in propack_b200.phase_ms().items()}
    propack_b200.set_profile(False)
    _roofline_bytes = (op.bytes_per_product(False), op.bytes_per_product(True))

    # e2e: this rank's shard from pinned host memory -> device, solve, its slices of U, V and sigma back to the host
    pin = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).pin_memory().numpy()
    if dense:   # the operator has no host copy: e2e = generate this rank's rows on the device + solve + result slices to the host
        harr = ()
    else:
        rows, colt = pdist.shard_csr(A, world, rank)
        harr = (pin(rows.indptr, np.int32), pin(rows.indices, np.int32), pin(rows.data, np.float64),
                pin(colt.indptr, np.int32), pin(colt.indices, np.int32), pin(colt.data, np.float64))
    sv.close(); op.close()
    # caller-owned result buffers in pinned memory (allocated once, outside the timed region), as in the 1-GPU e2e leg: the
    # copy back of this rank's slices of U and V then runs at PCIe speed instead of through pageable memory
    rb, cb = pdist.shard_bounds(m, world, rank), pdist.shard_bounds(n, world, rank)
    try:
        Upin = torch.empty((k, max(rb[1] - rb[0], 1)), dtype=torch.float64).pin_memory().numpy().T
        Vpin = torch.empty((k, max(cb[1] - cb[0], 1)), dtype=torch.float64).pin_memory().numpy().T
    except Exception:
        Upin = Vpin = None
    e2e_t, e2e_steps = [], 0
    for i in range(0 if args.no_e2e else 1 + max(1, min(args.steps, 3))):
        barrier()
        t0 = time.perf_counter()
        if dense:
            op2 = make_sharded()
        else:
            h = _lib.check(L.propack_b200_csr_create_sharded_d(C.c_int(m), C.c_int(n), *[a.ctypes.data_as(C.c_void_p) for a in harr], C.c_int(0)),
                           "csr_create_sharded")
            op2 = pdist.ShardedOperator.__new__(pdist.ShardedOperator)
            op2.handle, op2.dtype, op2.pfx, op2.shape, op2.rank, op2.world = h, np.dtype(np.float64), "d", (m, n), rank, world
            op2.rows, op2.cols = pdist.shard_bounds(m, world, rank), pdist.shard_bounds(n, world, rank)
        sv2 = pdist.Solver(op2, lanmax + 1, lanmax)
        sv2.set_start(u0)
        propack_b200.reset_counters()
        if args.workload in IRL_P:
            r = sv2.lansvd_irl("L", kmax, IRL_P[args.workload], k, IRL_MAXITER, tol=tol, cgs=True, U_out=Upin, V_out=Vpin)
        else:
            r = sv2.lansvd(k, kmax, tol=tol, cgs=True, U_out=Upin, V_out=Vpin)
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        if i >= 1:
            e2e_t.append(dt); e2e_steps += propack_b200.counters()["nsteps"]
        sv2.close(); op2.close()
    te = torch.tensor([float(np.sum(e2e_t)) if e2e_t else 1.0], device="cuda"); dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = e2e_steps / float(te.item())
    rows_b, cols_b = pdist.shard_bounds(m, world, rank), pdist.shard_bounds(n, world, rank)
    h2d = sum(a.nbytes for a in harr) + (rows_b[1] - rows_b[0]) * 8 + (A.table.nbytes if dense else 0)
    d2h = ((rows_b[1] - rows_b[0]) + (cols_b[1] - cols_b[0])) * k * 8 + 2 * k * 8
    nar, nag, agb = C.c_longlong(0), C.c_longlong(0), C.c_double(0)
    L.propack_b200_comm_stats(C.byref(nar), C.byref(nag), C.byref(agb))
    # roofline of the dominant phase (the products): HBM bytes streamed by all ranks / rank 0's aprod time of the profiled solve, against
    # N x the measured HBM peak; and the NVLink floor of the all-gather that sits on the same critical path (sparse operator)
    peak1, peak_src = peaks()
    w = 8.0
    try:
        per_rank = float(_roofline_ctr["nopx"]) / 2.0 * (_roofline_bytes[0] + _roofline_bytes[1])
        tb = torch.tensor([per_rank], device="cuda", dtype=torch.float64); dist.all_reduce(tb)
        aprod_ms = ph.get("aprod", 0.0)
        ach = float(tb.item()) / (aprod_ms * 1e-3) / 1e9 if aprod_ms > 0 else 0.0
        gather_in = 0.0 if dense else w * (m + n) * (world - 1) / world       # bytes a rank must receive per Lanczos step (two products)
        roofline = {"bound": "hbm", "kernel": ("dense APROD = gemv_n_kernel / gemv_t_tma_kernel over the local row block" if dense else
                                               "spmv_sell_kernel, phase-split over the source ranks of the gathered vector (all ranks)"),
                    "achieved": ach, "peak": peak1 * world, "unit": "GB/s", "frac": ach / (peak1 * world), "peak_source": peak_src + f" x {world} GPUs",
                    "traffic": None, "algorithmic_bytes": float(tb.item()), "kernel_ms_in_solve": aprod_ms,
                    "share_of_solve": aprod_ms / max(sum(ph.values()), 1e-9),
                    "nvlink_floor": None if dense else {"bytes_in_per_rank_per_step": gather_in, "measured_peer_gbs_per_direction": 770.0,
                                                        "floor_ms_per_solve": gather_in / 770e9 * 1e3 * float(_roofline_ctr["nsteps"]),
                                                        "note": "the all-gather of the SpMV input is on the critical path of every product "
                                                                "(SpMV -> all-gather -> SpMV); aprod time includes waiting for the slices"},
                    "how": "rank 0's CUDA-event phase timers in one extra profiled solve; bytes summed over ranks"}
    except Exception as e:   # never let the bookkeeping take the bench line down
        roofline = {"bound": "hbm", "error": repr(e)[:200]}
    if rank == 0:
        ctr, sigma, kc, info = last
        line = {
            "metric": "lanczos_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": wl_dtype(args.workload)[0], "data": "synthetic",
            "config": config_dict(args.workload, A, k, kmax, tol),
            "parallelism": ((f"rows of the dense A and of U, and V-vectors, block-sharded over {world} GPUs; A x = NCCL all-gather of the "
                             f"n-vector + local GEMV, A^T u = local GEMV^T + NCCL all-reduce of the n coefficients; reorthogonalisation "
                             f"coefficients and norm partials all-reduced inside the producing kernels over NVLink peer memory") if dense else
                            (f"rows of A and U, and V-vectors, block-sharded over {world} GPUs; the SpMV input is pushed slice by slice over "
                             f"NVLink peer memory by a thin side-stream kernel (64 CTAs, ring order, one arrival flag per slice) while the "
                             f"phase-split SpMV consumes the slices that have landed; all-reduce of reorthogonalisation coefficients and "
                             f"norm partials fused into the producing kernels (NCCL for the un-staged products); collectives_total counts "
                             f"the NCCL calls that remain")),
            "time_to_k_triplets_s": total_ms / args.steps / 1e3, "lanczos_steps_per_solve": ctr["nsteps"], "converged": kc,
            "info": info, "sigma_1": float(sigma[0]) if kc else None, "sigma_k": float(sigma[-1]) if kc else None,
            "gpu_launches": int(launches), "host_syncs_per_solve": ctr["host_syncs"],
            "collectives_total": {"allreduce": nar.value, "allgather": nag.value, "allgather_gbytes": agb.value / 1e9},
            "e2e": {"value": e2e_val, "unit": "steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "time_to_k_triplets_s": float(np.mean(e2e_t)) if e2e_t else None,
                    "path": ("per rank: propack_b200_dense_create_synthetic_sharded_d (this rank's rows generated on the device) + solver "
                             "session + local U,V slices and sigma copied back; bytes are per rank") if dense else
                            ("per rank: propack_b200_csr_create_sharded_d (pinned host shard) + solver session + local U,V slices (into pinned "
                             "result buffers) and sigma copied back; bytes are per rank")},
            "roofline": roofline, "phases_ms_profiled_solve": ph, "profiled_solve_ms": pms, "clocks": clk,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    pdist.finalize_comm()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "scipy"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scipy", action="store_true", help="(reference arm) skip the bounded SciPy _svdp cross-check")
    ap.add_argument("--no-e2e", action="store_true", help="(sharded arm, experiments only) skip the end-to-end leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "scipy":
        run_scipy(args)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_ours_sharded(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
