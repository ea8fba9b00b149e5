#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU:  torchrun --nproc-per-node N tests/dist_check.py

Runs the row-sharded DLANSVD / ZLANSVD / DLANSVD_IRL on N GPUs and checks, on rank 0, against dense LAPACK SVD and
the CPU oracle (test infrastructure): sigma to 1e-10 relative, residuals, orthogonality, and the same number of Lanczos
steps and restarts as the oracle -- the same bars as the single-GPU driver tests.  The last case is the BASELINE
configs[4] shape at DIST_CHECK_LARGE_ROWS rows (default 1M; 0 skips it).  Prints DIST_CHECK_OK on success.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from propack_b200 import dist as pdist  # noqa: E402
import propack_b200  # noqa: E402


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))))


def rand_sparse(rng, m, n, density, dtype):
    A = sp.random_array((m, n), density=density, format="csr", rng=rng, data_sampler=rng.standard_normal)
    if np.issubdtype(dtype, np.complexfloating):
        B = sp.random_array((m, n), density=density, format="csr", rng=rng, data_sampler=rng.standard_normal)
        A = A + 1j * B
    A = sp.csr_array(A.astype(dtype))
    A.sort_indices()
    return A


def run_case(name, A, k, kmax, rank, world, irl=None, tol=1e-12, dense_check=True, maxiter=200, golden=None):
    m, n = A.shape
    dtype = A.dtype
    u0 = np.random.default_rng(1).uniform(size=m).astype(dtype)
    dense_op = isinstance(A, np.ndarray)          # row-sharded dense operator (BASELINE configs[2] on N GPUs)
    op = pdist.ShardedDenseOperator(A, rank, world) if dense_op else pdist.ShardedOperator(A, rank, world)
    lanmax = min(m + 1, n + 1, kmax)
    sv = pdist.Solver(op, lanmax + 1, lanmax)
    sv.set_start(u0)
    propack_b200.reset_counters()
    if irl is None:
        r = sv.lansvd(k, kmax, tol=tol, cgs=True)
    else:
        r = sv.lansvd_irl("L", irl[0], irl[1], k, maxiter, tol=tol, cgs=True)
    ctr = propack_b200.counters()
    U = pdist.gather_rows(r["U"], m)
    V = pdist.gather_rows(r["V"], n)
    ok = True
    if rank == 0:
        S = r["sigma"]
        eps = np.finfo(dtype).eps
        e_sig = relerr(S, np.linalg.svd(A if dense_op else A.toarray(), compute_uv=False)[:k]) if dense_check else float("nan")
        res = float(np.max(np.linalg.norm(A @ V - U * S, axis=0)))
        orth = float(max(np.max(np.abs(U.conj().T @ U - np.eye(k))), np.max(np.abs(V.conj().T @ V - np.eye(k)))))
        # the CPU oracle on the same inputs (same start vector): parity of the algorithm, not only of the answer.
        # `golden`: the oracle's result for exactly this case, precomputed (tools/make_golden_atsize.py) so that the minutes
        # of CPU time are not spent on an N-GPU lease
        if golden is not None:
            ref = {"sigma": golden["sigma_oracle"], "k": int(golden["sigma_oracle"].size)}
            st = {"nsteps": int(golden["nsteps"]), "nrestart": int(golden["nrestart"])}
            assert int(golden["nnz"]) == A.nnz and int(golden["checksum_indices"]) == int(A.indices.astype(np.int64).sum())
        else:
            from oracle import oracle_py as O
            O.stats_reset()
            if irl is None:
                ref = O.lansvd(A, k, kmax, tol=tol, u0=u0, cgs=True, dtype=dtype, jobu=False, jobv=False)
            else:
                ref = O.lansvd_irl(A, k, irl[0], p=irl[1], which="L", maxiter=maxiter, tol=tol, u0=u0, cgs=True, dtype=dtype,
                                   jobu=False, jobv=False)
            st = O.stats()
        e_ref = relerr(S, ref["sigma"][:k]) if ref["k"] >= k else float("nan")
        same_path = ctr["nsteps"] == st["nsteps"] and ctr["nrestart"] == st["nrestart"]
        ok = (r["info"] == 0 and r["k"] == k and (np.isnan(e_sig) or e_sig < 1e-10) and res < 1e-8 * S[0]
              and orth < 200 * np.sqrt(eps) and ref["k"] >= k and e_ref < 1e-10 and same_path)
        print(f"[dist_check] {name}: world={world} info={r['info']} k={r['k']} sigma_relerr_dense={e_sig:.2e} "
              f"sigma_relerr_oracle={e_ref:.2e} max_residual={res:.2e} orth={orth:.2e} steps={ctr['nsteps']}/{st['nsteps']} "
              f"restarts={ctr['nrestart']}/{st['nrestart']} -> {'ok' if ok else 'FAIL'}", flush=True)
    sv.close(); op.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    return bool(flag.item())


def dense_planted_small(rng, m, n, dtype):
    """Dense test matrix with a decaying planted spectrum on top of noise (so that the leading triplets converge quickly)."""
    r = min(m, n, 24)
    X = rng.standard_normal((m, r)); Y = rng.standard_normal((n, r))
    A = (X * (2.0 ** -np.arange(r))) @ Y.T + 0.01 * rng.standard_normal((m, n))
    if np.issubdtype(dtype, np.complexfloating):
        A = A + 1j * ((rng.standard_normal((m, r)) * (2.0 ** -np.arange(r))) @ rng.standard_normal((n, r)).T)
    return np.asfortranarray(A.astype(dtype))


def c5_pattern(m, n, per=10):
    """BASELINE configs[4] pattern: exactly `per` uniform columns per row (bench.make_matrix 'c5')."""
    rng = np.random.default_rng(0)
    cols = rng.integers(0, n, size=(m, per), dtype=np.int32)
    cols.sort(axis=1)
    vals = rng.standard_normal(size=(m, per))
    A = sp.csr_array((vals.ravel(), cols.ravel(), np.arange(0, m * per + 1, per, dtype=np.int64)), shape=(m, n))
    A.sum_duplicates()
    A.sort_indices()
    return A


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = pdist.init_comm()
    rng = np.random.default_rng(0)
    ok = True
    ok &= run_case("d 3000x2000 lansvd", rand_sparse(rng, 3000, 2000, 0.005, np.float64), 8, 150, rank, world)
    ok &= run_case("d 500x4100 lansvd (fat)", rand_sparse(rng, 500, 4100, 0.01, np.float64), 6, 120, rank, world)
    ok &= run_case("d 4000x40 lansvd (ranks without columns)", rand_sparse(rng, 4000, 40, 0.2, np.float64), 5, 41, rank, world)
    ok &= run_case("z 1500x1200 zlansvd", rand_sparse(rng, 1500, 1200, 0.01, np.complex128), 6, 150, rank, world)
    ok &= run_case("d 3000x2000 lansvd_irl", rand_sparse(rng, 3000, 2000, 0.005, np.float64), 6, 60, rank, world, irl=(40, 20), tol=1e-10)
    # row-sharded dense operators: tall-skinny real with IRL (the config-3 shape in small), complex, and a fat one whose
    # trailing ranks own no rows
    ok &= run_case("d dense 5000x96 lansvd_irl", dense_planted_small(rng, 5000, 96, np.float64), 10, 40, rank, world, irl=(40, 20), tol=1e-10)
    ok &= run_case("z dense 1200x200 zlansvd", dense_planted_small(rng, 1200, 200, np.complex128), 6, 120, rank, world)
    ok &= run_case("d dense 40x900 lansvd (ranks without rows)", dense_planted_small(rng, 40, 900, np.float64), 5, 41, rank, world)
    # the config-5 shape of the north star (k=100, DLANSVD_IRL dim=300 p=200: restarts, the 101-column restart GEMM)
    rows_large = int(os.environ.get("DIST_CHECK_LARGE_ROWS", "1000000"))
    if rows_large > 0:
        gpath = os.path.join(ROOT, "tests", "golden", "c5_small_irl.npz")
        golden = dict(np.load(gpath)) if rows_large == 1_000_000 and os.path.exists(gpath) else None
        ok &= run_case(f"d {rows_large}x{rows_large} c5 pattern lansvd_irl dim=300 p=200" + (" (oracle result from tests/golden)" if golden else ""),
                       c5_pattern(rows_large, rows_large), 100, 300, rank, world, irl=(300, 200), tol=1e-10, dense_check=False, maxiter=50,
                       golden=golden)
    if rank == 0:
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
    pdist.finalize_comm()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
