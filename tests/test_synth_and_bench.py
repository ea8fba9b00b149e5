"""CPU: workload definitions used by bench.py and the parity tests (no device needed)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import synth_ref  # noqa: E402
from propack_b200 import synth  # noqa: E402


def test_planted_table_is_the_sum_of_signed_coefficients():
    c = synth.planted_coefficients(2000, 300)
    T = synth.planted_table(c)
    assert T.shape == (16, 256) and np.all(np.diff(c) < 0)
    for g in (0, 7, 15):
        for b in (0, 1, 0b10101010, 255):
            want = 0.0
            for t in range(8):
                want = want + (-c[8 * g + t] if (b >> t) & 1 else c[8 * g + t])
            assert T[g, b] == want   # same summation order => bit-identical


def test_dense_replica_rows_subset_and_spectrum():
    m, n = 1500, 120
    T = synth.planted_table(synth.planted_coefficients(m, n))
    A = synth_ref.dense_planted(m, n, 5, T)
    rows = np.array([0, 7, 255, 256, 1499])
    assert np.array_equal(synth_ref.dense_planted(m, n, 5, T, rows=rows), A[rows])      # counter-based: any row on its own
    assert not np.array_equal(A, synth_ref.dense_planted(m, n, 6, T))                     # the seed matters
    s = np.linalg.svd(A, compute_uv=False)
    edge = (np.sqrt(m) + np.sqrt(n)) / np.sqrt(3.0)
    # planted part: ~100x the noise edge at the top, decaying by 2^(-1/20) per value; the tail sits at the noise edge
    assert 60 * edge < s[0] < 140 * edge
    assert np.all(s[:60] > 5 * edge) and s[-1] < 1.2 * edge


@pytest.mark.parametrize("name", ["c2-tiny", "c4-small"])
def test_bench_workload_generators_are_deterministic(name):
    import bench
    if name == "c4-small":
        bench.WORKLOADS["c4-test"] = (3000, 3000, "powerlaw", 8, 100, 1e-10); bench.COMPLEX.add("c4-test"); name = "c4-test"
    A1, u1, k, kmax, tol = bench.make_matrix(name)
    A2, u2, _, _, _ = bench.make_matrix(name)
    assert A1.has_sorted_indices and A1.indices.dtype == np.int32 or A1.indices.max() < 2**31
    assert np.array_equal(A1.indptr, A2.indptr) and np.array_equal(A1.indices, A2.indices) and np.array_equal(A1.data, A2.data)
    assert np.array_equal(u1, u2) and u1.dtype == A1.dtype
    cfg = bench.config_dict(name, A1, k, kmax, tol)
    assert cfg["nnz"] == A1.nnz and cfg["rows"] == A1.shape[0] and "workload" in cfg
    assert bench.wl_dtype(name)[0] == ("c128" if np.iscomplexobj(A1.data) else "f64")


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU oracle's FULL solve of the workload) on the tiny replica: one JSON line with the keys the
    bench contract names, a converged solve, and SciPy's `_svdp` cross-check agreeing with the oracle on sigma_1."""
    import json
    import subprocess
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2-tiny", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["info"] == 0 and d["converged"] == 10 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    sc = d.get("scipy_svdp")
    if sc and "sigma_1" in sc:
        assert abs(sc["sigma_1"] - d["sigma_1"]) < 1e-10 * d["sigma_1"]


def test_at_size_fixture_is_consistent():
    """tests/golden/c5_small_irl.npz (tools/make_golden_atsize.py): the oracle and SciPy's PROPACK translation agree on all 100 values of
    the config-5 pattern at 1M rows, and the fixture describes the matrix bench.make_matrix('c5-small') builds."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "c5_small_irl.npz"))
    so, ss = g["sigma_oracle"], g["sigma_scipy_svdp"]
    assert so.size == 100 and np.all(np.diff(so) <= 0) and int(g["k"]) == 100 and int(g["dim"]) == 300 and int(g["p"]) == 200
    assert np.max(np.abs(so - ss) / so) < 1e-12
    assert tuple(g["shape"]) == (1_000_000, 1_000_000) and int(g["nsteps"]) == 1100 and int(g["nrestart"]) == 4
