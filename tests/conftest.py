import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running case")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def examples():
    """PROPACK example matrices + reference outputs (tools/make_golden.py)."""
    import scipy.sparse as sp
    g = np.load(os.path.join(GOLDEN, "propack_examples.npz"))

    def mat(prefix):
        shape = tuple(int(x) for x in g[f"{prefix}_shape"])
        A = sp.coo_array((g[f"{prefix}_data"], (g[f"{prefix}_row"], g[f"{prefix}_col"])), shape=shape).tocsr()
        A.sort_indices()
        return A

    return {"g": g, "illc1850": mat("illc1850"), "mhd1280b": mat("mhd1280b")}


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


DTYPES = [np.float32, np.float64, np.complex64, np.complex128]
TOL = {np.float32: 1e-4, np.float64: 1e-10, np.complex64: 1e-4, np.complex128: 1e-10}  # BASELINE.json parity bars


def rand_vec(rng, n, dtype):
    x = rng.standard_normal(n)
    if np.iscomplexobj(np.zeros(1, dtype=dtype)):
        x = x + 1j * rng.standard_normal(n)
    return x.astype(dtype)


def rand_sparse(rng, m, n, density, dtype, lengths=None):
    import scipy.sparse as sp
    A = sp.random_array((m, n), density=density, format="csr", rng=rng, data_sampler=rng.standard_normal)
    if np.iscomplexobj(np.zeros(1, dtype=dtype)):
        B = A.copy()
        B.data = rng.standard_normal(B.nnz)
        A = A + 1j * B
    A = sp.csr_array(A.astype(dtype))
    A.sort_indices()
    return A
