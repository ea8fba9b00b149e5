"""CPU: the matrix-file formats of the reference's example programs (README:103-121): coordinate, diagonal and dense, ASCII and
binary, and Harwell-Boeing -- every format round-trips the PROPACK example matrices to the bit, and the CSR arrays handed to the
library equal scipy's canonical (sorted, int32) ones exactly."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from propack_b200 import matio


@pytest.mark.parametrize("name", ["illc1850", "mhd1280b"])
@pytest.mark.parametrize("fmt", ["hb", "coord", "coord-bin", "diag", "diag-bin"])
def test_sparse_formats_round_trip_example_matrices(examples, tmp_path, name, fmt):
    A = examples[name]
    if fmt.startswith("diag") and name == "illc1850":
        A = sp.csr_array(A[:400, :300])          # keep the (very many) diagonals of the ASCII file small
    cz = np.iscomplexobj(A.data)
    p = os.path.join(tmp_path, "m." + fmt)
    matio.write_matrix(p, A, fmt)
    B = matio.read_matrix(p, fmt, complex_values=cz)
    ref = sp.csr_array(A); ref.sum_duplicates(); ref.sort_indices()
    assert B.shape == ref.shape and B.indices.dtype == np.int32 and B.indptr.dtype == np.int32
    assert np.array_equal(B.indptr, ref.indptr) and np.array_equal(B.indices, ref.indices) and np.array_equal(B.data, ref.data)


@pytest.mark.parametrize("cz", [False, True])
@pytest.mark.parametrize("fmt", ["dense", "dense-bin"])
def test_dense_formats_round_trip(tmp_path, cz, fmt):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((19, 7)) + (1j * rng.standard_normal((19, 7)) if cz else 0)
    p = os.path.join(tmp_path, "d." + fmt)
    matio.write_matrix(p, A, fmt)
    B = matio.read_matrix(p, fmt, complex_values=cz)
    assert B.flags.f_contiguous and np.array_equal(B, A)


def test_coordinate_ascii_accepts_fortran_style_input(tmp_path):
    """Free-format input as a Fortran list-directed READ would take it: D exponents, commas, unsorted entries, duplicates summed."""
    p = os.path.join(tmp_path, "t.coord")
    with open(p, "w") as f:
        f.write("3 4 5\n3 1 1.5D0\n1, 4, -2.0e0\n1 2 4\n3 1 0.5\n2 2 1d-1\n")
    A = matio.read_matrix(p)
    want = np.array([[0, 4, 0, -2.0], [0, 0.1, 0, 0], [2.0, 0, 0, 0]])
    assert np.array_equal(A.toarray(), want)
    with open(p, "w") as f:
        f.write("2 2 1\n3 1 1.0\n")
    with pytest.raises(ValueError):
        matio.read_matrix(p)


def test_make_example_data(tmp_path):
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_example_data", os.path.join(root, "examples", "make_example_data.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.main(os.path.join(tmp_path, "data"))
    Z = matio.read_matrix(os.path.join(out, "mhd1280b.cua"))
    assert Z.shape == (1280, 1280) and Z.nnz == 22778 and np.iscomplexobj(Z.data)
    C1 = matio.read_matrix(os.path.join(out, "illc1850.coord"))
    B = matio.read_matrix(os.path.join(out, "band4000.diag"))
    assert B.shape == (4000, 4000) and B.nnz == 3 * 4000 - 2 + 3960
    C3 = matio.read_matrix(os.path.join(out, "illc1850.cbin"))
    D = matio.read_matrix(os.path.join(out, "illc1850.bin"))
    assert C1.nnz == 8636 and (C1 != C3).nnz == 0 and np.array_equal(D, C1.toarray())
