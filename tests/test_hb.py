"""CPU: Harwell-Boeing reader / writer (the on-disk format of the reference's example matrices, README:103-121)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN
from propack_b200 import hb


def _example(g, name):
    return sp.coo_array((g[f"{name}_data"], (g[f"{name}_row"], g[f"{name}_col"])), shape=tuple(g[f"{name}_shape"])).tocsc()


@pytest.mark.parametrize("name", ["illc1850", "mhd1280b"])
def test_round_trip_is_exact(tmp_path, name):
    g = np.load(os.path.join(GOLDEN, "propack_examples.npz"))
    A = _example(g, name)
    A.sort_indices()
    p = str(tmp_path / (name + (".cua" if np.iscomplexobj(A.data) else ".rra")))
    hb.write_hb(p, A, title=name, key=name[:8].upper())
    B = hb.read_hb(p)
    assert B.shape == A.shape and B.dtype == A.dtype
    assert np.array_equal(B.indptr, A.indptr) and np.array_equal(B.indices, A.indices)   # integer work: bit-exact
    assert np.array_equal(B.data, A.data)                                               # 17 significant digits round-trip doubles


def test_committed_illc1850_rra_matches_fixture_and_scipy_reader():
    g = np.load(os.path.join(GOLDEN, "propack_examples.npz"))
    A = _example(g, "illc1850")
    B = hb.read_hb(os.path.join(GOLDEN, "illc1850.rra"))
    assert B.shape == (1850, 712) and B.nnz == 8636
    assert abs(A - B).max() == 0.0
    s = hb.read_sigma_ascii(os.path.join(GOLDEN, "Sigma_illc1850.ascii"))
    assert s.size == 200 and hb.compare(g["illc1850_svd"][:10], s) == 0.0


def test_writer_is_readable_by_scipy(tmp_path):
    """Independent reader: scipy.io.hb_read (it only accepts real unsymmetric square files, so RUA here)."""
    import scipy.io
    rng = np.random.default_rng(0)
    A = sp.random_array((57, 57), density=0.1, format="csc", rng=rng, data_sampler=rng.standard_normal)
    p = str(tmp_path / "sq.rua")
    hb.write_hb(p, A, title="square", key="SQ")
    C = sp.csc_array(scipy.io.hb_read(p))
    assert abs(C - A).max() == 0.0 and abs(hb.read_hb(p) - A).max() == 0.0


def test_classic_fortran_formats(tmp_path):
    """Files written by Fortran programs: (16I5) pointers, (1P,3D26.18) values with D exponents, blank-padded lines."""
    txt = ("tiny test matrix                                                        TINY    \n"
           "             4             1             1             2             0\n"
           "RUA                          3             3             4             0\n"
           "(16I5)          (16I5)          (1P,3D26.18)                            \n"
           "    1    3    4    5\n"
           "    1    3    2    3\n"
           "  1.000000000000000000D+00 -2.500000000000000000D+00  3.000000000000000000D-01\n"
           "  4.000000000000000000D+02\n")
    p = tmp_path / "tiny.rua"
    p.write_text(txt)
    A = hb.read_hb(str(p)).toarray()
    assert np.array_equal(A, np.array([[1.0, 0, 0], [0, 0.3, 0], [-2.5, 0, 400.0]]))
