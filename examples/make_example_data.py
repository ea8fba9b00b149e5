#!/usr/bin/env python
"""Write the example matrices the reference ships in ``Examples/`` (README:103-121) into examples/data/, from the golden
fixtures of the test-suite: ``mhd1280b.cua`` (complex Harwell-Boeing), ``illc1850.coord`` / ``illc1850.diag`` (coordinate and
diagonal ASCII) and binary / dense variants of illc1850.  Host-only (no GPU needed)."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from propack_b200 import matio  # noqa: E402


def main(out=None):
    out = out or os.path.join(ROOT, "examples", "data")
    os.makedirs(out, exist_ok=True)
    g = np.load(os.path.join(ROOT, "tests", "golden", "propack_examples.npz"))

    def mat(prefix):
        shape = tuple(int(x) for x in g[f"{prefix}_shape"])
        return sp.coo_array((g[f"{prefix}_data"], (g[f"{prefix}_row"], g[f"{prefix}_col"])), shape=shape).tocsr()

    real, cplx = mat("illc1850"), mat("mhd1280b")
    matio.write_matrix(os.path.join(out, "mhd1280b.cua"), cplx, "hb")
    matio.write_matrix(os.path.join(out, "illc1850.coord"), real, "coord")
    # (illc1850 has 2262 non-empty diagonals: a diagonal-format copy would be tens of MB, so the diagonal example is a banded
    # stencil with a decaying main diagonal instead)
    nb = 4000
    band = sp.diags_array([np.full(nb - 1, -1.3), 1.0 + 30.0 * 0.9 ** np.arange(nb), np.full(nb - 1, -0.7),
                           np.full(nb - 40, 0.1)], offsets=[-1, 0, 1, 40], shape=(nb, nb))
    matio.write_matrix(os.path.join(out, "band4000.diag"), band, "diag")
    matio.write_matrix(os.path.join(out, "illc1850.cbin"), real, "coord-bin")
    matio.write_matrix(os.path.join(out, "illc1850.bin"), real.toarray(), "dense-bin")
    np.savetxt(os.path.join(out, "Sigma_mhd1280b.ascii"), g["mhd1280b_svd"][:200], fmt="%.17e")
    print("wrote", sorted(os.listdir(out)))
    return out


if __name__ == "__main__":
    main()
