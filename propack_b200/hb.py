"""Harwell-Boeing files (``.rra`` real rectangular assembled, ``.rua``, ``.cua`` complex unsymmetric assembled ...).

The reference's example programs read their test matrices in this format (README:103-121: ``illc1850.rra``,
``mhd1280b.cua``; the readers lived in the ``Examples/`` directory that is absent from the checkout).  This is the
on-disk format on the input side of the hot path: the column-compressed arrays of the file are exactly the CSR arrays
of A^T, i.e. one of the two device operands.  Integer data round-trips bit-exactly; values are written with 17
significant digits, which round-trips IEEE doubles exactly.
"""
from __future__ import annotations

import re

import numpy as np

_FMT = re.compile(r"\(\s*(?:\d+P,?)?\s*(\d*)\s*([IEDFG])\s*(\d+)(?:\.(\d+))?(?:E\d+)?\s*\)", re.I)


def _parse_format(fmt: str):
    """'(16I5)' -> (16, 'I', 5);  '(1P,3E26.18)' -> (3, 'E', 26)."""
    m = _FMT.match(fmt.strip())
    if not m:
        raise ValueError(f"unsupported Fortran format {fmt!r}")
    return int(m.group(1) or 1), m.group(2).upper(), int(m.group(3))


def _read_fixed(lines, count, per_line, width, conv):
    out = []
    for ln in lines:
        ln = ln.rstrip("\n")
        for c in range(per_line):
            if len(out) == count:
                break
            field = ln[c * width:(c + 1) * width]
            if field.strip() == "":
                break
            out.append(conv(field))
    if len(out) != count:
        raise ValueError(f"expected {count} values, found {len(out)}")
    return out


def read_hb(path):
    """Read an assembled Harwell-Boeing matrix.  Returns ``scipy.sparse.csc_array`` (int32 indices, float64 / complex128)."""
    import scipy.sparse as sp
    with open(path, "r") as f:
        lines = f.readlines()
    # header counts are blank-separated in every file seen in the wild; tolerate columns that are slightly off
    counts = [int(t) for t in lines[1].split()] + [0] * 5
    totcrd, ptrcrd, indcrd, valcrd, rhscrd = counts[:5]
    mxtype = lines[2][:3].upper()
    nrow, ncol, nnz = (int(t) for t in lines[2][3:].split()[:3])
    ptrfmt, indfmt = lines[3][:16], lines[3][16:32]
    valfmt = lines[3][32:52]
    if mxtype[2] != "A":
        raise ValueError(f"only assembled matrices are supported, got {mxtype}")
    pos = 4 + (1 if rhscrd > 0 else 0)
    n, _, w = _parse_format(ptrfmt)
    colptr = np.array(_read_fixed(lines[pos:pos + ptrcrd], ncol + 1, n, w, int), dtype=np.int64) - 1
    pos += ptrcrd
    n, _, w = _parse_format(indfmt)
    rowind = np.array(_read_fixed(lines[pos:pos + indcrd], nnz, n, w, int), dtype=np.int64) - 1
    pos += indcrd
    if mxtype[0] == "P" or valcrd == 0:
        vals = np.ones(nnz)
    else:
        n, _, w = _parse_format(valfmt)
        conv = lambda s: float(s.replace("D", "E").replace("d", "e"))
        cplx = mxtype[0] == "C"
        raw = np.array(_read_fixed(lines[pos:pos + valcrd], nnz * (2 if cplx else 1), n, w, conv))
        vals = raw[0::2] + 1j * raw[1::2] if cplx else raw
    A = sp.csc_array((vals, rowind.astype(np.int32), colptr.astype(np.int32)), shape=(nrow, ncol))
    if mxtype[1] in "SH":   # symmetric / Hermitian storage: lower triangle only
        L = sp.tril(A, -1)
        A = sp.csc_array(A + (L.T.conj() if mxtype[1] == "H" else L.T))
    A.sort_indices()
    return A


def write_hb(path, A, title="", key="PROPACKB", mxtype=None):
    """Write ``A`` (any scipy sparse matrix) as an assembled Harwell-Boeing file with 1-based column-compressed arrays."""
    import scipy.sparse as sp
    A = sp.csc_array(A)
    A.sort_indices()
    nrow, ncol = A.shape
    cplx = np.iscomplexobj(A.data)
    if mxtype is None:
        mxtype = ("C" if cplx else "R") + ("U" if nrow == ncol else "R") + "A"
    ptr = A.indptr.astype(np.int64) + 1
    ind = A.indices.astype(np.int64) + 1
    vals = np.column_stack([A.data.real, A.data.imag]).ravel() if cplx else np.asarray(A.data, dtype=np.float64)
    pw = max(len(str(int(ptr.max()))) + 1, 8); pn = 80 // pw
    iw = max(len(str(int(max(ind.max(initial=1), 1)))) + 1, 8); inn = 80 // iw
    vw, vn = 26, 3

    def block(values, per_line, fmt):
        out = []
        for i in range(0, len(values), per_line):
            out.append("".join(fmt(v) for v in values[i:i + per_line]))
        return out

    pl = block(ptr, pn, lambda v: f"{int(v):>{pw}d}")
    il = block(ind, inn, lambda v: f"{int(v):>{iw}d}")
    vl = block(vals, vn, lambda v: f"{v:>{vw}.17E}")
    with open(path, "w") as f:
        f.write(f"{title[:72]:<72s}{key[:8]:<8s}\n")
        f.write(f"{len(pl) + len(il) + len(vl):>14d}{len(pl):>14d}{len(il):>14d}{len(vl):>14d}{0:>14d}\n")
        f.write(f"{mxtype:<3s}{'':11s}{nrow:>14d}{ncol:>14d}{A.nnz:>14d}{0:>14d}\n")
        f.write(f"{f'({pn}I{pw})':<16s}{f'({inn}I{iw})':<16s}{f'({vn}E{vw}.17)':<20s}{'':<20s}\n")
        for ln in pl + il + vl:
            f.write(ln + "\n")


def read_sigma_ascii(path):
    """Singular values as the reference's example programs store them (one value per line)."""
    return np.loadtxt(path, ndmin=1)


def compare(sigma, sigma_ref):
    """The check of the reference's ``compare`` program (README:143-157): max relative error of the singular values."""
    sigma, sigma_ref = np.asarray(sigma), np.asarray(sigma_ref)[:len(sigma)]
    return float(np.max(np.abs(sigma - sigma_ref) / np.abs(sigma_ref)))
