"""Matrix files of the reference's example programs: Harwell-Boeing, coordinate, diagonal and dense, ASCII or binary.

README:103-121 of the reference: "The example programs can also read matrices stored in diagonal, coordinate or dense
formats (binary or ASCII) ... See Examples/example.F and Examples/matvec.F"; examples ``illc1850.coord`` and
``illc1850.diag`` shipped next to ``illc1850.rra``.  That ``Examples/`` directory is absent from the checkout, so the
layouts below are this package's definition of those formats (a Fortran list-directed ``READ(*,*)`` parses every ASCII
variant; the binary variants are Fortran ``FORM='UNFORMATTED'`` sequential files with 4-byte record markers, native
endianness -- the README's warning about word size and endianness applies).  Everything here is host-side I/O on the
input side of the hot path; the integer arrays it produces (CSR of A) are bit-exact against ``scipy.sparse``.

  coordinate ASCII   line 1: ``m n nnz``; then ``i j value`` (1-based; complex: ``i j re im``), any order, duplicates summed
  coordinate binary  record 1: int32 (m, n, nnz); record 2: int32 i(nnz); record 3: int32 j(nnz); record 4: values(nnz)
  diagonal ASCII     line 1: ``m n ndiag``; then per diagonal one line ``offset`` and ``len`` values (len = length of that
                     diagonal inside the m x n matrix; offset k means entries A(i, i+k)), free format
  diagonal binary    record 1: int32 (m, n, ndiag); record 2: int32 offsets(ndiag); then one record of values per diagonal
  dense ASCII        line 1: ``m n``; then the entries column by column (Fortran order), free format
  dense binary       record 1: int32 (m, n); record 2: the m*n values in Fortran order
  Harwell-Boeing     ``propack_b200.hb`` (``.rra``, ``.rua``, ``.cua`` ...)

Complex values are written as ``re im`` pairs (ASCII) or as interleaved (re, im) doubles (binary).
"""
from __future__ import annotations

import os

import numpy as np

from . import hb

FORMATS = ("hb", "coord", "coord-bin", "diag", "diag-bin", "dense", "dense-bin")


def _tokens(path):
    with open(path, "r") as f:
        for line in f:
            line = line.split("!")[0].replace(",", " ").replace("D", "E").replace("d", "e")
            for t in line.split():
                yield t


def _values(tok, count, cplx):
    if cplx:
        raw = np.array([float(next(tok)) for _ in range(2 * count)])
        return raw[0::2] + 1j * raw[1::2]
    return np.array([float(next(tok)) for _ in range(count)])


def _write_record(f, *arrays):
    payload = b"".join(np.ascontiguousarray(a).tobytes() for a in arrays)
    mark = np.array([len(payload)], dtype=np.int32).tobytes()
    f.write(mark + payload + mark)


def _read_record(f):
    head = f.read(4)
    if len(head) != 4:
        raise ValueError("unexpected end of unformatted file")
    n = int(np.frombuffer(head, dtype=np.int32)[0])
    payload = f.read(n)
    tail = f.read(4)
    if len(payload) != n or tail != head:
        raise ValueError("corrupt Fortran unformatted record (wrong endianness or word size?)")
    return payload


def _csr(A):
    import scipy.sparse as sp
    A = sp.csr_array(A)
    A.sum_duplicates()
    A.sort_indices()
    A.indptr = A.indptr.astype(np.int32)
    A.indices = A.indices.astype(np.int32)
    return A


# ---- coordinate ------------------------------------------------------------------------------------------------------
def read_coord(path, complex_values=False, binary=False):
    import scipy.sparse as sp
    if binary:
        with open(path, "rb") as f:
            m, n, nnz = (int(v) for v in np.frombuffer(_read_record(f), dtype=np.int32)[:3])
            i = np.frombuffer(_read_record(f), dtype=np.int32)
            j = np.frombuffer(_read_record(f), dtype=np.int32)
            raw = np.frombuffer(_read_record(f), dtype=np.float64)
        v = raw[0::2] + 1j * raw[1::2] if raw.size == 2 * nnz else raw
    else:
        tok = _tokens(path)
        m, n, nnz = int(next(tok)), int(next(tok)), int(next(tok))
        i = np.empty(nnz, dtype=np.int64); j = np.empty(nnz, dtype=np.int64)
        v = np.empty(nnz, dtype=np.complex128 if complex_values else np.float64)
        for p in range(nnz):
            i[p], j[p] = int(next(tok)), int(next(tok))
            v[p] = complex(float(next(tok)), float(next(tok))) if complex_values else float(next(tok))
    if nnz and (i.min() < 1 or j.min() < 1 or i.max() > m or j.max() > n):
        raise ValueError("coordinate file: index out of range (indices are 1-based)")
    return _csr(sp.coo_array((v, (np.asarray(i) - 1, np.asarray(j) - 1)), shape=(m, n)))


def write_coord(path, A, binary=False):
    import scipy.sparse as sp
    C_ = sp.coo_array(_csr(A))
    cplx = np.iscomplexobj(C_.data)
    if binary:
        with open(path, "wb") as f:
            _write_record(f, np.array([C_.shape[0], C_.shape[1], C_.nnz], dtype=np.int32))
            _write_record(f, (C_.row + 1).astype(np.int32))
            _write_record(f, (C_.col + 1).astype(np.int32))
            _write_record(f, C_.data.astype(np.complex128 if cplx else np.float64))
        return
    with open(path, "w") as f:
        f.write(f"{C_.shape[0]} {C_.shape[1]} {C_.nnz}\n")
        for r, c, v in zip(C_.row, C_.col, C_.data):
            f.write(f"{r + 1} {c + 1} {v.real:.17E} {v.imag:.17E}\n" if cplx else f"{r + 1} {c + 1} {v:.17E}\n")


# ---- diagonal --------------------------------------------------------------------------------------------------------
def _diag_len(m, n, k):
    return max(0, min(m, n - k) if k >= 0 else min(m + k, n))


def read_diag(path, complex_values=False, binary=False):
    import scipy.sparse as sp
    rows, cols, vals = [], [], []
    if binary:
        with open(path, "rb") as f:
            m, n, nd = (int(v) for v in np.frombuffer(_read_record(f), dtype=np.int32)[:3])
            offs = np.frombuffer(_read_record(f), dtype=np.int32)
            for k in offs:
                ln = _diag_len(m, n, int(k))
                raw = np.frombuffer(_read_record(f), dtype=np.float64)
                d = raw[0::2] + 1j * raw[1::2] if raw.size == 2 * ln and ln > 0 else raw
                i0 = max(0, -int(k))
                rows.append(np.arange(i0, i0 + ln)); cols.append(np.arange(i0 + int(k), i0 + int(k) + ln)); vals.append(d)
    else:
        tok = _tokens(path)
        m, n, nd = int(next(tok)), int(next(tok)), int(next(tok))
        for _ in range(nd):
            k = int(next(tok))
            ln = _diag_len(m, n, k)
            d = _values(tok, ln, complex_values)
            i0 = max(0, -k)
            rows.append(np.arange(i0, i0 + ln)); cols.append(np.arange(i0 + k, i0 + k + ln)); vals.append(d)
    if not rows:
        return _csr(sp.coo_array((m, n)))
    r, c, v = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    keep = v != 0
    return _csr(sp.coo_array((v[keep], (r[keep], c[keep])), shape=(m, n)))


def write_diag(path, A, binary=False):
    import scipy.sparse as sp
    D = sp.dia_array(_csr(A))
    m, n = D.shape
    cplx = np.iscomplexobj(D.data)
    order = np.argsort(D.offsets)
    offs = D.offsets[order]
    diags = []
    for idx in order:
        k = int(D.offsets[idx])
        ln = _diag_len(m, n, k)
        j0 = max(0, k)                                 # scipy stores diagonal k with the column index as position
        diags.append(np.asarray(D.data[idx, j0:j0 + ln]))
    if binary:
        with open(path, "wb") as f:
            _write_record(f, np.array([m, n, len(offs)], dtype=np.int32))
            _write_record(f, offs.astype(np.int32))
            for d in diags:
                _write_record(f, d.astype(np.complex128 if cplx else np.float64))
        return
    with open(path, "w") as f:
        f.write(f"{m} {n} {len(offs)}\n")
        for k, d in zip(offs, diags):
            f.write(f"{int(k)}\n")
            for v in d:
                f.write(f"{v.real:.17E} {v.imag:.17E}\n" if cplx else f"{v:.17E}\n")


# ---- dense -----------------------------------------------------------------------------------------------------------
def read_dense(path, complex_values=False, binary=False):
    if binary:
        with open(path, "rb") as f:
            m, n = (int(v) for v in np.frombuffer(_read_record(f), dtype=np.int32)[:2])
            raw = np.frombuffer(_read_record(f), dtype=np.float64)
        v = raw[0::2] + 1j * raw[1::2] if raw.size == 2 * m * n else raw
    else:
        tok = _tokens(path)
        m, n = int(next(tok)), int(next(tok))
        v = _values(tok, m * n, complex_values)
    return np.asfortranarray(np.asarray(v).reshape((m, n), order="F"))


def write_dense(path, A, binary=False):
    A = np.asfortranarray(A)
    cplx = np.iscomplexobj(A)
    flat = A.ravel(order="F")
    if binary:
        with open(path, "wb") as f:
            _write_record(f, np.array(A.shape, dtype=np.int32))
            _write_record(f, flat.astype(np.complex128 if cplx else np.float64))
        return
    with open(path, "w") as f:
        f.write(f"{A.shape[0]} {A.shape[1]}\n")
        for v in flat:
            f.write(f"{v.real:.17E} {v.imag:.17E}\n" if cplx else f"{v:.17E}\n")


# ---- dispatch --------------------------------------------------------------------------------------------------------
def guess_format(path):
    ext = os.path.splitext(path)[1].lower()
    return {".coord": "coord", ".diag": "diag", ".dense": "dense", ".cbin": "coord-bin", ".dbin": "diag-bin", ".bin": "dense-bin"}.get(ext, "hb")


def read_matrix(path, fmt=None, complex_values=False):
    """Read a matrix file of any supported format.  Sparse formats return a CSR ``scipy.sparse`` array with sorted int32
    indices (the arrays ``propack_b200_csr_create_*`` takes), dense formats a Fortran-ordered ndarray."""
    fmt = fmt or guess_format(path)
    if fmt == "hb":
        return _csr(hb.read_hb(path))
    kind, _, b = fmt.partition("-")
    fn = {"coord": read_coord, "diag": read_diag, "dense": read_dense}[kind]
    return fn(path, complex_values=complex_values, binary=(b == "bin"))


def write_matrix(path, A, fmt=None):
    fmt = fmt or guess_format(path)
    if fmt == "hb":
        return hb.write_hb(path, A)
    kind, _, b = fmt.partition("-")
    return {"coord": write_coord, "diag": write_diag, "dense": write_dense}[kind](path, A, binary=(b == "bin"))
