"""numpy restatement of the sliced jagged-ELL construction (propack_b200/csrc/sell.cu: sell_lengths_kernel, sell_fill_kernel)
and of the column-panel split (csrc/csr_build.cu: k_csr_split_phases).

Test infrastructure (integer work must be bit-exact).  Rows keep their order; a slice is 32 consecutive rows; a row longer
than `long_thr` is left out (length byte 0xFF); inside a slice the stored entries are "k-major, compacted": first the 0-th
entry of every row that has one, in row order, then the 1-st entries, and so on.
"""
import numpy as np
import scipy.sparse as sp


def sell_ref(A, long_thr=64):
    indptr, indices, data = np.asarray(A.indptr, dtype=np.int64), np.asarray(A.indices), np.asarray(A.data)
    rows = A.shape[0]
    nslices = (rows + 31) // 32
    lens = np.diff(indptr)
    is_long = lens > long_thr
    len8 = np.where(is_long, 0xFF, lens).astype(np.uint8)
    eff = np.where(is_long, 0, lens)
    pad = np.zeros(nslices * 32, dtype=np.int64)
    pad[:rows] = eff
    count = pad.reshape(nslices, 32).sum(axis=1)
    joff = np.zeros(nslices + 1, dtype=np.int64)
    joff[1:] = np.cumsum(count)
    stored = int(joff[-1])
    ci = np.zeros(stored, dtype=np.int32)
    va = np.zeros(stored, dtype=data.dtype)
    for s in range(nslices):
        r0, r1 = s * 32, min(rows, s * 32 + 32)
        le = eff[r0:r1]
        if le.size == 0 or le.max() == 0:
            continue
        p = int(joff[s])
        for k in range(int(le.max())):
            act = np.nonzero(le > k)[0]
            src = indptr[r0 + act] + k
            ci[p:p + act.size] = indices[src]
            va[p:p + act.size] = data[src]
            p += act.size
    return dict(joff=joff, len8=len8, ci=ci, va=va)


def column_panels(A, G):
    """The G column blocks of a single-GPU operand, in panel order: block width = ceil(cols / G) rounded up to 32 columns,
    panel g = the block at ring distance g behind block 0 (block 0, then G-1, G-2, ..., 1), column ids stay global."""
    A = sp.csr_array(A)
    n = A.shape[1]
    per = (n + G - 1) // G
    ld = (per + 31) // 32 * 32
    out = []
    for g in range(G):
        owner = (G - g) % G
        lo, hi = min(owner * ld, n), min((owner + 1) * ld, n)
        keep = (A.indices >= lo) & (A.indices < hi)
        cs = np.concatenate([[0], np.cumsum(keep.astype(np.int64))])
        indptr = cs[A.indptr]
        out.append(sp.csr_array((A.data[keep], A.indices[keep], indptr), shape=A.shape))
    return out
