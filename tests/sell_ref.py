"""numpy restatement of the SELL-32-sigma construction (propack_b200/csrc/sell.cu: sell_sort_kernel, sell_fill_kernel).

Test infrastructure (integer work must be bit-exact): rows are sorted by effective length (descending, stable) inside
windows of `sigma` consecutive rows; a row longer than `long_thr` has effective length 0 and is flagged 0x40000000;
slots past the last row carry -1; a slice is 32 consecutive slots, as wide as its first (= longest) row; entry k of slot
l of slice s lives at soff[s] + 32*k + l; padding is (column -1, value 0).
"""
import numpy as np


def sell_ref(A, sigma=1024, long_thr=64):
    indptr, indices, data = np.asarray(A.indptr, dtype=np.int64), np.asarray(A.indices), np.asarray(A.data)
    rows = A.shape[0]
    nwin = (rows + sigma - 1) // sigma
    nslices = nwin * (sigma // 32)
    lens = np.diff(indptr)
    is_long = lens > long_thr
    eff = np.where(is_long, 0, lens)
    eff_pad = np.full(nwin * sigma, -1, dtype=np.int64)
    eff_pad[:rows] = eff
    perm = np.full(nwin * sigma, -1, dtype=np.int32)
    len_sorted = np.zeros(nwin * sigma, dtype=np.int64)
    for w in range(nwin):
        seg = eff_pad[w * sigma:(w + 1) * sigma]
        order = np.argsort(-seg, kind="stable")
        rows_w = w * sigma + order
        valid = seg[order] >= 0
        p = np.where(valid, rows_w, -1).astype(np.int64)
        flag = np.zeros(sigma, dtype=np.int64)
        flag[valid] = np.where(is_long[rows_w[valid]], 0x40000000, 0)
        perm[w * sigma:(w + 1) * sigma] = np.where(valid, p | flag, -1).astype(np.int32)
        len_sorted[w * sigma:(w + 1) * sigma] = np.maximum(seg[order], 0)
    width = len_sorted[::32]
    soff = np.zeros(nslices + 1, dtype=np.int64)
    soff[1:] = np.cumsum(32 * width)
    padded = int(soff[-1])
    ci = np.full(padded, -1, dtype=np.int32)
    va = np.zeros(padded, dtype=data.dtype)
    for s in range(nslices):
        w = int(width[s])
        if w == 0:
            continue
        for l in range(32):
            r = int(perm[s * 32 + l])
            if r < 0 or (r & 0x40000000):
                continue
            beg, ln = int(indptr[r]), int(lens[r])
            idx = soff[s] + 32 * np.arange(ln) + l
            ci[idx] = indices[beg:beg + ln]
            va[idx] = data[beg:beg + ln]
    return dict(soff=soff, perm=perm, ci=ci, va=va)
