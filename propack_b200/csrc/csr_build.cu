// propack_b200 -- device-side CSR analysis for operator registration (setup, not the Lanczos hot path).
//
// Reference context: PROPACK leaves the matrix to the user's APROD (dlansvd.F:20-33); here the built-in APROD needs
// A and A^T as CSR.  The transpose is the canonical one (row indices ascending inside every column -- identical to
// scipy's tocsc() / (A.T).tocsr() with sorted indices): a STABLE sort of the non-zeros by column index keeps the
// CSR order, which is ascending in the row index.  The sort is cub::DeviceRadixSort (library code, like cuBLAS for a
// plain GEMM); everything is integer work and bit-exact.  Also validates what the host loop used to validate:
// column range and sortedness inside each row.
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace pb {

namespace {

// row id of every non-zero (binary search in rp) + validation flags
__global__ void __launch_bounds__(kThreads)
expand_rows_kernel(int rows, int cols, long nnz, const int* __restrict__ rp, const int* __restrict__ ci, int* __restrict__ rowid,
                   int* __restrict__ pos, int* __restrict__ colcount, int* __restrict__ status) {
  for (long p = (long)blockIdx.x * kThreads + threadIdx.x; p < nnz; p += (long)gridDim.x * kThreads) {
    int lo = 0, hi = rows;  // largest r with rp[r] <= p
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(rp + mid) <= p) lo = mid; else hi = mid;
    }
    rowid[p] = lo;
    pos[p] = (int)p;
    const int c = ci[p];
    if (c < 0 || c >= cols) { atomicOr(status, 2); continue; }
    if (p > __ldg(rp + lo) && ci[p - 1] > c) atomicOr(status, 1);
    atomicAdd(colcount + c + 1, 1);
  }
}

template <class T>
__global__ void __launch_bounds__(kThreads)
permute_kernel(long nnz, const int* __restrict__ perm, const int* __restrict__ rowid, const T* __restrict__ va, int* __restrict__ tci,
               T* __restrict__ tva) {
  for (long q = (long)blockIdx.x * kThreads + threadIdx.x; q < nnz; q += (long)gridDim.x * kThreads) {
    const int p = perm[q];
    tci[q] = rowid[p];
    tva[q] = va[p];
  }
}

__global__ void rebase_kernel(long n, int* __restrict__ a, int base) {
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long)gridDim.x * kThreads) a[i] -= base;
}

}  // namespace

void k_rebase(Context& c, long n, int* a, int base) {
  if (n <= 0 || base == 0) return;
  rebase_kernel<<<c.grid_for(n, kThreads, 8), kThreads, 0, c.stream>>>(n, a, base);
  PB_LAUNCH_CHECK();
}

// CSR(A) -> CSR(A^T), all device pointers; trp has cols+1 entries.  Returns 0, or 1 (a row is not sorted) / 2 (column
// index out of range) / 3 (both) -- checked on the device, one int copied back.
template <class T>
int k_csr_transpose(Context& c, int rows, int cols, long nnz, const int* rp, const int* ci, const T* va, int* trp, int* tci, T* tva) {
  PB_CUDA(cudaMemsetAsync(trp, 0, sizeof(int) * ((size_t)cols + 1), c.stream));
  if (nnz <= 0) { c.sync(); return 0; }
  DeviceBuffer<int> rowid(nnz), pos(nnz), keys_out(nnz), perm(nnz), status(1);
  PB_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int), c.stream));
  const int grid = c.grid_for(nnz, kThreads, 8);
  expand_rows_kernel<<<grid, kThreads, 0, c.stream>>>(rows, cols, nnz, rp, ci, rowid.p, pos.p, trp, status.p);
  PB_LAUNCH_CHECK();
  int hstatus = 0;
  PB_CUDA(cudaMemcpyAsync(&hstatus, status.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  if (hstatus) return hstatus;
  // column pointers: inclusive scan of the counts (trp[0] = 0 stays)
  size_t tmp_scan = 0, tmp_sort = 0;
  int end_bit = 1;
  while ((1L << end_bit) < cols) ++end_bit;
  cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, trp, trp, cols + 1, c.stream);
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, ci, keys_out.p, pos.p, perm.p, (int)nnz, 0, end_bit, c.stream);
  DeviceBuffer<char> tmp(std::max(tmp_scan, tmp_sort) + 16);
  PB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_scan, trp, trp, cols + 1, c.stream));
  PB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_sort, ci, keys_out.p, pos.p, perm.p, (int)nnz, 0, end_bit, c.stream));
  permute_kernel<T><<<grid, kThreads, 0, c.stream>>>(nnz, perm.p, rowid.p, va, tci, tva);
  PB_LAUNCH_CHECK();
  c.sync();
  return 0;
}

#define PB_INST(T) template int k_csr_transpose<T>(Context&, int, int, long, const int*, const int*, const T*, int*, int*, T*);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
