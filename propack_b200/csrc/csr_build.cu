// propack_b200 -- device-side CSR analysis for operator registration (setup, not the Lanczos hot path).
//
// Reference context: PROPACK leaves the matrix to the user's APROD (dlansvd.F:20-33); here the built-in APROD needs
// A and A^T as CSR.  The transpose is the canonical one (row indices ascending inside every column -- identical to
// scipy's tocsc() / (A.T).tocsr() with sorted indices): a STABLE sort of the non-zeros by column index keeps the
// CSR order, which is ascending in the row index.  The sort is cub::DeviceRadixSort (library code, like cuBLAS for a
// plain GEMM); everything is integer work and bit-exact.  Also validates what the host loop used to validate:
// column range and sortedness inside each row.
#include <cub/cub.cuh>

#include <vector>

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

// rowptr must start at 0, end at nnz and be non-decreasing (status bit 4)
__global__ void __launch_bounds__(kThreads)
check_rowptr_kernel(int rows, long nnz, const int* __restrict__ rp, int* __restrict__ status) {
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i <= rows; i += (long)gridDim.x * kThreads) {
    const int v = rp[i];
    bool bad = v < 0 || (long)v > nnz;
    if (i == 0 && v != 0) bad = true;
    if (i == rows && (long)v != nnz) bad = true;
    if (i < rows && rp[i + 1] < v) bad = true;
    if (bad) atomicOr(status, 4);
  }
}

__global__ void __launch_bounds__(kThreads)
count_long_kernel(int rows, const int* __restrict__ rp, int threshold, int* __restrict__ counter, int* __restrict__ list) {
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < rows; i += (long)gridDim.x * kThreads)
    if (rp[i + 1] - rp[i] > threshold) {
      const int at = atomicAdd(counter, 1);
      if (list != nullptr) list[at] = (int)i;
    }
}

// ---- row-sharded operands: split a local CSR by the source rank of its columns into G phases -------------------------
struct PhasePtrs { int* rp[8]; int* ci[8]; void* va[8]; };
__device__ inline int phase_of(int col, long ld, int P, int rank, int per) {
  const int owner = (int)(col / ld);
  int dist = rank - owner;
  if (dist < 0) dist += P;
  return dist / per;
}
// thread per row: validate (column range, ascending columns), count the row's entries per phase into rp[g][row+1]
__global__ void __launch_bounds__(kThreads)
phase_count_kernel(int rows, long width, const int* __restrict__ rp, const int* __restrict__ ci, long ld, int P, int rank, int per, int G,
                   PhasePtrs out, int* __restrict__ status) {
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < rows; i += (long)gridDim.x * kThreads) {
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int prevc = -1;
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
      const int c = ci[p];
      if (c < 0 || c >= width) { atomicOr(status, 2); continue; }
      if (c < prevc) atomicOr(status, 1);
      prevc = c;
      const int g = phase_of(c, ld, P, rank, per);
#pragma unroll
      for (int q = 0; q < 8; ++q) cnt[q] += (q == g);
    }
    for (int g = 0; g < G; ++g) out.rp[g][i + 1] = cnt[g];
  }
}
template <class T>
__global__ void __launch_bounds__(kThreads)
phase_scatter_kernel(int rows, const int* __restrict__ rp, const int* __restrict__ ci, const T* __restrict__ va, long ld, int P, int rank,
                     int per, int G, PhasePtrs out) {
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < rows; i += (long)gridDim.x * kThreads) {
    int pos[8];
    for (int g = 0; g < 8; ++g) pos[g] = g < G ? out.rp[g][i] : 0;
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
      const int c = ci[p];
      const int g = phase_of(c, ld, P, rank, per);
      int at = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) if (q == g) at = pos[q]++;
      out.ci[g][at] = c;
      static_cast<T*>(out.va[g])[at] = va[p];
    }
  }
}

}  // namespace

// status bit 4 when the row pointers are not a valid CSR extent array for nnz entries
int k_csr_check_rowptr(Context& c, int rows, long nnz, const int* rp) {
  DeviceBuffer<int> status(1);
  PB_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int), c.stream));
  check_rowptr_kernel<<<c.grid_for((long)rows + 1, kThreads, 8), kThreads, 0, c.stream>>>(rows, nnz, rp, status.p);
  PB_LAUNCH_CHECK();
  int h = 0;
  PB_CUDA(cudaMemcpyAsync(&h, status.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  return h;
}

// rows with more than `threshold` non-zeros, ascending, into `list` (allocated here); returns their number
int k_csr_long_rows(Context& c, int rows, const int* rp, int threshold, DeviceBuffer<int>& list) {
  if (rows <= 0) return 0;
  DeviceBuffer<int> counter(1);
  PB_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(int), c.stream));
  const int grid = c.grid_for(rows, kThreads, 8);
  count_long_kernel<<<grid, kThreads, 0, c.stream>>>(rows, rp, threshold, counter.p, nullptr);
  PB_LAUNCH_CHECK();
  int n = 0;
  PB_CUDA(cudaMemcpyAsync(&n, counter.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  if (n == 0) return 0;
  DeviceBuffer<int> unsorted(n);
  list.alloc(n);
  PB_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(int), c.stream));
  count_long_kernel<<<grid, kThreads, 0, c.stream>>>(rows, rp, threshold, counter.p, unsorted.p);
  PB_LAUNCH_CHECK();
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, unsorted.p, list.p, n, 0, 32, c.stream);
  DeviceBuffer<char> tmp(tmp_bytes + 16);
  PB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, unsorted.p, list.p, n, 0, 32, c.stream));
  c.sync();
  return n;
}

// Split the CSR (rows x width, device pointers, 0-based) by phase_of(column) into G CSR matrices with the same row count
// (columns stay ascending inside a row).  out_rp[g] must hold rows+1 ints; out_ci/out_va are allocated here.
// Returns the validation status (0 ok / bit 0 unsorted row / bit 1 column out of range); nnz_out[g] = entries of phase g.
template <class T>
int k_csr_split_phases(Context& c, int rows, long width, const int* rp, const int* ci, const T* va, long ld, int P, int rank, int G,
                       int* const* out_rp, DeviceBuffer<int>* const* out_ci, DeviceBuffer<T>* const* out_va, long* nnz_out) {
  if (G < 1 || G > 8 || P % G) throw std::runtime_error("propack_b200: bad phase count");
  const int per = P / G;
  PhasePtrs pp{};
  for (int g = 0; g < G; ++g) { pp.rp[g] = out_rp[g]; PB_CUDA(cudaMemsetAsync(out_rp[g], 0, sizeof(int), c.stream)); }
  DeviceBuffer<int> status(1);
  PB_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int), c.stream));
  const int grid = c.grid_for(std::max(rows, 1), kThreads, 8);
  if (rows > 0) {
    phase_count_kernel<<<grid, kThreads, 0, c.stream>>>(rows, width, rp, ci, ld, P, rank, per, G, pp, status.p);
    PB_LAUNCH_CHECK();
  }
  int hstatus = 0;
  PB_CUDA(cudaMemcpyAsync(&hstatus, status.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  size_t tmp_bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, out_rp[0], out_rp[0], rows + 1, c.stream);
  DeviceBuffer<char> tmp(tmp_bytes + 16);
  for (int g = 0; g < G; ++g) PB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, out_rp[g], out_rp[g], rows + 1, c.stream));
  std::vector<int> last(G, 0);
  for (int g = 0; g < G; ++g) PB_CUDA(cudaMemcpyAsync(&last[g], out_rp[g] + rows, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  if (hstatus) return hstatus;
  for (int g = 0; g < G; ++g) {
    nnz_out[g] = last[g];
    out_ci[g]->alloc((size_t)std::max(last[g], 1));
    out_va[g]->alloc((size_t)std::max(last[g], 1));
    pp.ci[g] = out_ci[g]->p; pp.va[g] = out_va[g]->p;
  }
  if (rows > 0) {
    phase_scatter_kernel<T><<<grid, kThreads, 0, c.stream>>>(rows, rp, ci, va, ld, P, rank, per, G, pp);
    PB_LAUNCH_CHECK();
  }
  c.sync();
  return 0;
}

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

#define PB_INST(T)                                                                                                        \
  template int k_csr_transpose<T>(Context&, int, int, long, const int*, const int*, const T*, int*, int*, T*);            \
  template int k_csr_split_phases<T>(Context&, int, long, const int*, const int*, const T*, long, int, int, int, int* const*, \
                                     DeviceBuffer<int>* const*, DeviceBuffer<T>* const*, long*);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
