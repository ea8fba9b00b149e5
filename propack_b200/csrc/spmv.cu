// propack_b200 -- CSR SpMV with fused Lanczos epilogue.
//
// Reference: the user-supplied APROD (contract dlansvd.F:20-33), called at dlanbpro.F:288
// ('t': v = A^T u) and :420 ('n': u = A v), each followed by pdaxpy + pdnrm2 (:295-296, :423-424).
// Here one launch computes
//        y = op(A) x + coef * prev          and publishes ||y||_2,
// so the axpy and the norm cost no extra pass over y.  A^T x runs on a precomputed CSR copy of
// A^T (= CSC of A), so both directions are the same gather-style kernel: no atomics, no
// transpose-time scatter, bit-reproducible.
//
// Kernel organisation ("row blocks", computed once per matrix by csr_row_blocks()):
//   rows are cut into consecutive blocks of <= kSpmvRows rows and <= NB non-zeros (NB = what fits
//   the CTA's shared-memory product buffer); a row longer than NB is a block of its own.
//   phase 1  the CTA streams the block's (ci, va) slice with fully coalesced, 4-way unrolled loads,
//            gathers x through the read-only path and parks the products in shared memory --
//            every lane has 4 independent (ci -> x) chains in flight regardless of row lengths;
//   phase 2  rows are reduced out of shared memory by sub-warps whose width follows the block's
//            mean row length (4 lanes for ~10 nnz/row, a warp for power-law rows), then the
//            epilogue  y = sum + coef*prev  and the ||y||^2 partial.
//   long rows (> NB nnz) are swept by the whole CTA.
// (ci, va) are touched once per launch: evict-first loads keep them from displacing x in L2
// (x is 8 MB at 1M columns, 80 MB at 10M; L2 is 126 MB).
// HBM-bound: algorithmic bytes nnz*(w+4) + 4(rows+1) + w*cols + w*rows (+ w*rows for prev); the
// x gathers (one 32-byte sector per non-zero for random columns) ride on L2 bandwidth.
#include <algorithm>

#include "kernels.cuh"

namespace pb {

namespace {

template <class T> __device__ inline T ldcs_(const T* p) { return *p; }
template <> __device__ inline float ldcs_<float>(const float* p) { return __ldcs(p); }
template <> __device__ inline double ldcs_<double>(const double* p) { return __ldcs(p); }
template <> __device__ inline cplx<float> ldcs_<cplx<float>>(const cplx<float>* p) {
  float2 f = __ldcs(reinterpret_cast<const float2*>(p));
  return cplx<float>(f.x, f.y);
}
template <> __device__ inline cplx<double> ldcs_<cplx<double>>(const cplx<double>* p) {
  double2 f = __ldcs(reinterpret_cast<const double2*>(p));
  return cplx<double>(f.x, f.y);
}
template <class T> __device__ inline T ldg_(const T* p) { return *p; }
template <> __device__ inline float ldg_<float>(const float* p) { return __ldg(p); }
template <> __device__ inline double ldg_<double>(const double* p) { return __ldg(p); }
template <> __device__ inline cplx<float> ldg_<cplx<float>>(const cplx<float>* p) {
  float2 f = __ldg(reinterpret_cast<const float2*>(p));
  return cplx<float>(f.x, f.y);
}
template <> __device__ inline cplx<double> ldg_<cplx<double>>(const cplx<double>* p) {
  double2 f = __ldg(reinterpret_cast<const double2*>(p));
  return cplx<double>(f.x, f.y);
}

template <class T, bool CONJ> __device__ inline T mul_(T a, T x) {
  T acc = zero_<T>();
  if (CONJ) fma_conj(acc, a, x);
  else fma_(acc, a, x);
  return acc;
}

template <class T, bool CONJ>
__global__ void __launch_bounds__(kThreads)
spmv_kernel(CsrDevice<T> A, const T* __restrict__ x, T* __restrict__ y, real_t<T> coef, const T* __restrict__ prev,
            ReduceWs ws, int want_norm) {
  constexpr int NB = spmv_block_nnz<T>();
  __shared__ T prod[NB];
  __shared__ int srp[kSpmvRows + 1];
  __shared__ double red[32];
  __shared__ T redT[32];
  const int tid = threadIdx.x;
  double nrm = 0.0;

  for (int b = blockIdx.x; b < A.n_blocks; b += gridDim.x) {
    const int r0 = __ldg(A.block_row + b), r1 = __ldg(A.block_row + b + 1);
    const int nrows = r1 - r0;
    for (int i = tid; i <= nrows; i += kThreads) srp[i] = __ldg(A.rp + r0 + i);
    __syncthreads();
    const int p0 = srp[0];
    const int nnzb = srp[nrows] - p0;

    if (nnzb > NB) {
      // ---- a single long row: CTA-wide strided sweep -------------------------------------------
      T acc = zero_<T>();
      for (int p = p0 + tid; p < p0 + nnzb; p += kThreads) {
        const T a = ldcs_(A.va + p);
        const T xv = ldg_(x + __ldcs(A.ci + p));
        if (CONJ) fma_conj(acc, a, xv);
        else fma_(acc, a, xv);
      }
      acc = block_sum(acc, redT);
      if (tid == 0) {
        if (prev != nullptr) acc = acc + coef * prev[r0];
        y[r0] = acc;
        nrm += (double)abs2_(acc);
      }
      __syncthreads();
      continue;
    }

    // ---- phase 1: products into shared memory ------------------------------------------------------
    {
      const int* cip = A.ci + p0;
      const T* vap = A.va + p0;
      int i = tid;
      for (; i + 3 * kThreads < nnzb; i += 4 * kThreads) {
        const int c0 = __ldcs(cip + i), c1 = __ldcs(cip + i + kThreads), c2 = __ldcs(cip + i + 2 * kThreads),
                  c3 = __ldcs(cip + i + 3 * kThreads);
        const T a0 = ldcs_(vap + i), a1 = ldcs_(vap + i + kThreads), a2 = ldcs_(vap + i + 2 * kThreads),
                a3 = ldcs_(vap + i + 3 * kThreads);
        const T x0 = ldg_(x + c0), x1 = ldg_(x + c1), x2 = ldg_(x + c2), x3 = ldg_(x + c3);
        prod[i] = mul_<T, CONJ>(a0, x0);
        prod[i + kThreads] = mul_<T, CONJ>(a1, x1);
        prod[i + 2 * kThreads] = mul_<T, CONJ>(a2, x2);
        prod[i + 3 * kThreads] = mul_<T, CONJ>(a3, x3);
      }
      for (; i < nnzb; i += kThreads) prod[i] = mul_<T, CONJ>(ldcs_(vap + i), ldg_(x + __ldcs(cip + i)));
    }
    __syncthreads();

    // ---- phase 2: per-row reduction by sub-warps of `lpr` lanes ----------------------------------------
    {
      const int mean = (nnzb + nrows - 1) / max(nrows, 1);
      int lpr = 4;
      while (lpr < 32 && lpr * 4 < mean) lpr <<= 1;
      const int gpc = kThreads / lpr;             // rows per pass
      const int g = tid / lpr, lg = tid % lpr;
      for (int rbase = 0; rbase < nrows; rbase += gpc) {   // CTA-uniform trip count
        const int r = rbase + g;
        T acc = zero_<T>();
        if (r < nrows) {
          const int s = srp[r] - p0, e = srp[r + 1] - p0;
          for (int q = s + lg; q < e; q += lpr) acc = acc + prod[q];
        }
        for (int o = lpr >> 1; o > 0; o >>= 1) acc = acc + shfl_down_(acc, o, lpr);
        if (lg == 0 && r < nrows) {
          const long row = (long)r0 + r;
          if (prev != nullptr) acc = acc + coef * prev[row];
          y[row] = acc;
          nrm += (double)abs2_(acc);
        }
      }
    }
    __syncthreads();  // prod / srp are reused by the next block
  }
  if (want_norm) {
    double tot = block_sum(nrm, red);
    grid_publish(tot, 0.0, ws, 1, red);
  }
}

}  // namespace

template <class T>
void k_spmv(Context& c, const CsrDevice<T>& A, bool conj, const T* x, T* y, real_t<T> coef, const T* prev, Pending* nrm) {
  ReduceWs ws{};
  int want = 0;
  if (nrm) { ws = c.new_reduce(nrm); want = 1; }
  const int grid = c.grid_for(A.n_blocks, 1, A.ctas_per_sm);
  if (conj && scalar_traits<T>::is_complex)
    spmv_kernel<T, true><<<grid, kThreads, 0, c.stream>>>(A, x, y, coef, prev, ws, want);
  else
    spmv_kernel<T, false><<<grid, kThreads, 0, c.stream>>>(A, x, y, coef, prev, ws, want);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}

// Cut rows into blocks of <= kSpmvRows rows and <= nb non-zeros; longer rows stand alone.
std::vector<int> csr_row_blocks(const int* rp, int rows, int nb) {
  std::vector<int> blk;
  blk.push_back(0);
  int start = 0;
  while (start < rows) {
    int end = start;
    long nnz = 0;
    while (end < rows && end - start < kSpmvRows) {
      const long len = rp[end + 1] - rp[end];
      if (nnz + len > nb) break;
      nnz += len;
      ++end;
    }
    if (end == start) ++end;  // a single row longer than nb
    blk.push_back(end);
    start = end;
  }
  return blk;
}

#define PB_INST(T) \
  template void k_spmv<T>(Context&, const CsrDevice<T>&, bool, const T*, T*, real_t<T>, const T*, Pending*);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
