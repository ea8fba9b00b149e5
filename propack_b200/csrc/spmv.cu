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
// Row binning (computed once per matrix in csr_analyze):
//   short  rows (<= 4*LPR nnz)   : LPR-lane sub-warp per row, 256/LPR rows per CTA-iteration;
//                                  consecutive rows => the warp's (ci,va) reads are contiguous
//   medium rows (<= 2048 nnz)    : one warp per row (power-law matrices, BASELINE config 4)
//   long   rows                  : one CTA per row
// (ci,va) stream through with evict-first loads (each byte is used once); x goes through the
// read-only path and stays L2-resident (8 MB at 1M columns, 80 MB at 10M of the 126 MB L2).
// HBM-bound: algorithmic bytes nnz*(w+4) + 4(rows+1) + w*cols + w*rows (+ w*rows for prev).
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

template <class T, bool CONJ> __device__ inline void mac(T& acc, T a, T x) {
  if (CONJ) fma_conj(acc, a, x);
  else fma_(acc, a, x);
}

template <class T, int LPR, bool CONJ>
__global__ void __launch_bounds__(kThreads)
spmv_kernel(CsrDevice<T> A, const T* __restrict__ x, T* __restrict__ y,
            real_t<T> coef, const T* __restrict__ prev, ReduceWs ws, int want_norm) {
  constexpr int GPC = kThreads / LPR;  // row groups per CTA
  __shared__ double red[32];
  __shared__ T redT[32];
  const int g = threadIdx.x / LPR, lg = threadIdx.x % LPR;
  const int short_max = 4 * LPR;
  double nrm = 0.0;

  // ---- short rows ------------------------------------------------------------------------------
  for (long base = (long)blockIdx.x * GPC; base < A.rows; base += (long)gridDim.x * GPC) {  // warp-uniform trip count
    const long row = base + g;
    int p0 = 0, p1 = 0;
    if (row < A.rows) { p0 = __ldg(A.rp + row); p1 = __ldg(A.rp + row + 1); }
    const bool mine = (row < A.rows) && (p1 - p0 <= short_max);
    T acc = zero_<T>();
    if (mine) {
      for (int p = p0 + lg; p < p1; p += LPR) mac<T, CONJ>(acc, ldcs_(A.va + p), ldg_(x + __ldcs(A.ci + p)));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) acc = acc + shfl_down_(acc, o, LPR);
    if (lg == 0 && mine) {
      if (prev != nullptr) acc = acc + coef * prev[row];
      y[row] = acc;
      nrm += (double)abs2_(acc);
    }
  }
  // ---- medium rows: warp per row ----------------------------------------------------------------
  {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long i = (long)blockIdx.x * 8 + w; i < A.n_med; i += (long)gridDim.x * 8) {
      const int row = A.med_rows[i];
      const int p0 = A.rp[row], p1 = A.rp[row + 1];
      T acc = zero_<T>();
      for (int p = p0 + lane; p < p1; p += 32) mac<T, CONJ>(acc, ldcs_(A.va + p), ldg_(x + __ldcs(A.ci + p)));
      acc = warp_sum(acc);
      if (lane == 0) {
        if (prev != nullptr) acc = acc + coef * prev[row];
        y[row] = acc;
        nrm += (double)abs2_(acc);
      }
    }
  }
  // ---- long rows: CTA per row ---------------------------------------------------------------------
  for (int i = blockIdx.x; i < A.n_long; i += gridDim.x) {
    const int row = A.long_rows[i];
    const int p0 = A.rp[row], p1 = A.rp[row + 1];
    T acc = zero_<T>();
    for (int p = p0 + threadIdx.x; p < p1; p += kThreads) mac<T, CONJ>(acc, ldcs_(A.va + p), ldg_(x + __ldcs(A.ci + p)));
    acc = block_sum(acc, redT);
    if (threadIdx.x == 0) {
      if (prev != nullptr) acc = acc + coef * prev[row];
      y[row] = acc;
      nrm += (double)abs2_(acc);
    }
  }
  if (want_norm) {
    double tot = block_sum(nrm, red);
    grid_publish(tot, 0.0, ws, 1, red);
  }
}

template <class T, int LPR>
void launch(Context& c, const CsrDevice<T>& A, bool conj, const T* x, T* y,
            real_t<T> coef, const T* prev, Pending* nrm) {
  constexpr int GPC = kThreads / LPR;
  ReduceWs ws{};
  int want = 0;
  if (nrm) { ws = c.new_reduce(nrm); want = 1; }
  const int grid = c.grid_for(A.rows, GPC, 8);
  if (conj && scalar_traits<T>::is_complex)
    spmv_kernel<T, LPR, true><<<grid, kThreads, 0, c.stream>>>(A, x, y, coef, prev, ws, want);
  else
    spmv_kernel<T, LPR, false><<<grid, kThreads, 0, c.stream>>>(A, x, y, coef, prev, ws, want);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}

}  // namespace

template <class T>
void k_spmv(Context& c, const CsrDevice<T>& A, bool conj, const T* x, T* y, real_t<T> coef, const T* prev, Pending* nrm) {
  switch (A.lanes_per_row) {
    case 2: launch<T, 2>(c, A, conj, x, y, coef, prev, nrm); break;
    case 4: launch<T, 4>(c, A, conj, x, y, coef, prev, nrm); break;
    case 8: launch<T, 8>(c, A, conj, x, y, coef, prev, nrm); break;
    case 16: launch<T, 16>(c, A, conj, x, y, coef, prev, nrm); break;
    default: launch<T, 32>(c, A, conj, x, y, coef, prev, nrm); break;
  }
}

#define PB_INST(T) \
  template void k_spmv<T>(Context&, const CsrDevice<T>&, bool, const T*, T*, real_t<T>, const T*, Pending*);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
