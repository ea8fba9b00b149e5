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
// Organisation ("warp-private row groups"; chosen from the measurements in profiles/r01_spmv_lab.md):
//   * a row group = RPG = 32/LPR consecutive rows, LPR (a power of two, lanes per row) picked per matrix from
//     the mean row length so that a group's non-zeros fit the warp's 4 KB shared-memory slice;
//   * phase 1: the warp streams the group's (ci, va) slice with coalesced loads in batches of kSpmvU
//     independent (ci -> x) gather chains per lane and parks the products in its slice;
//   * phase 2: LPR lanes reduce each row out of shared memory, then y = sum + coef*prev and the ||y||^2 partial;
//   * no CTA-wide barrier anywhere in the loop: every warp runs its own load/gather/reduce pipeline, and the
//     next group's row pointers are fetched while the current gathers are in flight;
//   * rows of a group are packed greedily into slice-sized runs, so irregular row lengths cost extra runs, not a
//     slow path; rows longer than half a slice (power-law matrices) are done by spmv_long_kernel, one CTA per row,
//     ahead of the main kernel, which only folds their |y|^2 into the norm.
// The random x gathers make this kernel L1TEX-wavefront bound, not HBM bound (one wavefront per gathered lane:
// 10 M gathers take >= 44 us on a B200 however the rest is organised).  L1 also tracks the outstanding misses,
// so the shared-memory footprint is capped (3 CTAs x 33 KB per SM, carve-out hint 50% = the 132 KB configuration): with the carve-out at
// the maximum the same gathers run 3x slower.
// HBM traffic: nnz*(w+4) + 4(rows+1) + w*cols + w*rows (+ w*rows for prev); (ci, va) are touched once per
// launch and loaded evict-first so they do not displace x in L2 (x is 8 MB at 1M columns, 80 MB at 10M).
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

constexpr int kSpmvWarps = kThreads / 32;
constexpr int kLongThreads = 128;

// Rows longer than half a slice ("long rows": power-law matrices, BASELINE config 4) get a CTA each, before the main
// kernel: y[row] = sum + coef*prev[row].  The main kernel skips them and only folds |y[row]|^2 into its norm partial,
// so the result is still deterministic (no atomics, fixed reduction order).
template <class T, bool CONJ>
__global__ void __launch_bounds__(kLongThreads)
spmv_long_kernel(CsrDevice<T> A, const T* __restrict__ x, T* y, real_t<T> coef, const T* prev, int accumulate) {
  __shared__ T red[32];
  const int row = __ldg(A.long_rows + blockIdx.x);
  const int s = __ldg(A.rp + row), e = __ldg(A.rp + row + 1);
  T acc = zero_<T>();
  for (int p = s + threadIdx.x; p < e; p += kLongThreads) {
    const T a = ldcs_(A.va + p);
    const T xv = ldg_(x + __ldcs(A.ci + p));
    if (CONJ) fma_conj(acc, a, xv);
    else fma_(acc, a, xv);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    if (accumulate) acc = acc + y[row];
    if (prev != nullptr) acc = acc + coef * prev[row];
    y[row] = acc;
  }
}

template <class T, bool CONJ>
__global__ void __launch_bounds__(kThreads)
spmv_kernel(CsrDevice<T> A, const T* __restrict__ x, T* y, real_t<T> coef, const T* prev, ReduceWs ws, int want_norm) {
  constexpr int NB = spmv_group_nnz<T>();
  constexpr int LONG = NB / 2;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ T prod[kSpmvWarps][NB];
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int lpr_log2 = A.lpr_log2, lpr = 1 << lpr_log2, rpg = 32 >> lpr_log2;
  const int sub = lane >> lpr_log2, lg = lane & (lpr - 1);
  T* pw = prod[w];
  double nrm = 0.0;
  const long ngroups = ((long)A.rows + rpg - 1) / rpg;
  const long gstride = (long)gridDim.x * kSpmvWarps;
  long g = (long)blockIdx.x * kSpmvWarps + w;
  // lane (sub, lg) carries the extent [s, e) of row g*rpg + sub; lanes past the last row carry an empty extent
  auto load_extent = [&](long gg, int& s, int& e) {
    const long row = gg * rpg + sub;
    const bool ok = gg < ngroups && row < A.rows;
    s = __ldg(A.rp + (ok ? row : (long)A.rows));
    e = __ldg(A.rp + (ok ? row + 1 : (long)A.rows));
  };
  int s, e;
  load_extent(g, s, e);
  for (; g < ngroups; g += gstride) {
    int sn, en;
    load_extent(g + gstride, sn, en);  // prefetch the next group's extents
    const long row = g * rpg + sub;
    const bool ok = row < A.rows;
    const bool is_long = (e - s) > LONG;
    T pv = zero_<T>();
    if (ok && prev != nullptr && lg == 0 && !is_long) pv = ldcs_(prev + row);
    if (is_long && lg == 0) nrm += (double)abs2_(y[row]);  // produced by spmv_long_kernel (earlier launch)
    // Greedy packing of the group's rows into slice-sized runs [a, b): consecutive short rows whose non-zeros fit
    // the warp's slice.  Uniform short rows give one run per group; a long row splits the group around it.
    int a = 0;
    while (a < rpg) {
      const int first_long = (int)__reduce_min_sync(FULL, (unsigned)((is_long && sub >= a) ? sub : rpg));
      if (first_long == a) { ++a; continue; }
      const int pa = __shfl_sync(FULL, s, a << lpr_log2);
      const bool fits = sub >= a && sub < first_long && (e - pa) <= NB;
      const int b = (int)__reduce_max_sync(FULL, (unsigned)(fits ? sub + 1 : 0));   // >= a+1: a short row always fits
      const int nn = __shfl_sync(FULL, e, ((b - 1) << lpr_log2)) - pa;
      // ---- phase 1: products of the run's slice into the warp's shared-memory slice -----------------------
      const int* cip = A.ci + pa;
      const T* vap = A.va + pa;
      for (int i0 = 0; i0 < nn; i0 += kSpmvU * 32) {
        int c[kSpmvU];
        T av[kSpmvU], xv[kSpmvU];
#pragma unroll
        for (int u = 0; u < kSpmvU; ++u) { const int i = i0 + u * 32 + lane; c[u] = i < nn ? __ldcs(cip + i) : -1; }
#pragma unroll
        for (int u = 0; u < kSpmvU; ++u) { const int i = i0 + u * 32 + lane; av[u] = i < nn ? ldcs_(vap + i) : zero_<T>(); }
#pragma unroll
        for (int u = 0; u < kSpmvU; ++u) xv[u] = c[u] >= 0 ? ldg_(x + c[u]) : zero_<T>();
#pragma unroll
        for (int u = 0; u < kSpmvU; ++u) { const int i = i0 + u * 32 + lane; if (i < nn) pw[i] = mul_<T, CONJ>(av[u], xv[u]); }
      }
      __syncwarp();
      // ---- phase 2: LPR lanes per row ------------------------------------------------------------------------
      const bool mine = sub >= a && sub < b;
      T acc = zero_<T>();
      if (mine)
        for (int q = s - pa + lg; q < e - pa; q += lpr) acc = acc + pw[q];
      for (int o = lpr >> 1; o > 0; o >>= 1) acc = acc + shfl_down_(acc, o, lpr);
      __syncwarp();  // the slice is rewritten by the next run
      if (mine && ok && lg == 0) {
        if (prev != nullptr) acc = acc + coef * pv;
        y[row] = acc;
        nrm += (double)abs2_(acc);
      }
      a = b;
    }
    s = sn; e = en;
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
  const int rpg = 32 >> A.lpr_log2;
  const long ngroups = ((long)A.rows + rpg - 1) / rpg;
  const int grid = c.grid_for(ngroups, kSpmvWarps, kSpmvCtasPerSm);
  static bool attr_set = false;
  if (!attr_set) {  // keep most of the 228 KB for L1: it tracks the outstanding gather misses
    PB_CUDA(cudaFuncSetAttribute(spmv_kernel<T, true>, cudaFuncAttributePreferredSharedMemoryCarveout, kSpmvCarveoutPct));
    PB_CUDA(cudaFuncSetAttribute(spmv_kernel<T, false>, cudaFuncAttributePreferredSharedMemoryCarveout, kSpmvCarveoutPct));
    attr_set = true;
  }
  const bool cj = conj && scalar_traits<T>::is_complex;
  k_spmv_long<T>(c, A, cj, x, y, coef, prev, false);
  if (cj) spmv_kernel<T, true><<<grid, kThreads, 0, c.stream>>>(A, x, y, coef, prev, ws, want);
  else spmv_kernel<T, false><<<grid, kThreads, 0, c.stream>>>(A, x, y, coef, prev, ws, want);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
  if (nrm) c.complete_reduce(*nrm, 1);
}

template <class T>
void k_spmv_long(Context& c, const CsrDevice<T>& A, bool cj, const T* x, T* y, real_t<T> coef, const T* prev, bool accumulate) {
  if (A.n_long <= 0) return;
  if (cj) spmv_long_kernel<T, true><<<A.n_long, kLongThreads, 0, c.stream>>>(A, x, y, coef, prev, accumulate ? 1 : 0);
  else spmv_long_kernel<T, false><<<A.n_long, kLongThreads, 0, c.stream>>>(A, x, y, coef, prev, accumulate ? 1 : 0);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}

// Lanes per row for a matrix with `nnz` non-zeros in `rows` rows: the largest row group (32/LPR rows) whose
// expected non-zero count, with 30% head-room for row-length spread, fits the warp's slice of `nb` entries.
int csr_lanes_per_row_log2(long nnz, int rows, int nb) {
  const double mean = rows > 0 ? (double)nnz / rows : 0.0;
  int rpg = 32;
  while (rpg > 1 && rpg * 1.3 * mean > nb) rpg >>= 1;
  int l = 0;
  while ((32 >> l) > rpg) ++l;
  return l;
}

#define PB_INST(T) \
  template void k_spmv<T>(Context&, const CsrDevice<T>&, bool, const T*, T*, real_t<T>, const T*, Pending*); \
  template void k_spmv_long<T>(Context&, const CsrDevice<T>&, bool, const T*, T*, real_t<T>, const T*, bool);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
