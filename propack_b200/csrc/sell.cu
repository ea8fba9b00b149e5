// propack_b200 -- sliced-ELL (SELL-32-sigma) SpMV with the fused Lanczos epilogue, and its device-side builder.
//
// Reference: the user-supplied APROD (contract dlansvd.F:20-33), called at dlanbpro.F:288 ('t': v = A^T u) and
// :420 ('n': u = A v), each followed by pdaxpy + pdnrm2 (:295-296, :423-424).  One launch computes
//        y = op(A) x + coef * prev          and publishes ||y||_2.
//
// Layout (built once per operator from the device CSR, sell_build below):
//   * rows are sorted by length (descending, stable) inside windows of kSellSigma = 1024 consecutive rows, so the
//     32 rows of a slice have (nearly) equal lengths and the padding stays at a few per cent for any row-length
//     distribution; perm[slot] = original row of a slot;
//   * a slice = 32 slots = one warp, lane = slot; entry k of slot `lane` lives at  soff[s] + 32*k + lane, so every
//     (ci, va) load of a warp is one fully coalesced 128 / 256-byte line and is touched exactly once;
//   * padding entries carry column -1 and are predicated off (no gather wavefront is spent on them);
//   * rows longer than kSellLong (power-law matrices, BASELINE config 4) keep width 0 here and are done by
//     spmv_long_kernel (spmv.cu: one CTA per row, from the CSR arrays) BEFORE this kernel, which only folds their
//     |y|^2 into the norm.
// Kernel: thread per row, no shared memory, no shuffles, no barrier in the loop: all of the SM's 228 KB stays L1 and
// tracks the outstanding (ci -> x) gather misses (profiles/r01_spmv_lab.md: the x gathers bound this kernel by the
// L1TEX wavefront rate, and every KB of shared-memory carve-out costs gather throughput).  U independent gather
// chains per lane per batch; the next slice's offsets / row ids are prefetched while the gathers are in flight.
// Reduction order inside a row is the column order, lane-private => bit-reproducible.
//
// Row-sharded runs (one process per GPU): the local operand is split by SOURCE RANK of the gathered vector into
// phases, each phase its own SELL matrix (own permutation): y = sum_g A_g x_g.  Phase g waits (in-kernel) for the
// arrival flags of its sources only, accumulates into y, and the last phase applies the epilogue -- so the SpMV of the
// slices that have landed overlaps the NVLink transfer of the ones still in flight (engine.hpp: ShardedCsrOperator).
#include <cub/cub.cuh>

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

template <class T> constexpr int sell_unroll() { return sizeof(T) <= 8 ? 8 : 4; }

constexpr int kSellAcc = kSellModeAcc;      // y += A_g x   (a later phase of a split product)
constexpr int kSellFinal = kSellModeFinal;  // apply  + coef*prev, publish ||y||

template <class T, bool CONJ>
__global__ void __launch_bounds__(kThreads)
spmv_sell_kernel(SellDevice<T> S, const T* __restrict__ x, T* y, real_t<T> coef, const T* __restrict__ prev, ReduceWs ws,
                 int want_norm, int mode, const unsigned long long* flags, unsigned int src_mask, unsigned long long epoch) {
  constexpr int U = sell_unroll<T>();
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  if (src_mask != 0u) {
    // row-sharded run: the slices of the ranks in src_mask must have landed (epoch `epoch`) before x is gathered
    if (threadIdx.x < 32 && ((src_mask >> threadIdx.x) & 1u)) {
      const long long t0 = clock64();
      while (*reinterpret_cast<const volatile unsigned long long*>(flags + threadIdx.x) < epoch) {
        if (clock64() - t0 > ws.timeout_cycles) { *ws.host_err = 1u; break; }   // a peer died; do not hang the GPU
      }
      __threadfence();
    }
    __syncthreads();
  }
  const long nwarps = (long)gridDim.x * (kThreads / 32);
  long s = (long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  double nrm = 0.0;
  const bool acc_mode = (mode & kSellAcc) != 0, final_mode = (mode & kSellFinal) != 0;
  long long off = 0, end = 0;
  int row = -1;
  if (s < S.nslices) { off = __ldg(S.soff + s); end = __ldg(S.soff + s + 1); row = __ldg(S.perm + s * 32 + lane); }
  for (; s < S.nslices; s += nwarps) {
    // prefetch the next slice's extent and row ids
    const long sn = s + nwarps;
    long long offn = 0, endn = 0;
    int rown = -1;
    if (sn < S.nslices) { offn = __ldg(S.soff + sn); endn = __ldg(S.soff + sn + 1); rown = __ldg(S.perm + sn * 32 + lane); }
    const int w = (int)((end - off) >> 5);
    const int* cip = S.ci + off + lane;
    const T* vap = S.va + off + lane;
    T acc = zero_<T>();
    int k0 = 0;
    for (; k0 + U <= w; k0 += U) {   // full batches: U unconditional, independent (ci -> x) gather chains per lane
      int c[U];
      T av[U], xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) c[u] = __ldcs(cip + (long)(k0 + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) av[u] = ldcs_(vap + (long)(k0 + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u] = ldg_(x + max(c[u], 0));   // padding (-1) reads x[0] and is discarded below
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const T xe = c[u] >= 0 ? xv[u] : zero_<T>();
        if (CONJ) fma_conj(acc, av[u], xe);
        else fma_(acc, av[u], xe);
      }
    }
    if (k0 < w) {                    // remainder batch (warp-uniform count)
      const int rem = w - k0;
      int c[U];
      T av[U], xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) c[u] = u < rem ? __ldcs(cip + (long)(k0 + u) * 32) : -1;
#pragma unroll
      for (int u = 0; u < U; ++u) av[u] = u < rem ? ldcs_(vap + (long)(k0 + u) * 32) : zero_<T>();
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u] = ldg_(x + max(c[u], 0));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const T xe = c[u] >= 0 ? xv[u] : zero_<T>();
        if (CONJ) fma_conj(acc, av[u], xe);
        else fma_(acc, av[u], xe);
      }
    }
    if (row >= 0) {
      const int r = row & 0x3fffffff;
      if (row & 0x40000000) {            // long row: produced by spmv_long_kernel (earlier launch), epilogue included
        if (final_mode) nrm += (double)abs2_(y[r]);
      } else {
        if (acc_mode) acc = acc + y[r];
        if (final_mode) {
          if (prev != nullptr) acc = acc + coef * ldcs_(prev + r);
          nrm += (double)abs2_(acc);
        }
        y[r] = acc;
      }
    }
    off = offn; end = endn; row = rown;
  }
  if (want_norm) {
    double tot = block_sum(nrm, red);
    grid_publish(tot, 0.0, ws, 1, red);
  }
}

// -----------------------------------------------------------------------------------------------------------
// builder (setup, integer work: bit-exact against the numpy restatement in tests/sell_ref.py)
// -----------------------------------------------------------------------------------------------------------
constexpr int kSortThreads = kSellSigma;     // one thread per row of a window
constexpr int kBins = kSellLong + 2;         // lengths kSellLong .. 0 (descending), then "no such row"

// One CTA per window: stable counting sort of the window's rows by effective length, descending.
// perm[window*sigma + rank] = row | long flag (or -1 past the last row); width[slice] = length of the slice's first
// (= longest) row.
__global__ void __launch_bounds__(kSortThreads)
sell_sort_kernel(int rows, const int* __restrict__ rp, int* __restrict__ perm, long long* __restrict__ width32) {
  __shared__ int cnt[kSortThreads / 32][kBins];
  __shared__ int bin_off[kBins];
  __shared__ int len_sorted[kSortThreads];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const long row = (long)blockIdx.x * kSellSigma + t;
  int len = -1;
  bool is_long = false;
  if (row < rows) {
    len = __ldg(rp + row + 1) - __ldg(rp + row);
    if (len > kSellLong) { is_long = true; len = 0; }
  }
  const int bin = len < 0 ? kBins - 1 : kSellLong - len;
  for (int i = t; i < (kSortThreads / 32) * kBins; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const unsigned same = __match_any_sync(0xffffffffu, bin);
  const int rank_in_warp = __popc(same & ((1u << lane) - 1u));
  if (rank_in_warp == 0) cnt[w][bin] = __popc(same);
  __syncthreads();
  if (t < kBins) {  // exclusive prefix over the warps of this bin; total left in bin_off
    int run = 0;
    for (int ww = 0; ww < kSortThreads / 32; ++ww) { const int c = cnt[ww][t]; cnt[ww][t] = run; run += c; }
    bin_off[t] = run;
  }
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int b = 0; b < kBins; ++b) { const int c = bin_off[b]; bin_off[b] = run; run += c; }
  }
  __syncthreads();
  const int rank = bin_off[bin] + cnt[w][bin] + rank_in_warp;
  perm[(long)blockIdx.x * kSellSigma + rank] = len < 0 ? -1 : ((int)row | (is_long ? 0x40000000 : 0));
  len_sorted[rank] = len < 0 ? 0 : len;
  __syncthreads();
  if (lane == 0) {
    const long slice = (long)blockIdx.x * (kSellSigma / 32) + w;
    width32[slice] = 32LL * len_sorted[w * 32];
  }
}

template <class T>
__global__ void __launch_bounds__(kThreads)
sell_fill_kernel(long nslices, const int* __restrict__ rp, const int* __restrict__ ci, const T* __restrict__ va,
                 const int* __restrict__ perm, const long long* __restrict__ soff, int* __restrict__ sci, T* __restrict__ sva) {
  const int lane = threadIdx.x & 31;
  const long nwarps = (long)gridDim.x * (kThreads / 32);
  for (long s = (long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); s < nslices; s += nwarps) {
    const long long off = soff[s];
    const int w = (int)((soff[s + 1] - off) >> 5);
    const int row = perm[s * 32 + lane];
    int beg = 0, len = 0;
    if (row >= 0 && !(row & 0x40000000)) { beg = rp[row]; len = rp[row + 1] - beg; }
    for (int k = 0; k < w; ++k) {
      const bool ok = k < len;
      sci[off + (long)k * 32 + lane] = ok ? ci[beg + k] : -1;
      sva[off + (long)k * 32 + lane] = ok ? va[beg + k] : zero_<T>();
    }
  }
}

}  // namespace

template <class T>
void sell_build(Context& c, int rows, int cols, long nnz, const int* rp, const int* ci, const T* va, SellStorage<T>& out) {
  const long nwin = ((long)rows + kSellSigma - 1) / kSellSigma;
  const long nslices = nwin * (kSellSigma / 32);
  out.perm.alloc((size_t)std::max<long>(nslices * 32, 1));
  out.soff.alloc((size_t)nslices + 1);
  SellDevice<T>& D = out.dev;
  D.rows = rows; D.cols = cols; D.nnz = nnz; D.nslices = nslices;
  D.perm = out.perm.p; D.soff = out.soff.p;
  if (nwin == 0) { PB_CUDA(cudaMemsetAsync(out.soff.p, 0, sizeof(long long), c.stream)); D.padded = 0; c.sync(); return; }
  DeviceBuffer<long long> width((size_t)nslices + 1);
  PB_CUDA(cudaMemsetAsync(width.p, 0, sizeof(long long) * (nslices + 1), c.stream));
  sell_sort_kernel<<<(unsigned)nwin, kSortThreads, 0, c.stream>>>(rows, rp, out.perm.p, width.p);
  PB_LAUNCH_CHECK();
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, width.p, out.soff.p, (int)(nslices + 1), c.stream);
  DeviceBuffer<char> tmp(tmp_bytes + 16);
  PB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, width.p, out.soff.p, (int)(nslices + 1), c.stream));
  long long padded = 0;
  PB_CUDA(cudaMemcpyAsync(&padded, out.soff.p + nslices, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  D.padded = padded;
  out.ci.alloc((size_t)std::max<long long>(padded, 1));
  out.va.alloc((size_t)std::max<long long>(padded, 1));
  D.ci = out.ci.p; D.va = out.va.p;
  if (padded > 0) {
    sell_fill_kernel<T><<<c.grid_for(nslices, kThreads / 32, 8), kThreads, 0, c.stream>>>(nslices, rp, ci, va, out.perm.p, out.soff.p,
                                                                                         out.ci.p, out.va.p);
    PB_LAUNCH_CHECK();
  }
  c.sync();
}

template <class T>
void k_spmv_sell(Context& c, const SellDevice<T>& S, const CsrDevice<T>* long_src, bool conj, const T* x, T* y, real_t<T> coef,
                 const T* prev, Pending* nrm, int mode, const unsigned long long* flags, unsigned int src_mask,
                 unsigned long long epoch) {
  ReduceWs ws{};
  int want = 0;
  if (nrm && (mode & kSellFinal)) { ws = c.new_reduce(nrm); want = 1; }
  ws.host_err = c.host_err_dev;
  ws.timeout_cycles = c.peer_timeout_cycles;
  // persistent grid; one CTA slot per SM is left free so that a concurrent NVLink push kernel (row-sharded runs) can
  // always become resident while these CTAs spin on arrival flags
  const int per_sm = src_mask ? kSellCtasPerSm - 1 : kSellCtasPerSm;
  const int grid = c.grid_for(S.nslices, kThreads / 32, per_sm);
  const bool cj = conj && scalar_traits<T>::is_complex;
  if (long_src != nullptr && long_src->n_long > 0)
    k_spmv_long<T>(c, *long_src, cj, x, y, coef, (mode & kSellFinal) ? prev : nullptr, (mode & kSellAcc) != 0);
  if (cj)
    spmv_sell_kernel<T, true><<<grid, kThreads, 0, c.stream>>>(S, x, y, coef, prev, ws, want, mode, flags, src_mask, epoch);
  else
    spmv_sell_kernel<T, false><<<grid, kThreads, 0, c.stream>>>(S, x, y, coef, prev, ws, want, mode, flags, src_mask, epoch);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
  if (want) c.complete_reduce(*nrm, 1);
}

#define PB_INST(T)                                                                                                       \
  template void sell_build<T>(Context&, int, int, long, const int*, const int*, const T*, SellStorage<T>&);              \
  template void k_spmv_sell<T>(Context&, const SellDevice<T>&, const CsrDevice<T>*, bool, const T*, T*, real_t<T>,       \
                               const T*, Pending*, int, const unsigned long long*, unsigned int, unsigned long long);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
