// propack_b200 -- sliced jagged-ELL SpMV with the fused Lanczos epilogue, and its device-side builder.
//
// Reference: the user-supplied APROD (contract dlansvd.F:20-33), called at dlanbpro.F:288 ('t': v = A^T u) and
// :420 ('n': u = A v), each followed by pdaxpy + pdnrm2 (:295-296, :423-424).  One launch computes
//        y = op(A) x + coef * prev          and publishes ||y||_2.
//
// Layout (built once per operand from the device CSR, sell_build below): rows keep their order; a slice = 32
// consecutive rows = one warp, lane = row.  Inside a slice the entries are stored "k-major, compacted": first the 0-th
// entry of every row that has one (in row order), then the 1-st entries, ... -- i.e. column k of the slice holds
// popc(ballot(len > k)) entries and lane l finds its own at  base_k + popc(ballot & lanes_below(l)).  So
//   * storage is exactly nnz entries (no padding) and needs no permutation: y, prev and the row lengths are read and
//     written fully coalesced, in natural row order;
//   * every (ci, va) load of a warp is one contiguous run (one or two 128-byte lines), touched exactly once;
//   * ragged rows cost idle lanes in the late k-steps, not memory traffic: the kernel is bound by the x gathers, and
//     inactive lanes issue none;
//   * rows longer than kSellLong (power-law matrices, BASELINE config 4) are left out of the slices (length byte 0xFF)
//     and done by spmv_long_kernel (spmv.cu: one CTA per row, from the CSR arrays) BEFORE this kernel, which only folds
//     their |y|^2 into the norm.
// Kernel: thread per row, no shared memory in the loop, no barrier: all of the SM's 228 KB stays L1 and tracks the
// outstanding (ci -> x) gather misses (profiles/r01_spmv_lab.md: the x gathers bound this kernel by the L1TEX wavefront
// rate, and shared-memory carve-out costs gather throughput).  U independent gather chains per lane per batch.  Warps
// own contiguous slice ranges of equal weight (planned at build time for the launch grid), so ragged slices do not pile
// up on some warps.  Reduction order inside a row is the column order, lane-private => bit-reproducible.
//
// Column blocking.  A gathered vector larger than the L2 can keep (x = 80 MB on BASELINE config 5) makes most gathers
// miss to DRAM at 32-byte sector granularity: ncu measured 3.1 GB of DRAM reads per product against 1.4 GB algorithmic
// with the operand in one piece, 1.68 GB with two panels (profiles/r02_ncu_extract_c5.txt: algorithmic + the second panel's
// read-modify-write of y).  Operands are therefore split by COLUMN RANGE into panels ("phases") whose slice of x
// stays L2-resident: y = sum_g A_g x_g, one launch per panel, accumulating into y (coalesced read-modify-write), the
// last panel applying the epilogue.  Row-sharded runs (one process per GPU) use the same mechanism with the panels
// = the source ranks of the gathered vector: panel g waits (in-kernel) only for the arrival flags of its sources, so the
// SpMV over the slices that have landed overlaps the NVLink transfer of the ones still in flight (engine.hpp).
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace pb {

namespace {

// predicated streaming / read-only loads as single PTX instructions, so that a batch of U of them stays a batch of U
// independent loads in flight (the compiler turns a C++ "p ? load : default" chain into serialised branches)
__device__ inline int ld_cs_pred(const int* p, bool pred) {
  int v = -1;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.cs.s32 %0, [%1]; }" : "+r"(v) : "l"(p), "r"((int)pred));
  return v;
}
__device__ inline float ld_cs_pred(const float* p, bool pred) {
  float v = 0.f;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.cs.f32 %0, [%1]; }" : "+f"(v) : "l"(p), "r"((int)pred));
  return v;
}
__device__ inline double ld_cs_pred(const double* p, bool pred) {
  double v = 0.0;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.cs.f64 %0, [%1]; }" : "+d"(v) : "l"(p), "r"((int)pred));
  return v;
}
__device__ inline cplx<float> ld_cs_pred(const cplx<float>* p, bool pred) {
  float a = 0.f, b = 0.f;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q ld.global.cs.v2.f32 {%0, %1}, [%2]; }" : "+f"(a), "+f"(b) : "l"(p), "r"((int)pred));
  return cplx<float>(a, b);
}
__device__ inline cplx<double> ld_cs_pred(const cplx<double>* p, bool pred) {
  double a = 0.0, b = 0.0;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q ld.global.cs.v2.f64 {%0, %1}, [%2]; }" : "+d"(a), "+d"(b) : "l"(p), "r"((int)pred));
  return cplx<double>(a, b);
}
__device__ inline float ld_nc_pred(const float* p, bool pred) {
  float v = 0.f;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.nc.f32 %0, [%1]; }" : "+f"(v) : "l"(p), "r"((int)pred));
  return v;
}
__device__ inline double ld_nc_pred(const double* p, bool pred) {
  double v = 0.0;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.global.nc.f64 %0, [%1]; }" : "+d"(v) : "l"(p), "r"((int)pred));
  return v;
}
__device__ inline cplx<float> ld_nc_pred(const cplx<float>* p, bool pred) {
  float a = 0.f, b = 0.f;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q ld.global.nc.v2.f32 {%0, %1}, [%2]; }" : "+f"(a), "+f"(b) : "l"(p), "r"((int)pred));
  return cplx<float>(a, b);
}
__device__ inline cplx<double> ld_nc_pred(const cplx<double>* p, bool pred) {
  double a = 0.0, b = 0.0;
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q ld.global.nc.v2.f64 {%0, %1}, [%2]; }" : "+d"(a), "+d"(b) : "l"(p), "r"((int)pred));
  return cplx<double>(a, b);
}
template <class T> __device__ inline T ldcs_(const T* p) { return ld_cs_pred(p, true); }

constexpr int kSellAcc = kSellModeAcc;      // y += A_g x   (a later panel of a split product)
constexpr int kSellFinal = kSellModeFinal;  // apply  + coef*prev, publish ||y||

// One warp's pass over the slices [s, s_end) of one operand S: lane = row, U independent (ci -> x) gather chains per lane
// per batch.  acc_mode: y += (a later panel of a split product); final_mode: + coef*prev and |y|^2 into nrm.
template <class T, bool CONJ, int U>
__device__ __forceinline__ void sell_range(const SellDevice<T>& S, long s, const long s_end, const T* __restrict__ x, T* y,
                                           real_t<T> coef, const T* __restrict__ prev, const bool acc_mode, const bool final_mode,
                                           const int lane, const unsigned below, double& nrm) {
  // lane state of the current slice: row length byte (0xFF = long row, done elsewhere) and the slice's first entry
  int lenb = 0;
  long long base = 0;
  if (s < s_end) {
    const long row = s * 32 + lane;
    lenb = row < S.rows ? (int)__ldg(S.len8 + row) : 0;
    base = __ldg(S.joff + s);
  }
  for (; s < s_end; ++s) {
    const long row = s * 32 + lane;
    const bool is_long = lenb == 0xFF;
    const int len = is_long ? 0 : lenb;
    // prefetch the next slice's row lengths and base
    int lenb_n = 0;
    long long base_n = 0;
    if (s + 1 < s_end) {
      const long rn = row + 32;
      lenb_n = rn < S.rows ? (int)__ldg(S.len8 + rn) : 0;
      base_n = __ldg(S.joff + s + 1);
    }
    const int maxlen = (int)__reduce_max_sync(0xffffffffu, (unsigned)len);
    const int* cis = S.ci + base;      // this slice's entries; offsets inside a slice fit 32 bits (<= 32 * kSellLong)
    const T* vas = S.va + base;
    // the epilogue operands are requested up front, under the gathers: with the short rows of a panel of a row-sharded
    // operand (2-3 entries) the slice is latency-bound, and a y / prev load issued after the loop would add a full
    // memory round trip to every slice
    const bool live = row < S.rows && !is_long;
    T y_in = zero_<T>(), p_in = zero_<T>();
    if (live && acc_mode) y_in = y[row];
    if (live && final_mode && prev != nullptr) p_in = ldcs_(prev + row);
    T acc = zero_<T>();
    int pos0 = 0;
    for (int k0 = 0; k0 < maxlen; k0 += U) {
      bool act[U];
      int pos[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        act[u] = (k0 + u) < len;
        const unsigned m = __ballot_sync(0xffffffffu, act[u]);
        pos[u] = pos0 + __popc(m & below);
        pos0 += __popc(m);
      }
      int c[U];
      T av[U], xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) c[u] = ld_cs_pred(cis + pos[u], act[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) av[u] = ld_cs_pred(vas + pos[u], act[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u] = ld_nc_pred(x + (act[u] ? c[u] : 0), act[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (CONJ) fma_conj(acc, av[u], xv[u]);
        else fma_(acc, av[u], xv[u]);
      }
    }
    if (row < S.rows) {
      if (is_long) {                       // produced by spmv_long_kernel (earlier launch), epilogue included
        if (final_mode) nrm += (double)abs2_(y[row]);
      } else {
        if (acc_mode) acc = acc + y_in;
        if (final_mode) {
          if (prev != nullptr) acc = acc + coef * p_in;
          nrm += (double)abs2_(acc);
        }
        y[row] = acc;
      }
    }
    lenb = lenb_n; base = base_n;
  }
}

// row-sharded run: the slices of the ranks in src_mask must have landed (epoch `epoch`) before x is gathered.
// Called by every thread of the CTA (contains a barrier).
__device__ __forceinline__ void sell_wait_sources(const unsigned long long* flags, unsigned int src_mask, unsigned long long epoch,
                                                  const ReduceWs& ws) {
  if (src_mask == 0u) return;
  if (threadIdx.x < 32 && ((src_mask >> threadIdx.x) & 1u)) {
    const long long t0 = clock64();
    while (*reinterpret_cast<const volatile unsigned long long*>(flags + threadIdx.x) < epoch) {
      if (clock64() - t0 > ws.timeout_cycles) { *ws.host_err = 1u; break; }   // a peer died; do not hang the GPU
    }
    __threadfence();
  }
  __syncthreads();
}

// U = independent (ci -> x) gather chains per lane per batch; MINB = CTAs per SM the register budget is capped for
// (U = 8: <= 51 registers, 5 CTAs = 40 warps per SM; U = 4: <= 32 registers, 8 CTAs = 64 warps per SM)
template <class T, bool CONJ, int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
spmv_sell_kernel(SellDevice<T> S, const T* __restrict__ x, T* y, real_t<T> coef, const T* __restrict__ prev, ReduceWs ws,
                 int want_norm, int mode, const unsigned long long* flags, unsigned int src_mask, unsigned long long epoch) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  sell_wait_sources(flags, src_mask, epoch, ws);
  const int wid = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  double nrm = 0.0;
  sell_range<T, CONJ, U>(S, __ldg(S.wstart + wid), __ldg(S.wstart + wid + 1), x, y, coef, prev, (mode & kSellAcc) != 0,
                         (mode & kSellFinal) != 0, lane, below, nrm);
  if (want_norm) {
    double tot = block_sum(nrm, red);
    grid_publish(tot, 0.0, ws, 1, red);
  }
}

// kernel variant: PROPACK_B200_SELL_VARIANT = 0 (U = 8, 5 CTAs/SM), 1 (U = 8, 4 CTAs/SM: no register cap; the default -- measured
// best on configs 2 and 5: 32 warps x 8 gather chains beat more warps with spills or shorter batches), 2 (U = 4, 8 CTAs/SM),
// 3 (U = 12, 3 CTAs/SM), 4 (U = 16, 2 CTAs/SM).  Measured (us, configs 5 / 2 / 4, A x and A^H x): variant 1: 534 588 / 68 68 / 272 136;
// variant 0: 621 652 / 74 73; variant 2: 778 798 / 88 89 / 394 209; variant 3: 570 587 / 71 72 / 306 154; variant 4: 771 793 / 79 78 / 416 243.
template <class T> int sell_variant() {   // read at every call: cheap, and lets one process compare settings
  const char* e = std::getenv("PROPACK_B200_SELL_VARIANT");
  const int d = 1;
  const int u = e ? std::atoi(e) : d;
  return (u < 0 || u > 4) ? d : u;
}
template <class T, int V> struct SellCfg;
template <class T> struct SellCfg<T, 0> { static constexpr int U = 8, MINB = 5; };
template <class T> struct SellCfg<T, 1> { static constexpr int U = 8, MINB = 4; };
template <class T> struct SellCfg<T, 2> { static constexpr int U = 4, MINB = sizeof(T) <= 8 ? 8 : 5; };
template <class T> struct SellCfg<T, 3> { static constexpr int U = 12, MINB = sizeof(T) <= 8 ? 3 : 2; };
template <class T> struct SellCfg<T, 4> { static constexpr int U = 16, MINB = sizeof(T) <= 8 ? 2 : 1; };
template <class T, bool CONJ, int V>
void sell_launch(const SellDevice<T>& S, cudaStream_t st, const T* x, T* y, real_t<T> coef, const T* prev, const ReduceWs& ws, int want,
                 int mode, const unsigned long long* flags, unsigned int src_mask, unsigned long long epoch) {
  spmv_sell_kernel<T, CONJ, SellCfg<T, V>::U, SellCfg<T, V>::MINB><<<S.grid, kThreads, 0, st>>>(S, x, y, coef, prev, ws, want, mode, flags,
                                                                                                 src_mask, epoch);
}
template <class T, bool CONJ>
void sell_dispatch(const SellDevice<T>& S, cudaStream_t st, const T* x, T* y, real_t<T> coef, const T* prev, const ReduceWs& ws, int want,
                   int mode, const unsigned long long* flags, unsigned int src_mask, unsigned long long epoch) {
  switch (sell_variant<T>()) {
    case 0: sell_launch<T, CONJ, 0>(S, st, x, y, coef, prev, ws, want, mode, flags, src_mask, epoch); break;
    case 1: sell_launch<T, CONJ, 1>(S, st, x, y, coef, prev, ws, want, mode, flags, src_mask, epoch); break;
    case 3: sell_launch<T, CONJ, 3>(S, st, x, y, coef, prev, ws, want, mode, flags, src_mask, epoch); break;
    case 4: sell_launch<T, CONJ, 4>(S, st, x, y, coef, prev, ws, want, mode, flags, src_mask, epoch); break;
    default: sell_launch<T, CONJ, 2>(S, st, x, y, coef, prev, ws, want, mode, flags, src_mask, epoch); break;
  }
}
// resident CTAs per SM of the variant in use (the persistent grid must not exceed what is co-resident: the warp
// partition is static, and in row-sharded runs the CTAs spin on arrival flags)
template <class T> int sell_occupancy() {
  int occ = 0;
  switch (sell_variant<T>()) {
    case 0: PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_sell_kernel<T, false, SellCfg<T, 0>::U, SellCfg<T, 0>::MINB>, kThreads, 0)); break;
    case 1: PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_sell_kernel<T, false, SellCfg<T, 1>::U, SellCfg<T, 1>::MINB>, kThreads, 0)); break;
    case 3: PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_sell_kernel<T, false, SellCfg<T, 3>::U, SellCfg<T, 3>::MINB>, kThreads, 0)); break;
    case 4: PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_sell_kernel<T, false, SellCfg<T, 4>::U, SellCfg<T, 4>::MINB>, kThreads, 0)); break;
    default: PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_sell_kernel<T, false, SellCfg<T, 2>::U, SellCfg<T, 2>::MINB>, kThreads, 0)); break;
  }
  return std::max(1, occ);
}

template <class T> int sell_registers() {
  cudaFuncAttributes a{};
  switch (sell_variant<T>()) {
    case 0: PB_CUDA(cudaFuncGetAttributes(&a, spmv_sell_kernel<T, false, SellCfg<T, 0>::U, SellCfg<T, 0>::MINB>)); break;
    case 1: PB_CUDA(cudaFuncGetAttributes(&a, spmv_sell_kernel<T, false, SellCfg<T, 1>::U, SellCfg<T, 1>::MINB>)); break;
    case 3: PB_CUDA(cudaFuncGetAttributes(&a, spmv_sell_kernel<T, false, SellCfg<T, 3>::U, SellCfg<T, 3>::MINB>)); break;
    case 4: PB_CUDA(cudaFuncGetAttributes(&a, spmv_sell_kernel<T, false, SellCfg<T, 4>::U, SellCfg<T, 4>::MINB>)); break;
    default: PB_CUDA(cudaFuncGetAttributes(&a, spmv_sell_kernel<T, false, SellCfg<T, 2>::U, SellCfg<T, 2>::MINB>)); break;
  }
  return a.numRegs;
}

// -----------------------------------------------------------------------------------------------------------
// builder (setup, integer work: bit-exact against the numpy restatement in tests/sell_ref.py)
// -----------------------------------------------------------------------------------------------------------
// per row: length byte; per slice: stored entries (long rows excluded) and the weight used for the warp partition
__global__ void __launch_bounds__(kThreads)
sell_lengths_kernel(int rows, long nslices, const int* __restrict__ rp, unsigned char* __restrict__ len8,
                    long long* __restrict__ count, long long* __restrict__ weight) {
  const int lane = threadIdx.x & 31;
  const long nwarps = (long)gridDim.x * (kThreads / 32);
  for (long s = (long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); s < nslices; s += nwarps) {
    const long row = s * 32 + lane;
    int len = 0;
    if (row < rows) {
      const int l = rp[row + 1] - rp[row];
      len8[row] = l > kSellLong ? (unsigned char)0xFF : (unsigned char)l;
      len = l > kSellLong ? 0 : l;
    }
    const int maxlen = (int)__reduce_max_sync(0xffffffffu, (unsigned)len);
    const int tot = (int)__reduce_add_sync(0xffffffffu, (unsigned)len);
    if (lane == 0) {
      count[s] = tot;
      weight[s] = (long long)tot + (long long)kSellStepCost * maxlen + kSellSliceCost;
    }
  }
}

template <class T>
__global__ void __launch_bounds__(kThreads)
sell_fill_kernel(int rows, long nslices, const int* __restrict__ rp, const int* __restrict__ ci, const T* __restrict__ va,
                 const unsigned char* __restrict__ len8, const long long* __restrict__ joff, int* __restrict__ sci, T* __restrict__ sva) {
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  const long nwarps = (long)gridDim.x * (kThreads / 32);
  for (long s = (long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); s < nslices; s += nwarps) {
    const long row = s * 32 + lane;
    int len = 0, beg = 0;
    if (row < rows) {
      const int lb = len8[row];
      len = lb == 0xFF ? 0 : lb;
      beg = rp[row];
    }
    const int maxlen = (int)__reduce_max_sync(0xffffffffu, (unsigned)len);
    long long pos0 = joff[s];
    for (int k = 0; k < maxlen; ++k) {
      const bool act = k < len;
      const unsigned m = __ballot_sync(0xffffffffu, act);
      if (act) {
        const long long p = pos0 + __popc(m & below);
        sci[p] = ci[beg + k];
        sva[p] = va[beg + k];
      }
      pos0 += __popc(m);
    }
  }
}

// wstart[w] = first slice s whose weight prefix reaches w/nwarps of the total (w = 0..nwarps)
__global__ void __launch_bounds__(kThreads)
sell_partition_kernel(long nslices, const long long* __restrict__ wpre, int nwarps, int* __restrict__ wstart) {
  const int w = blockIdx.x * kThreads + threadIdx.x;
  if (w > nwarps) return;
  if (w == nwarps) { wstart[w] = (int)nslices; return; }
  const long long total = wpre[nslices];
  const long long target = (long long)(((__int128)total * w) / nwarps);
  long lo = 0, hi = nslices;   // smallest s in [0, nslices] with wpre[s] >= target
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if (wpre[mid] >= target) hi = mid; else lo = mid + 1;
  }
  wstart[w] = (int)lo;
}

}  // namespace

template <class T>
void sell_build(Context& c, int rows, int cols, long nnz, const int* rp, const int* ci, const T* va, SellStorage<T>& out,
                int ctas_per_sm) {
  const long nslices = ((long)rows + 31) / 32;
  SellDevice<T>& D = out.dev;
  D.rows = rows; D.cols = cols; D.nnz = nnz; D.nslices = nslices;
  out.len8.alloc((size_t)std::max<long>(rows, 1));
  out.joff.alloc((size_t)nslices + 1);
  D.len8 = out.len8.p; D.joff = out.joff.p;
  // ctas_per_sm <= 0: every resident slot but -ctas_per_sm (row-sharded operands keep one free for the NVLink push kernel)
  const int occ = sell_occupancy<T>();
  int per_sm = ctas_per_sm > 0 ? std::min(ctas_per_sm, occ) : std::max(1, occ + ctas_per_sm);
  if (ctas_per_sm < 0) {
    // room for the push kernel means registers too: it runs 256 threads x up to 64 registers (level1.cu), and a slot freed
    // by a 32-register SpMV variant would not hold it
    const int regs = sell_registers<T>();
    if (regs > 0) per_sm = std::max(1, std::min(per_sm, (65536 - 64 * kThreads) / (regs * kThreads)));
  }
  D.grid = c.grid_for(std::max<long>(nslices, 1), kThreads / 32, per_sm);
  const int nwarps = D.grid * (kThreads / 32);
  out.wstart.alloc((size_t)nwarps + 1);
  D.wstart = out.wstart.p;
  if (nslices == 0) {
    PB_CUDA(cudaMemsetAsync(out.joff.p, 0, sizeof(long long), c.stream));
    PB_CUDA(cudaMemsetAsync(out.wstart.p, 0, sizeof(int) * ((size_t)nwarps + 1), c.stream));
    D.stored = 0; c.sync(); return;
  }
  DeviceBuffer<long long> count((size_t)nslices + 1), weight((size_t)nslices + 1), wpre((size_t)nslices + 1);
  PB_CUDA(cudaMemsetAsync(count.p + nslices, 0, sizeof(long long), c.stream));
  PB_CUDA(cudaMemsetAsync(weight.p + nslices, 0, sizeof(long long), c.stream));
  const int grid = c.grid_for(nslices, kThreads / 32, 8);
  sell_lengths_kernel<<<grid, kThreads, 0, c.stream>>>(rows, nslices, rp, out.len8.p, count.p, weight.p);
  PB_LAUNCH_CHECK();
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, count.p, out.joff.p, (int)(nslices + 1), c.stream);
  DeviceBuffer<char> tmp(tmp_bytes + 16);
  PB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, count.p, out.joff.p, (int)(nslices + 1), c.stream));
  PB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, weight.p, wpre.p, (int)(nslices + 1), c.stream));
  sell_partition_kernel<<<ceil_div(nwarps + 1, kThreads), kThreads, 0, c.stream>>>(nslices, wpre.p, nwarps, out.wstart.p);
  PB_LAUNCH_CHECK();
  long long stored = 0;
  PB_CUDA(cudaMemcpyAsync(&stored, out.joff.p + nslices, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  D.stored = stored;
  out.ci.alloc((size_t)std::max<long long>(stored, 1));
  out.va.alloc((size_t)std::max<long long>(stored, 1));
  D.ci = out.ci.p; D.va = out.va.p;
  if (stored > 0) {
    sell_fill_kernel<T><<<grid, kThreads, 0, c.stream>>>(rows, nslices, rp, ci, va, out.len8.p, out.joff.p, out.ci.p, out.va.p);
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
  const bool cj = conj && scalar_traits<T>::is_complex;
  if (long_src != nullptr && long_src->n_long > 0)
    k_spmv_long<T>(c, *long_src, cj, x, y, coef, (mode & kSellFinal) ? prev : nullptr, (mode & kSellAcc) != 0);
  if (cj) sell_dispatch<T, true>(S, c.stream, x, y, coef, prev, ws, want, mode, flags, src_mask, epoch);
  else sell_dispatch<T, false>(S, c.stream, x, y, coef, prev, ws, want, mode, flags, src_mask, epoch);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
  if (want) c.complete_reduce(*nrm, 1);
}

#define PB_INST(T)                                                                                                       \
  template void sell_build<T>(Context&, int, int, long, const int*, const int*, const T*, SellStorage<T>&, int);         \
  template void k_spmv_sell<T>(Context&, const SellDevice<T>&, const CsrDevice<T>*, bool, const T*, T*, real_t<T>,       \
                               const T*, Pending*, int, const unsigned long long*, unsigned int, unsigned long long);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
