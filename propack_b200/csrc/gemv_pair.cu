// propack_b200 -- the tall-skinny GEMV pair behind reorthogonalisation.
//
// Reference: dcgs (double/dreorth.F:106-210) does, per index interval [p,q],
//     y = V(:,p:q)^T * vnew        (dgemv 'T', :174; zgemv 'C' in zreorth.F:169)
//     vnew = vnew - V(:,p:q) * y   (dgemv 'N', :199-205)
// and dreorth (:85-95) follows with pdnrm2.  Both GEMVs are HBM streams over the L x l basis
// block (algorithmic bytes  w*L*(2l+3)  per pass); nothing here is GEMM-shaped, so the
// kernels are organised around 128-bit coalesced loads, many loads in flight per lane, and
// fixed-order reductions (bit-reproducible run to run):
//
//   gemv_t_kernel    h_part[cta][c] = sum over the CTA's rows of conj(V[r,c]) * q[r]
//                    - a warp owns 32*VEC*S consecutive rows per iteration, q held in registers
//                    - CG columns at a time => S*CG independent LDG.128 in flight per lane
//                    - CG accumulators folded with a transposing butterfly (CG+2 shuffles, not 5*CG)
//                    - per-warp column sums accumulate in shared memory, no __syncthreads in the loop
//   gemv_t_tma_kernel  the same product for long vectors, TMA-staged: one elected producer thread streams 16 KB column
//                    segments HBM -> shared memory with cp.async.bulk (SASS: UBLKCP) into a 4-stage ring guarded by
//                    full / empty mbarriers (SYNCS), 8 consumer warps multiply the stage with the CTA's block of q held in
//                    shared memory; ~64 KB per CTA x 2 CTAs/SM in flight with the register file free
//                    (tools/gemv_lab.cu measured 7.2-7.3 TB/s against 6.8-7.1 TB/s register-staged)
//   gemv_t_finalize  h[c] = sum_cta h_part[cta][c]   (fixed order; the hook for the multi-GPU
//                    all-reduce of the l coefficients)
//   gemv_n_kernel    out = cin*in -/+ V*h, h staged in shared memory, CU columns unrolled,
//                    fused ||out||^2 -> last-CTA publication (dreorth's pdnrm2 for free)
#include <algorithm>
#include <cstdlib>

#include "comm.hpp"
#include "kernels.cuh"

namespace pb {

namespace {

constexpr int GT_S = 4;        // strips (packs) per lane per warp-iteration in gemv_t
constexpr int GT_CG = 4;       // columns per group in gemv_t
constexpr int GT_CHUNK = 256;  // columns per blockIdx.y slice (bounds shared memory)
constexpr int GN_S = 2;        // packs per thread per tile in gemv_n
constexpr int GN_CU = 8;       // column unroll in gemv_n
constexpr int GN_CHUNK = 1024; // columns per launch in gemv_n (h staged in smem)

// Fold 4 per-lane accumulators so that every lane of octet o (lanes 8o..8o+7) ends with the
// warp-wide total of accumulator o.
template <class T> __device__ inline T fold4(T a0, T a1, T a2, T a3, int lane) {
  const bool hi16 = lane & 16;
  // exchange halves: lanes <16 keep {a0,a1}, lanes >=16 keep {a2,a3}
  T s0 = hi16 ? a0 : a2, s1 = hi16 ? a1 : a3;
  T k0 = hi16 ? a2 : a0, k1 = hi16 ? a3 : a1;
  k0 = k0 + shfl_xor_(s0, 16);
  k1 = k1 + shfl_xor_(s1, 16);
  const bool hi8 = lane & 8;
  T s = hi8 ? k0 : k1, k = hi8 ? k1 : k0;
  k = k + shfl_xor_(s, 8);
  k = k + shfl_xor_(k, 4);
  k = k + shfl_xor_(k, 2);
  k = k + shfl_xor_(k, 1);
  return k;  // total of accumulator index (lane >> 3)
}

template <class T>
__global__ void __launch_bounds__(kThreads, 2)
gemv_t_kernel(long L, int l, const T* __restrict__ V, long ldv, const T* __restrict__ q, T* __restrict__ part, int lpad, int chunk) {
  constexpr int VEC = Pack<T>::N;
  constexpr int WROWS = 32 * VEC * GT_S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* hs = reinterpret_cast<T*>(smem_raw);  // [8][GT_CHUNK]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c_begin = blockIdx.y * chunk;   // chunk <= GT_CHUNK: the l columns are split EVENLY over gridDim.y slices
  const int c_end = min(l, c_begin + chunk);
  const int nc = c_end - c_begin;
  for (int i = threadIdx.x; i < 8 * GT_CHUNK; i += kThreads) hs[i] = zero_<T>();
  __syncthreads();
  T* hw = hs + w * GT_CHUNK;

  const long gw = (long)blockIdx.x * 8 + w, GW = (long)gridDim.x * 8;
  for (long r0 = gw * WROWS; r0 < L; r0 += GW * WROWS) {
    Pack<T> qv[GT_S];
    bool ok[GT_S];
#pragma unroll
    for (int s = 0; s < GT_S; ++s) {
      const long row = r0 + (long)s * 32 * VEC + (long)lane * VEC;
      ok[s] = row < L;  // padding rows [L, ld) of q and V are zero by construction
      if (ok[s]) qv[s] = ld_pack(q + row);
      else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) qv[s].v[e] = zero_<T>();
      }
    }
    const T* Vr = V + (long)c_begin * ldv + r0 + (long)lane * VEC;
    int c0 = 0;
    for (; c0 + GT_CG <= nc; c0 += GT_CG) {
      Pack<T> pv[GT_CG][GT_S];
#pragma unroll
      for (int cc = 0; cc < GT_CG; ++cc)
#pragma unroll
        for (int s = 0; s < GT_S; ++s)
          if (ok[s]) pv[cc][s] = ld_pack_stream(Vr + (long)(c0 + cc) * ldv + (long)s * 32 * VEC);
      T acc[GT_CG];
#pragma unroll
      for (int cc = 0; cc < GT_CG; ++cc) {
        acc[cc] = zero_<T>();
#pragma unroll
        for (int s = 0; s < GT_S; ++s)
          if (ok[s]) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) fma_conj(acc[cc], pv[cc][s].v[e], qv[s].v[e]);
          }
      }
      T tot = fold4(acc[0], acc[1], acc[2], acc[3], lane);
      if ((lane & 7) == 0) hw[c0 + (lane >> 3)] = hw[c0 + (lane >> 3)] + tot;
    }
    for (; c0 < nc; ++c0) {  // remainder columns
      T acc = zero_<T>();
#pragma unroll
      for (int s = 0; s < GT_S; ++s)
        if (ok[s]) {
          Pack<T> pv = ld_pack_stream(Vr + (long)c0 * ldv + (long)s * 32 * VEC);
#pragma unroll
          for (int e = 0; e < VEC; ++e) fma_conj(acc, pv.v[e], qv[s].v[e]);
        }
      acc = warp_sum(acc);
      if (lane == 0) hw[c0] = hw[c0] + acc;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < nc; c += kThreads) {
    T s = hs[c];
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) s = s + hs[ww * GT_CHUNK + c];
    part[(long)blockIdx.x * lpad + c_begin + c] = s;
  }
}

// ---- TMA-staged variant ------------------------------------------------------------------------------------------
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void mbar_init(uint64_t* b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ inline void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ inline void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ inline void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)),
               "r"(parity)
               : "memory");
}
__device__ inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

constexpr int GTT_STAGES = 4;             // ring depth
constexpr int GTT_STAGE_BYTES = 16384;    // one stage = one column segment of RS = 16 KB / sizeof(T) rows = the CTA's q block
constexpr int GTT_THREADS = 288;          // 8 consumer warps + 1 producer warp
template <class T> constexpr int gtt_chunk() { return sizeof(T) <= 8 ? 256 : 128; }   // columns per blockIdx.y slice (hs = 16 KB)

// CTA (bx, by): rows [bx*RC, (bx+1)*RC) in blocks of RS rows, columns [by*chunk, ...).  For every row block the consumers
// first stage q's block in shared memory, then column after column multiply the 16 KB stage the producer delivered.
template <class T>
__global__ void __launch_bounds__(GTT_THREADS)
gemv_t_tma_kernel(long Lp, int l, const T* __restrict__ V, long ldv, const T* __restrict__ q, T* __restrict__ part, int lpad, int chunk,
                  long RC) {
  constexpr int VEC = Pack<T>::N;
  constexpr int RS = GTT_STAGE_BYTES / (int)sizeof(T);
  constexpr int S = GTT_STAGES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* stage = reinterpret_cast<T*>(smem_raw);          // [S][RS]
  T* qs = stage + (size_t)S * RS;                     // [RS]
  T* hs = qs + RS;                                    // [8][chunk]
  __shared__ __align__(8) uint64_t full[S], empty[S];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int c_begin = blockIdx.y * chunk, c_end = min(l, c_begin + chunk), nc = c_end - c_begin;
  const long r_lo = (long)blockIdx.x * RC, r_hi = min(Lp, r_lo + RC);
  const long nrows = r_hi > r_lo ? r_hi - r_lo : 0;   // multiple of VEC (Lp and RC are)
  const int nblk = (int)((nrows + RS - 1) / RS);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 8 * chunk; i += GTT_THREADS) hs[i] = zero_<T>();
  __syncthreads();
  if (w == 8) {
    // ---- producer: one thread, one bulk copy per (row block, column) --------------------------------------------
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int b = 0; b < nblk; ++b) {
        const long rows = min((long)RS, nrows - (long)b * RS);
        const T* src = V + (long)c_begin * ldv + r_lo + (long)b * RS;
        for (int c = 0; c < nc; ++c) {
          mbar_wait(&empty[st], ph ^ 1);              // slot free (first pass: passes immediately)
          mbar_expect(&full[st], (uint32_t)(rows * sizeof(T)));
          bulk_g2s(stage + (size_t)st * RS, src + (long)c * ldv, (uint32_t)(rows * sizeof(T)), &full[st]);
          if (++st == S) { st = 0; ph ^= 1; }
        }
      }
    }
    return;
  }
  // ---- consumers (warps 0..7) ---------------------------------------------------------------------------------------
  int st = 0; uint32_t ph = 0;
  T* hw = hs + w * chunk;
  for (int b = 0; b < nblk; ++b) {
    const long rows = min((long)RS, nrows - (long)b * RS);
    const int npk = (int)(rows / VEC);
    asm volatile("bar.sync 1, 256;" ::: "memory");   // everybody is done with the previous q block
    for (int u = tid; u < npk; u += 256) st_pack(qs + (long)u * VEC, ld_pack(q + r_lo + (long)b * RS + (long)u * VEC));
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int c = 0; c < nc; ++c) {
      mbar_wait(&full[st], ph);
      const T* sv = stage + (size_t)st * RS;
      T acc = zero_<T>();
#pragma unroll 4
      for (int u = tid; u < npk; u += 256) {
        const Pack<T> v = ld_pack(sv + (long)u * VEC), qq = ld_pack(qs + (long)u * VEC);
#pragma unroll
        for (int e = 0; e < VEC; ++e) fma_conj(acc, v.v[e], qq.v[e]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == S) { st = 0; ph ^= 1; }
      acc = warp_sum(acc);
      if (lane == 0) hw[c] = hw[c] + acc;
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  for (int c = tid; c < nc; c += 256) {
    T s = hs[c];
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) s = s + hs[ww * chunk + c];
    part[(long)blockIdx.x * lpad + c_begin + c] = s;
  }
}

// h[c] = sum_g part[g][c], g ascending within each of 8 interleaved sub-sums, then a fixed tree.
template <class T>
__global__ void __launch_bounds__(kThreads)
gemv_t_finalize(int l, int G, const T* __restrict__ part, int lpad, T* __restrict__ h) {
  __shared__ T sm[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  T s = zero_<T>();
  if (c < l)
    for (int g = w; g < G; g += 8) s = s + part[(long)g * lpad + c];
  sm[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < l) {
    T t = sm[0][lane];
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) t = t + sm[ww][lane];
    h[c] = t;
  }
}

// Row-sharded run with mapped peers: the finalize kernel also performs the cross-rank all-reduce of the l coefficients
// (the OpenMP build's CRITICAL sum over threads, dreorth.F:177-197, across GPUs).  Every CTA pushes its 32 local sums into
// every rank's coefficient window over NVLink; the last CTA raises this rank's flag everywhere, waits for all ranks'
// flags in its own window and sums the P contributions in rank order (identical on every rank) into h.
template <class T>
__global__ void __launch_bounds__(kThreads)
gemv_t_finalize_fused(int l, int G, const T* __restrict__ part, int lpad, T* __restrict__ h, void** bases, int rank, int world,
                      int buf, unsigned long long seq, unsigned int* ticket, volatile unsigned int* host_err, long long timeout_cycles) {
  __shared__ T sm[8][33];
  __shared__ bool is_last;
  constexpr size_t kBuf = (size_t)Comm::kMaxRanks * Comm::kCoefMax * 16;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  T s = zero_<T>();
  if (c < l)
    for (int g = w; g < G; g += 8) s = s + part[(long)g * lpad + c];
  sm[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < l) {
    T t = sm[0][lane];
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) t = t + sm[ww][lane];
    for (int r = 0; r < world; ++r) {
      T* dst = reinterpret_cast<T*>(static_cast<char*>(bases[r]) + (size_t)buf * kBuf + ((size_t)rank * Comm::kCoefMax) * 16);
      dst[c] = t;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  if ((int)threadIdx.x < world) {
    unsigned long long* f = reinterpret_cast<unsigned long long*>(static_cast<char*>(bases[threadIdx.x]) + 2 * kBuf);
    *reinterpret_cast<volatile unsigned long long*>(f + buf * Comm::kMaxRanks + rank) = seq;
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(static_cast<char*>(bases[rank]) + 2 * kBuf);
    const long long t0 = clock64();
    while (*reinterpret_cast<const volatile unsigned long long*>(mine + buf * Comm::kMaxRanks + threadIdx.x) != seq) {
      if (clock64() - t0 > timeout_cycles) { *host_err = 1u; break; }   // a peer died; do not hang the GPU
    }
  }
  __syncthreads();
  __threadfence();
  const char* my = static_cast<const char*>(bases[rank]) + (size_t)buf * kBuf;
  for (int cc = threadIdx.x; cc < l; cc += kThreads) {
    T t = zero_<T>();
    for (int q = 0; q < world; ++q) {
      const volatile T* src = reinterpret_cast<const volatile T*>(my + ((size_t)q * Comm::kCoefMax) * 16);
      t = t + ld_volatile(src + cc);
    }
    h[cc] = t;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

template <class T, bool SUB>
__global__ void __launch_bounds__(kThreads, 2)
gemv_n_kernel(long L, int l, const T* __restrict__ V, long ldv, const T* __restrict__ h, real_t<T> cin,
              const T* in, T* out, ReduceWs ws, int want_norm) {
  constexpr int VEC = Pack<T>::N;
  constexpr long TROWS = (long)kThreads * VEC * GN_S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* hsm = reinterpret_cast<T*>(smem_raw);  // [l]
  __shared__ double red[32];
  for (int i = threadIdx.x; i < l; i += kThreads) hsm[i] = h[i];
  __syncthreads();
  double nrm = 0.0;
  const long Lp = (L + VEC - 1) / VEC * VEC;
  for (long t0 = (long)blockIdx.x * TROWS; t0 < L; t0 += (long)gridDim.x * TROWS) {
    Pack<T> acc[GN_S];
    bool ok[GN_S];
    long row[GN_S];
#pragma unroll
    for (int s = 0; s < GN_S; ++s) {
      row[s] = t0 + (long)s * kThreads * VEC + (long)threadIdx.x * VEC;
      ok[s] = row[s] < Lp;
      if (ok[s] && in != nullptr) {
        acc[s] = ld_pack(in + row[s]);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[s].v[e] = cin * acc[s].v[e];
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[s].v[e] = zero_<T>();
      }
    }
    const T* Vr = V + t0 + (long)threadIdx.x * VEC;
    int c0 = 0;
    for (; c0 + GN_CU <= l; c0 += GN_CU) {
      Pack<T> pv[GN_CU][GN_S];
#pragma unroll
      for (int cc = 0; cc < GN_CU; ++cc)
#pragma unroll
        for (int s = 0; s < GN_S; ++s)
          if (ok[s]) pv[cc][s] = ld_pack_stream(Vr + (long)(c0 + cc) * ldv + (long)s * kThreads * VEC);
#pragma unroll
      for (int cc = 0; cc < GN_CU; ++cc) {
        const T hc = hsm[c0 + cc];
#pragma unroll
        for (int s = 0; s < GN_S; ++s)
          if (ok[s]) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              if (SUB) fnma_(acc[s].v[e], pv[cc][s].v[e], hc);
              else fma_(acc[s].v[e], pv[cc][s].v[e], hc);
            }
          }
      }
    }
    for (; c0 < l; ++c0) {
      const T hc = hsm[c0];
#pragma unroll
      for (int s = 0; s < GN_S; ++s)
        if (ok[s]) {
          Pack<T> pv = ld_pack_stream(Vr + (long)c0 * ldv + (long)s * kThreads * VEC);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            if (SUB) fnma_(acc[s].v[e], pv.v[e], hc);
            else fma_(acc[s].v[e], pv.v[e], hc);
          }
        }
    }
#pragma unroll
    for (int s = 0; s < GN_S; ++s)
      if (ok[s]) {
        // rows in [L, Lp) are padding: keep them zero so later dot products ignore them
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (row[s] + e >= L) acc[s].v[e] = zero_<T>();
          nrm += (double)abs2_(acc[s].v[e]);
        }
        st_pack(out + row[s], acc[s]);
      }
  }
  if (want_norm) {
    double tot = block_sum(nrm, red);
    grid_publish(tot, 0.0, ws, 1, red);
  }
}

}  // namespace

// The TMA-staged kernel pays off once every CTA streams many 16 KB stages (long vectors, a few columns at least);
// PROPACK_B200_GEMV_TMA=0 / 1 forces the choice.
inline bool gemv_t_use_tma(long L, int l) {
  static const int forced = [] { const char* e = std::getenv("PROPACK_B200_GEMV_TMA"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  if (forced >= 0) return forced == 1;
  return L >= 262144 && l >= 8;
}

template <class T> void k_gemv_t(Context& c, long L, int l, const T* V, long ldv, const T* q, T* h) {
  if (l <= 0) return;
  constexpr int VEC = Pack<T>::N;
  const long wrows = 32L * VEC * GT_S;
  // column slices of equal width (a 256 + 44 split of l = 300 would leave half the CTAs with 15% of the work)
  const int nchunks = ceil_div(l, GT_CHUNK);
  const int chunk = ceil_div(ceil_div(l, nchunks), GT_CG) * GT_CG;
  int gx = (int)std::min<long>(ceil_div(L, wrows * 8), std::max(1, (2 * c.num_sms) / nchunks));
  if (gx < 1) gx = 1;
  const int lpad = (l + 3) / 4 * 4;
  T* part;
  if (gemv_t_use_tma(L, l)) {
    // TMA-staged: 2 CTAs per SM over (row ranges) x (column slices); a CTA's rows are a multiple of the pack size
    const int tchunks = ceil_div(l, gtt_chunk<T>());
    const int tchunk = ceil_div(l, tchunks);
    const long Lp = (L + VEC - 1) / VEC * VEC;
    gx = std::max(1, (2 * c.num_sms) / tchunks);
    const long RC = (ceil_div(Lp, gx) + 31) / 32 * 32;
    gx = ceil_div(Lp, RC);
    part = static_cast<T*>(c.scratch(sizeof(T) * (size_t)gx * lpad));
    const size_t smem = (size_t)GTT_STAGES * GTT_STAGE_BYTES + GTT_STAGE_BYTES + sizeof(T) * 8 * (size_t)gtt_chunk<T>() + 128;
    static bool tma_attr_set = false;
    if (!tma_attr_set) {
      PB_CUDA(cudaFuncSetAttribute(gemv_t_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tma_attr_set = true;
    }
    gemv_t_tma_kernel<T><<<dim3(gx, tchunks), GTT_THREADS, smem, c.stream>>>(Lp, l, V, ldv, q, part, lpad, tchunk, RC);
    PB_LAUNCH_CHECK();
  } else {
    part = static_cast<T*>(c.scratch(sizeof(T) * (size_t)gx * lpad));
    const size_t smem = sizeof(T) * 8 * GT_CHUNK;
    static bool attr_set = false;
    if (!attr_set) {
      PB_CUDA(cudaFuncSetAttribute(gemv_t_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    gemv_t_kernel<T><<<dim3(gx, nchunks), kThreads, smem, c.stream>>>(L, l, V, ldv, q, part, lpad, chunk);
    PB_LAUNCH_CHECK();
  }
  if (c.dist_reduce && c.coef_table != nullptr && l <= Comm::kCoefMax) {
    // fused: finalize + cross-rank all-reduce of the coefficients over NVLink peer memory, one launch
    c.coef_seq += 1;
    gemv_t_finalize_fused<T><<<ceil_div(l, 32), kThreads, 0, c.stream>>>(l, gx, part, lpad, h, c.coef_table, c.peer_rank, c.peer_world,
                                                                         (int)(c.coef_seq & 1ull), c.coef_seq, c.ticket, c.host_err_dev, c.peer_timeout_cycles);
    PB_LAUNCH_CHECK();
  } else {
    gemv_t_finalize<T><<<ceil_div(l, 32), kThreads, 0, c.stream>>>(l, gx, part, lpad, h);
    PB_LAUNCH_CHECK();
    if (c.dist_reduce)
      Comm::get().allreduce_sum(reinterpret_cast<real_t<T>*>(h), (size_t)l * (scalar_traits<T>::is_complex ? 2 : 1), c.stream);
  }
  c.ctr.launches += 2;
}

template <class T>
void k_gemv_n(Context& c, long L, int l, const T* V, long ldv, const T* h, real_t<T> cin, const T* in, int sgn, T* out,
              Pending* nrm) {
  constexpr int VEC = Pack<T>::N;
  const long trows = (long)kThreads * VEC * GN_S;
  const int grid = c.grid_for(L, (int)trows, 2);
  int c0 = 0;
  do {
    const int lc = std::min(GN_CHUNK, l - c0);
    const bool last = (c0 + lc >= l);
    ReduceWs ws{};
    int want = 0;
    if (last && nrm) { ws = c.new_reduce(nrm); want = 1; }
    const T* in_c = (c0 == 0) ? in : out;
    const real_t<T> cin_c = (c0 == 0) ? cin : real_t<T>(1);
    const size_t smem = sizeof(T) * (size_t)std::max(lc, 1);
    if (sgn < 0)
      gemv_n_kernel<T, true><<<grid, kThreads, smem, c.stream>>>(L, lc, V + (long)c0 * ldv, ldv, h + c0, cin_c, in_c, out, ws, want);
    else
      gemv_n_kernel<T, false><<<grid, kThreads, smem, c.stream>>>(L, lc, V + (long)c0 * ldv, ldv, h + c0, cin_c, in_c, out, ws, want);
    PB_LAUNCH_CHECK();
    c.ctr.launches += 1;
    if (want) c.complete_reduce(*nrm, 1);
    c0 += lc;
  } while (c0 < l);
}

#define PB_INST(T)                                                                                       \
  template void k_gemv_t<T>(Context&, long, int, const T*, long, const T*, T*);                          \
  template void k_gemv_n<T>(Context&, long, int, const T*, long, const T*, real_t<T>, const T*, int, T*, Pending*);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
