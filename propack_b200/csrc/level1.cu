// propack_b200 -- level-1 vector kernels and the LAPACK-compatible random start vector.
//
// Reference: the "blasext" layer double/dblasext.F (pdnrm2 :6, pdscal :38, pdaxpy :92, pddot :121,
// pdzero :202), dsafescal.F:4-55, and dgetu0.F:66-70 (dlarnv(idist=2, iseed=(1,3,5,7)) + pdnrm2).
// All are pure HBM streams; every vector access is a 128-bit pack, norms / dots are reduced in a
// fixed order and published by the last CTA (common.cuh: grid_publish).
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace pb {

namespace {

constexpr int L1_S = 2;  // packs per thread per iteration

template <class T>
__global__ void __launch_bounds__(kThreads)
scal_kernel(long n, T* __restrict__ x, real_t<T> a) {
  constexpr int VEC = Pack<T>::N;
  const long np = (n + VEC - 1) / VEC;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < np; i += (long)gridDim.x * kThreads) {
    Pack<T> p = ld_pack(x + i * VEC);
#pragma unroll
    for (int e = 0; e < VEC; ++e) p.v[e] = a * p.v[e];  // padding stays 0
    st_pack(x + i * VEC, p);
  }
}

// Fused "normalise + all-gather" of a row-sharded run, in two kernels:
//   scal_local_kernel (main stream)  x <- a*x, and a copy into this rank's own slice of its gather buffer;
//   push_kernel (side stream, a few CTAs)  the scaled slice goes to every other rank's gather buffer over NVLink peer
//   memory (bases[r] = rank r's buffer), one destination after the other in ring order (rank+1, rank+2, ...), and each
//   destination's arrival flag is raised as soon as ITS copy is complete system-wide -- so slices land at a consumer in
//   the order rank-1, rank-2, ... and its phase-split SpMV (sell.cu) starts on the early ones while the rest is in flight.
// The push is NVLink-bound, not SM-bound: a thin grid leaves the SMs to the SpMV that runs concurrently.
// (at most 64 registers: the kernel must fit the one CTA slot per SM that the flag-polling SpMV CTAs leave free, or the
// consumers would spin on slices nobody can send)
template <class T, int PU>   // PU = packs in flight per thread
__global__ void __launch_bounds__(kThreads, 4)
push_kernel(long n, long ld, const T* __restrict__ x, void** bases, int rank, int world, unsigned int* tickets,
            unsigned long long epoch) {
  constexpr int VEC = Pack<T>::N;
  __shared__ bool is_last;
  const long np = (n + VEC - 1) / VEC;
  const long stride = (long)gridDim.x * kThreads;
  for (int d = 1; d < world; ++d) {
    const int dest = (rank + d) % world;
    T* dst = static_cast<T*>(bases[dest]) + (long)rank * ld;
    for (long i0 = (long)blockIdx.x * kThreads + threadIdx.x; i0 < np; i0 += stride * PU) {
      Pack<T> p[PU];
#pragma unroll
      for (int u = 0; u < PU; ++u) if (i0 + u * stride < np) p[u] = ld_pack(x + (i0 + u * stride) * VEC);
#pragma unroll
      for (int u = 0; u < PU; ++u) if (i0 + u * stride < np) st_pack(dst + (i0 + u * stride) * VEC, p[u]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(tickets + d, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
      unsigned long long* flags = reinterpret_cast<unsigned long long*>(static_cast<T*>(bases[dest]) + (long)world * ld);
      *reinterpret_cast<volatile unsigned long long*>(flags + rank) = epoch;
      tickets[d] = 0u;
    }
  }
}

// x <- a*x and a copy into this rank's own slice of its gather buffer (main-stream half of the push)
template <class T>
__global__ void __launch_bounds__(kThreads)
scal_local_kernel(long n, T* __restrict__ x, real_t<T> a, T* __restrict__ self, unsigned long long* epoch_out, unsigned long long epoch) {
  constexpr int VEC = Pack<T>::N;
  const long np = (n + VEC - 1) / VEC;
  // copy-engine transport: the arrival flags the peers receive are copies of this word (stream-ordered behind the slice)
  if (epoch_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *epoch_out = epoch;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < np; i += (long)gridDim.x * kThreads) {
    Pack<T> p = ld_pack(x + i * VEC);
#pragma unroll
    for (int e = 0; e < VEC; ++e) p.v[e] = a * p.v[e];
    st_pack(x + i * VEC, p);
    st_pack(self + i * VEC, p);
  }
}

// Consumer side of the fused all-gather: returns once the slices of the ranks in `src_mask` (epoch `epoch`) have landed.
__global__ void wait_flags_kernel(const unsigned long long* flags, unsigned int src_mask, unsigned long long epoch,
                                  volatile unsigned int* host_err, long long timeout_cycles) {
  if (!((src_mask >> threadIdx.x) & 1u)) return;
  const long long t0 = clock64();
  while (*reinterpret_cast<const volatile unsigned long long*>(flags + threadIdx.x) < epoch) {
    if (clock64() - t0 > timeout_cycles) { *host_err = 1u; break; }   // a peer died; do not hang the GPU
  }
}

template <class T>
__global__ void __launch_bounds__(kThreads)
zero_kernel(long n, T* __restrict__ x) {
  constexpr int VEC = Pack<T>::N;
  const long np = (n + VEC - 1) / VEC;
  Pack<T> z;
#pragma unroll
  for (int e = 0; e < VEC; ++e) z.v[e] = zero_<T>();
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < np; i += (long)gridDim.x * kThreads) st_pack(x + i * VEC, z);
}

// MODE 0: y += a*x, publish ||y||.  MODE 1: publish conj(x).y.  MODE 2: publish ||x||.
template <class T, int MODE>
__global__ void __launch_bounds__(kThreads)
reduce_kernel(long n, T a, const T* __restrict__ x, T* y, ReduceWs ws) {
  constexpr int VEC = Pack<T>::N;
  __shared__ double red[32];
  const long np = (n + VEC - 1) / VEC;
  double sr = 0.0, si = 0.0;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < np; i += (long)gridDim.x * kThreads) {
    Pack<T> px = ld_pack(x + i * VEC);
    if (MODE == 0) {
      Pack<T> py = ld_pack(y + i * VEC);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        fma_(py.v[e], a, px.v[e]);
        sr += (double)abs2_(py.v[e]);
      }
      st_pack(y + i * VEC, py);
    } else if (MODE == 1) {
      Pack<T> py = ld_pack(y + i * VEC);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        T acc = zero_<T>();
        fma_conj(acc, px.v[e], py.v[e]);
        sr += (double)real_(acc);
        si += (double)imag_(acc);
      }
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) sr += (double)abs2_(px.v[e]);
    }
  }
  double tr = block_sum(sr, red);
  double ti = block_sum(si, red);
  grid_publish(tr, ti, ws, MODE == 1 ? 0 : 1, red);
}

// --- xLARNV(idist = 2) ---------------------------------------------------------------------------
// DLARUV (Lapack_Util/dlaruv.f:335-363) multiplies the 48-bit seed by a^i mod 2^48 for the i-th value
// of a call (a = 33952834046453) and returns the last state, so element i (1-based) of the stream
// is s0 * a^i mod 2^48: every thread jumps straight to its element with a square-and-multiply.
// Value = R*(IT1 + R*(IT2 + R*(IT3 + R*IT4))), R = 2^-12, evaluated in working precision without
// fma contraction (exact in double; three roundings in single, as in slaruv.f:354-355).
__device__ inline unsigned long long laruv_pow(unsigned long long e) {
  const unsigned long long M = (1ull << 48) - 1;
  unsigned long long base = 33952834046453ull, r = 1ull;
  while (e) {
    if (e & 1ull) r = (r * base) & M;
    base = (base * base) & M;
    e >>= 1;
  }
  return r;
}
__device__ inline double laruv_val(unsigned long long s, double) {
  const double r = 1.0 / 4096.0;
  double t = (double)(s & 4095ull);
  t = __dadd_rn((double)((s >> 12) & 4095ull), __dmul_rn(r, t));
  t = __dadd_rn((double)((s >> 24) & 4095ull), __dmul_rn(r, t));
  t = __dadd_rn((double)((s >> 36) & 4095ull), __dmul_rn(r, t));
  return __dmul_rn(r, t);
}
__device__ inline float laruv_val(unsigned long long s, float) {
  const float r = 1.0f / 4096.0f;
  float t = (float)(s & 4095ull);
  t = __fadd_rn((float)((s >> 12) & 4095ull), __fmul_rn(r, t));
  t = __fadd_rn((float)((s >> 24) & 4095ull), __fmul_rn(r, t));
  t = __fadd_rn((float)((s >> 36) & 4095ull), __fmul_rn(r, t));
  return __fmul_rn(r, t);
}
__device__ inline void larnv_elem(float& out, unsigned long long s0, long i) {
  const unsigned long long M = (1ull << 48) - 1;
  out = __fadd_rn(__fmul_rn(2.0f, laruv_val((s0 * laruv_pow((unsigned long long)i + 1)) & M, 0.0f)), -1.0f);
}
__device__ inline void larnv_elem(double& out, unsigned long long s0, long i) {
  const unsigned long long M = (1ull << 48) - 1;
  out = __dadd_rn(__dmul_rn(2.0, laruv_val((s0 * laruv_pow((unsigned long long)i + 1)) & M, 0.0)), -1.0);
}
template <class R> __device__ inline void larnv_elem(cplx<R>& out, unsigned long long s0, long i) {
  // zlarnv draws 2 reals per element: (2*u(2i-1)-1, 2*u(2i)-1)   (complex16/Lapack_Util/zlarnv.f)
  R re, im;
  larnv_elem(re, s0, 2 * i);
  larnv_elem(im, s0, 2 * i + 1);
  out = cplx<R>(re, im);
}

template <class T>
__global__ void __launch_bounds__(kThreads)
larnv_kernel(long n, T* __restrict__ x, unsigned long long s0, long offset, ReduceWs ws) {
  constexpr int VEC = Pack<T>::N;
  __shared__ double red[32];
  const long np = (n + VEC - 1) / VEC;
  double sr = 0.0;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < np; i += (long)gridDim.x * kThreads) {
    Pack<T> p;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const long idx = i * VEC + e;
      if (idx < n) larnv_elem(p.v[e], s0, offset + idx);
      else p.v[e] = zero_<T>();
      sr += (double)abs2_(p.v[e]);
    }
    st_pack(x + i * VEC, p);
  }
  double tr = block_sum(sr, red);
  grid_publish(tr, 0.0, ws, 1, red);
}

template <class T> int l1_grid(Context& c, long n) {  // >= 1 CTA even for n = 0 (a rank that owns no rows still publishes)
  return c.grid_for((n + Pack<T>::N - 1) / Pack<T>::N, kThreads * L1_S, 8);
}

}  // namespace

template <class T> void k_scal(Context& c, long n, T* x, real_t<T> a) {
  if (n <= 0) return;
  scal_kernel<T><<<l1_grid<T>(c, n), kThreads, 0, c.stream>>>(n, x, a);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}
inline int push_ctas() {   // read at every call: cheap, and lets one process compare settings
  const char* e = std::getenv("PROPACK_B200_PUSH_CTAS");
  const int v = e ? std::atoi(e) : 64;   // measured on 8 GPUs, config 5: 32 -> 1050 ms, 48/64 -> 954-977, 96 -> 990, 128/148 -> 1005 (pre y-prefetch)
  return std::min(148, std::max(1, v));
}
template <class T>
void k_scal_push(Context& c, long n, long ld, T* x, real_t<T> a, void** bases_dev, int rank, int world, unsigned long long epoch,
                 T* self_slice) {
  if (n > 0) {
    scal_local_kernel<T><<<l1_grid<T>(c, n), kThreads, 0, c.stream>>>(n, x, a, self_slice, nullptr, 0ull);
    PB_LAUNCH_CHECK();
    c.ctr.launches += 1;
  }
  // (a rank that owns no rows still raises its flags: the consumers wait for every source)
  PB_CUDA(cudaEventRecord(c.ev_fork, c.stream));
  PB_CUDA(cudaStreamWaitEvent(c.stream2, c.ev_fork, 0));
  const int grid = (int)std::min<long>(push_ctas(), std::max<long>(1, ((n + Pack<T>::N - 1) / Pack<T>::N + kThreads - 1) / kThreads));
  push_kernel<T, 4><<<grid, kThreads, 0, c.stream2>>>(n, ld, x, bases_dev, rank, world, c.tickets8, epoch);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}
// Main-stream half of the copy-engine all-gather: x <- a*x, copy into this rank's own slice of its gather buffer, and
// the epoch word the flag copies will carry (launched even for n = 0: the flags must still be raised).
template <class T>
void k_scal_local(Context& c, long n, T* x, real_t<T> a, T* self_slice, unsigned long long* epoch_dev, unsigned long long epoch) {
  scal_local_kernel<T><<<l1_grid<T>(c, std::max<long>(n, 1)), kThreads, 0, c.stream>>>(n, x, a, self_slice, epoch_dev, epoch);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}
void k_wait_flags(Context& c, const unsigned long long* flags, unsigned int src_mask, unsigned long long epoch) {
  if (src_mask == 0u) return;
  wait_flags_kernel<<<1, 32, 0, c.stream>>>(flags, src_mask, epoch, c.host_err_dev, c.peer_timeout_cycles);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}
template <class T> void k_zero(Context& c, long n, T* x) {
  if (n <= 0) return;
  zero_kernel<T><<<l1_grid<T>(c, n), kThreads, 0, c.stream>>>(n, x);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}
template <class T> void k_axpy_nrm(Context& c, long n, T a, const T* x, T* y, Pending* nrm) {
  ReduceWs ws = c.new_reduce(nrm);
  reduce_kernel<T, 0><<<l1_grid<T>(c, n), kThreads, 0, c.stream>>>(n, a, x, y, ws);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
  c.complete_reduce(*nrm, 1);
}
template <class T> void k_dotc(Context& c, long n, const T* x, const T* y, Pending* out) {
  ReduceWs ws = c.new_reduce(out);
  reduce_kernel<T, 1><<<l1_grid<T>(c, n), kThreads, 0, c.stream>>>(n, zero_<T>(), x, const_cast<T*>(y), ws);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
  c.complete_reduce(*out, 0);
}
template <class T> void k_nrm2(Context& c, long n, const T* x, Pending* out) {
  ReduceWs ws = c.new_reduce(out);
  reduce_kernel<T, 2><<<l1_grid<T>(c, n), kThreads, 0, c.stream>>>(n, zero_<T>(), x, nullptr, ws);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
  c.complete_reduce(*out, 1);
}
template <class T> void k_larnv_nrm(Context& c, long n, T* x, const int iseed[4], Pending* nrm, long offset) {
  const unsigned long long s0 = ((unsigned long long)iseed[0] << 36) | ((unsigned long long)iseed[1] << 24) |
                                ((unsigned long long)iseed[2] << 12) | (unsigned long long)iseed[3];
  ReduceWs ws = c.new_reduce(nrm);
  larnv_kernel<T><<<l1_grid<T>(c, std::max<long>(n, 1)), kThreads, 0, c.stream>>>(n, x, s0, offset, ws);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
  c.complete_reduce(*nrm, 1);
}

#define PB_INST(T)                                                                \
  template void k_scal<T>(Context&, long, T*, real_t<T>);                         \
  template void k_zero<T>(Context&, long, T*);                                    \
  template void k_scal_push<T>(Context&, long, long, T*, real_t<T>, void**, int, int, unsigned long long, T*); \
  template void k_scal_local<T>(Context&, long, T*, real_t<T>, T*, unsigned long long*, unsigned long long); \
  template void k_axpy_nrm<T>(Context&, long, T, const T*, T*, Pending*);         \
  template void k_dotc<T>(Context&, long, const T*, const T*, Pending*);          \
  template void k_nrm2<T>(Context&, long, const T*, Pending*);                    \
  template void k_larnv_nrm<T>(Context&, long, T*, const int*, Pending*, long);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
