// propack_b200 -- common device/host definitions for the sm_100a kernels.
//
// Scalar model: T in {float, double, cplx<float>, cplx<double>} mirrors the four PROPACK
// precision directories (single/ double/ complex8/ complex16/ of the reference); R = real
// type of T (Sigma, bnd, B, doption stay real in the complex variants: zlansvd.F:98-109).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>

namespace pb {

// --------------------------------------------------------------------------------------------
// errors: every CUDA failure throws; the C-ABI layer turns it into info = -100 - cudaError.
// --------------------------------------------------------------------------------------------
struct CudaError : std::runtime_error {
  cudaError_t code;
  CudaError(cudaError_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void cuda_check(cudaError_t e, const char* expr, const char* file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "propack_b200: CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, expr);
    throw CudaError(e, buf);
  }
}
#define PB_CUDA(expr) ::pb::cuda_check((expr), #expr, __FILE__, __LINE__)
#define PB_LAUNCH_CHECK() PB_CUDA(cudaGetLastError())

// --------------------------------------------------------------------------------------------
// complex scalar (binary compatible with Fortran COMPLEX / std::complex)
// --------------------------------------------------------------------------------------------
template <class R> struct alignas(2 * sizeof(R)) cplx {
  R x, y;
  __host__ __device__ cplx() {}
  __host__ __device__ cplx(R re, R im = R(0)) : x(re), y(im) {}
};
template <class R> __host__ __device__ inline cplx<R> operator+(cplx<R> a, cplx<R> b) { return cplx<R>(a.x + b.x, a.y + b.y); }
template <class R> __host__ __device__ inline cplx<R> operator-(cplx<R> a, cplx<R> b) { return cplx<R>(a.x - b.x, a.y - b.y); }
template <class R> __host__ __device__ inline cplx<R> operator-(cplx<R> a) { return cplx<R>(-a.x, -a.y); }
template <class R> __host__ __device__ inline cplx<R> operator*(cplx<R> a, cplx<R> b) {
  return cplx<R>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <class R> __host__ __device__ inline cplx<R> operator*(R a, cplx<R> b) { return cplx<R>(a * b.x, a * b.y); }
template <class R> __host__ __device__ inline cplx<R> operator*(cplx<R> b, R a) { return cplx<R>(a * b.x, a * b.y); }
template <class R> __host__ __device__ inline cplx<R>& operator+=(cplx<R>& a, cplx<R> b) { a.x += b.x; a.y += b.y; return a; }
template <class R> __host__ __device__ inline cplx<R>& operator-=(cplx<R>& a, cplx<R> b) { a.x -= b.x; a.y -= b.y; return a; }

template <class T> struct scalar_traits;
template <> struct scalar_traits<float> { using real = float; static constexpr bool is_complex = false; static constexpr char prefix = 's'; };
template <> struct scalar_traits<double> { using real = double; static constexpr bool is_complex = false; static constexpr char prefix = 'd'; };
template <> struct scalar_traits<cplx<float>> { using real = float; static constexpr bool is_complex = true; static constexpr char prefix = 'c'; };
template <> struct scalar_traits<cplx<double>> { using real = double; static constexpr bool is_complex = true; static constexpr char prefix = 'z'; };
template <class T> using real_t = typename scalar_traits<T>::real;

__host__ __device__ inline float conj_(float a) { return a; }
__host__ __device__ inline double conj_(double a) { return a; }
template <class R> __host__ __device__ inline cplx<R> conj_(cplx<R> a) { return cplx<R>(a.x, -a.y); }
__host__ __device__ inline float abs2_(float a) { return a * a; }
__host__ __device__ inline double abs2_(double a) { return a * a; }
template <class R> __host__ __device__ inline R abs2_(cplx<R> a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ inline float real_(float a) { return a; }
__host__ __device__ inline double real_(double a) { return a; }
template <class R> __host__ __device__ inline R real_(cplx<R> a) { return a.x; }
__host__ __device__ inline float imag_(float) { return 0.f; }
__host__ __device__ inline double imag_(double) { return 0.0; }
template <class R> __host__ __device__ inline R imag_(cplx<R> a) { return a.y; }
template <class T> __host__ __device__ inline T zero_() { return T(real_t<T>(0)); }

// acc += conj(a) * b   /   acc += a * b  (explicit fma chains so real and complex share kernels)
__device__ inline void fma_conj(float& acc, float a, float b) { acc = fmaf(a, b, acc); }
__device__ inline void fma_conj(double& acc, double a, double b) { acc = fma(a, b, acc); }
template <class R> __device__ inline void fma_conj(cplx<R>& acc, cplx<R> a, cplx<R> b) {
  acc.x += a.x * b.x + a.y * b.y;
  acc.y += a.x * b.y - a.y * b.x;
}
__device__ inline void fma_(float& acc, float a, float b) { acc = fmaf(a, b, acc); }
__device__ inline void fma_(double& acc, double a, double b) { acc = fma(a, b, acc); }
template <class R> __device__ inline void fma_(cplx<R>& acc, cplx<R> a, cplx<R> b) {
  acc.x += a.x * b.x - a.y * b.y;
  acc.y += a.x * b.y + a.y * b.x;
}
// acc -= a * b
__device__ inline void fnma_(float& acc, float a, float b) { acc = fmaf(-a, b, acc); }
__device__ inline void fnma_(double& acc, double a, double b) { acc = fma(-a, b, acc); }
template <class R> __device__ inline void fnma_(cplx<R>& acc, cplx<R> a, cplx<R> b) {
  acc.x -= a.x * b.x - a.y * b.y;
  acc.y -= a.x * b.y + a.y * b.x;
}

// --------------------------------------------------------------------------------------------
// 128-bit packs: every bulk HBM access in the reorthogonalisation / level-1 kernels is one
// LDG.128 / STG.128 per lane (device columns are 256-byte aligned, leading dims padded).
// --------------------------------------------------------------------------------------------
__device__ inline float ld_volatile(const volatile float* p) { return *p; }
__device__ inline double ld_volatile(const volatile double* p) { return *p; }
template <class R> __device__ inline cplx<R> ld_volatile(const volatile cplx<R>* p) {
  return cplx<R>(*reinterpret_cast<const volatile R*>(&p->x), *reinterpret_cast<const volatile R*>(&p->y));
}

template <class T> struct alignas(16) Pack {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};
template <class T> __device__ inline Pack<T> ld_pack(const T* p) { return *reinterpret_cast<const Pack<T>*>(p); }
// read-only (LDG.E.128.CONSTANT) variant for basis / matrix data a kernel only reads
template <class T> __device__ inline Pack<T> ld_pack_stream(const T* p) {
  Pack<T> r;
  float4 f = __ldg(reinterpret_cast<const float4*>(p));  // read-only path, default L2 policy
  *reinterpret_cast<float4*>(&r) = f;
  return r;
}
template <class T> __device__ inline void st_pack(T* p, const Pack<T>& v) { *reinterpret_cast<Pack<T>*>(p) = v; }

// --------------------------------------------------------------------------------------------
// warp / block reductions (fixed order => run-to-run deterministic)
// --------------------------------------------------------------------------------------------
__device__ inline float shfl_xor_(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ inline double shfl_xor_(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <class R> __device__ inline cplx<R> shfl_xor_(cplx<R> v, int m) { return cplx<R>(shfl_xor_(v.x, m), shfl_xor_(v.y, m)); }
__device__ inline float shfl_down_(float v, int d, int w = 32) { return __shfl_down_sync(0xffffffffu, v, d, w); }
__device__ inline double shfl_down_(double v, int d, int w = 32) { return __shfl_down_sync(0xffffffffu, v, d, w); }
template <class R> __device__ inline cplx<R> shfl_down_(cplx<R> v, int d, int w = 32) {
  return cplx<R>(shfl_down_(v.x, d, w), shfl_down_(v.y, d, w));
}
__device__ inline float shfl_idx_(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ inline double shfl_idx_(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <class R> __device__ inline cplx<R> shfl_idx_(cplx<R> v, int src) { return cplx<R>(shfl_idx_(v.x, src), shfl_idx_(v.y, src)); }
template <class T> __device__ inline T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + shfl_xor_(v, o);
  return v;
}
// Sum over the CTA; result valid in thread 0.  `red` = shared scratch of >= 32 T.
template <class T> __device__ inline T block_sum(T v, T* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  T s = zero_<T>();
  if (wid == 0) {
    s = lane < nw ? red[lane] : zero_<T>();
    s = warp_sum(s);
  }
  return s;
}

// --------------------------------------------------------------------------------------------
// Result slots.  Each norm / dot kernel ends with the "last CTA" summing the per-CTA partials
// in index order and publishing the scalar (a) in device memory, for kernels queued behind
// it, and (b) in host-mapped pinned memory followed by a sequence number the host spins on --
// one PCIe posted write instead of a cudaStreamSynchronize per scalar (dlanbpro.F branches on
// every pdnrm2/pddot result: SURVEY 3.3).
// --------------------------------------------------------------------------------------------
struct ScalarSlot {
  double re, im;               // value (always stored as double; exact for float)
  volatile unsigned long long seq;  // bumped after the value is visible
  unsigned long long pad;
};

struct ReduceWs {              // workspace of one in-flight grid reduction
  double* partials;            // [2 * max_ctas]
  unsigned int* ticket;        // arrival counter, self-resetting
  ScalarSlot* dev_slot;        // device copy of the result
  ScalarSlot* host_slot;       // mapped pinned copy (device address)
  unsigned long long seq;      // sequence number this launch publishes
  int local_only;              // row-sharded run: store this rank's raw partial in dev_slot only; the host-side
                               // Context::complete_reduce() all-reduces it across ranks and publishes
  // fused cross-rank reduction over NVLink peer memory (row-sharded run with mapped peers): the last CTA writes this
  // rank's partial into every rank's window, waits for the other ranks' partials in its own, sums them in rank order
  // (identical on every rank) and publishes -- no NCCL call, no extra launch.
  void** peer_table;           // device array: peer_table[r] = rank r's PeerSlot window, or null
  int rank, world, slot;
  volatile unsigned int* host_err;  // mapped pinned error word (set when a peer never shows up)
  long long timeout_cycles;         // give up on a peer after this many SM clocks (do not hang the GPU)
};

struct PeerSlotDev {           // layout of comm.hpp::PeerSlot
  double re, im;
  unsigned long long seq;
  unsigned long long pad;
};
constexpr int kPeerSlotsPerRank = 64;

// Called by every thread of every CTA; `v` = this CTA's partial (valid in thread 0; real part
// and imag part for complex dots).  kind: 0 = store sum, 1 = store sqrt(sum) (norms).
__device__ inline void grid_publish(double vre, double vim, const ReduceWs& ws, int kind, double* red_smem) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    ws.partials[2 * blockIdx.x] = vre;
    ws.partials[2 * blockIdx.x + 1] = vim;
    __threadfence();
    unsigned int t = atomicAdd(ws.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // fixed-order sum: thread t takes partials t, t+B, ...; then block_sum (fixed tree)
  double sr = 0.0, si = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    sr += __ldcg(ws.partials + 2 * i);
    si += __ldcg(ws.partials + 2 * i + 1);
  }
  sr = block_sum(sr, red_smem);
  si = block_sum(si, red_smem);
  if (threadIdx.x == 0) {
    if (ws.peer_table != nullptr) {
      // ---- one-shot all-reduce over peer memory ----------------------------------------------------------
      for (int p = 0; p < ws.world; ++p) {
        PeerSlotDev* d = static_cast<PeerSlotDev*>(ws.peer_table[p]) + ws.rank * kPeerSlotsPerRank + ws.slot;
        d->re = sr; d->im = si;
      }
      __threadfence_system();
      for (int p = 0; p < ws.world; ++p) {
        PeerSlotDev* d = static_cast<PeerSlotDev*>(ws.peer_table[p]) + ws.rank * kPeerSlotsPerRank + ws.slot;
        *reinterpret_cast<volatile unsigned long long*>(&d->seq) = ws.seq;
      }
      double tr = 0.0, ti = 0.0;
      const long long t0 = clock64();
      bool dead = false;
      for (int q = 0; q < ws.world; ++q) {
        PeerSlotDev* m = static_cast<PeerSlotDev*>(ws.peer_table[ws.rank]) + q * kPeerSlotsPerRank + ws.slot;
        while (*reinterpret_cast<volatile unsigned long long*>(&m->seq) != ws.seq) {
          if (clock64() - t0 > ws.timeout_cycles) { dead = true; break; }   // a peer died; do not hang the GPU
        }
        if (dead) break;
        __threadfence();
        tr += *reinterpret_cast<volatile double*>(&m->re);
        ti += *reinterpret_cast<volatile double*>(&m->im);
      }
      if (dead) { *ws.host_err = 1u; tr = ti = 0.0; }
      if (kind == 1) tr = sqrt(tr);
      ws.dev_slot->re = tr; ws.dev_slot->im = ti;
      ws.host_slot->re = tr; ws.host_slot->im = ti;
      __threadfence_system();
      ws.host_slot->seq = ws.seq;
    } else if (ws.local_only) {
      ws.dev_slot->re = sr; ws.dev_slot->im = si;
    } else {
      if (kind == 1) sr = sqrt(sr);
      ws.dev_slot->re = sr; ws.dev_slot->im = si;
      ws.host_slot->re = sr; ws.host_slot->im = si;
      __threadfence_system();
      ws.host_slot->seq = ws.seq;
    }
    *ws.ticket = 0u;
  }
}


constexpr int kThreads = 256;   // CTA size of the streaming kernels
constexpr int kMaxCtas = 4096;  // upper bound on any grid that uses grid_publish

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace pb
