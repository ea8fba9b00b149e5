// spmv_lab -- stand-alone micro-benchmark of CSR SpMV kernel organisations on B200 (development tool,
// not part of the library).  It answers: what is the floor for a random-column gather at 10 nnz/row
// (L1TEX wavefront-bound), and which kernel organisation gets closest to it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gpurun_out/spmv_lab tools/spmv_lab.cu
//   ./spmv_lab [rows=1000000] [mean_nnz=10] [iters=20]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Csr { int rows, cols; long nnz; int* rp; int* ci; double* va; };

static inline uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

static Csr make_random(int rows, int cols, double mean, uint64_t seed) {
  std::vector<int> rp(rows + 1); rp[0] = 0;
  uint64_t s = seed;
  for (int r = 0; r < rows; ++r) {  // ~binomial row lengths: mean +- sqrt(mean)
    double u1 = (splitmix(s) >> 11) * (1.0 / 9007199254740992.0), u2 = (splitmix(s) >> 11) * (1.0 / 9007199254740992.0);
    double g = std::sqrt(-2.0 * std::log(u1 + 1e-300)) * std::cos(6.283185307179586 * u2);
    int len = (int)std::lround(mean + std::sqrt(mean) * g);
    if (len < 0) len = 0;
    rp[r + 1] = rp[r] + len;
  }
  long nnz = rp[rows];
  std::vector<int> ci(nnz); std::vector<double> va(nnz);
  for (int r = 0; r < rows; ++r) {
    for (int p = rp[r]; p < rp[r + 1]; ++p) { ci[p] = (int)(splitmix(s) % (uint64_t)cols); va[p] = ((splitmix(s) >> 11) * (1.0 / 9007199254740992.0)) - 0.5; }
    std::sort(ci.begin() + rp[r], ci.begin() + rp[r + 1]);
  }
  Csr A; A.rows = rows; A.cols = cols; A.nnz = nnz;
  CK(cudaMalloc(&A.rp, 4 * (rows + 1) + 64)); CK(cudaMalloc(&A.ci, 4 * nnz + 64)); CK(cudaMalloc(&A.va, 8 * nnz + 64));
  CK(cudaMemcpy(A.rp, rp.data(), 4 * (rows + 1), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(A.ci, ci.data(), 4 * nnz, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(A.va, va.data(), 8 * nnz, cudaMemcpyHostToDevice));
  return A;
}

// cache-operator experiments: streamed operands / gathers that do not allocate L1 lines
__device__ inline int ld_na_i32(const int* p) { int v; asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ inline double ld_na_f64(const double* p) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ inline double ld_gather_na(const double* p) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ inline double ld_gather_el(const double* p) { double v; asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

__device__ inline double warp_sum(double v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }

// ---------------------------------------------------------------------------------------------------
// floors
// ---------------------------------------------------------------------------------------------------
template <int MODE>  // 0: stream only (no gather); 1: gather via ld.global.nc; 2: gather via ld.global.cg ; 3: default ld
__global__ void __launch_bounds__(256) k_floor(long nnz, const int* __restrict__ ci, const double* __restrict__ va, const double* x, double* out) {
  double acc = 0;
  const long stride = (long)gridDim.x * 256 * 4;
  for (long p = (long)blockIdx.x * 1024 + threadIdx.x; p < nnz; p += stride) {
    int c[4]; double a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      long q = p + u * 256;
      if (MODE >= 5) { c[u] = q < nnz ? ld_na_i32(ci + q) : 0; a[u] = q < nnz ? ld_na_f64(va + q) : 0.0; }
      else { c[u] = q < nnz ? __ldcs(ci + q) : 0; a[u] = q < nnz ? __ldcs(va + q) : 0.0; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double xv;
      if (MODE == 0) xv = (double)c[u];
      else if (MODE == 1) xv = __ldg(x + c[u]);
      else if (MODE == 2) xv = __ldcg(x + c[u]);
      else if (MODE == 4) xv = ld_gather_na(x + c[u]);
      else if (MODE == 6) xv = ld_gather_el(x + c[u]);
      else xv = x[c[u]];
      acc = fma(a[u], xv, acc);
    }
  }
  out[(long)blockIdx.x * 256 + threadIdx.x] = acc;
}

// ---------------------------------------------------------------------------------------------------
// K2: sub-warp per row, straight from global
// ---------------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256) k_subwarp(Csr A, const double* __restrict__ x, double* __restrict__ y, double coef, const double* __restrict__ prev, double* nrm2) {
  const int rows_per_cta = 256 / LPR, rows_per_warp = 32 / LPR;
  const int lane = threadIdx.x % LPR;
  double nrm = 0;
  // warp-uniform loop bound: every lane of a warp runs the same number of iterations (full-mask shuffles)
  for (long wbase = (long)blockIdx.x * rows_per_cta + (threadIdx.x >> 5) * rows_per_warp; wbase < A.rows; wbase += (long)gridDim.x * rows_per_cta) {
    const long row = wbase + (threadIdx.x & 31) / LPR;
    double acc = 0;
    const bool ok = row < A.rows;
    if (ok) {
      const int s = __ldg(A.rp + row), e = __ldg(A.rp + row + 1);
      for (int p = s + lane; p < e; p += LPR) acc = fma(__ldcs(A.va + p), __ldg(x + __ldcs(A.ci + p)), acc);
    }
#pragma unroll
    for (int o = LPR >> 1; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LPR);
    if (ok && lane == 0) { if (prev) acc += coef * prev[row]; y[row] = acc; nrm += acc * acc; }
  }
  nrm = warp_sum(nrm);
  if ((threadIdx.x & 31) == 0 && nrm != 0) atomicAdd(nrm2, nrm);
}

// K3: thread per row
__global__ void __launch_bounds__(256) k_thread_row(Csr A, const double* __restrict__ x, double* __restrict__ y, double coef, const double* __restrict__ prev, double* nrm2) {
  double nrm = 0;
  for (long row = (long)blockIdx.x * 256 + threadIdx.x; row < A.rows; row += (long)gridDim.x * 256) {
    const int s = __ldg(A.rp + row), e = __ldg(A.rp + row + 1);
    double acc = 0;
    int p = s;
    for (; p + 3 < e; p += 4) {
      int c0 = A.ci[p], c1 = A.ci[p + 1], c2 = A.ci[p + 2], c3 = A.ci[p + 3];
      double a0 = A.va[p], a1 = A.va[p + 1], a2 = A.va[p + 2], a3 = A.va[p + 3];
      double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
      acc = fma(a0, x0, acc); acc = fma(a1, x1, acc); acc = fma(a2, x2, acc); acc = fma(a3, x3, acc);
    }
    for (; p < e; ++p) acc = fma(A.va[p], __ldg(x + A.ci[p]), acc);
    if (prev) acc += coef * prev[row];
    y[row] = acc; nrm += acc * acc;
  }
  nrm = warp_sum(nrm);
  if ((threadIdx.x & 31) == 0 && nrm != 0) atomicAdd(nrm2, nrm);
}

// ---------------------------------------------------------------------------------------------------
// K4: warp-private groups of 32 rows; products staged in the warp's own shared-memory slice (no CTA barriers)
// ---------------------------------------------------------------------------------------------------
template <int WARPS, int MAXNNZ, int U>
__global__ void __launch_bounds__(WARPS * 32) k_warp_group(Csr A, const double* __restrict__ x, double* __restrict__ y, double coef, const double* __restrict__ prev, double* nrm2) {
  __shared__ double prod[WARPS][MAXNNZ];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* pw = prod[w];
  double nrm = 0;
  const long ngroups = ((long)A.rows + 31) / 32;
  for (long g = (long)blockIdx.x * WARPS + w; g < ngroups; g += (long)gridDim.x * WARPS) {
    const long row = g * 32 + lane;
    const bool ok = row < A.rows;
    const int s = __ldg(A.rp + (ok ? row : A.rows)), e = __ldg(A.rp + (ok ? row + 1 : A.rows));
    const int p0 = __shfl_sync(0xffffffffu, s, 0), pe = __shfl_sync(0xffffffffu, e, 31);
    const int nn = pe - p0;
    double pv = (ok && prev) ? __ldcs(prev + row) : 0.0;
    if (nn <= MAXNNZ) {
      const int* cip = A.ci + p0; const double* vap = A.va + p0;
      int i = lane;
      for (; i + (U - 1) * 32 < nn; i += U * 32) {
        int c[U]; double a[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = __ldcs(cip + i + u * 32);
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = __ldcs(vap + i + u * 32);
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = __ldg(x + c[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) pw[i + u * 32] = a[u] * xv[u];
      }
      for (; i < nn; i += 32) pw[i] = __ldcs(vap + i) * __ldg(x + __ldcs(cip + i));
      __syncwarp();
      double acc = 0;
      for (int q = s - p0; q < e - p0; ++q) acc += pw[q];
      __syncwarp();
      if (ok) { acc += coef * pv; y[row] = acc; nrm += acc * acc; }
    } else {  // long rows: warp per row
      for (int r = 0; r < 32; ++r) {
        const int rs = __shfl_sync(0xffffffffu, s, r), re = __shfl_sync(0xffffffffu, e, r);
        double acc = 0;
        for (int p = rs + lane; p < re; p += 32) acc = fma(__ldcs(A.va + p), __ldg(x + __ldcs(A.ci + p)), acc);
        acc = warp_sum(acc);
        const double pr = __shfl_sync(0xffffffffu, pv, r);
        if (lane == 0 && g * 32 + r < A.rows) { acc += coef * pr; y[g * 32 + r] = acc; nrm += acc * acc; }
      }
    }
  }
  nrm = warp_sum(nrm);
  if (lane == 0 && nrm != 0) atomicAdd(nrm2, nrm);
}

// ---------------------------------------------------------------------------------------------------
// K5: as K4 but the x gathers are issued as 16-byte cp.async.bulk copies (TMA unit, UBLKCP) into shared memory
// ---------------------------------------------------------------------------------------------------
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void mbar_init(uint64_t* b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ inline void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ inline void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ inline void bulk16(void* dst, const void* src, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(smem_u32(dst)), "l"(src), "r"(smem_u32(b)) : "memory");
}
template <int WARPS, int MAXNNZ>
__global__ void __launch_bounds__(WARPS * 32) k_warp_group_tma(Csr A, const double* __restrict__ x, double* __restrict__ y, double coef, const double* __restrict__ prev, double* nrm2) {
  __shared__ __align__(16) double2 stage[WARPS][MAXNNZ];
  __shared__ __align__(8) uint64_t bars[WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double2* sw = stage[w];
  uint64_t* bar = &bars[w];
  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t parity = 0;
  double nrm = 0;
  const long ngroups = ((long)A.rows + 31) / 32;
  for (long g = (long)blockIdx.x * WARPS + w; g < ngroups; g += (long)gridDim.x * WARPS) {
    const long row = g * 32 + lane;
    const bool ok = row < A.rows;
    const int s = __ldg(A.rp + (ok ? row : A.rows)), e = __ldg(A.rp + (ok ? row + 1 : A.rows));
    const int p0 = __shfl_sync(0xffffffffu, s, 0), pe = __shfl_sync(0xffffffffu, e, 31);
    const int nn = pe - p0;
    double pv = (ok && prev) ? __ldcs(prev + row) : 0.0;
    if (nn <= MAXNNZ && nn > 0) {
      const int* cip = A.ci + p0; const double* vap = A.va + p0;
      if (lane == 0) mbar_expect(bar, (uint32_t)nn * 16u);
      __syncwarp();
      for (int i = lane; i < nn; i += 32) {
        const int c = __ldcs(cip + i);
        bulk16(&sw[i], x + (c & ~1), bar);
      }
      mbar_wait(bar, parity); parity ^= 1;
      // products in place: sw[i].x <- a * x[c]
      for (int i = lane; i < nn; i += 32) {
        const int c = __ldcs(cip + i);   // L1/L2 hit (just loaded)
        const double2 v = sw[i];
        sw[i].x = __ldcs(vap + i) * ((c & 1) ? v.y : v.x);
      }
      __syncwarp();
      double acc = 0;
      for (int q = s - p0; q < e - p0; ++q) acc += sw[q].x;
      __syncwarp();
      if (ok) { acc += coef * pv; y[row] = acc; nrm += acc * acc; }
    } else {
      for (int r = 0; r < 32; ++r) {
        const int rs = __shfl_sync(0xffffffffu, s, r), re = __shfl_sync(0xffffffffu, e, r);
        double acc = 0;
        for (int p = rs + lane; p < re; p += 32) acc = fma(__ldcs(A.va + p), __ldg(x + __ldcs(A.ci + p)), acc);
        acc = warp_sum(acc);
        const double pr = __shfl_sync(0xffffffffu, pv, r);
        if (lane == 0 && g * 32 + r < A.rows) { acc += coef * pr; y[g * 32 + r] = acc; nrm += acc * acc; }
      }
    }
  }
  nrm = warp_sum(nrm);
  if (lane == 0 && nrm != 0) atomicAdd(nrm2, nrm);
}

// ---------------------------------------------------------------------------------------------------
// K6: CTA-wide row panel (the library's round-1 organisation, simplified): products of a 256-row panel into
// shared memory with one barrier, then 4 lanes per row.
// ---------------------------------------------------------------------------------------------------
template <int NB, int LPR>
__global__ void __launch_bounds__(256) k_cta_panel(Csr A, const double* __restrict__ x, double* __restrict__ y, double coef, const double* __restrict__ prev, double* nrm2) {
  __shared__ double prod[NB];
  __shared__ int srp[257];
  const int tid = threadIdx.x;
  double nrm = 0;
  const int npanels = (A.rows + 255) / 256;
  for (int b = blockIdx.x; b < npanels; b += gridDim.x) {
    const int r0 = b * 256, nrows = min(256, A.rows - r0);
    for (int i = tid; i <= nrows; i += 256) srp[i] = __ldg(A.rp + r0 + i);
    __syncthreads();
    const int p0 = srp[0], nnzb = min(srp[nrows] - p0, NB);  // (lab: panels longer than NB are truncated)
    for (int i = tid; i < nnzb; i += 256) prod[i] = __ldcs(A.va + p0 + i) * __ldg(x + __ldcs(A.ci + p0 + i));
    __syncthreads();
    const int g = tid / LPR, lg = tid % LPR;
    for (int rb = 0; rb < nrows; rb += 256 / LPR) {
      const int r = rb + g;
      double acc = 0;
      if (r < nrows) { const int s = srp[r] - p0, e = min(srp[r + 1] - p0, NB); for (int q = s + lg; q < e; q += LPR) acc += prod[q]; }
#pragma unroll
      for (int o = LPR >> 1; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LPR);
      if (lg == 0 && r < nrows) { if (prev) acc += coef * prev[r0 + r]; y[r0 + r] = acc; nrm += acc * acc; }
    }
    __syncthreads();
  }
  nrm = warp_sum(nrm);
  if ((tid & 31) == 0 && nrm != 0) atomicAdd(nrm2, nrm);
}


// ---------------------------------------------------------------------------------------------------
// K8: TMA-staged panel pipeline.  One elected thread streams each row panel's (ci, va, rp, prev) slices into a
// shared-memory ring with cp.async.bulk (UBLKCP) + mbarrier; the LSU pipe only carries the x gathers.
// ---------------------------------------------------------------------------------------------------
__device__ inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
template <int NB, int RMAX> struct alignas(128) PanelStage {
  double va[NB + 2];
  double prev[RMAX + 2];
  int ci[NB + 4];
  int rp[RMAX + 8];
};
struct PanelDesc { int r0, p0; };   // panel b covers rows [d[b].r0, d[b+1].r0), non-zeros [d[b].p0, d[b+1].p0)

template <int THREADS, int NB, int RMAX, int STAGES, int U>
__global__ void __launch_bounds__(THREADS) k_tma_panel(Csr A, const PanelDesc* __restrict__ desc, int npanels, const double* __restrict__ x,
                                                       double* __restrict__ y, double coef, const double* __restrict__ prev, double* nrm2) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Stage = PanelStage<NB, RMAX>;
  Stage* st = reinterpret_cast<Stage*>(smem_raw);
  __shared__ __align__(8) uint64_t full[STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nt = (npanels - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // panels of this CTA
  auto issue = [&](int t) {  // thread 0
    if (t >= nt) return;
    const int b = blockIdx.x + t * gridDim.x;
    const PanelDesc d0 = desc[b], d1 = desc[b + 1];
    Stage& S = st[t % STAGES];
    uint64_t* bar = &full[t % STAGES];
    const int nn = d1.p0 - d0.p0, nr = d1.r0 - d0.r0;
    const int pc = d0.p0 & ~3, pv = d0.p0 & ~1, rr = d0.r0 & ~3, rv = d0.r0 & ~1;
    const uint32_t bci = (uint32_t)(((d0.p0 - pc) + nn + 3) & ~3) * 4u, bva = (uint32_t)(((d0.p0 - pv) + nn + 1) & ~1) * 8u;
    const uint32_t brp = (uint32_t)(((d0.r0 - rr) + nr + 1 + 3) & ~3) * 4u, bpr = prev ? (uint32_t)(((d0.r0 - rv) + nr + 1) & ~1) * 8u : 0u;
    mbar_expect(bar, (nn > 0 ? bci + bva : 0u) + brp + bpr);
    if (nn > 0) { bulk_g2s(S.ci, A.ci + pc, bci, bar); bulk_g2s(S.va, A.va + pv, bva, bar); }
    bulk_g2s(S.rp, A.rp + rr, brp, bar);
    if (prev) bulk_g2s(S.prev, prev + rv, bpr, bar);
  };
  if (tid == 0) for (int t = 0; t < STAGES; ++t) issue(t);
  double nrm = 0;
  for (int t = 0; t < nt; ++t) {
    const int b = blockIdx.x + t * gridDim.x;
    const PanelDesc d0 = desc[b];
    const PanelDesc d1 = desc[b + 1];
    Stage& S = st[t % STAGES];
    mbar_wait(&full[t % STAGES], (uint32_t)((t / STAGES) & 1));
    const int nn = d1.p0 - d0.p0, nr = d1.r0 - d0.r0, r0 = d0.r0, p0 = d0.p0;
    const int* sci = S.ci + (p0 & 3);
    double* sva = S.va + (p0 & 1);
    const int* srp = S.rp + (r0 & 3);
    const double* spv = S.prev + (r0 & 1);
    // phase 1: gathers; products overwrite the staged values
    for (int i0 = 0; i0 < nn; i0 += THREADS * U) {
      int c[U]; double xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const int i = i0 + u * THREADS + tid; c[u] = i < nn ? sci[i] : -1; }
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u] = c[u] >= 0 ? __ldg(x + c[u]) : 0.0;
#pragma unroll
      for (int u = 0; u < U; ++u) { const int i = i0 + u * THREADS + tid; if (i < nn) sva[i] *= xv[u]; }
    }
    __syncthreads();
    // phase 2: LPR lanes per row
    {
      const int mean = (nn + nr - 1) / max(nr, 1);
      int lpr = 4; while (lpr < 32 && lpr * 4 < mean) lpr <<= 1;
      const int g = tid / lpr, lg = tid % lpr, gpc = THREADS / lpr;
      for (int rb = 0; rb < nr; rb += gpc) {
        const int r = rb + g;
        double acc = 0;
        if (r < nr) { const int s = srp[r] - p0, e = srp[r + 1] - p0; for (int q = s + lg; q < e; q += lpr) acc += sva[q]; }
        for (int o = lpr >> 1; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, lpr);
        if (lg == 0 && r < nr) { if (prev) acc += coef * spv[r]; y[r0 + r] = acc; nrm += acc * acc; }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes to the stage precede its refill by the async proxy
    __syncthreads();
    if (tid == 0) issue(t + STAGES);
  }
  nrm = warp_sum(nrm);
  if ((tid & 31) == 0 && nrm != 0) atomicAdd(nrm2, nrm);
}

// K2b: sub-warp per row, R rows per sub-warp in flight (independent load chains), persistent grid
template <int LPR, int R>
__global__ void __launch_bounds__(256) k_subwarp_ilp(Csr A, const double* __restrict__ x, double* __restrict__ y, double coef, const double* __restrict__ prev, double* nrm2) {
  const int rows_per_warp = (32 / LPR) * R, rows_per_cta = 8 * rows_per_warp;
  const int lane = threadIdx.x % LPR, sub = (threadIdx.x & 31) / LPR;
  double nrm = 0;
  for (long wbase = (long)blockIdx.x * rows_per_cta + (threadIdx.x >> 5) * rows_per_warp; wbase < A.rows; wbase += (long)gridDim.x * rows_per_cta) {
    int s[R], e[R]; double acc[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const long row = wbase + k * (32 / LPR) + sub;
      const bool ok = row < A.rows;
      s[k] = ok ? __ldg(A.rp + row) + lane : 0; e[k] = ok ? __ldg(A.rp + row + 1) : 0; acc[k] = 0;
    }
    bool any = true;
    while (any) {
      int c[R]; double a[R];
      any = false;
#pragma unroll
      for (int k = 0; k < R; ++k) if (s[k] < e[k]) { c[k] = __ldg(A.ci + s[k]); a[k] = __ldg(A.va + s[k]); }
#pragma unroll
      for (int k = 0; k < R; ++k) if (s[k] < e[k]) { acc[k] = fma(a[k], __ldg(x + c[k]), acc[k]); s[k] += LPR; any |= s[k] < e[k]; }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      double v = acc[k];
#pragma unroll
      for (int o = LPR >> 1; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, LPR);
      const long row = wbase + k * (32 / LPR) + sub;
      if (lane == 0 && row < A.rows) { if (prev) v += coef * prev[row]; y[row] = v; nrm += v * v; }
    }
  }
  nrm = warp_sum(nrm);
  if ((threadIdx.x & 31) == 0 && nrm != 0) atomicAdd(nrm2, nrm);
}


// ---------------------------------------------------------------------------------------------------
// K4g: generalised warp-private row groups.  A group = RPG consecutive rows (RPG = 32/LPR, LPR a power of two
// chosen from the mean row length); phase 1 streams the group's (ci,va) slice coalesced with U-deep batches of
// independent gathers, products go to the warp's shared-memory slice; phase 2 reduces each row with LPR lanes.
// The next group's row pointers are prefetched while the current group's gathers are in flight.
// ---------------------------------------------------------------------------------------------------
template <int WARPS, int MAXNNZ, int U, int NA = 0>
__global__ void __launch_bounds__(WARPS * 32) k_warp_group_g(Csr A, int lpr_log2, const double* __restrict__ x, double* __restrict__ y, double coef,
                                                             const double* __restrict__ prev, double* nrm2) {
  __shared__ double prod[WARPS][MAXNNZ];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int lpr = 1 << lpr_log2, rpg = 32 >> lpr_log2;
  const int sub = lane >> lpr_log2, lg = lane & (lpr - 1);
  double* pw = prod[w];
  double nrm = 0;
  const long ngroups = ((long)A.rows + rpg - 1) / rpg;
  const long gstride = (long)gridDim.x * WARPS;
  long g = (long)blockIdx.x * WARPS + w;
  auto load_rp = [&](long gg, int& s, int& e) {
    const long row = gg * rpg + sub;
    const bool ok = gg < ngroups && row < A.rows;
    s = __ldg(A.rp + (ok ? row : A.rows)); e = __ldg(A.rp + (ok ? row + 1 : A.rows));
  };
  int s, e;
  load_rp(g, s, e);
  for (; g < ngroups; g += gstride) {
    int sn, en;
    load_rp(g + gstride, sn, en);                         // prefetch
    const long row = g * rpg + sub;
    const bool ok = row < A.rows;
    const int p0 = __shfl_sync(0xffffffffu, s, 0), pe = __shfl_sync(0xffffffffu, e, 31);
    const int nn = pe - p0;
    const double pv = (ok && prev && lg == 0) ? __ldcs(prev + row) : 0.0;
    if (nn <= MAXNNZ) {
      const int* cip = A.ci + p0; const double* vap = A.va + p0;
      for (int i0 = 0; i0 < nn; i0 += U * 32) {
        int c[U]; double a[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * 32 + lane; c[u] = i < nn ? (NA >= 1 ? ld_na_i32(cip + i) : __ldcs(cip + i)) : -1; }
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * 32 + lane; a[u] = i < nn ? (NA >= 1 ? ld_na_f64(vap + i) : __ldcs(vap + i)) : 0.0; }
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = c[u] >= 0 ? (NA == 2 ? ld_gather_na(x + c[u]) : NA == 3 ? ld_gather_el(x + c[u]) : __ldg(x + c[u])) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * 32 + lane; if (i < nn) pw[i] = a[u] * xv[u]; }
      }
      __syncwarp();
      double acc = 0;
      for (int q = s - p0 + lg; q < e - p0; q += lpr) acc += pw[q];
      for (int o = lpr >> 1; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, lpr);
      __syncwarp();
      if (ok && lg == 0) { acc += coef * pv; y[row] = acc; nrm += acc * acc; }
    } else {  // a long row in the group: warp per row straight from global memory
      for (int r = 0; r < rpg; ++r) {
        const int rs = __shfl_sync(0xffffffffu, s, r << lpr_log2), re = __shfl_sync(0xffffffffu, e, r << lpr_log2);
        double acc = 0;
        for (int p = rs + lane; p < re; p += 32) acc = fma(__ldcs(A.va + p), __ldg(x + __ldcs(A.ci + p)), acc);
        acc = warp_sum(acc);
        const double pr = __shfl_sync(0xffffffffu, pv, r << lpr_log2);
        if (lane == 0 && g * rpg + r < A.rows) { acc += coef * pr; y[g * rpg + r] = acc; nrm += acc * acc; }
      }
    }
    s = sn; e = en;
  }
  nrm = warp_sum(nrm);
  if (lane == 0 && nrm != 0) atomicAdd(nrm2, nrm);
}

// ---------------------------------------------------------------------------------------------------
struct Timer { cudaEvent_t a, b; Timer() { cudaEventCreate(&a); cudaEventCreate(&b); } };

int main(int argc, char** argv) {
  const int rows = argc > 1 ? atoi(argv[1]) : 1000000;
  const double mean = argc > 2 ? atof(argv[2]) : 10.0;
  const int iters = argc > 3 ? atoi(argv[3]) : 20;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs; rows=%d mean nnz/row=%.1f\n", prop.name, sms, rows, mean);
  Csr A = make_random(rows, rows, mean, 1), B = make_random(rows, rows, mean, 2);
  double *x, *y, *prev, *out, *nrm, *ref;
  CK(cudaMalloc(&x, 8L * rows)); CK(cudaMalloc(&y, 8L * rows)); CK(cudaMalloc(&prev, 8L * rows + 64)); CK(cudaMalloc(&ref, 8L * rows));
  CK(cudaMalloc(&out, 8L * 4096 * 256)); CK(cudaMalloc(&nrm, 8));
  std::vector<double> hx(rows);
  uint64_t s = 7; for (auto& v : hx) v = ((splitmix(s) >> 11) * (1.0 / 9007199254740992.0)) - 0.5;
  CK(cudaMemcpy(x, hx.data(), 8L * rows, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(prev, hx.data(), 8L * rows, cudaMemcpyHostToDevice));
  const double bytes = A.nnz * 12.0 + 4.0 * (rows + 1) + 8.0 * rows * 3;   // ci,va,rp, x, y, prev
  Timer T;
  std::vector<double> href(rows), hy(rows);
  bool have_ref = false;

  auto run = [&](const char* name, auto launch, bool check) {
    // alternate the two matrices so neither stays L2-resident between launches (as A / A^T alternate in a solve)
    for (int w = 0; w < 3; ++w) { launch(A, x, y); launch(B, y, ref); }
    CK(cudaDeviceSynchronize());
    cudaEventRecord(T.a);
    for (int it = 0; it < iters; ++it) { launch(A, x, y); launch(B, y, ref); }
    cudaEventRecord(T.b);
    CK(cudaEventSynchronize(T.b));
    float ms; cudaEventElapsedTime(&ms, T.a, T.b);
    const double us = ms * 1e3 / (2 * iters);
    double err = -1;
    if (check) {
      launch(A, x, y); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hy.data(), y, 8L * rows, cudaMemcpyDeviceToHost));
      if (!have_ref) { href = hy; have_ref = true; err = 0; }
      else { err = 0; for (int i = 0; i < rows; ++i) err = std::max(err, std::fabs(hy[i] - href[i])); }
    }
    printf("%-34s %8.1f us  %7.1f GB/s (algorithmic)  maxdiff=%g\n", name, us, bytes / us * 1e-3, err);
    CK(cudaGetLastError());
  };

  const int g4 = sms * 8;
  run("floor stream-only", [&](Csr& M, double* xi, double* yo) { k_floor<0><<<g4, 256>>>(M.nnz, M.ci, M.va, xi, out); }, false);
  run("floor gather ld.nc", [&](Csr& M, double* xi, double* yo) { k_floor<1><<<g4, 256>>>(M.nnz, M.ci, M.va, xi, out); }, false);
  run("floor gather ld.nc.L1::no_allocate", [&](Csr& M, double* xi, double* yo) { k_floor<4><<<g4, 256>>>(M.nnz, M.ci, M.va, xi, out); }, false);
  run("floor streams no_allocate + gather ld", [&](Csr& M, double* xi, double* yo) { k_floor<5><<<g4, 256>>>(M.nnz, M.ci, M.va, xi, out); }, false);
  run("floor streams no_allocate + gather evict_last", [&](Csr& M, double* xi, double* yo) { k_floor<6><<<g4, 256>>>(M.nnz, M.ci, M.va, xi, out); }, false);
  int lprl = 0; { double mn = (double)A.nnz / A.rows; int rpg = 32; while (rpg > 1 && rpg * 1.3 * mn > 512) rpg >>= 1; while ((32 >> lprl) > rpg) ++lprl; }
#define RUN_K4G(W, N, U, NA, CPS, CARVE)                                                                               \
  {                                                                                                                    \
    CK(cudaFuncSetAttribute(k_warp_group_g<W, N, U, NA>, cudaFuncAttributePreferredSharedMemoryCarveout, CARVE));      \
    char nm[96]; snprintf(nm, 96, "K4g W%d N%d U%d NA%d x%d carve=%d", W, N, U, NA, CPS, CARVE);                       \
    run(nm, [&](Csr& M, double* xi, double* yo) { k_warp_group_g<W, N, U, NA><<<sms * CPS, W * 32>>>(M, lprl, xi, yo, -0.5, prev, nrm); }, true); \
  }
  RUN_K4G(8, 512, 4, 0, 3, 45)
  RUN_K4G(8, 512, 4, 1, 3, 45)
  RUN_K4G(8, 512, 4, 2, 3, 45)
  RUN_K4G(8, 512, 4, 3, 3, 45)
  RUN_K4G(8, 512, 4, 1, 4, 60)
  RUN_K4G(8, 512, 8, 1, 3, 45)
  RUN_K4G(8, 512, 6, 0, 3, 45)
  RUN_K4G(8, 512, 6, 1, 3, 45)
  return 0;

}
