// gemv_lab -- stand-alone micro-benchmark for the tall-skinny GEMV^T behind reorthogonalisation
// (h = V(:,0:l)^T q, V column-major L x l): the register-staged organisation of the library's gemv_t_kernel
// against a TMA-staged one (cp.async.bulk ring + mbarrier, one producer thread, 8 consumer warps).
// Motivation (profiles/r01_summary.md): at l = 537 the library kernel reaches 5.9 TB/s with DRAM 73 % busy, 82 % of the
// stall samples waiting on loads and only 16 warps/SM because the 16 x 16 B in flight per lane live in registers;
// staging through shared memory lets ~100-190 KB per SM be in flight with the register file free.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/bin/gemv_lab tools/gemv_lab.cu
//   ./gemv_lab [L=1000000] [l=300] [iters=10]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ inline double warp_sum(double v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void mbar_init(uint64_t* b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ inline void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ inline void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ inline void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// A: register-staged (the library's organisation, simplified to double): warp owns 256 rows, 4 strips x 4 columns of
// LDG.128 in flight per lane, transposing fold, per-warp column sums in shared memory.
// ---------------------------------------------------------------------------------------------------
__device__ inline double fold4(double a0, double a1, double a2, double a3, int lane) {
  const bool hi16 = lane & 16;
  double s0 = hi16 ? a0 : a2, s1 = hi16 ? a1 : a3;
  double k0 = hi16 ? a2 : a0, k1 = hi16 ? a3 : a1;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  const bool hi8 = lane & 8;
  double s = hi8 ? k0 : k1, k = hi8 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, s, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}
__global__ void __launch_bounds__(256, 2) k_regs(long L, int l, const double* __restrict__ V, long ldv, const double* __restrict__ q,
                                                 double* __restrict__ part, int chunk) {
  extern __shared__ double hs[];  // [8][chunk]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c_begin = blockIdx.y * chunk, c_end = min(l, c_begin + chunk), nc = c_end - c_begin;
  for (int i = threadIdx.x; i < 8 * chunk; i += 256) hs[i] = 0.0;
  __syncthreads();
  double* hw = hs + w * chunk;
  const long gw = (long)blockIdx.x * 8 + w, GW = (long)gridDim.x * 8;
  for (long r0 = gw * 256; r0 < L; r0 += GW * 256) {
    double2 qv[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) qv[s] = *reinterpret_cast<const double2*>(q + r0 + s * 64 + lane * 2);
    const double* Vr = V + (long)c_begin * ldv + r0 + lane * 2;
    for (int c0 = 0; c0 + 4 <= nc; c0 += 4) {
      double2 pv[4][4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int s = 0; s < 4; ++s) pv[cc][s] = __ldg(reinterpret_cast<const double2*>(Vr + (long)(c0 + cc) * ldv + s * 64));
      double acc[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        acc[cc] = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s) { acc[cc] = fma(pv[cc][s].x, qv[s].x, acc[cc]); acc[cc] = fma(pv[cc][s].y, qv[s].y, acc[cc]); }
      }
      const double tot = fold4(acc[0], acc[1], acc[2], acc[3], lane);
      if ((lane & 7) == 0) hw[c0 + (lane >> 3)] += tot;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < nc; c += 256) {
    double s = 0;
    for (int ww = 0; ww < 8; ++ww) s += hs[ww * chunk + c];
    part[(long)blockIdx.x * l + c_begin + c] = s;
  }
}

// ---------------------------------------------------------------------------------------------------
// B: TMA-staged.  A CTA owns a contiguous block of RC rows (its share of L); q for those rows sits in shared memory
// for the whole kernel.  The producer (thread 0 of an extra warp) walks columns (outer) x row chunks of RS rows (inner) and
// issues one cp.async.bulk of RS*8 bytes per stage into a ring of S stages; the 8 consumer warps multiply the stage with
// their q values, keep one accumulator per lane for the current column, and at each column's end add the warp sum to the
// warp's slice of the column sums (no CTA barrier in the loop).  Stage reuse: an "empty" mbarrier with 8 arrivals.
// ---------------------------------------------------------------------------------------------------
template <int RS, int S>
__global__ void __launch_bounds__(288) k_tma(long L, int l, const double* __restrict__ V, long ldv, const double* __restrict__ q,
                                             double* __restrict__ part, long RC) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage = reinterpret_cast<double*>(smem_raw);                 // [S][RS]
  double* qs = stage + (size_t)S * RS;                                 // [RC]
  double* hs = qs + RC;                                                // [8][l]
  __shared__ __align__(8) uint64_t full[S], empty[S];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const long r_lo = (long)blockIdx.x * RC, r_hi = min(L, r_lo + RC);
  const long nrows = r_hi > r_lo ? r_hi - r_lo : 0;                    // multiple of 32 (L and RC are)
  const int nchunk = (int)((nrows + RS - 1) / RS);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (long i = tid; i < nrows; i += 288) qs[i] = q[r_lo + i];
  for (int i = tid; i < 8 * l; i += 288) hs[i] = 0.0;
  __syncthreads();
  if (w == 8) {
    // ---- producer ------------------------------------------------------------------------------------
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int c = 0; c < l; ++c)
        for (int k = 0; k < nchunk; ++k) {
          mbar_wait(&empty[st], ph ^ 1);                                // slot free (first pass: passes immediately)
          const long rows = min((long)RS, nrows - (long)k * RS);
          mbar_expect(&full[st], (uint32_t)(rows * 8));
          bulk_g2s(stage + (size_t)st * RS, V + (long)c * ldv + r_lo + (long)k * RS, (uint32_t)(rows * 8), &full[st]);
          if (++st == S) { st = 0; ph ^= 1; }
        }
    }
    return;
  }
  // ---- consumers (warps 0..7): thread t < 256 reads 16-byte units t, t+256, ... of a stage -----------------
  int st = 0; uint32_t ph = 0;
  double* hw = hs + w * l;
  for (int c = 0; c < l; ++c) {
    double acc = 0.0;
    for (int k = 0; k < nchunk; ++k) {
      mbar_wait(&full[st], ph);
      const long rows = min((long)RS, nrows - (long)k * RS);
      const double2* sv = reinterpret_cast<const double2*>(stage + (size_t)st * RS);
      const double2* sq = reinterpret_cast<const double2*>(qs + (size_t)k * RS);
#pragma unroll 4
      for (int u = tid; u < rows / 2; u += 256) {
        const double2 v = sv[u], qq = sq[u];
        acc = fma(v.x, qq.x, acc); acc = fma(v.y, qq.y, acc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      if (++st == S) { st = 0; ph ^= 1; }
    }
    acc = warp_sum(acc);
    if (lane == 0) hw[c] = acc;
  }
  // the 8 consumer warps meet on a named barrier (the producer warp has left)
  asm volatile("bar.sync 1, 256;" ::: "memory");
  for (int c = tid; c < l; c += 256) {
    double s = 0;
    for (int ww = 0; ww < 8; ++ww) s += hs[ww * l + c];
    part[(long)blockIdx.x * l + c] = s;
  }
}

__global__ void k_finalize(int l, int G, const double* __restrict__ part, double* __restrict__ h) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= l) return;
  double s = 0;
  for (int g = 0; g < G; ++g) s += part[(long)g * l + c];
  h[c] = s;
}

int main(int argc, char** argv) {
  const long L = argc > 1 ? atol(argv[1]) : 1000000;
  const int l = argc > 2 ? atoi(argv[2]) : 300;
  const int iters = argc > 3 ? atoi(argv[3]) : 10;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const long ld = (L + 255) / 256 * 256;   // zero-padded so that whole 256-row warp blocks can be read
  printf("device %s, %d SMs; L=%ld l=%d (%.2f GB per pass)\n", prop.name, sms, L, l, 8.0 * ld * l / 1e9);
  double *V, *q, *part, *h, *href;
  CK(cudaMalloc(&V, 8L * ld * l)); CK(cudaMalloc(&q, 8L * ld)); CK(cudaMalloc(&part, 8L * 4096 * l)); CK(cudaMalloc(&h, 8L * l)); CK(cudaMalloc(&href, 8L * l));
  {
    std::vector<double> hv((size_t)ld);
    uint64_t s = 1;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return ((s >> 11) * (1.0 / 9007199254740992.0)) - 0.5; };
    for (int c = 0; c < l; ++c) { for (long i = 0; i < ld; ++i) hv[i] = i < L ? rnd() : 0.0; CK(cudaMemcpy(V + (long)c * ld, hv.data(), 8L * ld, cudaMemcpyHostToDevice)); }
    for (long i = 0; i < ld; ++i) hv[i] = i < L ? rnd() : 0.0;
    CK(cudaMemcpy(q, hv.data(), 8L * ld, cudaMemcpyHostToDevice));
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<double> ref(l), got(l);
  bool have_ref = false;
  const double bytes = 8.0 * ld * l + 8.0 * ld;
  auto run = [&](const char* name, auto launch, int G) {
    for (int w = 0; w < 2; ++w) launch();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int it = 0; it < iters; ++it) launch();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    k_finalize<<<(l + 127) / 128, 128>>>(l, G, part, h);
    CK(cudaMemcpy(got.data(), h, 8L * l, cudaMemcpyDeviceToHost));
    double err = 0, nrm = 0;
    if (!have_ref) { ref = got; have_ref = true; }
    for (int c = 0; c < l; ++c) { err = fmax(err, fabs(got[c] - ref[c])); nrm = fmax(nrm, fabs(ref[c])); }
    const double us = ms * 1e3 / iters;
    printf("%-44s %9.1f us  %7.1f GB/s   max|diff|/max|h| = %.2e\n", name, us, bytes / us * 1e-3, err / nrm);
    CK(cudaGetLastError());
  };
  {
    const int nchunks = (l + 255) / 256, chunk = ((l + nchunks - 1) / nchunks + 3) / 4 * 4;
    const int gx = (2 * sms) / nchunks;
    const size_t smem = 8UL * 8 * chunk;
    run("A register-staged (library organisation)", [&] { k_regs<<<dim3(gx, nchunks), 256, smem>>>(ld, l, V, ld, q, part, chunk); }, gx);
  }
#define RUN_TMA(RS, S, CPS)                                                                                       \
  {                                                                                                                \
    const int grid = sms * CPS;                                                                                    \
    const long RC = ((ld + grid - 1) / grid + 31) / 32 * 32;                                                       \
    const size_t smem = 8UL * ((size_t)S * RS + RC + 8UL * l) + 128;                                               \
    if (smem <= 227 * 1024 / CPS) {                                                                                \
      CK(cudaFuncSetAttribute(k_tma<RS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
      char nm[96]; snprintf(nm, 96, "B TMA-staged RS=%d S=%d x%d (%zu KB smem/CTA)", RS, S, CPS, smem / 1024);      \
      run(nm, [&] { k_tma<RS, S><<<grid, 288, smem>>>(ld, l, V, ld, q, part, RC); }, grid);                        \
    } else printf("B TMA-staged RS=%d S=%d x%d: %zu KB smem does not fit\n", RS, S, CPS, smem / 1024);              \
  }
  RUN_TMA(2048, 6, 1)
  RUN_TMA(2048, 8, 1)
  RUN_TMA(1024, 12, 1)
  RUN_TMA(4096, 4, 1)
  RUN_TMA(1024, 6, 2)
  RUN_TMA(2048, 4, 2)
  RUN_TMA(1024, 8, 2)
  return 0;
}
