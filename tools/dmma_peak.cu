// FP64 tensor-pipe ceiling on this GPU: a register-only loop of independent mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) chains,
// no memory traffic.  Prints TFLOP/s for a few warps-per-SM settings; the best one is the denominator used for
// gemm_tall_kernel's "fraction of the FP64 tensor peak" (profiles/r02_summary.md).  Also times a plain DFMA loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/dmma_peak tools/dmma_peak.cu && tools/bin/dmma_peak
#include <cuda_runtime.h>
#include <cstdio>

template <int CH>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double a0, double b0) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double a0, double b0) {
  double c[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a0, b0);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int ctas = 1; ctas <= 8; ctas *= 2) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      dmma_loop<16><<<sms * ctas, 256>>>(out, iters, 1.0000001, 0.9999999);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 2.0 * 8 * 8 * 4 * 16.0 * iters * (double)sms * ctas * 8;
    printf("{\"kernel\": \"dmma.8x8x4 register loop\", \"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", ctas, ctas * 8, best,
           flops / best / 1e9);
  }
  for (int ctas = 2; ctas <= 8; ctas *= 2) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      dfma_loop<16><<<sms * ctas, 256>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 2.0 * 16.0 * iters * (double)sms * ctas * 256;
    printf("{\"kernel\": \"dfma register loop\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", ctas, best, flops / best / 1e9);
  }
  return 0;
}
