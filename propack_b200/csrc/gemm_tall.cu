// propack_b200 -- tall in-place GEMM  A(:,0:N) <- A(:,0:K) * W  on the FP64 tensor pipe (DMMA).
//
// Reference: dgemm_ovwr_left (double/dgemm_ovwr.F:56-87), called for Ritz vectors
// (dritzvec.F:160,193) and for the implicit-restart basis update (dlansvd_irl.F:387,394); complex
// variant zdgemm_ovwr_left (complex16/zgemm_ovwr.F:6-61) multiplies a complex basis by a REAL
// small matrix, which is the same real GEMM on the interleaved (re,im) rows (2M rows, ld 2*lda).
// The reference streams row blocks through a scratch buffer because BLAS dgemm cannot work in
// place.  Here a warp owns 8 rows for the whole K sweep and holds its 8 x N result in DMMA
// accumulator registers, so the product is written back over the same rows once every column of
// those rows has been read: in place, no scratch, M*(K+N)*w bytes of HBM traffic.
//
//   - mma.sync.aligned.m8n8k4.row.col.f64 (SASS: DMMA.8x8x4); tcgen05 has no FP64 kind.
//   - A fragments: one LDG.64 per lane per k-step straight from HBM (8 consecutive rows x 4
//     columns = four 64-byte segments, every 32-byte sector fully used); no reuse across warps
//     exists for the tall operand, so shared-memory staging would add nothing.
//   - W is pre-packed on the host in fragment order (pack_w below), so a B fragment is one
//     coalesced 256-byte warp load served by L1/L2 (W is K*N*8 <= ~0.6 MB, shared by every warp).
//   - N > 128 (more accumulators than registers): column slabs of 128 go through a per-warp
//     L2-resident scratch strip and are copied over A after the last slab.
//   - float / complex-float bases are widened to FP64 on load (same kernel, half the bytes).
// Arithmetic intensity 2KN/(w(K+N)) ~ N/4 flop/B: FP64-pipe bound for N >~ 25.
#include <algorithm>
#include <vector>

#include "kernels.cuh"

namespace pb {

namespace {

__device__ inline void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// NT = n-tiles (of 8 columns) held in registers by one warp.
template <class R, int NT>
__global__ void __launch_bounds__(kThreads)
gemm_tall_kernel(long Mr, int N, int K, R* A, long lda, const double* __restrict__ Wp, int nt_total, int nt0,
                 R* dst, long ldd) {
  const int lane = threadIdx.x & 31;
  const int kk = lane & 3, rr = lane >> 2;
  const long nwarps = (long)gridDim.x * (kThreads / 32);
  const int ksteps = (K + 3) / 4;
  for (long blk = (long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); blk * 8 < Mr; blk += nwarps) {
    const long row = blk * 8 + rr;
    const bool rok = row < Mr;
    double c[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) { c[j][0] = 0.0; c[j][1] = 0.0; }
    const R* ap = A + row + (long)kk * lda;
    const double* wp = Wp + (long)nt0 * 32 + lane;
#pragma unroll 4
    for (int s = 0; s < ksteps; ++s) {
      const int k = 4 * s + kk;
      double a = 0.0;
      if (rok && k < K) a = (double)ap[(long)(4 * s) * lda];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const double b = __ldg(wp + ((long)s * nt_total + j) * 32);
        dmma(c[j][0], c[j][1], a, b);
      }
    }
    // all K columns of these 8 rows have been read by this warp: safe to overwrite in place
    if (rok) {
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int col = (nt0 + j) * 8 + 2 * kk + i;
          if (col < N) dst[(long)col * ldd + row] = (R)c[j][i];
        }
    }
  }
}

template <class R>
__global__ void __launch_bounds__(kThreads)
copy_cols_kernel(long Mr, int N, const R* __restrict__ src, long lds, R* __restrict__ dst, long ldd) {
  for (int col = blockIdx.y; col < N; col += gridDim.y)
    for (long r = (long)blockIdx.x * kThreads + threadIdx.x; r < Mr; r += (long)gridDim.x * kThreads)
      dst[(long)col * ldd + r] = src[(long)col * lds + r];
}

template <class R, int NT>
void launch_slab(Context& c, long Mr, int N, int K, R* A, long lda, const double* Wp, int nt_total, int nt0, R* dst, long ldd) {
  const int grid = c.grid_for((Mr + 7) / 8, kThreads / 32, 4);
  gemm_tall_kernel<R, NT><<<grid, kThreads, 0, c.stream>>>(Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}

template <class R>
void slab(Context& c, int nt, long Mr, int N, int K, R* A, long lda, const double* Wp, int nt_total, int nt0, R* dst, long ldd) {
  if (nt <= 2) launch_slab<R, 2>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  else if (nt <= 4) launch_slab<R, 4>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  else if (nt <= 7) launch_slab<R, 7>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  else if (nt <= 10) launch_slab<R, 10>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  else if (nt <= 13) launch_slab<R, 13>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  else launch_slab<R, 16>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
}

// real GEMM on Mr real rows
template <class R> void gemm_real(Context& c, long Mr, int N, int K, R* A, long lda, const R* W_dev_unused, const std::vector<double>& Wp_host, int nt_total) {
  (void)W_dev_unused;
  double* Wp = static_cast<double*>(c.scratch(Wp_host.size() * sizeof(double)));
  PB_CUDA(cudaMemcpyAsync(Wp, Wp_host.data(), Wp_host.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  PB_CUDA(cudaStreamSynchronize(c.stream));  // Wp_host is pageable and dies with the caller
  if (nt_total <= 16) {
    slab<R>(c, nt_total, Mr, N, K, A, lda, Wp, nt_total, 0, A, lda);
    return;
  }
  // wide result: slabs of 128 columns into a scratch panel, then copy over A
  DeviceBuffer<R> panel((size_t)Mr * N);
  for (int nt0 = 0; nt0 < nt_total; nt0 += 16) {
    const int nt = std::min(16, nt_total - nt0);
    slab<R>(c, nt, Mr, N, K, A, lda, Wp, nt_total, nt0, panel.p, Mr);
  }
  dim3 g(c.grid_for(Mr, kThreads, 4), std::min(N, 64));
  copy_cols_kernel<R><<<g, kThreads, 0, c.stream>>>(Mr, N, panel.p, Mr, A, lda);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
  PB_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace

// Pack a real K x N matrix W (column-major, leading dim ldw, given on the HOST) into DMMA
// B-fragment order: Wp[s][j][t] = W[4s + t%4][8j + t/4], zero padded.
template <class R> std::vector<double> pack_w(int K, int N, const R* W, int ldw, int* nt_total) {
  const int ksteps = (K + 3) / 4, nt = (N + 7) / 8;
  std::vector<double> out((size_t)ksteps * nt * 32, 0.0);
  for (int s = 0; s < ksteps; ++s)
    for (int j = 0; j < nt; ++j)
      for (int t = 0; t < 32; ++t) {
        const int k = 4 * s + (t & 3), n = 8 * j + (t >> 2);
        if (k < K && n < N) out[((size_t)s * nt + j) * 32 + t] = (double)W[(size_t)n * ldw + k];
      }
  *nt_total = nt;
  return out;
}
template std::vector<double> pack_w<float>(int, int, const float*, int, int*);
template std::vector<double> pack_w<double>(int, int, const double*, int, int*);

// W here is a HOST pointer (the small matrix comes from the host bidiagonal SVD / QR sweeps).
template <class T> void k_gemm_tall(Context& c, long M, int N, int K, T* A, long lda, const real_t<T>* W) {
  using R = real_t<T>;
  if (M <= 0 || N <= 0 || K <= 0) return;
  int nt_total = 0;
  std::vector<double> Wp = pack_w<R>(K, N, W, K, &nt_total);
  constexpr int f = scalar_traits<T>::is_complex ? 2 : 1;
  gemm_real<R>(c, M * f, N, K, reinterpret_cast<R*>(A), lda * f, nullptr, Wp, nt_total);
}

#define PB_INST(T) template void k_gemm_tall<T>(Context&, long, int, int, T*, long, const real_t<T>*);
PB_INST(float)
PB_INST(double)
PB_INST(cplx<float>)
PB_INST(cplx<double>)
#undef PB_INST

}  // namespace pb
