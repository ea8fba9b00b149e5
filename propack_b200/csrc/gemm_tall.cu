// propack_b200 -- tall in-place GEMM  A(:,0:N) <- A(:,0:K) * W  on the FP64 tensor pipe (DMMA).
//
// Reference: dgemm_ovwr_left (double/dgemm_ovwr.F:56-87), called for Ritz vectors
// (dritzvec.F:160,193) and for the implicit-restart basis update (dlansvd_irl.F:387,394); complex
// variant zdgemm_ovwr_left (complex16/zgemm_ovwr.F:6-61) multiplies a complex basis by a REAL
// small matrix, which is the same real GEMM on the interleaved (re,im) rows (2M rows, ld 2*lda).
// The reference streams row blocks through a scratch buffer because BLAS dgemm cannot work in
// place.  Here a warp owns 8*MT rows for the whole K sweep and holds its (8*MT) x N result in DMMA
// accumulator registers, so the product is written back over the same rows once every column of
// those rows has been read: in place, no scratch, M*(K+N)*w bytes of HBM traffic.
//
//   - mma.sync.aligned.m8n8k4.row.col.f64 (SASS: DMMA.8x8x4); tcgen05 has no FP64 kind.
//   - A fragments: one LDG.64 per lane per k-step straight from HBM (8 consecutive rows x 4
//     columns = four 64-byte segments, every 32-byte sector fully used), prefetched one K chunk
//     ahead; the tall operand has no reuse across warps, so it is not staged in shared memory.
//   - W is pre-packed on the host in DMMA B-fragment order (pack_w below); the CTA double-buffers
//     K chunks of it in shared memory with cp.async (conflict-free 256-byte rows) and all 8 warps x MT
//     m-tiles reuse it, so W traffic from L2 is |W| per 64*MT rows instead of per 8 rows.
//   - N > 128 (more accumulators than registers): column slabs of 128 go through a scratch panel
//     and are copied over A after the last slab.
//   - float / complex-float bases are widened to FP64 on load (same kernel, half the bytes).
// Arithmetic intensity 2KN/(w(K+N)) ~ N/4 flop/B: FP64-pipe bound for N >~ 25.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "kernels.cuh"

namespace pb {

namespace {

__device__ inline void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// NT = n-tiles (8 columns each) and MT = m-tiles (8 rows each) held in DMMA accumulators by one warp.
// The CTA (8 warps) owns 64*MT consecutive rows; all warps walk K in lock step, GK_KS k-steps (4 columns each) at a
// time.  The packed W chunk is double-buffered in shared memory with cp.async (one barrier per chunk: wait for chunk
// ch, barrier, start the copy of chunk ch+1 into the buffer everybody just finished reading, multiply chunk ch), and
// the A fragments of chunk ch+1 are prefetched from HBM into registers while chunk ch is multiplied.
constexpr int GK_KS_MAX = 8;                                        // pack_w pads W to whole chunks of this many k-steps
template <int MT> constexpr int gk_ks() { return MT >= 2 ? 4 : 8; }  // k-steps per chunk: 2*KS*MT A registers per lane
__device__ inline void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ inline void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ inline void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class R, int NT, int MT>
__global__ void __launch_bounds__(kThreads)
gemm_tall_kernel(long Mr, int N, int K, R* A, long lda, const double* __restrict__ Wp, int nt_total, int nt0,
                 R* dst, long ldd) {
  extern __shared__ __align__(16) double wsm[];   // [2][GK_KS * NT * 32]
  constexpr int GK_KS = gk_ks<MT>();
  constexpr int CHUNK = GK_KS * NT * 32;          // doubles per W chunk
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int kk = lane & 3, rr = lane >> 2;
  const int ksteps = (K + 3) / 4;
  const int nchunks = (ksteps + GK_KS - 1) / GK_KS;   // Wp is zero-padded to nchunks*GK_KS k-steps (pack_w)
  const long tile_rows = 64L * MT;
  // W chunk ch -> buffer b: per k-step a contiguous run of NT*32 doubles (tiles nt0 .. nt0+NT-1), 16-byte pieces
  auto stage_w = [&](int ch, int b) {
    double* dstw = wsm + b * CHUNK;
    for (int i = threadIdx.x; i < CHUNK / 2; i += kThreads) {
      const int s = i / (NT * 16), rem = i - s * (NT * 16);
      cp_async16(dstw + s * NT * 32 + rem * 2, Wp + ((long)(ch * GK_KS + s) * nt_total + nt0) * 32 + rem * 2);
    }
    cp_async_commit();
  };
  for (long t0 = (long)blockIdx.x * tile_rows; t0 < Mr; t0 += (long)gridDim.x * tile_rows) {
    long row[MT];
    bool rok[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { row[mt] = t0 + (long)(wid * MT + mt) * 8 + rr; rok[mt] = row[mt] < Mr; }
    double c[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int j = 0; j < NT; ++j) { c[mt][j][0] = 0.0; c[mt][j][1] = 0.0; }
    double a_cur[GK_KS][MT], a_nxt[GK_KS][MT];
    auto load_a = [&](int chunk, double (&a)[GK_KS][MT]) {
#pragma unroll
      for (int s = 0; s < GK_KS; ++s) {
        const int k = 4 * (chunk * GK_KS + s) + kk;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a[s][mt] = (rok[mt] && k < K) ? (double)A[(long)k * lda + row[mt]] : 0.0;
      }
    };
    __syncthreads();            // the previous row tile's last chunk is fully consumed before buffer 0 is refilled
    stage_w(0, 0);
    load_a(0, a_cur);
    for (int ch = 0; ch < nchunks; ++ch) {
      cp_async_wait<0>();       // this thread's pieces of chunk ch have landed ...
      __syncthreads();          // ... and everybody's; also: everybody is done reading the other buffer (chunk ch-1)
      if (ch + 1 < nchunks) { stage_w(ch + 1, (ch + 1) & 1); load_a(ch + 1, a_nxt); }
      const double* wb = wsm + (ch & 1) * CHUNK;
      const int ks_here = min(GK_KS, ksteps - ch * GK_KS);   // the last chunk is not padded with zero k-steps (warp-uniform bound)
#pragma unroll
      for (int s = 0; s < GK_KS; ++s)
        if (s < ks_here) {
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const double b = wb[(s * NT + j) * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) dmma(c[mt][j][0], c[mt][j][1], a_cur[s][mt], b);
          }
        }
#pragma unroll
      for (int s = 0; s < GK_KS; ++s)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a_cur[s][mt] = a_nxt[s][mt];
    }
    // every column of this warp's rows has been read (the last mma.sync joins the warp): overwrite in place
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
      if (rok[mt]) {
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int col = (nt0 + j) * 8 + 2 * kk + i;
            if (col < N) dst[(long)col * ldd + row[mt]] = (R)c[mt][j][i];
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

template <class R, int NT, int MT>
void launch_slab(Context& c, long Mr, int N, int K, R* A, long lda, const double* Wp, int nt_total, int nt0, R* dst, long ldd) {
  constexpr size_t smem = sizeof(double) * 2 * gk_ks<MT>() * NT * 32;
  static int occ = 0;            // resident CTAs per SM of this instantiation (registers / shared memory), at most 2
  if (occ == 0) {
    PB_CUDA(cudaFuncSetAttribute(gemm_tall_kernel<R, NT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemm_tall_kernel<R, NT, MT>, kThreads, smem));
    occ = std::min(2, std::max(1, occ));
  }
  const int grid = c.grid_for(Mr, 64 * MT, occ);
  gemm_tall_kernel<R, NT, MT><<<grid, kThreads, smem, c.stream>>>(Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  PB_LAUNCH_CHECK();
  c.ctr.launches += 1;
}

template <class R>
void slab(Context& c, int nt, long Mr, int N, int K, R* A, long lda, const double* Wp, int nt_total, int nt0, R* dst, long ldd) {
  if (nt <= 2) launch_slab<R, 2, 4>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  else if (nt <= 4) launch_slab<R, 4, 4>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  else if (nt <= 7) launch_slab<R, 7, 2>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  // (two m-tiles per warp for the 10- and 13-tile widths -- each W fragment feeding two DMMAs -- need 233 registers, i.e. one
  // CTA per SM: measured 21.4 against 25.8 TFLOP/s at N = 101, K = 301, so the wide slabs keep one m-tile per warp)
  else if (nt <= 10) launch_slab<R, 10, 1>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  else if (nt <= 13) launch_slab<R, 13, 1>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
  else launch_slab<R, 16, 1>(c, Mr, N, K, A, lda, Wp, nt_total, nt0, dst, ldd);
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
  const int ksteps_pad = (ksteps + GK_KS_MAX - 1) / GK_KS_MAX * GK_KS_MAX;   // whole chunks: the kernel copies them unconditionally
  // + 16 zero tiles of slack: a slab launched with NT rounded up (11 -> 13 tiles, ...) copies NT tiles per k-step starting at its
  // first tile, i.e. up to NT-1 tiles past the end of the last k-step (their products land in columns >= N and are discarded)
  std::vector<double> out(((size_t)ksteps_pad * nt + 16) * 32, 0.0);
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
