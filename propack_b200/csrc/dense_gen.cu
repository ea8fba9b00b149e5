// propack_b200 -- on-device generator of the synthetic dense tall-skinny operator (BASELINE config 3).
//
// 2M x 4096 doubles are 65.5 GB: too large to build on the host and upload, so the matrix is defined by a
// counter-based formula that the device evaluates tile-wise and that a numpy replica (propack_b200/synth.py) evaluates
// bit-identically for the small parity cases:
//     A(i,j) = u(i,j) + sum_{g=0..15} T[g][ byte_g( X(i) xor Y(j) ) ]
//   u(i,j)  = ((mix(seed, i, j) >> 40) - 2^23 + 0.5) / 2^23                 uniform in (-1,1), exact in fp64
//   X(i), Y(j) = 128 pseudo-random sign bits per row / column (two splitmix64 words each)
//   T[g][b]  = sum_{t=0..7} (bit t of b ? -1 : +1) * c_{8g+t}                built on the host, passed in
// i.e. a planted rank-128 part sum_r c_r x_r y_r^T with +-1 vectors x_r, y_r (singular values ~ c_r sqrt(m n)) on top
// of a uniform noise bulk (largest singular value ~ (sqrt(m)+sqrt(n))/sqrt(3)).  Setup code, not a timed path.
#include "kernels.cuh"

namespace pb {

namespace {

__host__ __device__ inline unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
dense_synth_kernel(long m, int n, long lda, unsigned long long seed, const double* __restrict__ table, double* __restrict__ A, long row0) {
  __shared__ double T[16 * 256];
  for (int i = threadIdx.x; i < 16 * 256; i += 256) T[i] = table[i];
  __syncthreads();
  // a CTA owns a strip of 256 consecutive rows (coalesced column-major stores) and sweeps a slab of columns
  // (row0 = global index of local row 0: a row-sharded operator evaluates only its own rows of the same matrix)
  const long il = (long)blockIdx.x * 256 + threadIdx.x;
  const int j0 = blockIdx.y * 64, j1 = min(n, j0 + 64);
  if (il >= m) return;
  const long i = row0 + il;
  const unsigned long long x0 = splitmix64(seed ^ (0x1000000000000000ull + 2ull * (unsigned long long)i));
  const unsigned long long x1 = splitmix64(seed ^ (0x1000000000000000ull + 2ull * (unsigned long long)i + 1ull));
  for (int j = j0; j < j1; ++j) {
    const unsigned long long y0 = splitmix64(seed ^ (0x2000000000000000ull + 2ull * (unsigned long long)j));
    const unsigned long long y1 = splitmix64(seed ^ (0x2000000000000000ull + 2ull * (unsigned long long)j + 1ull));
    const unsigned long long h = splitmix64(splitmix64(seed ^ (unsigned long long)i) ^ (0x3000000000000000ull + (unsigned long long)j));
    double a = __dmul_rn(__dadd_rn((double)(long long)(h >> 40), -8388607.5), 1.0 / 8388608.0);
    unsigned long long w = x0 ^ y0;
#pragma unroll
    for (int g = 0; g < 8; ++g) a = __dadd_rn(a, T[g * 256 + (int)((w >> (8 * g)) & 255ull)]);
    w = x1 ^ y1;
#pragma unroll
    for (int g = 0; g < 8; ++g) a = __dadd_rn(a, T[(8 + g) * 256 + (int)((w >> (8 * g)) & 255ull)]);
    A[(long)j * lda + il] = a;
  }
}

}  // namespace

void k_dense_synth(Context& c, long m, int n, long lda, unsigned long long seed, const double* table_host, double* A, long row0) {
  DeviceBuffer<double> tab(16 * 256);
  PB_CUDA(cudaMemcpyAsync(tab.p, table_host, sizeof(double) * 16 * 256, cudaMemcpyHostToDevice, c.stream));
  if (m <= 0) { c.sync(); return; }
  dim3 grid((unsigned)((m + 255) / 256), (unsigned)((n + 63) / 64));
  dense_synth_kernel<<<grid, 256, 0, c.stream>>>(m, n, lda, seed, tab.p, A, row0);
  PB_LAUNCH_CHECK();
  c.sync();
}

}  // namespace pb
