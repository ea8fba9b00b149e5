// propack_b200 -- launch wrappers of the hand-written sm_100a kernels (declarations).
// Each wrapper enqueues on Context::stream and never synchronises; results the host must branch
// on come back through Context::Pending slots (see common.cuh).
#pragma once
#include <vector>

#include "common.cuh"
#include "context.hpp"

namespace pb {

using Pending = Context::Pending;

// --- level-1 (reference: blasext, double/dblasext.F:6-255; dsafescal.F) ---------------------------
// x <- a * x                                                     (pdscal dblasext.F:38)
template <class T> void k_scal(Context& c, long n, T* x, real_t<T> a);
// y <- y + a*x ; publish ||y||_2                                  (pdaxpy + pdnrm2 fused, dlanbpro.F:295-296)
template <class T> void k_axpy_nrm(Context& c, long n, T a, const T* x, T* y, Pending* nrm);
// publish conj(x).y                                              (pddot dblasext.F:121 / pzdotc)
template <class T> void k_dotc(Context& c, long n, const T* x, const T* y, Pending* out);
// publish ||x||_2                                                (pdnrm2 dblasext.F:6)
template <class T> void k_nrm2(Context& c, long n, const T* x, Pending* out);
template <class T> void k_zero(Context& c, long n, T* x);          // pdzero dblasext.F:202
// row-sharded run, fused normalise + all-gather: x <- a*x and the slice is pushed into every rank's gather buffer
// (bases_dev[r], peer memory) at offset rank*ld; arrival flags sit behind the world*ld elements of each buffer
template <class T>
void k_scal_push(Context& c, long n, long ld, T* x, real_t<T> a, void** bases_dev, int rank, int world, unsigned long long epoch,
                 T* self_slice);
template <class T>
void k_scal_local(Context& c, long n, T* x, real_t<T> a, T* self_slice, unsigned long long* epoch_dev, unsigned long long epoch);
void k_wait_flags(Context& c, const unsigned long long* flags, unsigned int src_mask, unsigned long long epoch);
// x(i) <- LAPACK xLARNV(idist=2, iseed) stream element offset+i, i=0..n-1 ; publish ||x||  (dgetu0.F:69-70).
// `offset` = global index of this rank's first element in a row-sharded run (0 on one GPU).
template <class T> void k_larnv_nrm(Context& c, long n, T* x, const int iseed[4], Pending* nrm, long offset = 0);

// --- tall-skinny GEMV pair (reference: dcgs, double/dreorth.F:174 and :199-205) --------------------
// h(0:l) <- V(:,0:l)^H q   (column-major V, leading dim ldv, L rows).  h is a device buffer.
template <class T> void k_gemv_t(Context& c, long L, int l, const T* V, long ldv, const T* q, T* h);
// out <- cin*in + sgn * V(:,0:l) h ; optionally publish ||out||_2.  `in` may alias `out` or be null (cin ignored).
template <class T>
void k_gemv_n(Context& c, long L, int l, const T* V, long ldv, const T* h, real_t<T> cin, const T* in, int sgn, T* out,
              Pending* nrm);

// --- CSR SpMV (reference: the user's APROD, dlansvd.F:20-33; call sites dlanbpro.F:288,420) --------
constexpr int kSpmvU = 4;            // independent (ci -> x) gather chains per lane per batch
constexpr int kSpmvCtasPerSm = 3;    // persistent CTAs per SM (3 x 32 KB of shared memory: L1 keeps ~130 KB)
constexpr int kSpmvCarveoutPct = 50; // cudaFuncAttributePreferredSharedMemoryCarveout: the 132 KB configuration (3 CTAs x ~34 KB need > 100 KB)
// non-zeros per row group = a 4 KB shared-memory slice per warp
template <class T> constexpr int spmv_group_nnz() { return 4096 / (int)sizeof(T); }
template <class T> struct CsrDevice {
  int rows = 0, cols = 0;
  long nnz = 0;
  const int* rp = nullptr;     // [rows+1]
  const int* ci = nullptr;     // [nnz], sorted within a row
  const T* va = nullptr;       // [nnz]
  int lpr_log2 = 0;            // lanes per row in the reduce phase = 1 << lpr_log2 (csr_lanes_per_row_log2)
  const int* long_rows = nullptr;  // rows done by spmv_long_kernel (k_csr_long_rows), ascending
  int n_long = 0;
};
// host-side analysis: lanes per row so that a group of 32/LPR rows fits a slice of nb products
int csr_lanes_per_row_log2(long nnz, int rows, int nb);
// y <- op(A) x + coef*prev (prev may be null) ; optionally publish ||y||_2.  conj: use conj(values).
template <class T>
void k_spmv(Context& c, const CsrDevice<T>& A, bool conj, const T* x, T* y, real_t<T> coef, const T* prev, Pending* nrm);

// The rows listed in A.long_rows, one CTA per row: y[row] = sum [+ y[row]] [+ coef*prev[row]] (spmv.cu).
template <class T>
void k_spmv_long(Context& c, const CsrDevice<T>& A, bool conj, const T* x, T* y, real_t<T> coef, const T* prev, bool accumulate);

// --- sliced jagged-ELL SpMV, the default sparse APROD kernel (sell.cu) ------------------------------------------------
constexpr int kSellLong = 64;        // rows longer than this go to spmv_long_kernel (one CTA per row)
constexpr int kSellCtasPerSm = 8;    // upper bound on persistent CTAs per SM; the builder caps it by the kernel's real occupancy
constexpr int kSellStepCost = 4;     // weight of one k-step of a slice (coalesced ci / va loads), in gathered entries
constexpr int kSellSliceCost = 8;    // fixed weight of a slice (lengths, y / prev access, epilogue)
template <class T> struct SellDevice {
  int rows = 0, cols = 0;
  long nnz = 0;
  long nslices = 0;                   // 32-row slices: ceil(rows / 32)
  long long stored = 0;               // entries in the slices (= nnz minus the long rows' entries)
  const long long* joff = nullptr;    // [nslices+1] first stored entry of each slice
  const unsigned char* len8 = nullptr;// [rows] row length, 0xFF = long row (done by spmv_long_kernel)
  const int* ci = nullptr;            // [stored] columns, k-major compacted inside a slice
  const T* va = nullptr;              // [stored]
  // static load balance: warp w of the planned grid owns the contiguous slices [wstart[w], wstart[w+1]) -- equal shares of
  // sum(entries + kSellStepCost*longest row + kSellSliceCost) -- so ragged slices do not pile up on some warps
  const int* wstart = nullptr;        // [grid*8 + 1]
  int grid = 0;                       // CTAs the partition was made for (the kernel is launched with exactly this grid)
};
template <class T> struct SellStorage {
  DeviceBuffer<long long> joff;
  DeviceBuffer<unsigned char> len8;
  DeviceBuffer<int> ci, wstart;
  DeviceBuffer<T> va;
  SellDevice<T> dev;
};
// CSR (device pointers, 0-based, sorted rows) -> sliced jagged ELL.  Integer work, deterministic.  ctas_per_sm: size of
// the persistent grid the warp partition is planned for (> 0: at most that many, capped by the kernel's occupancy;
// <= 0: every resident slot but -ctas_per_sm).
template <class T>
void sell_build(Context& c, int rows, int cols, long nnz, const int* rp, const int* ci, const T* va, SellStorage<T>& out,
                int ctas_per_sm = kSellCtasPerSm);
// y <- [y +] op(S) x [+ coef*prev ; publish ||y||].  mode: bit 0 = accumulate into y, bit 1 = final phase (epilogue).
// long_src: the CSR whose long_rows list spmv_long_kernel serves (final phase only), or null.
// flags/src_mask/epoch: row-sharded runs -- wait in-kernel until the gather-buffer slices of the ranks in src_mask
// have arrived (flags[r] >= epoch); src_mask = 0: no wait.
constexpr int kSellModeAcc = 1, kSellModeFinal = 2;
template <class T>
void k_spmv_sell(Context& c, const SellDevice<T>& S, const CsrDevice<T>* long_src, bool conj, const T* x, T* y, real_t<T> coef,
                 const T* prev, Pending* nrm, int mode, const unsigned long long* flags, unsigned int src_mask,
                 unsigned long long epoch);

// --- operator registration (setup): device-side transpose + validation, csr_build.cu ---------------------------
// CSR(A) -> canonical CSR(A^T) (all device pointers).  Returns 0 / 1 (unsorted row) / 2 (index out of range).
template <class T>
int k_csr_transpose(Context& c, int rows, int cols, long nnz, const int* rp, const int* ci, const T* va, int* trp, int* tci, T* tva);
void k_rebase(Context& c, long n, int* a, int base);  // a[i] -= base (Fortran 1-based index arrays)
int k_csr_check_rowptr(Context& c, int rows, long nnz, const int* rp);   // 0, or 4 when rp is not monotone from 0 to nnz
int k_csr_long_rows(Context& c, int rows, const int* rp, int threshold, DeviceBuffer<int>& list);  // rows longer than threshold
// row-sharded operands: CSR -> G CSRs by the ring distance of each column's owner rank (csr_build.cu)
template <class T>
int k_csr_split_phases(Context& c, int rows, long width, const int* rp, const int* ci, const T* va, long ld, int P, int rank, int G,
                       int* const* out_rp, DeviceBuffer<int>* const* out_ci, DeviceBuffer<T>* const* out_va, long* nnz_out);

// synthetic dense operator A(i,j) = u(i,j) + planted rank-128 part, evaluated on the device (dense_gen.cu)
void k_dense_synth(Context& c, long m, int n, long lda, unsigned long long seed, const double* table_host, double* A, long row0 = 0);

// --- tall in-place GEMM (reference: dgemm_ovwr_left, double/dgemm_ovwr.F:56-87) --------------------
// A(:,0:N) <- A(:,0:K) * W,  W real K x N column-major (ld = K) in HOST memory (it comes from the host
// bidiagonal SVD); it is packed into DMMA fragment order and uploaded by the wrapper.
template <class T> void k_gemm_tall(Context& c, long M, int N, int K, T* A, long lda, const real_t<T>* W);

}  // namespace pb
