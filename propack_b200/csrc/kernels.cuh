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
// x <- x / (*slot).re, scalar taken from a device-resident result slot (speculative dsafescal)
template <class T> void k_scal_inv_slot(Context& c, long n, T* x, const ScalarSlot* slot);
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
                 bool staggered, T* self_slice);
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
  const int* long_rows = nullptr;  // rows with more than spmv_group_nnz/2 non-zeros (csr_long_rows), ascending
  int n_long = 0;
};
// host-side analysis: lanes per row so that a group of 32/LPR rows fits a slice of nb products
int csr_lanes_per_row_log2(long nnz, int rows, int nb);
std::vector<int> csr_long_rows(const int* rp, int rows, int nb);
// y <- op(A) x + coef*prev (prev may be null) ; optionally publish ||y||_2.  conj: use conj(values).
template <class T>
void k_spmv(Context& c, const CsrDevice<T>& A, bool conj, const T* x, T* y, real_t<T> coef, const T* prev, Pending* nrm);

// --- operator registration (setup): device-side transpose + validation, csr_build.cu ---------------------------
// CSR(A) -> canonical CSR(A^T) (all device pointers).  Returns 0 / 1 (unsorted row) / 2 (index out of range).
template <class T>
int k_csr_transpose(Context& c, int rows, int cols, long nnz, const int* rp, const int* ci, const T* va, int* trp, int* tci, T* tva);
void k_rebase(Context& c, long n, int* a, int base);  // a[i] -= base (Fortran 1-based index arrays)

// synthetic dense operator A(i,j) = u(i,j) + planted rank-128 part, evaluated on the device (dense_gen.cu)
void k_dense_synth(Context& c, long m, int n, long lda, unsigned long long seed, const double* table_host, double* A);

// --- tall in-place GEMM (reference: dgemm_ovwr_left, double/dgemm_ovwr.F:56-87) --------------------
// A(:,0:N) <- A(:,0:K) * W,  W real K x N column-major (ld = K) in HOST memory (it comes from the host
// bidiagonal SVD); it is packed into DMMA fragment order and uploaded by the wrapper.
template <class T> void k_gemm_tall(Context& c, long M, int N, int K, T* A, long lda, const real_t<T>* W);

}  // namespace pb
