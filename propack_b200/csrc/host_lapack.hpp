// propack_b200 -- host LAPACK for the O(k^2) bidiagonal SVD that stays on the CPU
// (reference: dbdsqr call at double/dlansvd.F:198 / dlansvd_irl.F:228, dbdsdc at dritzvec.F:123;
// upstream vendors LAPACK 3.0 under */Lapack_Util, here the image's LAPACK is bound at run time).
#pragma once
#include <cstddef>
#include <vector>

namespace pb {
namespace host {

// Binds xBDSQR / xBDSDC from a shared object.  Search order: explicit path argument,
// $PROPACK_B200_LAPACK, liblapack.so.3, libopenblas.so.0, then a scipy_openblas found next to the
// running Python's scipy (symbols carry a "scipy_" prefix there).  Throws std::runtime_error.
void bind_lapack(const char* path = nullptr);
bool lapack_bound();

// singular values of an upper bidiagonal + one row-vector of left rotations (nru = 1)
void bdsqr_row(int n, double* d, double* e, double* urow, int* info);
void bdsqr_row(int n, float* d, float* e, float* urow, int* info);
// divide & conquer SVD of an upper bidiagonal: B = U diag(d) VT  (compq = 'I')
void bdsdc_full(int n, double* d, double* e, double* U, int ldu, double* VT, int ldvt, int* info);
void bdsdc_full(int n, float* d, float* e, float* U, int ldu, float* VT, int ldvt, int* info);

// Leading K Ritz values of the (j+1) x j lower bidiagonal B (diag alpha, sub-diag beta) and the last components of
// their left singular vectors -- the quantities the non-restarted driver needs per outer iteration (error bounds
// bnd(i) = |rnorm * u_i(j+1)|, dlansvd.F:196-209) -- without the O(j^2) QR sweep over all j values: singular values by
// dqds (xLASQ1, high relative accuracy) on the QR-reduced bidiagonal, vectors by inverse iteration (xSTEIN, with its
// reorthogonalisation inside clusters) on the Golub-Kahan tridiagonal of B, whose eigenvector for +sigma_i interleaves
// (u_i, v_i).  theta[0:K] descending, last[0:K] = |u_i(j+1)| for unit u_i.  Returns false (nothing written) when the
// routines are not available, B has a zero / negligible entry (reducible problem) or xSTEIN reports a failure: the
// caller then takes the reference's xBDSQR route.  The float overload always returns false.
bool ritz_leading(int j, const double* alpha, const double* beta, int K, double* theta, double* last);
bool ritz_leading(int j, const float* alpha, const float* beta, int K, float* theta, float* last);

// The k leading singular vector pairs of the same bidiagonal by the same route (dqds + xSTEIN on the Golub-Kahan
// tridiagonal): WU ((dim+1) x k, leading dimension dim+1) = left vectors u_i, WV (dim x k, ld dim) = right vectors v_i,
// unit norm, B v_i = sigma_i u_i.  These are the small matrices the Ritz-vector GEMMs multiply the Lanczos bases with
// (dritzvec.F:116-193 obtains them from dbdqr + dbdsdc of ALL dim vectors).  Same decline conditions as ritz_leading.
bool ritz_vectors_leading(int dim, const double* alpha, const double* beta, int k, std::vector<double>& WU, std::vector<double>& WV);
bool ritz_vectors_leading(int dim, const float* alpha, const float* beta, int k, std::vector<float>& WU, std::vector<float>& WV);

}  // namespace host
}  // namespace pb
