/* propack_b200.h -- C-ABI of libpropack_b200.so
 *
 * A B200-native (sm_100a) implementation of PROPACK's Lanczos-bidiagonalisation hot path behind the
 * reference's own calling convention.  Three groups of entry points:
 *
 *  (1) the Fortran-77 ABI symbols of the reference library (gfortran mangling: lower case + '_',
 *      every argument by reference, default INTEGER = 32 bit, hidden CHARACTER lengths appended by
 *      value as size_t).  A program linked against PROPACK's libdpropack/… resolves the same names
 *      here.  Each prototype cites the reference interface it replaces.
 *  (2) device-resident operators: the caller registers a CSR / dense matrix once, receives an
 *      integer handle, stores it in IPARM(1) and passes `propack_b200_aprod_<p>_` as APROD.  The
 *      drivers recognise that function pointer and keep A.x and A^T.x on the GPU (no callback, no
 *      PCIe traffic per product).  Any other APROD is honoured through a host-staged path.
 *  (3) a solver-session API that keeps the Lanczos bases in HBM across calls (used by the Python
 *      binding and by bench.py's device-resident measurement), plus counters / timers.
 *
 * All bulk arrays in groups (1)-(2) are HOST pointers unless stated.  No torch / CUDA types appear in
 * any signature; a CUDA stream is passed as void*.  There is no CPU fallback: without a usable
 * sm_100 device every compute entry point fails (info = -100 - cudaError, or a negative return).
 */
#ifndef PROPACK_B200_H
#define PROPACK_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } pb200_complex8;    /* Fortran COMPLEX    */
typedef struct { double re, im; } pb200_complex16;  /* Fortran COMPLEX*16 */

/* ---- user APROD contract: reference double/dlansvd.F:20-33 (SUBROUTINE APROD(TRANSA,M,N,X,Y,DPARM,IPARM)).
 * transa = 'n': y = A x (x length n, y length m); 't' (real) / 'c' (complex, zlanbpro.F:297): y = A^H x. */
typedef void (*pb200_aprod_s_t)(const char* transa, const int* m, const int* n, const float* x, float* y, float* sparm, int* iparm, size_t transa_len);
typedef void (*pb200_aprod_d_t)(const char* transa, const int* m, const int* n, const double* x, double* y, double* dparm, int* iparm, size_t transa_len);
typedef void (*pb200_aprod_c_t)(const char* transa, const int* m, const int* n, const pb200_complex8* x, pb200_complex8* y, pb200_complex8* cparm, int* iparm, size_t transa_len);
typedef void (*pb200_aprod_z_t)(const char* transa, const int* m, const int* n, const pb200_complex16* x, pb200_complex16* y, pb200_complex16* zparm, int* iparm, size_t transa_len);

/* =====================================================================================================
 * (1) Fortran-ABI drivers
 * ===================================================================================================== */

/* xLANSVD -- reference double/dlansvd.F:1-3 (args :96-104), single/slansvd.F:1-3.
 * k in: wanted / out: converged; U(ldu,kmax+1), V(ldv,kmax); U(:,1) = start vector (all zero => random,
 * dlansvd.F:165-169); doption(3) = delta, eta, anorm (anorm in/out, dlanbpro.F:547); ioption(2) = cgs, elr;
 * info: 0 ok, j>0 invariant subspace of dimension j, -1 kmax exhausted (dlansvd.F:79-83),
 * <= -100: CUDA / setup failure (new; see propack_b200_last_error).  work/iwork are accepted for
 * compatibility and not used for bulk data. */
void slansvd_(const char* jobu, const char* jobv, const int* m, const int* n, int* k, const int* kmax, pb200_aprod_s_t aprod,
              float* U, const int* ldu, float* sigma, float* bnd, float* V, const int* ldv, const float* tolin, float* work,
              const int* lwork, int* iwork, const int* liwork, float* soption, int* ioption, int* info, float* sparm, int* iparm,
              size_t jobu_len, size_t jobv_len);
void dlansvd_(const char* jobu, const char* jobv, const int* m, const int* n, int* k, const int* kmax, pb200_aprod_d_t aprod,
              double* U, const int* ldu, double* sigma, double* bnd, double* V, const int* ldv, const double* tolin, double* work,
              const int* lwork, int* iwork, const int* liwork, double* doption, int* ioption, int* info, double* dparm, int* iparm,
              size_t jobu_len, size_t jobv_len);
/* complex: reference complex16/zlansvd.F:1-3 (extra zwork,lzwrk after lwork; args :98-109), complex8/clansvd.F */
void clansvd_(const char* jobu, const char* jobv, const int* m, const int* n, int* k, const int* kmax, pb200_aprod_c_t aprod,
              pb200_complex8* U, const int* ldu, float* sigma, float* bnd, pb200_complex8* V, const int* ldv, const float* tolin,
              float* work, const int* lwork, pb200_complex8* cwork, const int* lcwrk, int* iwork, const int* liwork, float* soption,
              int* ioption, int* info, pb200_complex8* cparm, int* iparm, size_t jobu_len, size_t jobv_len);
void zlansvd_(const char* jobu, const char* jobv, const int* m, const int* n, int* k, const int* kmax, pb200_aprod_z_t aprod,
              pb200_complex16* U, const int* ldu, double* sigma, double* bnd, pb200_complex16* V, const int* ldv, const double* tolin,
              double* work, const int* lwork, pb200_complex16* zwork, const int* lzwrk, int* iwork, const int* liwork, double* doption,
              int* ioption, int* info, pb200_complex16* zparm, int* iparm, size_t jobu_len, size_t jobv_len);

/* xLANSVD_IRL -- reference double/dlansvd_irl.F:1-3 (args :113-121).  dim is clamped in place (:170),
 * neig in: wanted / out: converged (:417); doption(4) adds the minimum relative gap for shifts. */
void slansvd_irl_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, int* dim, const int* p, int* neig,
                  const int* maxiter, pb200_aprod_s_t aprod, float* U, const int* ldu, float* sigma, float* bnd, float* V,
                  const int* ldv, const float* tolin, float* work, const int* lwork, int* iwork, const int* liwork, float* soption,
                  int* ioption, int* info, float* sparm, int* iparm, size_t which_len, size_t jobu_len, size_t jobv_len);
void dlansvd_irl_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, int* dim, const int* p, int* neig,
                  const int* maxiter, pb200_aprod_d_t aprod, double* U, const int* ldu, double* sigma, double* bnd, double* V,
                  const int* ldv, const double* tolin, double* work, const int* lwork, int* iwork, const int* liwork, double* doption,
                  int* ioption, int* info, double* dparm, int* iparm, size_t which_len, size_t jobu_len, size_t jobv_len);
/* complex: reference complex16/zlansvd_irl.F:1-3.  NOTE: the reference's complex restart multiplies by
 * P^T/Q^T instead of P/Q (zgemm_ovwr.F:41-61 ignores transb); this library applies P/Q. */
void clansvd_irl_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, int* dim, const int* p, int* neig,
                  const int* maxiter, pb200_aprod_c_t aprod, pb200_complex8* U, const int* ldu, float* sigma, float* bnd,
                  pb200_complex8* V, const int* ldv, const float* tolin, float* work, const int* lwork, pb200_complex8* cwork,
                  const int* lcwrk, int* iwork, const int* liwork, float* soption, int* ioption, int* info, pb200_complex8* cparm,
                  int* iparm, size_t which_len, size_t jobu_len, size_t jobv_len);
void zlansvd_irl_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, int* dim, const int* p, int* neig,
                  const int* maxiter, pb200_aprod_z_t aprod, pb200_complex16* U, const int* ldu, double* sigma, double* bnd,
                  pb200_complex16* V, const int* ldv, const double* tolin, double* work, const int* lwork, pb200_complex16* zwork,
                  const int* lzwrk, int* iwork, const int* liwork, double* doption, int* ioption, int* info, pb200_complex16* zparm,
                  int* iparm, size_t which_len, size_t jobu_len, size_t jobv_len);

/* ---- secondary public routines of the reference library (host arrays, staged through the GPU) ------- */

/* xLANBPRO -- reference double/dlanbpro.F:1-2 (complex: zlanbpro.F:1-3 takes dwork, zwork).  B(ldb,2). */
void slanbpro_(const int* m, const int* n, const int* k0, int* k, pb200_aprod_s_t aprod, float* U, const int* ldu, float* V,
               const int* ldv, float* B, const int* ldb, float* rnorm, float* soption, int* ioption, float* work, int* iwork,
               float* sparm, int* iparm, int* ierr);
void dlanbpro_(const int* m, const int* n, const int* k0, int* k, pb200_aprod_d_t aprod, double* U, const int* ldu, double* V,
               const int* ldv, double* B, const int* ldb, double* rnorm, double* doption, int* ioption, double* work, int* iwork,
               double* dparm, int* iparm, int* ierr);
void clanbpro_(const int* m, const int* n, const int* k0, int* k, pb200_aprod_c_t aprod, pb200_complex8* U, const int* ldu,
               pb200_complex8* V, const int* ldv, float* B, const int* ldb, float* rnorm, float* soption, int* ioption, float* swork,
               pb200_complex8* cwork, int* iwork, pb200_complex8* cparm, int* iparm, int* ierr);
void zlanbpro_(const int* m, const int* n, const int* k0, int* k, pb200_aprod_z_t aprod, pb200_complex16* U, const int* ldu,
               pb200_complex16* V, const int* ldv, double* B, const int* ldb, double* rnorm, double* doption, int* ioption,
               double* dwork, pb200_complex16* zwork, int* iwork, pb200_complex16* zparm, int* iparm, int* ierr);

/* xREORTH -- reference double/dreorth.F:5-6: iterated Gram-Schmidt of vnew against V(:,index intervals).
 * index = [s1,e1,...,T] 1-based inclusive, T > k terminates.  iflag 1 = CGS, 0 = MGS. */
void sreorth_(const int* n, const int* k, const float* V, const int* ldv, float* vnew, float* normvnew, const int* index,
              const float* alpha, float* work, const int* iflag);
void dreorth_(const int* n, const int* k, const double* V, const int* ldv, double* vnew, double* normvnew, const int* index,
              const double* alpha, double* work, const int* iflag);
void creorth_(const int* n, const int* k, const pb200_complex8* V, const int* ldv, pb200_complex8* vnew, float* normvnew,
              const int* index, const float* alpha, pb200_complex8* work, const int* iflag);
void zreorth_(const int* n, const int* k, const pb200_complex16* V, const int* ldv, pb200_complex16* vnew, double* normvnew,
              const int* index, const double* alpha, pb200_complex16* work, const int* iflag);

/* xGETU0 -- reference double/dgetu0.F:11-12: random vector in range(op(A)) orthogonal to U(:,1:j). */
void sgetu0_(const char* transa, const int* m, const int* n, const int* j, const int* ntry, float* u0, float* u0norm, const float* U,
             const int* ldu, pb200_aprod_s_t aprod, float* sparm, int* iparm, int* ierr, const int* icgs, float* anormest, float* work,
             size_t transa_len);
void dgetu0_(const char* transa, const int* m, const int* n, const int* j, const int* ntry, double* u0, double* u0norm, const double* U,
             const int* ldu, pb200_aprod_d_t aprod, double* dparm, int* iparm, int* ierr, const int* icgs, double* anormest, double* work,
             size_t transa_len);
void cgetu0_(const char* transa, const int* m, const int* n, const int* j, const int* ntry, pb200_complex8* u0, float* u0norm,
             const pb200_complex8* U, const int* ldu, pb200_aprod_c_t aprod, pb200_complex8* cparm, int* iparm, int* ierr,
             const int* icgs, float* anormest, pb200_complex8* work, size_t transa_len);
void zgetu0_(const char* transa, const int* m, const int* n, const int* j, const int* ntry, pb200_complex16* u0, double* u0norm,
             const pb200_complex16* U, const int* ldu, pb200_aprod_z_t aprod, pb200_complex16* zparm, int* iparm, int* ierr,
             const int* icgs, double* anormest, pb200_complex16* work, size_t transa_len);

/* xSAFESCAL -- reference double/dsafescal.F:4: x <- x / alpha. */
void ssafescal_(const int* n, const float* alpha, float* x);
void dsafescal_(const int* n, const double* alpha, double* x);
void csafescal_(const int* n, const float* alpha, pb200_complex8* x);
void zsafescal_(const int* n, const double* alpha, pb200_complex16* x);

/* xGEMM_OVWR_LEFT -- reference double/dgemm_ovwr.F:56-57: A(m x k) <- alpha * A * op(B), result m x n.
 * complex variants: reference complex16/zgemm_ovwr.F:6 (zdgemm_ovwr_left: complex A, REAL B, no alpha/beta). */
void sgemm_ovwr_left_(const char* transb, const int* m, const int* n, const int* k, const float* alpha, float* A, const int* lda,
                      const float* beta, const float* B, const int* ldb, float* work, const int* lwork, size_t transb_len);
void dgemm_ovwr_left_(const char* transb, const int* m, const int* n, const int* k, const double* alpha, double* A, const int* lda,
                      const double* beta, const double* B, const int* ldb, double* dwork, const int* ldwork, size_t transb_len);
void csgemm_ovwr_left_(const char* transb, const int* m, const int* n, const int* k, pb200_complex8* A, const int* lda, const float* B,
                       const int* ldb, pb200_complex8* cwork, const int* lcwork, size_t transb_len);
void zdgemm_ovwr_left_(const char* transb, const int* m, const int* n, const int* k, pb200_complex16* A, const int* lda, const double* B,
                       const int* ldb, pb200_complex16* zwork, const int* lzwork, size_t transb_len);

/* xRITZVEC -- reference double/dritzvec.F:1-2 (args :11-55), single/sritzvec.F; complex16/zritzvec.F:1-2 and
 * complex8/critzvec.F add (zwork, lzwrk) after in_lwrk.  U(ldu,dim+1), V(ldv,dim) hold the Lanczos bases on entry and the
 * k Ritz vectors in their first k columns on return; D(dim), E(dim) = diagonal / sub-diagonal of B on entry, D returns
 * the singular values of B (descending), E is destroyed (dbdqr + dbdsdc, :116-123).  S is not referenced (the
 * reference never writes it either).  work/iwork are accepted for compatibility. */
void sritzvec_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, const int* k, const int* dim,
               float* D, float* E, float* S, float* U, const int* ldu, float* V, const int* ldv, float* work, const int* in_lwrk,
               int* iwork, size_t which_len, size_t jobu_len, size_t jobv_len);
void dritzvec_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, const int* k, const int* dim,
               double* D, double* E, double* S, double* U, const int* ldu, double* V, const int* ldv, double* work,
               const int* in_lwrk, int* iwork, size_t which_len, size_t jobu_len, size_t jobv_len);
void critzvec_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, const int* k, const int* dim,
               float* D, float* E, float* S, pb200_complex8* U, const int* ldu, pb200_complex8* V, const int* ldv, float* work,
               const int* in_lwrk, pb200_complex8* cwork, const int* lcwrk, int* iwork, size_t which_len, size_t jobu_len,
               size_t jobv_len);
void zritzvec_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, const int* k, const int* dim,
               double* D, double* E, double* S, pb200_complex16* U, const int* ldu, pb200_complex16* V, const int* ldv,
               double* work, const int* in_lwrk, pb200_complex16* zwork, const int* lzwrk, int* iwork, size_t which_len,
               size_t jobu_len, size_t jobv_len);

/* xGEMM_OVWR -- reference double/dgemm_ovwr.F:5-6 (single/sgemm_ovwr.F): B <- alpha*op(A)*B + beta*B, op(A) m x k,
 * B k x n on entry and m x n on return (ldb >= max(m,k)).  Used by dritzvec on the small (dim+1)-square factors. */
void sgemm_ovwr_(const char* transa, const int* m, const int* n, const int* k, const float* alpha, const float* A, const int* lda,
                 const float* beta, float* B, const int* ldb, float* work, const int* lwork, size_t transa_len);
void dgemm_ovwr_(const char* transa, const int* m, const int* n, const int* k, const double* alpha, const double* A, const int* lda,
                 const double* beta, double* B, const int* ldb, double* dwork, const int* ldwork, size_t transa_len);

/* blasext level-1 -- reference double/dblasext.F: pdnrm2 :6, pdscal :38, pdaxpy :92, pddot :121, pdzero :202;
 * complex16/zblasext.F: pdznrm2 :6, pzdscal :60, pzaxpy :113, pzdaxpy :138, pzdotc :167, pzdotu :197, pzzero :344;
 * single/sblasext.F and complex8/cblasext.F likewise.  Host vectors (any increment) staged through the device kernels
 * the Lanczos loop uses; complex functions return by value like gfortran's COMPLEX functions. */
float psnrm2_(const int* n, const float* x, const int* incx);
double pdnrm2_(const int* n, const double* x, const int* incx);
float pscnrm2_(const int* n, const pb200_complex8* x, const int* incx);
double pdznrm2_(const int* n, const pb200_complex16* x, const int* incx);
float psdot_(const int* n, const float* x, const int* incx, const float* y, const int* incy);
double pddot_(const int* n, const double* x, const int* incx, const double* y, const int* incy);
pb200_complex8 pcdotc_(const int* n, const pb200_complex8* x, const int* incx, const pb200_complex8* y, const int* incy);
pb200_complex16 pzdotc_(const int* n, const pb200_complex16* x, const int* incx, const pb200_complex16* y, const int* incy);
pb200_complex8 pcdotu_(const int* n, const pb200_complex8* x, const int* incx, const pb200_complex8* y, const int* incy);
pb200_complex16 pzdotu_(const int* n, const pb200_complex16* x, const int* incx, const pb200_complex16* y, const int* incy);
void psaxpy_(const int* n, const float* alpha, const float* x, const int* incx, float* y, const int* incy);
void pdaxpy_(const int* n, const double* alpha, const double* x, const int* incx, double* y, const int* incy);
void pcaxpy_(const int* n, const pb200_complex8* alpha, const pb200_complex8* x, const int* incx, pb200_complex8* y, const int* incy);
void pzaxpy_(const int* n, const pb200_complex16* alpha, const pb200_complex16* x, const int* incx, pb200_complex16* y, const int* incy);
void pcsaxpy_(const int* n, const float* alpha, const pb200_complex8* x, const int* incx, pb200_complex8* y, const int* incy);
void pzdaxpy_(const int* n, const double* alpha, const pb200_complex16* x, const int* incx, pb200_complex16* y, const int* incy);
void psscal_(const int* n, const float* alpha, float* x, const int* incx);
void pdscal_(const int* n, const double* alpha, double* x, const int* incx);
void pcsscal_(const int* n, const float* alpha, pb200_complex8* x, const int* incx);
void pzdscal_(const int* n, const double* alpha, pb200_complex16* x, const int* incx);
void pszero_(const int* n, float* x, const int* incx);
void pdzero_(const int* n, double* x, const int* incx);
void pczero_(const int* n, pb200_complex8* x, const int* incx);
void pzzero_(const int* n, pb200_complex16* x, const int* incx);

/* Host bidiagonal algebra -- reference double/dbsvd.F: dbsvdstep :5, dbdqr :87, drefinebounds :162;
 * omega-recurrence helpers double/dlanbpro.F: dset_mu :555, dcompute_int :581, dupdate_mu :628, dupdate_nu :684. */
void sbsvdstep_(const char* jobu, const char* jobv, const int* m, const int* n, const int* k, const float* sigma, float* D, float* E,
                float* U, const int* ldu, float* V, const int* ldv, size_t, size_t);
void dbsvdstep_(const char* jobu, const char* jobv, const int* m, const int* n, const int* k, const double* sigma, double* D, double* E,
                double* U, const int* ldu, double* V, const int* ldv, size_t, size_t);
void sbdqr_(const int* ignorelast, const char* jobq, const int* n, float* D, float* E, float* c1, float* c2, float* Qt, const int* ldq, size_t);
void dbdqr_(const int* ignorelast, const char* jobq, const int* n, double* D, double* E, double* c1, double* c2, double* Qt, const int* ldq, size_t);
void srefinebounds_(const int* n, const int* k, const float* theta, float* bound, const float* tol, const float* eps34);
void drefinebounds_(const int* n, const int* k, const double* theta, double* bound, const double* tol, const double* eps34);
void sset_mu_(const int* k, float* mu, const int* index, const float* val);
void dset_mu_(const int* k, double* mu, const int* index, const double* val);
void scompute_int_(const float* mu, const int* j, const float* delta, const float* eta, int* index);
void dcompute_int_(const double* mu, const int* j, const double* delta, const double* eta, int* index);
void supdate_mu_(float* mumax, float* mu, const float* nu, const int* j, const float* alpha, const float* beta, const float* anorm, const float* eps1);
void dupdate_mu_(double* mumax, double* mu, const double* nu, const int* j, const double* alpha, const double* beta, const double* anorm, const double* eps1);
void supdate_nu_(float* numax, const float* mu, float* nu, const int* j, const float* alpha, const float* beta, const float* anorm, const float* eps1);
void dupdate_nu_(double* numax, const double* mu, double* nu, const int* j, const double* alpha, const double* beta, const double* anorm, const double* eps1);

/* COMMON /timing/ -- reference double/stat.h:7-15 (same member order), clearstat/printstat double/printstat.F:4,35 */
struct pb200_timing_common {
  int nopx, nreorth, ndot, nreorthu, nreorthv, nitref, nrestart, nbsvd;
  float tmvopx, tgetu0, tupdmu, tupdnu, tintv, tlanbpro, treorth, treorthu, treorthv, telru, telrv, tbsvd, tnorm2, tlansvd;
  int nlandim;
  float tritzvec, trestart, tdot;
  int nsing;
};
extern struct pb200_timing_common timing_;
void clearstat_(void);
void printstat_(void);

/* =====================================================================================================
 * (2) device-resident operators (new; the GPU addendum of SURVEY.md section 8b)
 * ===================================================================================================== */

/* Register an m x n CSR matrix (0- or 1-based indices, int32; the three arrays may be host pointers or pointers into this
 * device's memory, e.g. a torch CSR tensor).  The library copies it and builds the CSR of A^T and the kernel-side
 * layouts on the device.  Returns a handle > 0, or a negative error code. */
int propack_b200_csr_create_s(int m, int n, const int* rowptr, const int* colind, const float* values, int index_base);
int propack_b200_csr_create_d(int m, int n, const int* rowptr, const int* colind, const double* values, int index_base);
int propack_b200_csr_create_c(int m, int n, const int* rowptr, const int* colind, const pb200_complex8* values, int index_base);
int propack_b200_csr_create_z(int m, int n, const int* rowptr, const int* colind, const pb200_complex16* values, int index_base);
/* Copy back the device-built transpose (CSR of A^T, 0-based, sorted) -- integer work is bit-exact and testable. */
int propack_b200_csr_get_transpose(int handle, int* t_rowptr, int* t_colind, void* t_values);
/* The sliced jagged-ELL copy the default SpMV kernel streams (adjoint = 1: the copy of A^T; panel = column block, 0 for
 * operators whose gathered vector fits L2): info4 = slices (32 rows each), stored entries, number of panels, long-row
 * threshold; then the arrays (slice_offsets[slices+1]; row_len[rows], 0xFF = long row kept in CSR; colind/values[stored],
 * inside a slice first the 0-th entries of its rows in row order, then the 1-st entries, ...). */
int propack_b200_csr_sell_info(int handle, int adjoint, int panel, long long* info4);
int propack_b200_csr_get_sell(int handle, int adjoint, int panel, long long* slice_offsets, unsigned char* row_len, int* colind,
                              void* values);
/* Dense column-major m x n operator; host array is copied (…_create) or a device array is adopted, not copied
 * (…_adopt_device; it must stay alive; base 16-byte aligned, lda in elements and a multiple of the 128-bit pack (2 doubles),
 * rows m..lda-1 of every column zero -- the GEMV kernels read whole packs). */
int propack_b200_dense_create_s(int m, int n, const float* A, long lda);
int propack_b200_dense_create_d(int m, int n, const double* A, long lda);
int propack_b200_dense_create_c(int m, int n, const pb200_complex8* A, long lda);
int propack_b200_dense_create_z(int m, int n, const pb200_complex16* A, long lda);
int propack_b200_dense_adopt_device_s(int m, int n, const float* A_device, long lda);
int propack_b200_dense_adopt_device_d(int m, int n, const double* A_device, long lda);
int propack_b200_dense_adopt_device_c(int m, int n, const pb200_complex8* A_device, long lda);
int propack_b200_dense_adopt_device_z(int m, int n, const pb200_complex16* A_device, long lda);
/* synthetic dense operator generated on the device (BASELINE config 3: 2M x 4096 = 65.5 GB never exists on the host):
 * A(i,j) = u(i,j) + sum_g table[g][byte_g(X(i)^Y(j))], see csrc/dense_gen.cu and propack_b200/synth.py (bit-identical
 * numpy replica for parity tests).  Plays the role of the user's APROD data (double/dlansvd.F:20-33). */
int propack_b200_dense_create_synthetic_d(int m, int n, unsigned long long seed, const double* table16x256);
int propack_b200_op_destroy(int handle);   /* drops the handle; solver sessions created on it keep the operator alive */
/* HBM bytes one product moves by the SURVEY 8(d) model (adjoint = 0: A x, 1: A^H x) */
double propack_b200_op_bytes(int handle, int adjoint);

/* The APROD to pass to the drivers with the handle in iparm[0].  Called directly it is also a complete
 * host-pointer APROD (x up, product on the GPU, y down). */
void propack_b200_aprod_s_(const char* transa, const int* m, const int* n, const float* x, float* y, float* sparm, int* iparm, size_t);
void propack_b200_aprod_d_(const char* transa, const int* m, const int* n, const double* x, double* y, double* dparm, int* iparm, size_t);
void propack_b200_aprod_c_(const char* transa, const int* m, const int* n, const pb200_complex8* x, pb200_complex8* y, pb200_complex8* cparm, int* iparm, size_t);
void propack_b200_aprod_z_(const char* transa, const int* m, const int* n, const pb200_complex16* x, pb200_complex16* y, pb200_complex16* zparm, int* iparm, size_t);

/* =====================================================================================================
 * (3) solver sessions: bases stay in HBM; only sigma/bnd/scalars cross PCIe unless vectors are fetched
 * ===================================================================================================== */
int propack_b200_solver_create(int op_handle, int ucols, int vcols);           /* returns solver id > 0 */
int propack_b200_solver_destroy(int solver);
int propack_b200_solver_set_start(int solver, const void* u0_host);            /* NULL => zero => random start */
int propack_b200_solver_lansvd(int solver, int jobu, int jobv, int* k, int kmax, void* sigma, void* bnd, double tolin,
                               void* option3, int* ioption, int* info);         /* sigma/bnd/option in the real type of the operator */
int propack_b200_solver_lansvd_irl(int solver, int which_smallest, int jobu, int jobv, int* dim, int p, int* neig, int maxiter,
                                   void* sigma, void* bnd, double tolin, void* option4, int* ioption, int* info);
int propack_b200_solver_get_u(int solver, int ncols, void* U_host, long ldu);
int propack_b200_solver_get_v(int solver, int ncols, void* V_host, long ldv);

/* ---- runtime --------------------------------------------------------------------------------------- */
/* ---- multi-GPU: one process per GPU.  The reference has no distributed layer; its only
 * parallelism is OpenMP row-chunking of the same loops (double/dreorth.F:147-208, double/dritzvec.F:145-196), which
 * is the scheme these entry points lift across GPUs.  Rank 0 calls _comm_unique_id, the caller broadcasts the 128
 * bytes (MPI, torch.distributed, a file ...), every rank calls _comm_init (collective). ---- */
int propack_b200_comm_unique_id(void* id128_out);
int propack_b200_comm_init(int rank, int world, const void* id128);
int propack_b200_comm_finalize(void);
int propack_b200_comm_rank(void);
int propack_b200_comm_world(void);
/* block partition used for rows of A / U (dim = m) and for V-vectors (dim = n): rank r owns [lo, hi) */
long propack_b200_shard_slice(long dim, int world);
void propack_b200_shard_bounds(long dim, int world, int rank, long* lo, long* hi);
void propack_b200_comm_stats(long long* n_allreduce, long long* n_allgather, double* allgather_bytes);
/* Row-sharded operator of THIS rank: row_* = CSR of A[r0:r1, :] (global column ids); colt_* = CSR of (A[:, c0:c1])^T
 * (global row ids); [r0,r1) and [c0,c1) from _shard_bounds for this rank.  Replaces the user's APROD (contract
 * double/dlansvd.F:20-33) by: all-gather the input vector, then a local SpMV (both directions). */
int propack_b200_csr_create_sharded_s(int m_global, int n_global, const int* row_rowptr, const int* row_colind, const float* row_values,
                                      const int* colt_rowptr, const int* colt_rowind, const float* colt_values, int index_base);
int propack_b200_csr_create_sharded_d(int m_global, int n_global, const int* row_rowptr, const int* row_colind, const double* row_values,
                                      const int* colt_rowptr, const int* colt_rowind, const double* colt_values, int index_base);
int propack_b200_csr_create_sharded_c(int m_global, int n_global, const int* row_rowptr, const int* row_colind, const pb200_complex8* row_values,
                                      const int* colt_rowptr, const int* colt_rowind, const pb200_complex8* colt_values, int index_base);
int propack_b200_csr_create_sharded_z(int m_global, int n_global, const int* row_rowptr, const int* row_colind, const pb200_complex16* row_values,
                                      const int* colt_rowptr, const int* colt_rowind, const pb200_complex16* colt_values, int index_base);
/* Row-sharded DENSE operator of THIS rank (BASELINE configs[2] on N GPUs): A_rows = A[r0:r1, :] column-major with leading
 * dimension lda (host or device memory), [r0,r1) from _shard_bounds(m_global).  A x = all-gather of the n-vector + local GEMV;
 * A^H u = local GEMV^T + all-reduce of the n coefficients, every rank keeping its slice.  Replaces the user's APROD
 * (double/dlansvd.F:20-33) the way the reference's OpenMP build chunks rows (double/dreorth.F:147-208). */
int propack_b200_dense_create_sharded_s(int m_global, int n_global, const float* A_rows, long lda);
int propack_b200_dense_create_sharded_d(int m_global, int n_global, const double* A_rows, long lda);
int propack_b200_dense_create_sharded_c(int m_global, int n_global, const pb200_complex8* A_rows, long lda);
int propack_b200_dense_create_sharded_z(int m_global, int n_global, const pb200_complex16* A_rows, long lda);
/* this rank's rows of the synthetic dense matrix of _dense_create_synthetic_d (same matrix, evaluated on the device) */
int propack_b200_dense_create_synthetic_sharded_d(int m_global, int n_global, unsigned long long seed, const double* table16x256);
/* local slice sizes of a solver session (U is m_local x ucols with leading dimension ldu on the device, ...) */
int propack_b200_solver_local_rows(int solver, int* m_local, int* n_local, long* ldu, long* ldv);

/* host-only test hook: leading Ritz values and |last component of the left singular vector| of the (j+1) x j lower
 * bidiagonal (alpha, beta) by the reference route (method 0: dbdqr + dbdsqr, double/dlansvd.F:193-199) or the fast route
 * used for large j (method 1: dqds + inverse iteration); returns 1 when method 1 declines */
int propack_b200_host_ritz_bounds_d(int j, const double* alpha, const double* beta, int K, int method, double* theta, double* last);
/* host-only test hook: the (dim+1) x k and dim x k matrices that dritzvec (double/dritzvec.F:116-193) multiplies the Lanczos
 * bases with, by the reference route (method 0) or the fast route (method 1; returns 1 when it declines) */
int propack_b200_host_ritz_vectors_d(int dim, const double* alpha, const double* beta, int k, int method, double* WU, double* WV);
/* host-only test hook: the shifted QR sweeps of one implicit restart (double/dlansvd_irl.F:350-363, dbsvd.F:5-82) accumulating
 * P ((dim+1) x (dim+1)) and Q (dim x dim); nthreads = 0 is the sequential reference order, >= 1 the row-parallel route */
int propack_b200_host_restart_sweeps_d(int dim, int k, const double* shift, double* alpha, double* beta, double* P, double* Q, int nthreads);
int propack_b200_init(void);                         /* create the context on the current device; 0 or negative */
int propack_b200_set_stream(void* cuda_stream);      /* run on a caller stream (e.g. torch's current stream) */
int propack_b200_set_lapack(const char* path);       /* shared object providing {d,s}bdsqr / {d,s}bdsdc */
int propack_b200_set_option(const char* name, int value); /* "peer_timeout_s": give up on a silent peer rank after this many seconds (default 30); "bench_skip_gather": measurement only */
void propack_b200_release_cache(void);               /* free the Krylov-basis buffers parked for reuse by the next driver call of the same shape */
void propack_b200_set_profile(int on);               /* per-phase CUDA-event timers (adds synchronisation) */
void propack_b200_reset_counters(void);
/* out[0..15] = nopx nreorth ndot nitref nrestart nbsvd nlandim nsing nsteps reorth_passes reorth_cols
 *              reorth_elems reorth_vec_elems launches host_syncs reserved */
void propack_b200_get_counters(long long* out);
/* out[0..6] = ms in aprod, reorth, level1, getu0, ritzvec, restart, host_bsvd (profile mode); launches[0..6] likewise */
void propack_b200_get_phase_ms(double* out_ms, long long* out_launches);
const char* propack_b200_last_error(void);
int propack_b200_device_sms(void);

/* Micro-benchmark hooks used by bench.py / profiles: run one kernel `reps` times on device-resident
 * synthetic data and return the mean CUDA-event milliseconds per launch (negative on error). */
double propack_b200_bench_reorth_d(long L, int l, int reps, int flush_l2);          /* one GEMV pair + norm */
double propack_b200_bench_spmv(int op_handle, int adjoint, int reps, int flush_l2); /* fused SpMV */
double propack_b200_bench_gemm_d(long M, int N, int K, int reps);                    /* tall in-place GEMM */

#ifdef __cplusplus
}
#endif
#endif /* PROPACK_B200_H */
