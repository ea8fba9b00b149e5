// =============================================================================
// ORACLE -- TEST INFRASTRUCTURE ONLY (see propack_oracle.hpp header).
// extern "C" entry points over the templated CPU restatement, for ctypes.
// =============================================================================
#include "propack_oracle.hpp"

#include <dlfcn.h>

namespace oracle {
Lapack& lapack() { static Lapack l; return l; }
}  // namespace oracle

using namespace oracle;
typedef std::complex<float> cfloat;
typedef std::complex<double> cdouble;

extern "C" {

// Resolve {d,s}bdsqr / {d,s}bdsdc from a LAPACK shared object (scipy_openblas uses a
// "scipy_" symbol prefix; plain LAPACK uses none).  Returns 0 on success.
int oracle_init_lapack(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) { std::fprintf(stderr, "oracle: dlopen(%s) failed: %s\n", path, dlerror()); return -1; }
  auto sym = [&](const char* name) -> void* {
    std::string a = std::string("scipy_") + name + "_", b = std::string(name) + "_";
    void* p = dlsym(h, a.c_str());
    if (!p) p = dlsym(h, b.c_str());
    return p;
  };
  Lapack& l = lapack();
  l.dbdsqr = (bdsqr_d_t)sym("dbdsqr"); l.sbdsqr = (bdsqr_s_t)sym("sbdsqr");
  l.dbdsdc = (bdsdc_d_t)sym("dbdsdc"); l.sbdsdc = (bdsdc_s_t)sym("sbdsdc");
  return (l.dbdsqr && l.sbdsqr && l.dbdsdc && l.sbdsdc) ? 0 : -2;
}

void oracle_stats_reset() { stats() = Stats(); }
// out[0..9] = nopx nreorth ndot nitref nrestart nbsvd nlandim nsing nsteps reorth_cols
void oracle_stats_get(long long* out) {
  Stats& s = stats();
  out[0] = s.nopx; out[1] = s.nreorth; out[2] = s.ndot; out[3] = s.nitref; out[4] = s.nrestart;
  out[5] = s.nbsvd; out[6] = s.nlandim; out[7] = s.nsing; out[8] = s.nsteps; out[9] = s.reorth_cols;
}
int oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

}  // extern "C"

// A user callback in C form: transa is the ASCII code of 'n' / 't' / 'c'.
template <class T> struct Op {
  CsrOp<T> csr;
  void (*cb)(int transa, int m, int n, const T* x, T* y);
};
template <class T> static void op_aprod(char transa, int m, int n, const T* x, T* y, void* ctx) {
  Op<T>* op = static_cast<Op<T>*>(ctx);
  if (op->cb) op->cb((int)transa, m, n, x, y);
  else csr_aprod<T>(transa, m, n, x, y, &op->csr);
}

#define ORACLE_API(SFX, T, R)                                                                                         \
  extern "C" {                                                                                                        \
  void oracle_larnv_##SFX(int* iseed, long n, T* x) { larnv2(iseed, n, x); }                                          \
  R oracle_nrm2_##SFX(long n, const T* x) { return pnrm2<T>(n, x); }                                                  \
  void oracle_reorth_##SFX(long n, int k, const T* V, long ldv, T* vnew, R* normvnew, const int* index, R alpha,     \
                           int iflag) {                                                                               \
    std::vector<T> work(k + 1);                                                                                       \
    reorth<T>(n, k, V, ldv, vnew, *normvnew, index, alpha, work.data(), iflag);                                      \
  }                                                                                                                   \
  void oracle_gemm_ovwr_left_##SFX(int transb, long m, int n, int k, T* A, long lda, const R* B, int ldb) {          \
    gemm_ovwr_left<T>((char)transb, m, n, k, A, lda, B, ldb);                                                         \
  }                                                                                                                   \
  void oracle_safescal_##SFX(long n, R alpha, T* x) { safescal<T>(n, alpha, x); }                                     \
  void oracle_dotc_##SFX(long n, const T* x, const T* y, T* out) { *out = pdotc<T>(n, x, y); }                        \
  void oracle_axpy_##SFX(long n, const T* alpha, const T* x, T* y) { paxpy<T, T>(n, *alpha, x, y); }                  \
  void oracle_scal_##SFX(long n, R alpha, T* x) { pscal<T, R>(n, alpha, x); }                                         \
  void oracle_ritzvec_##SFX(int which, int jobu, int jobv, int m, int n, int k, int dim, R* D, R* E, T* U, long ldu, \
                            T* V, long ldv) {                                                                         \
    ritzvec<T>((char)which, jobu != 0, jobv != 0, m, n, k, dim, D, E, U, ldu, V, ldv);                                \
  }                                                                                                                   \
  void oracle_csr_aprod_##SFX(int transa, int m, int n, const int* rp, const int* ci, const T* va, const int* trp,   \
                              const int* tci, const T* tva, const T* x, T* y) {                                       \
    CsrOp<T> A{m, n, rp, ci, va, trp, tci, tva};                                                                      \
    csr_aprod<T>((char)transa, m, n, x, y, &A);                                                                       \
  }                                                                                                                   \
  void oracle_getu0_##SFX(int transa, int m, int n, int j, int ntry, T* u0, R* u0norm, const T* U, long ldu,         \
                          const int* rp, const int* ci, const T* va, const int* trp, const int* tci, const T* tva,   \
                          void (*cb)(int, int, int, const T*, T*), int* ierr, int icgs, R* anormest) {                \
    Op<T> op{{m, n, rp, ci, va, trp, tci, tva}, cb};                                                                  \
    std::vector<T> work(size_t(std::max(m, n)) + 1);                                                                  \
    getu0<T>((char)transa, m, n, j, ntry, u0, *u0norm, U, ldu, op_aprod<T>, &op, *ierr, icgs, *anormest,              \
             work.data());                                                                                            \
  }                                                                                                                   \
  void oracle_lanbpro_##SFX(int m, int n, int k0, int* k, const int* rp, const int* ci, const T* va, const int* trp, \
                            const int* tci, const T* tva, void (*cb)(int, int, int, const T*, T*), T* U, long ldu,    \
                            T* V, long ldv, R* B, int ldb, R* rnorm, R* doption, const int* ioption, int* ierr) {     \
    Op<T> op{{m, n, rp, ci, va, trp, tci, tva}, cb};                                                                  \
    lanbpro<T>(m, n, k0, *k, op_aprod<T>, &op, U, ldu, V, ldv, B, B + ldb, *rnorm, doption, ioption, *ierr);          \
  }                                                                                                                   \
  void oracle_lansvd_##SFX(int jobu, int jobv, int m, int n, int* k, int kmax, const int* rp, const int* ci,          \
                           const T* va, const int* trp, const int* tci, const T* tva,                                 \
                           void (*cb)(int, int, int, const T*, T*), T* U, long ldu, R* sigma, R* bnd, T* V,           \
                           long ldv, R tolin, R* doption, const int* ioption, int* info) {                            \
    Op<T> op{{m, n, rp, ci, va, trp, tci, tva}, cb};                                                                  \
    lansvd<T>(jobu != 0, jobv != 0, m, n, *k, kmax, op_aprod<T>, &op, U, ldu, sigma, bnd, V, ldv, tolin, doption,     \
              ioption, *info);                                                                                        \
  }                                                                                                                   \
  void oracle_lansvd_irl_##SFX(int which, int jobu, int jobv, int m, int n, int* dim, int p, int* neig, int maxiter, \
                               const int* rp, const int* ci, const T* va, const int* trp, const int* tci,             \
                               const T* tva, void (*cb)(int, int, int, const T*, T*), T* U, long ldu, R* sigma,       \
                               R* bnd, T* V, long ldv, R tolin, R* doption, const int* ioption, int* info) {          \
    Op<T> op{{m, n, rp, ci, va, trp, tci, tva}, cb};                                                                  \
    lansvd_irl<T>((char)which, jobu != 0, jobv != 0, m, n, *dim, p, *neig, maxiter, op_aprod<T>, &op, U, ldu, sigma,  \
                  bnd, V, ldv, tolin, doption, ioption, *info);                                                       \
  }                                                                                                                   \
  }

ORACLE_API(s, float, float)
ORACLE_API(d, double, double)
ORACLE_API(c, cfloat, float)
ORACLE_API(z, cdouble, double)

// real-only host algebra, exposed so the product's host logic can be checked against it
#define ORACLE_REAL_API(SFX, R)                                                                                      \
  extern "C" {                                                                                                       \
  void oracle_compute_int_##SFX(const R* mu, int j, R delta, R eta, int* index) {                                    \
    compute_int<R>(mu - 1, j, delta, eta, index - 1);                                                                \
  }                                                                                                                  \
  void oracle_set_mu_##SFX(int k, R* mu, const int* index, R val) { set_mu<R>(k, mu - 1, index - 1, val); }          \
  void oracle_update_mu_##SFX(R* mumax, R* mu, const R* nu, int j, const R* alpha, const R* beta, R anorm, R eps1) { \
    update_mu<R>(*mumax, mu - 1, nu - 1, j, alpha - 1, beta - 1, anorm, eps1);                                       \
  }                                                                                                                  \
  void oracle_update_nu_##SFX(R* numax, const R* mu, R* nu, int j, const R* alpha, const R* beta, R anorm, R eps1) { \
    update_nu<R>(*numax, mu - 1, nu - 1, j, alpha - 1, beta - 1, anorm, eps1);                                       \
  }                                                                                                                  \
  void oracle_bdqr_##SFX(int ignorelast, int jobq, int n, R* D, R* E, R* c1, R* c2, R* Qt, int ldq) {                \
    bdqr<R>(ignorelast != 0, jobq != 0, n, D, E, *c1, *c2, Qt, ldq);                                                 \
  }                                                                                                                  \
  void oracle_bsvdstep_##SFX(int jobu, int jobv, int m, int n, int k, R sigma, R* D, R* E, R* U, int ldu, R* V,      \
                             int ldv) {                                                                              \
    bsvdstep<R>(jobu != 0, jobv != 0, m, n, k, sigma, D, E, U, ldu, V, ldv);                                         \
  }                                                                                                                  \
  void oracle_refinebounds_##SFX(int n, int k, const R* theta, R* bound, R tol, R eps34) {                           \
    refinebounds<R>(n, k, theta, bound, tol, eps34);                                                                 \
  }                                                                                                                  \
  void oracle_lartg_##SFX(R f, R g, R* cs, R* sn, R* r) { lartg<R>(f, g, *cs, *sn, *r); }                            \
  }

ORACLE_REAL_API(s, float)
ORACLE_REAL_API(d, double)
