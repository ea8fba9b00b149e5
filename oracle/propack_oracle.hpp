// =============================================================================
// ORACLE -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
//
// CPU restatement (C++17 + OpenMP) of PROPACK's Lanczos-bidiagonalization hot
// path, following the Fortran in /root/reference line by line.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build, load or call anything in this directory.
//
// PARITY STATUS: "parity unpinned" at the Fortran level -- the reference ships
// no golden vectors, no tests and cannot be compiled (no Fortran compiler in
// the image, Examples/ + second.F missing).  The oracle is
// pinned instead against (1) scipy 1.18.1's `_propack` C translation of this
// same code base on the two PROPACK example matrices (illc1850, mhd1280b from
// scipy's propack_test_data.npz) and on seeded random matrices -- fixtures
// under tests/golden/ made by tools/make_golden.py; (2) the LAPACK DLARNV
// known-answer vector for iseed=(1,3,5,7); (3) dense LAPACK SVD.
//
// Every routine cites the reference file:line it follows (paths relative to
// /root/reference/double unless stated; s/c/z variants are the mechanical
// copies listed in SURVEY.md section 2.3 and are produced here by templates,
// honouring the non-mechanical complex differences of that section).
//
// Third-party arithmetic not in the reference tree: BLAS level 1/2/3 (chosen at
// link time upstream, no version pin) is restated as plain OpenMP loops in the
// style of the reference's own _OPENMP branch (dblasext.F:5-150); LAPACK
// DBDSQR/DBDSDC (vendored LAPACK 3.0 upstream) are taken from the image's
// scipy_openblas (LAPACK 3.11 semantics, same algorithms) through dlopen.
// =============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

// ----------------------------------------------------------------------------
// COMMON /timing/ counters (stat.h:7-15) -- only the live integer ones.
// ----------------------------------------------------------------------------
struct Stats {
  int nopx = 0, nreorth = 0, ndot = 0, nitref = 0, nrestart = 0, nbsvd = 0;
  int nlandim = 0, nsing = 0;
  int nsteps = 0;          // iterations of the dlanbpro.F:283 loop (not in the reference)
  long long reorth_cols = 0; // sum over passes of interval lengths (byte model)
};
inline Stats& stats() { static Stats s; return s; }

// ----------------------------------------------------------------------------
// scalar traits
// ----------------------------------------------------------------------------
template <class T> struct traits {
  using real = T;
  static constexpr bool is_complex = false;
  static T conj(T x) { return x; }
  static T from_real(real x) { return x; }
  static real abs2(T x) { return x * x; }
};
template <class R> struct traits<std::complex<R>> {
  using real = R;
  static constexpr bool is_complex = true;
  static std::complex<R> conj(std::complex<R> x) { return std::conj(x); }
  static std::complex<R> from_real(R x) { return std::complex<R>(x, 0); }
  static R abs2(std::complex<R> x) { return x.real() * x.real() + x.imag() * x.imag(); }
};

// dlamch('e') / dlamch('s')  (Lapack_Util/dlamch.f:79-84: eps = relative machine
// precision *with rounding* = 2^-53 for IEEE double, 2^-24 for single)
template <class R> inline R lamch_e();
template <> inline double lamch_e<double>() { return 1.1102230246251565e-16; }
template <> inline float lamch_e<float>() { return 5.9604644775390625e-8f; }
template <class R> inline R lamch_s();
template <> inline double lamch_s<double>() { return 2.2250738585072014e-308; }
template <> inline float lamch_s<float>() { return 1.17549435e-38f; }

// dlapy2 (Lapack_Util/dlapy2.f): sqrt(x^2+y^2) without unnecessary overflow
template <class R> inline R lapy2(R x, R y) {
  R xa = std::fabs(x), ya = std::fabs(y);
  R w = std::max(xa, ya), z = std::min(xa, ya);
  if (z == R(0)) return w;
  return w * std::sqrt(R(1) + (z / w) * (z / w));
}

// dlartg (Lapack_Util/dlartg.f, LAPACK 3.0 sign convention)
template <class R> inline void lartg(R f, R g, R& cs, R& sn, R& r) {
  static const R safmin = lamch_s<R>();
  static const R eps = lamch_e<R>();
  static const R safmn2 = std::pow(R(2), R(int(std::log(safmin / eps) / std::log(R(2)) / R(2))));
  static const R safmx2 = R(1) / safmn2;
  if (g == R(0)) { cs = 1; sn = 0; r = f; return; }
  if (f == R(0)) { cs = 0; sn = 1; r = g; return; }
  R f1 = f, g1 = g;
  R scale = std::max(std::fabs(f1), std::fabs(g1));
  if (scale >= safmx2) {
    int count = 0;
    do { ++count; f1 *= safmn2; g1 *= safmn2; scale = std::max(std::fabs(f1), std::fabs(g1)); } while (scale >= safmx2);
    r = std::sqrt(f1 * f1 + g1 * g1); cs = f1 / r; sn = g1 / r;
    for (int i = 0; i < count; ++i) r *= safmx2;
  } else if (scale <= safmn2) {
    int count = 0;
    do { ++count; f1 *= safmx2; g1 *= safmx2; scale = std::max(std::fabs(f1), std::fabs(g1)); } while (scale <= safmn2);
    r = std::sqrt(f1 * f1 + g1 * g1); cs = f1 / r; sn = g1 / r;
    for (int i = 0; i < count; ++i) r *= safmn2;
  } else {
    r = std::sqrt(f1 * f1 + g1 * g1); cs = f1 / r; sn = g1 / r;
  }
  if (std::fabs(f) > std::fabs(g) && cs < R(0)) { cs = -cs; sn = -sn; r = -r; }
}

// ----------------------------------------------------------------------------
// LAPACK bidiagonal SVD entry points resolved from scipy_openblas at run time
// (dlansvd.F:198 dbdsqr; dritzvec.F:123 dbdsdc).  Fortran ABI, LP64.
// ----------------------------------------------------------------------------
extern "C" {
typedef void (*bdsqr_d_t)(const char*, const int*, const int*, const int*, const int*, double*, double*,
                          double*, const int*, double*, const int*, double*, const int*, double*, int*, size_t);
typedef void (*bdsqr_s_t)(const char*, const int*, const int*, const int*, const int*, float*, float*,
                          float*, const int*, float*, const int*, float*, const int*, float*, int*, size_t);
typedef void (*bdsdc_d_t)(const char*, const char*, const int*, double*, double*, double*, const int*, double*,
                          const int*, double*, int*, double*, int*, int*, size_t, size_t);
typedef void (*bdsdc_s_t)(const char*, const char*, const int*, float*, float*, float*, const int*, float*,
                          const int*, float*, int*, float*, int*, int*, size_t, size_t);
}
struct Lapack {
  bdsqr_d_t dbdsqr = nullptr; bdsqr_s_t sbdsqr = nullptr;
  bdsdc_d_t dbdsdc = nullptr; bdsdc_s_t sbdsdc = nullptr;
};
Lapack& lapack();  // defined in propack_oracle.cpp (dlopen)

inline void bdsqr(const char* uplo, int n, int ncvt, int nru, int ncc, double* d, double* e, double* vt, int ldvt,
                  double* u, int ldu, double* c, int ldc, double* work, int* info) {
  lapack().dbdsqr(uplo, &n, &ncvt, &nru, &ncc, d, e, vt, &ldvt, u, &ldu, c, &ldc, work, info, 1);
}
inline void bdsqr(const char* uplo, int n, int ncvt, int nru, int ncc, float* d, float* e, float* vt, int ldvt,
                  float* u, int ldu, float* c, int ldc, float* work, int* info) {
  lapack().sbdsqr(uplo, &n, &ncvt, &nru, &ncc, d, e, vt, &ldvt, u, &ldu, c, &ldc, work, info, 1);
}
inline void bdsdc(const char* uplo, const char* compq, int n, double* d, double* e, double* u, int ldu, double* vt,
                  int ldvt, double* q, int* iq, double* work, int* iwork, int* info) {
  lapack().dbdsdc(uplo, compq, &n, d, e, u, &ldu, vt, &ldvt, q, iq, work, iwork, info, 1, 1);
}
inline void bdsdc(const char* uplo, const char* compq, int n, float* d, float* e, float* u, int ldu, float* vt,
                  int ldvt, float* q, int* iq, float* work, int* iwork, int* info) {
  lapack().sbdsdc(uplo, compq, &n, d, e, u, &ldu, vt, &ldvt, q, iq, work, iwork, info, 1, 1);
}

// ----------------------------------------------------------------------------
// dlarnv(idist=2) / zlarnv(idist=2)  (Lapack_Util/dlarnv.f, dlaruv.f:335-363,
// complex16/Lapack_Util/zlarnv.f).  DLARUV multiplies the 48-bit seed by the
// i-th power of a = 494*2^36+322*2^24+2508*2^12+2549 modulo 2^48 (its MM table
// holds a^i, i=1..128) and returns the last state as the new seed, so the
// stream is s_i = s_0 * a^i mod 2^48, restated here with 64-bit integers.
// The value is R*(IT1+R*(IT2+R*(IT3+R*IT4))), R=1/4096, evaluated in the
// working precision exactly as written (double: exact; single: 3 roundings).
// ----------------------------------------------------------------------------
static constexpr uint64_t LARUV_A = 33952834046453ull;
static constexpr uint64_t MASK48 = (1ull << 48) - 1;
template <class R> inline R laruv_value(uint64_t s) {
  const R r = R(1) / R(4096);
  volatile R t = R(double(s & 4095));              // volatile: forbid fma contraction / reassociation
  t = r * t; t = R(double((s >> 12) & 4095)) + t;
  t = r * t; t = R(double((s >> 24) & 4095)) + t;
  t = r * t; t = R(double((s >> 36) & 4095)) + t;
  t = r * t;
  return t;
}
inline uint64_t seed_to_u64(const int iseed[4]) {
  return (uint64_t(iseed[0]) << 36) | (uint64_t(iseed[1]) << 24) | (uint64_t(iseed[2]) << 12) | uint64_t(iseed[3]);
}
inline void u64_to_seed(uint64_t s, int iseed[4]) {
  iseed[0] = int((s >> 36) & 4095); iseed[1] = int((s >> 24) & 4095);
  iseed[2] = int((s >> 12) & 4095); iseed[3] = int(s & 4095);
}
template <class R> inline void larnv2(int iseed[4], long n, R* x) {  // dlarnv idist=2: uniform(-1,1)
  uint64_t s = seed_to_u64(iseed);
  for (long i = 0; i < n; ++i) { s = (s * LARUV_A) & MASK48; x[i] = R(2) * laruv_value<R>(s) - R(1); }
  u64_to_seed(s, iseed);
}
template <class R> inline void larnv2(int iseed[4], long n, std::complex<R>* x) {  // zlarnv idist=2
  uint64_t s = seed_to_u64(iseed);
  for (long i = 0; i < n; ++i) {
    s = (s * LARUV_A) & MASK48; R re = R(2) * laruv_value<R>(s) - R(1);
    s = (s * LARUV_A) & MASK48; R im = R(2) * laruv_value<R>(s) - R(1);
    x[i] = std::complex<R>(re, im);
  }
  u64_to_seed(s, iseed);
}

// ----------------------------------------------------------------------------
// blasext level-1 ops, _OPENMP branch (dblasext.F:6-150; zblasext.F)
// ----------------------------------------------------------------------------
template <class T> typename traits<T>::real pnrm2(long n, const T* x) {  // pdnrm2 dblasext.F:6-30 / pdznrm2
  using R = typename traits<T>::real;
  R sum = 0;
#pragma omp parallel for reduction(+ : sum) schedule(static)
  for (long i = 0; i < n; ++i) sum += traits<T>::abs2(x[i]);
  return std::sqrt(sum);
}
template <class T, class S> void pscal(long n, S alpha, T* x) {  // pdscal dblasext.F:38-60 / pzdscal
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) x[i] = alpha * x[i];
}
template <class T, class S> void paxpy(long n, S alpha, const T* x, T* y) {  // pdaxpy dblasext.F:92-115 / pzaxpy / pzdaxpy
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) y[i] = alpha * x[i] + y[i];
}
template <class T> T pdotc(long n, const T* x, const T* y) {  // pddot dblasext.F:121-147 / pzdotc (conj(x).y)
  using R = typename traits<T>::real;
  if constexpr (traits<T>::is_complex) {
    R sr = 0, si = 0;
#pragma omp parallel for reduction(+ : sr, si) schedule(static)
    for (long i = 0; i < n; ++i) { T p = std::conj(x[i]) * y[i]; sr += p.real(); si += p.imag(); }
    return T(sr, si);
  } else {
    R s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (long i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
  }
}
template <class T> void pzero(long n, T* x) {  // pdzero dblasext.F:202-225
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) x[i] = T(0);
}

// dsafescal (dsafescal.F:4-55): x <- x/alpha; dlascl branch for |alpha| < sfmin
template <class T> void safescal(long n, typename traits<T>::real alpha, T* x) {
  using R = typename traits<T>::real;
  const R sfmin = lamch_s<R>();
  if (std::fabs(alpha) >= sfmin) {
    pscal(n, R(1) / alpha, x);
  } else {
    // dlascl('General',..,cfrom=alpha,cto=1,..) (Lapack_Util/dlascl.f:113-137): multiply by
    // cto/cfrom in safe steps of smlnum / bignum.
    const R smlnum = sfmin, bignum = R(1) / smlnum;
    R cfromc = alpha, ctoc = 1;
    bool done = false;
    while (!done) {
      R cfrom1 = cfromc * smlnum, cto1 = ctoc / bignum, mul;
      if (std::fabs(cfrom1) > std::fabs(ctoc) && ctoc != R(0)) { mul = smlnum; done = false; cfromc = cfrom1; }
      else if (std::fabs(cto1) > std::fabs(cfromc)) { mul = bignum; done = false; ctoc = cto1; }
      else { mul = ctoc / cfromc; done = true; }
      pscal(n, mul, x);
    }
  }
}

// ----------------------------------------------------------------------------
// APROD contract (dlansvd.F:20-33): y = op(A) x, transa in {'n','t'} (complex: {'n','c'})
// ----------------------------------------------------------------------------
template <class T> using aprod_t = void (*)(char transa, int m, int n, const T* x, T* y, void* ctx);

// Built-in OpenMP CSR operator: holds CSR(A) and CSR(A^T) -- stand-in for the user's matvec.F
template <class T> struct CsrOp {
  int m, n;
  const int *rp, *ci; const T* va;     // CSR of A (m rows)
  const int *trp, *tci; const T* tva;  // CSR of A^T (n rows), values NOT conjugated
};
template <class T> void csr_aprod(char transa, int m, int n, const T* x, T* y, void* ctx) {
  const CsrOp<T>* A = static_cast<const CsrOp<T>*>(ctx);
  (void)m; (void)n;
  if (transa == 'n' || transa == 'N') {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < A->m; ++i) {
      T s = 0;
      for (int p = A->rp[i]; p < A->rp[i + 1]; ++p) s += A->va[p] * x[A->ci[p]];
      y[i] = s;
    }
  } else {
    const bool cj = (transa == 'c' || transa == 'C');
#pragma omp parallel for schedule(static)
    for (int i = 0; i < A->n; ++i) {
      T s = 0;
      if (cj) for (int p = A->trp[i]; p < A->trp[i + 1]; ++p) s += traits<T>::conj(A->tva[p]) * x[A->tci[p]];
      else    for (int p = A->trp[i]; p < A->trp[i + 1]; ++p) s += A->tva[p] * x[A->tci[p]];
      y[i] = s;
    }
  }
}

// ----------------------------------------------------------------------------
// dcgs (dreorth.F:106-210): block classical Gram-Schmidt over index intervals.
// SPMD over row chunks like the reference's OpenMP region (:147-208), with the
// per-thread partial coefficient vectors summed in thread order (the reference
// sums them in CRITICAL-section arrival order; SURVEY 2.4 notes its latent race).
// index is 1-based inclusive [s1,e1,...,T], consumers stop at index(i)>k or <=0.
// ----------------------------------------------------------------------------
template <class T> void cgs(long n, int k, const T* V, long ldv, T* vnew, const int* index, T* work) {
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  int i = 0;
  std::vector<T> ylocal;
  while (index[i] <= k && index[i] > 0) {
    const int p = index[i], q = index[i + 1];
    const int l = q - p + 1;
    stats().ndot += l;
    if (l > 0) {
      stats().reorth_cols += l;
      ylocal.assign(size_t(nt) * l, T(0));
      // y = V(:,p:q)^H vnew   (dgemv 'T' dreorth.F:174 / zgemv 'C' zreorth.F:169)
#pragma omp parallel num_threads(nt)
      {
        int tid = 0, nth = 1;
#ifdef _OPENMP
        tid = omp_get_thread_num(); nth = omp_get_num_threads();
#endif
        long cnk = n / nth, st = tid * cnk;
        if (tid == nth - 1) cnk = n - st;
        T* yl = ylocal.data() + size_t(tid) * l;
        for (int c = 0; c < l; ++c) {
          const T* col = V + (long)(p - 1 + c) * ldv + st;
          T s = 0;
          for (long r = 0; r < cnk; ++r) s += traits<T>::conj(col[r]) * vnew[st + r];
          yl[c] = s;
        }
      }
      for (int c = 0; c < l; ++c) { T s = ylocal[c]; for (int t = 1; t < nt; ++t) s += ylocal[size_t(t) * l + c]; work[c] = s; }
      // vnew += -V(:,p:q) y   (dgemv 'N' dreorth.F:199-205)
#pragma omp parallel num_threads(nt)
      {
        int tid = 0, nth = 1;
#ifdef _OPENMP
        tid = omp_get_thread_num(); nth = omp_get_num_threads();
#endif
        long cnk = n / nth, st = tid * cnk;
        if (tid == nth - 1) cnk = n - st;
        const long RB = 2048;  // row block so the vnew slice stays in L1/L2 across columns
        for (long r0 = 0; r0 < cnk; r0 += RB) {
          const long r1 = std::min(cnk, r0 + RB);
          for (int c = 0; c < l; ++c) {
            const T* col = V + (long)(p - 1 + c) * ldv + st;
            const T w = work[c];
            for (long r = r0; r < r1; ++r) vnew[st + r] -= col[r] * w;
          }
        }
      }
    }
    i += 2;
  }
}

// ----------------------------------------------------------------------------
// dmgs, "risc" variant (dmgs.risc.F:11-84; parallel pdmgs :87-160; zmgs.risc.F):
// column-sequential MGS with the software-pipelined fused loop
//   vn0 = vnew - coef*V(:,i-1); newcoef += conj(V(:,i))*vn0     (:65-75)
// ----------------------------------------------------------------------------
template <class T> void mgs(long n, int k, const T* V, long ldv, T* vnew, const int* index) {
  using R = typename traits<T>::real;
  if (k <= 0 || n <= 0) return;
  int iblck = 0;
  int p = index[iblck], q = index[iblck + 1];
  while (p <= k && p > 0 && p <= q) {
    stats().ndot += (q - p + 1);
    stats().reorth_cols += (q - p + 1);
    T coef = pdotc(n, V + (long)(p - 1) * ldv, vnew);
    for (int i = p + 1; i <= q; ++i) {
      const T* vm = V + (long)(i - 2) * ldv;
      const T* vi = V + (long)(i - 1) * ldv;
      T newcoef;
      if constexpr (traits<T>::is_complex) {
        R sr = 0, si = 0;
#pragma omp parallel for reduction(+ : sr, si) schedule(static)
        for (long j = 0; j < n; ++j) {
          T vn0 = vnew[j] - coef * vm[j];
          T pr = std::conj(vi[j]) * vn0; sr += pr.real(); si += pr.imag();
          vnew[j] = vn0;
        }
        newcoef = T(sr, si);
      } else {
        R s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (long j = 0; j < n; ++j) {
          T vn0 = vnew[j] - coef * vm[j];
          s += vn0 * vi[j];
          vnew[j] = vn0;
        }
        newcoef = s;
      }
      coef = newcoef;
    }
    paxpy(n, -coef, V + (long)(q - 1) * ldv, vnew);
    iblck += 2;
    p = index[iblck]; q = index[iblck + 1];
  }
}

// ----------------------------------------------------------------------------
// dreorth (dreorth.F:5-101): iterated CGS (iflag=1) / MGS (iflag=0), NTRY=5,
// DGKS test ||v'|| > alpha*||v||, else zero the vector.
// ----------------------------------------------------------------------------
template <class T>
void reorth(long n, int k, const T* V, long ldv, T* vnew, typename traits<T>::real& normvnew, const int* index,
            typename traits<T>::real alpha, T* work, int iflag) {
  using R = typename traits<T>::real;
  const int NTRY = 5;
  if (k <= 0 || n <= 0) return;
  for (int itry = 1; itry <= NTRY; ++itry) {
    R normvnew_0 = normvnew;
    if (iflag == 1) cgs(n, k, V, ldv, vnew, index, work);
    else mgs(n, k, V, ldv, vnew, index);
    stats().ndot += k;
    normvnew = pnrm2(n, vnew);
    if (normvnew > alpha * normvnew_0) { stats().nreorth += 1; return; }
  }
  normvnew = 0;
  pzero(n, vnew);
  stats().nreorth += 1;
}

// ----------------------------------------------------------------------------
// dgetu0 (dgetu0.F:11-89): random vector in range(op(A)) orthogonal to U(:,1:j).
// iseed is reset to (1,3,5,7) on EVERY call (:41-44) => deterministic.
// kappa is the single-precision literal 0.717 (:28-29).
// ----------------------------------------------------------------------------
template <class T>
void getu0(char transa, int m, int n, int j, int ntry, T* u0, typename traits<T>::real& u0norm, const T* U, long ldu,
           aprod_t<T> aprod, void* ctx, int& ierr, int icgs, typename traits<T>::real& anormest, T* work) {
  using R = typename traits<T>::real;
  const R kappa = R(0.717f);
  int iseed[4] = {1, 3, 5, 7};
  long rsize, usize;
  if (transa == 'n' || transa == 'N') { rsize = n; usize = m; } else { rsize = m; usize = n; }
  ierr = 0;
  for (int itry = 1; itry <= ntry; ++itry) {
    larnv2(iseed, rsize, work);
    R nrm = pnrm2(rsize, work);
    aprod(transa, m, n, work, u0, ctx);
    stats().nopx += 1;
    u0norm = pnrm2(usize, u0);
    anormest = u0norm / nrm;
    if (j >= 1) {
      int index[3] = {1, j, j + 1};
      reorth(usize, j, U, ldu, u0, u0norm, index, kappa, work, icgs);
    }
    if (u0norm > 0) return;
  }
  ierr = -1;
}

// ----------------------------------------------------------------------------
// omega-recurrence helpers (dlanbpro.F:555-723).  Arrays are passed 1-based:
// callers hand in pointers already offset by -1 (mu[1..], index[1..]).
// ----------------------------------------------------------------------------
template <class R> void set_mu(int k, R* mu, const int* index, R val) {  // dlanbpro.F:555-577
  int i = 1;
  while (index[i] <= k && index[i] > 0) {
    int p = index[i], q = index[i + 1];
    for (int j = p; j <= q; ++j) mu[j] = val;
    i += 2;
  }
}
template <class R> void compute_int(const R* mu, int j, R delta, R eta, int* index) {  // dlanbpro.F:581-624
  if (delta < eta) { std::fprintf(stderr, "Warning delta<eta in dcompute_int\n"); return; }
  int ip = 0, i = 0, k, s;
  index[1] = 0;
  while (i < j) {
    for (k = i + 1; k <= j; ++k) if (std::fabs(mu[k]) > delta) break;
    if (k > j) break;                                   // goto 40
    const int lo = std::max(i, 1);
    for (s = k; s >= lo; --s) if (std::fabs(mu[s]) < eta) break;   // normal exit leaves s = lo-1
    ip += 1; index[ip] = s + 1;
    for (i = s + 1; i <= j; ++i) if (std::fabs(mu[i]) < eta) break; // normal exit leaves i = j+1
    ip += 1; index[ip] = i - 1;
  }
  ip += 1; index[ip] = j + 1;
}
template <class R> void update_mu(R& mumax, R* mu, const R* nu, int j, const R* alpha, const R* beta, R anorm, R eps1) {
  // dlanbpro.F:628-680
  R d;
  if (j == 1) {
    d = eps1 * (lapy2(alpha[j], beta[j]) + alpha[1]) + eps1 * anorm;
    mu[1] = eps1 / beta[1];
    mumax = std::fabs(mu[1]);
  } else {
    mu[1] = alpha[1] * nu[1] - alpha[j] * mu[1];
    d = eps1 * (lapy2(alpha[j], beta[j]) + alpha[1]) + eps1 * anorm;
    mu[1] = (mu[1] + std::copysign(d, mu[1])) / beta[j];
    mumax = std::fabs(mu[1]);
    for (int k = 2; k <= j - 1; ++k) {
      mu[k] = alpha[k] * nu[k] + beta[k - 1] * nu[k - 1] - alpha[j] * mu[k];
      d = eps1 * (lapy2(alpha[j], beta[j]) + lapy2(alpha[k], beta[k - 1])) + eps1 * anorm;
      mu[k] = (mu[k] + std::copysign(d, mu[k])) / beta[j];
      mumax = std::max(mumax, std::fabs(mu[k]));
    }
    mu[j] = beta[j - 1] * nu[j - 1];
    d = eps1 * (lapy2(alpha[j], beta[j]) + lapy2(alpha[j], beta[j - 1])) + eps1 * anorm;
    mu[j] = (mu[j] + std::copysign(d, mu[j])) / beta[j];
    mumax = std::max(mumax, std::fabs(mu[j]));
  }
  mu[j + 1] = 1;
}
template <class R> void update_nu(R& numax, const R* mu, R* nu, int j, const R* alpha, const R* beta, R anorm, R eps1) {
  // dlanbpro.F:684-723
  if (j > 1) {
    numax = 0;
    for (int k = 1; k <= j - 1; ++k) {
      nu[k] = beta[k] * mu[k + 1] + alpha[k] * mu[k] - beta[j - 1] * nu[k];
      R d = eps1 * (lapy2(alpha[k], beta[k]) + lapy2(alpha[j], beta[j - 1])) + eps1 * anorm;
      nu[k] = (nu[k] + std::copysign(d, nu[k])) / alpha[j];
      numax = std::max(numax, std::fabs(nu[k]));
    }
    nu[j] = 1;
  }
}

// ----------------------------------------------------------------------------
// dlanbpro (dlanbpro.F:1-549; complex deltas zlanbpro.F:297,318-331,451-464).
// U(ldu,k+1), V(ldv,k) column-major.  B: alpha_[1..k] = B(:,1), beta_[1..k] = B(:,2)
// (0-based storage Ba[j-1], Bb[j-1]).  k is in/out.  doption(3) in/out.
// ----------------------------------------------------------------------------
template <class T>
void lanbpro(int m, int n, int k0, int& k, aprod_t<T> aprod, void* ctx, T* U, long ldu, T* V, long ldv,
             typename traits<T>::real* Ba, typename traits<T>::real* Bb, typename traits<T>::real& rnorm,
             typename traits<T>::real* doption, const int* ioption, int& ierr) {
  using R = typename traits<T>::real;
  const R one = 1, zero = 0, FUDGE = R(1.01), kappa = R(0.717);
  auto Ucol = [&](int j) { return U + (long)(j - 1) * ldu; };
  auto Vcol = [&](int j) { return V + (long)(j - 1) * ldv; };
  R* alpha_ = Ba - 1; R* beta_ = Bb - 1;  // 1-based views of B(:,1), B(:,2)

  const R eps = lamch_e<R>();
  const R epsn = R(std::max(m, n)) * eps;
  const R epsn2 = std::sqrt(R(std::max(m, n))) * eps;
  const R eps34 = std::pow(eps, R(0.75));
  R delta, eta, anorm, anormest = 0;
  bool force_reorth, full_reorth;
  if (doption[0] < zero) delta = std::sqrt(eps / R(k)); else delta = doption[0];
  if (doption[1] < zero) eta = eps34 / std::sqrt(R(k)); else eta = doption[1];
  full_reorth = (delta <= eta || delta == zero);
  if (doption[2] > zero) anorm = doption[2];
  else if (k0 > 0) {
    anorm = lapy2(alpha_[1], beta_[1]);
    if (anorm <= zero) { ierr = -1; doption[2] = anorm; return; }
  } else anorm = zero;
  ierr = 0;

  // work = mu(k+1) | nu(k+1) | s(max(m,n)) (+ probe space), iwork(2k+1)   (:196-201)
  std::vector<R> mu_(k + 2, zero), nu_(k + 2, zero);
  std::vector<T> swork(size_t(std::max(m, n)) + size_t(std::max(m, n)), T(0));
  std::vector<int> iwork_(2 * k + 4, 0);
  R* mu = mu_.data() - 1 + 1 - 1;  // mu[1..k+1] -> mu_[0..k]
  mu = mu_.data() - 1; R* nu = nu_.data() - 1;
  int* iwork = iwork_.data() - 1;  // iwork[1..2k+1]
  T* s_ = swork.data();

  if (rnorm == zero) {  // :186-191
    getu0<T>('n', m, n, k0, 3, Ucol(k0 + 1), rnorm, U, ldu, aprod, ctx, ierr, ioption[0], anormest, s_);
    anorm = std::max(anorm, anormest);
  }

  R alpha, beta, amax, a1, b1, numax, mumax;
  int j0;
  if (k0 == 0) {  // :206-230
    amax = zero; alpha = zero; beta = rnorm; force_reorth = false;
    R sn; int ierr2;
    if (n > m) getu0<T>('n', m, n, 0, 1, s_, sn, U, ldu, aprod, ctx, ierr2, ioption[0], anormest, s_ + m);
    else getu0<T>(traits<T>::is_complex ? 'c' : 't', m, n, 0, 1, s_, sn, V, ldv, aprod, ctx, ierr2, ioption[0], anormest, s_ + n);
    ierr = ierr2;
    anorm = std::max(anorm, FUDGE * anormest);
    j0 = 1;
    if (beta != zero) safescal(m, beta, Ucol(1));
    mu[1] = one; nu[1] = one;
  } else {  // :231-275
    force_reorth = true;
    alpha = alpha_[k0]; beta = rnorm;
    if (k0 < k && beta * delta < anorm * eps) { full_reorth = true; ierr = k0; }
    iwork[1] = 1; iwork[2] = k0; iwork[3] = k0 + 1;
    pscal(m, rnorm, Ucol(k0 + 1));
    reorth(m, k0, U, ldu, Ucol(k0 + 1), rnorm, iwork + 1, kappa, s_, ioption[0]);
    safescal(m, rnorm, Ucol(k0 + 1));
    set_mu(k0, mu, iwork, epsn2);
    set_mu(k0, nu, iwork, epsn2);
    beta = rnorm;
    beta_[k0] = beta;
    amax = zero;
    for (int j = 1; j <= k0; ++j) {
      amax = std::max(amax, std::max(alpha_[j], beta_[j]));
      if (j == 1) anorm = std::max(anorm, FUDGE * alpha);
      else if (j == 2) {
        a1 = beta_[1] / amax;
        a1 = FUDGE * amax * std::sqrt((alpha_[1] / amax) * (alpha_[1] / amax) + a1 * a1 + alpha_[2] / amax * a1);
        anorm = std::max(anorm, a1);
      } else {
        a1 = alpha_[j - 1] / amax; b1 = beta_[j - 1] / amax;
        a1 = FUDGE * amax * std::sqrt(a1 * a1 + b1 * b1 + a1 * beta_[j - 2] / amax + alpha_[j] / amax * b1);
        anorm = std::max(anorm, a1);
      }
    }
    j0 = k0 + 1;
  }
  numax = zero; mumax = zero;

  for (int j = j0; j <= k; ++j) {  // :283-546
    stats().nsteps += 1;
    // alpha_j v_j = A^H u_j - beta_j v_{j-1}   (:288-296)
    aprod(traits<T>::is_complex ? 'c' : 't', m, n, Ucol(j), Vcol(j), ctx);
    stats().nopx += 1;
    if (j == 1) {
      alpha = pnrm2(n, Vcol(j));
      anorm = std::max(anorm, FUDGE * alpha);
    } else {
      paxpy(n, -beta, Vcol(j - 1), Vcol(j));
      alpha = pnrm2(n, Vcol(j));
      // extended local reorthogonalization (:301-316; complex: zlanbpro.F:318-331)
      if (j > 1 && ioption[1] > 0 && alpha < kappa * beta) {
        R nrm = alpha;
        for (int i = 1; i <= ioption[1]; ++i) {
          T s = pdotc(n, Vcol(j - 1), Vcol(j));
          paxpy(n, -s, Vcol(j - 1), Vcol(j));
          if constexpr (!traits<T>::is_complex) {
            if (beta != zero) { beta = beta + s; beta_[j - 1] = beta; }
          }
          nrm = pnrm2(n, Vcol(j));
          if (nrm >= kappa * alpha) break;
          alpha = nrm;
        }
        nu[j - 1] = eps;
        alpha = nrm;
      }
      alpha_[j] = alpha;
      amax = std::max(amax, alpha);
      if (j == 2) {  // :321-334
        a1 = beta_[1] / amax;
        a1 = FUDGE * amax * std::sqrt((alpha_[1] / amax) * (alpha_[1] / amax) + a1 * a1 + alpha_[2] / amax * a1);
      } else {
        a1 = alpha_[j - 1] / amax; b1 = beta_[j - 1] / amax;
        a1 = FUDGE * amax * std::sqrt(a1 * a1 + b1 * b1 + a1 * beta_[j - 2] / amax + alpha_[j] / amax * b1);
      }
      anorm = std::max(anorm, a1);
    }
    // nu recurrence (:340-343)
    if (!full_reorth && alpha != zero) update_nu(numax, mu, nu, j, alpha_, beta_, anorm, epsn2);
    // reorthogonalize v_j (:348-367)
    if ((full_reorth || numax > delta || force_reorth) && alpha != zero) {
      if (full_reorth || eta == zero) { iwork[1] = 1; iwork[2] = j - 1; iwork[3] = j; }
      else if (!force_reorth) compute_int(nu, j - 1, delta, eta, iwork);
      reorth(n, j - 1, V, ldv, Vcol(j), alpha, iwork + 1, kappa, s_, ioption[0]);
      set_mu(j - 1, nu, iwork, eps);
      numax = eta;
      force_reorth = !force_reorth;
    }
    // invariant subspace check (:372-408)
    if (alpha < anorm * epsn && j < k) {
      rnorm = alpha; alpha = zero;
      getu0<T>(traits<T>::is_complex ? 'c' : 't', m, n, j - 1, 3, Vcol(j), alpha, V, ldv, aprod, ctx, ierr, ioption[0], anormest, s_);
      if (alpha == zero) { k = j - 1; ierr = -j; doption[2] = anorm; return; }
      safescal(n, alpha, Vcol(j));
      alpha = zero; force_reorth = true;
      if (delta > zero) full_reorth = false;
    } else if (j > 1 && !full_reorth && j < k && (delta * alpha < anorm * eps)) {
      ierr = j;
    }
    alpha_[j] = alpha;
    if (alpha != zero) safescal(n, alpha, Vcol(j));

    // beta_{j+1} u_{j+1} = A v_j - alpha_j u_j   (:420-424)
    aprod('n', m, n, Vcol(j), Ucol(j + 1), ctx);
    stats().nopx += 1;
    paxpy(m, -alpha, Ucol(j), Ucol(j + 1));
    beta = pnrm2(m, Ucol(j + 1));
    // extended local reorthogonalization (:429-443; complex: zlanbpro.F:451-464)
    if (ioption[1] > 0 && beta < kappa * alpha) {
      R nrm = beta;
      for (int i = 1; i <= ioption[1]; ++i) {
        T s = pdotc(m, Ucol(j), Ucol(j + 1));
        paxpy(m, -s, Ucol(j), Ucol(j + 1));
        if constexpr (!traits<T>::is_complex) {
          if (alpha != zero) { alpha = alpha + s; alpha_[j] = alpha; }
        }
        nrm = pnrm2(m, Ucol(j + 1));
        if (nrm >= kappa * beta) break;
        beta = nrm;
      }
      mu[j] = eps;
      beta = nrm;
    }
    beta_[j] = beta;
    amax = std::max(amax, beta);
    // ||A|| estimate (:451-458)
    if (j <= 1) a1 = lapy2(alpha_[1], beta_[1]);
    else {
      a1 = alpha_[j] / amax;
      a1 = amax * std::sqrt(a1 * a1 + (beta_[j] / amax) * (beta_[j] / amax) + a1 * beta_[j - 1] / amax);
    }
    anorm = std::max(anorm, a1);
    // mu recurrence (:463-466)
    if (!full_reorth && beta != zero) update_mu(mumax, mu, nu, j, alpha_, beta_, anorm, epsn2);
    // reorthogonalize u_{j+1} (:471-498)
    if ((full_reorth || mumax > delta || force_reorth) && beta != zero) {
      if (full_reorth || eta == zero) { iwork[1] = 1; iwork[2] = j; iwork[3] = j + 1; }
      else if (!force_reorth) compute_int(mu, j, delta, eta, iwork);
      else {
        for (int i = 1; i <= 2 * j + 1; ++i) if (iwork[i] == j) { iwork[i] = j + 1; break; }
      }
      reorth(m, j, U, ldu, Ucol(j + 1), beta, iwork + 1, kappa, s_, ioption[0]);
      set_mu(j, mu, iwork, eps);
      mumax = eta;
      force_reorth = !force_reorth;
    }
    // invariant subspace check (:503-539)
    if (beta < anorm * epsn && j < k) {
      rnorm = beta; beta = zero;
      getu0<T>('n', m, n, j, 3, Ucol(j + 1), beta, U, ldu, aprod, ctx, ierr, ioption[0], anormest, s_);
      if (beta == zero) { k = j; ierr = -j; doption[2] = anorm; return; }
      safescal(m, beta, Ucol(j + 1));
      beta = zero; force_reorth = true;
      if (delta > zero) full_reorth = false;
    } else if (!full_reorth && j < k && (delta * beta < anorm * eps)) {
      ierr = j;
    }
    beta_[j] = beta;
    if (beta != zero && beta != one) safescal(m, beta, Ucol(j + 1));
    rnorm = beta;
  }
  doption[2] = anorm;  // :547
}

// ----------------------------------------------------------------------------
// dbsvd.F: dbsvdstep :5-82, dbdqr :87-157, drefinebounds :162-231 (real, host)
// ----------------------------------------------------------------------------
template <class R> inline void rot(int n, R* x, R* y, R c, R s) {  // BLAS drot
  for (int i = 0; i < n; ++i) { R t = c * x[i] + s * y[i]; y[i] = c * y[i] - s * x[i]; x[i] = t; }
}
template <class R>
void bsvdstep(bool dou, bool dov, int m, int n, int k, R sigma, R* D_, R* E_, R* U, int ldu, R* V, int ldv) {
  if (k <= 1) return;
  R* D = D_ - 1; R* E = E_ - 1;
  auto Uc = [&](int j) { return U + (size_t)(j - 1) * ldu; };
  auto Vc = [&](int j) { return V + (size_t)(j - 1) * ldv; };
  R c, s, r, x, y;
  x = D[1] * D[1] - sigma * sigma;
  y = E[1] * D[1];
  for (int i = 1; i <= k - 1; ++i) {
    if (i > 1) lartg(x, y, c, s, E[i - 1]); else lartg(x, y, c, s, r);
    x = c * D[i] + s * E[i];
    E[i] = -s * D[i] + c * E[i];
    D[i] = x;
    y = s * D[i + 1];
    D[i + 1] = c * D[i + 1];
    if (dou && m > 0) rot(m, Uc(i), Uc(i + 1), c, s);
    lartg(x, y, c, s, D[i]);
    x = c * E[i] + s * D[i + 1];
    D[i + 1] = -s * E[i] + c * D[i + 1];
    E[i] = x;
    y = s * E[i + 1];
    E[i + 1] = c * E[i + 1];
    if (dov && n > 0) rot(n, Vc(i), Vc(i + 1), c, s);
  }
  lartg(x, y, c, s, E[k - 1]);
  x = c * D[k] + s * E[k];
  E[k] = -s * D[k] + c * E[k];
  D[k] = x;
  if (dou && m > 0) rot(m, Uc(k), Uc(k + 1), c, s);
}
template <class R> void bdqr(bool ignorelast, bool jobq, int n, R* D_, R* E_, R& c1, R& c2, R* Qt, int ldq) {
  if (n < 1) return;
  R* d = D_ - 1; R* e = E_ - 1;
  auto Q = [&](int i, int j) -> R& { return Qt[(size_t)(j - 1) * ldq + (i - 1)]; };
  if (jobq) {
    for (int j = 1; j <= n + 1; ++j) { for (int i = 1; i <= n + 1; ++i) Q(i, j) = 0; Q(j, j) = 1; }
  }
  R cs, sn, r;
  int i;
  for (i = 1; i <= n - 1; ++i) {
    lartg(d[i], e[i], cs, sn, r);
    d[i] = r; e[i] = sn * d[i + 1]; d[i + 1] = cs * d[i + 1];
    if (jobq) {
      for (int j = 1; j <= i; ++j) { Q(i + 1, j) = -sn * Q(i, j); Q(i, j) = cs * Q(i, j); }
      Q(i, i + 1) = sn; Q(i + 1, i + 1) = cs;
    }
  }
  // after the DO loop the Fortran variable i == n (dbsvd.F:128-141; for n==1 the loop is skipped and i==1==n)
  if (!ignorelast) {
    lartg(d[n], e[n], cs, sn, r);
    d[n] = r; e[n] = 0; c1 = sn; c2 = cs;
    if (jobq) {
      for (int j = 1; j <= i; ++j) { Q(i + 1, j) = -sn * Q(i, j); Q(i, j) = cs * Q(i, j); }
      Q(i, i + 1) = sn; Q(i + 1, i + 1) = cs;
    }
  }
}
template <class R> void refinebounds(int n, int k, const R* theta_, R* bound_, R tol, R eps34) {
  if (k <= 1) return;
  const R* theta = theta_ - 1; R* bound = bound_ - 1;
  for (int i = 1; i <= k; ++i)
    for (int l = -1; l <= 1; l += 2)
      if ((l == 1 && i < k) || (l == -1 && i > 1))
        if (std::fabs(theta[i] - theta[i + l]) < eps34 * theta[i])
          if (bound[i] > tol && bound[i + l] > tol) { bound[i + l] = lapy2(bound[i], bound[i + l]); bound[i] = 0; }
  for (int i = 1; i <= k; ++i) {
    if (i < k || k == n) {
      R gap;
      if (i == 1) gap = std::fabs(theta[i] - theta[i + 1]) - std::max(bound[i], bound[i + 1]);
      else if (i == n) gap = std::fabs(theta[i - 1] - theta[i]) - std::max(bound[i - 1], bound[i]);
      else {
        gap = std::fabs(theta[i] - theta[i + 1]) - std::max(bound[i], bound[i + 1]);
        gap = std::min(gap, std::fabs(theta[i - 1] - theta[i]) - std::max(bound[i - 1], bound[i]));
      }
      if (gap > bound[i]) bound[i] = bound[i] * (bound[i] / gap);
    }
  }
}

// ----------------------------------------------------------------------------
// dgemm_ovwr_left (dgemm_ovwr.F:56-87) / zdgemm_ovwr_left (zgemm_ovwr.F:6-61):
// in place A(m x k) <- A * op(B), result m x n, B small and REAL, row-blocked
// through a scratch buffer; OpenMP over row chunks as in dritzvec.F:145-163.
// transb='n': op(B)=B (k x n, ldb>=k); transb='t': op(B)=B^T (B is n x k, ldb>=n).
// NOTE: the reference's complex zdgemm ignores transb and always
// forms A*B^T; the oracle implements the mathematically intended product so the
// complex IRL restart is correct -- a documented deviation from a reference bug.
// ----------------------------------------------------------------------------
template <class T>
void gemm_ovwr_left(char transb, long m, int n, int k, T* A, long lda, const typename traits<T>::real* B, int ldb) {
  using R = typename traits<T>::real;
  if (m <= 0 || n <= 0 || k <= 0) return;
  const bool tr = (transb == 't' || transb == 'T');
  const long RB = 256;
#pragma omp parallel
  {
    std::vector<T> buf(size_t(RB) * n);
#pragma omp for schedule(static)
    for (long i0 = 0; i0 < m; i0 += RB) {
      const long rb = std::min(RB, m - i0);
      std::fill(buf.begin(), buf.begin() + size_t(rb) * n, T(0));
      for (int jn = 0; jn < n; ++jn) {
        T* c = buf.data() + size_t(jn) * rb;
        for (int l = 0; l < k; ++l) {
          const R b = tr ? B[(size_t)l * ldb + jn] : B[(size_t)jn * ldb + l];
          const T* a = A + (size_t)l * lda + i0;
          for (long r = 0; r < rb; ++r) c[r] += a[r] * b;
        }
      }
      for (int jn = 0; jn < n; ++jn) std::memcpy(A + (size_t)jn * lda + i0, buf.data() + size_t(jn) * rb, sizeof(T) * rb);
    }
  }
}

// ----------------------------------------------------------------------------
// dritzvec (dritzvec.F:1-199; zritzvec.F)
// D,E: copies of alpha,beta (overwritten); S receives sigma (aliases D in callers).
// ----------------------------------------------------------------------------
template <class T>
void ritzvec(char which, bool jobu, bool jobv, int m, int n, int k, int dim, typename traits<T>::real* D,
             typename traits<T>::real* E, T* U, long ldu, T* V, long ldv) {
  using R = typename traits<T>::real;
  std::vector<R> Mt((size_t)(dim + 1) * (dim + 1), R(0)), Qt((size_t)dim * dim, R(0)), P((size_t)dim * dim, R(0));
  std::vector<R> wrk((size_t)3 * dim * dim + 4 * dim + 16);
  std::vector<int> iwork(8 * dim + 8);
  R c1 = 0, c2 = 0, dd[1]; int id[1], info = 0;
  bdqr(dim == std::min(m, n), jobu, dim, D, E, c1, c2, Mt.data(), dim + 1);          // :116
  bdsdc("U", "I", dim, D, E, P.data(), dim, Qt.data(), dim, dd, id, wrk.data(), iwork.data(), &info);  // :123
  // X = P^T * M^T(1:dim,:)  (dgemm_ovwr('t',dim,dim+1,dim,..) :130) -- small, host
  std::vector<R> X((size_t)dim * (dim + 1), R(0));  // X(i,j), ld = dim
  if (jobu) {
    for (int j = 0; j < dim + 1; ++j)
      for (int i = 0; i < dim; ++i) {
        R s = 0;
        for (int l = 0; l < dim; ++l) s += P[(size_t)i * dim + l] * Mt[(size_t)j * (dim + 1) + l];
        X[(size_t)j * dim + i] = s;
      }
  }
  const int mstart = (which == 's' || which == 'S') ? dim - k + 1 : 1;
  if (jobu)  // U(:,1:k) = U(:,1:dim+1) * X(mstart:mstart+k-1,:)^T  (:139-161)
    gemm_ovwr_left<T>('t', m, k, dim + 1, U, ldu, X.data() + (mstart - 1), dim);
  if (jobv)  // V(:,1:k) = V(:,1:dim) * Qt(mstart:mstart+k-1,:)^T  (:167-194)
    gemm_ovwr_left<T>('t', n, k, dim, V, ldv, Qt.data() + (mstart - 1), dim);
}

// ----------------------------------------------------------------------------
// dlansvd (dlansvd.F:1-291)
// ----------------------------------------------------------------------------
template <class T>
void lansvd(bool jobu, bool jobv, int m, int n, int& k, int kmax, aprod_t<T> aprod, void* ctx, T* U, long ldu,
            typename traits<T>::real* sigma, typename traits<T>::real* bnd, T* V, long ldv,
            typename traits<T>::real tolin, typename traits<T>::real* doption, const int* ioption, int& info) {
  using R = typename traits<T>::real;
  const R one = 1, zero = 0;
  const R eps = lamch_e<R>();
  const R eps34 = std::pow(eps, R(0.75));
  const R epsn = R(std::max(m, n)) * eps / R(2);
  const int lanmax = std::min(std::min(n + 1, m + 1), kmax);
  const R tol = std::min(one, std::max(R(16) * eps, tolin));
  R anorm = zero, rnorm;
  std::vector<R> wbnd(lanmax + 2, zero), B(2 * (size_t)lanmax, zero), B1(2 * (size_t)lanmax, zero), wrk(4 * (size_t)lanmax + 16);
  std::vector<T> swork(size_t(m) + n);
  int ierr = 0, lapinfo = 0;

  rnorm = pnrm2(m, U);
  if (rnorm == zero)
    getu0<T>('n', m, n, 0, 1, U, rnorm, U, ldu, aprod, ctx, ierr, ioption[0], anorm, swork.data());  // :165-169
  stats().nsing = k;
  info = 0;
  int neig = 0, jold = 0;
  int j = std::min(k + std::max(8, k) + 1, lanmax);
  R dummy[1];
  while (neig < k) {
    lanbpro<T>(m, n, jold, j, aprod, ctx, U, ldu, V, ldv, B.data(), B.data() + lanmax, rnorm, doption, ioption, ierr);  // :185
    jold = j;
    B1 = B;
    std::fill(wbnd.begin(), wbnd.begin() + j + 1, zero);
    R* b1a = B1.data(); R* b1b = B1.data() + lanmax;
    bdqr(j == std::min(m, n), false, j, b1a, b1b, wbnd[j - 1], wbnd[j], (R*)nullptr, lanmax + 1);  // :196
    bdsqr("u", j, 0, 1, 0, b1a, b1b, dummy, 1, wbnd.data(), 1, dummy, 1, wrk.data(), &lapinfo);       // :198
    stats().nbsvd += 1;
    if (j > 5) anorm = b1a[0]; else anorm = std::max(anorm, b1a[0]);
    for (int i = 0; i < j; ++i) wbnd[i] = std::fabs(rnorm * wbnd[i]);
    refinebounds(std::min(m, n), j, b1a, wbnd.data(), epsn * anorm, eps34);  // :215
    for (int i = 0; i < std::min(j, k); ++i) bnd[i] = wbnd[i];
    int i = 0; neig = 0;
    while (i < std::min(j, k)) {
      if (wbnd[i] <= tol * b1a[i]) { sigma[neig] = b1a[i]; neig += 1; i += 1; }
      else i = k;
    }
    if (ierr < 0) { if (j < k) info = j; break; }      // :242-249
    if (j >= lanmax) { if (neig < k) info = -1; break; }  // :250-259
    int dj;
    if (neig > 1) { dj = std::min(j / 2, ((k - neig) * (j - 6)) / (2 * neig + 1)); dj = std::min(100, std::max(2, dj)); }
    else { dj = j / 2; dj = std::min(100, std::max(10, dj)); }
    j = std::min(j + dj, lanmax);
  }
  if ((neig >= k || info > 0) && (jobu || jobv)) {  // :278-288
    std::vector<R> D(B.begin(), B.begin() + lanmax), E(B.begin() + lanmax, B.end());
    ritzvec<T>('L', jobu, jobv, m, n, neig, jold, D.data(), E.data(), U, ldu, V, ldv);
  }
  k = neig;
  stats().nlandim = j;
}

// ----------------------------------------------------------------------------
// dlansvd_irl (dlansvd_irl.F:1-419)
// ----------------------------------------------------------------------------
template <class T>
void lansvd_irl(char which, bool jobu, bool jobv, int m, int n, int& dim, int p, int& neig, int maxiter,
                aprod_t<T> aprod, void* ctx, T* U, long ldu, typename traits<T>::real* sigma,
                typename traits<T>::real* bnd, T* V, long ldv, typename traits<T>::real tolin,
                typename traits<T>::real* doption, const int* ioption, int& info) {
  using R = typename traits<T>::real;
  const R one = 1, zero = 0;
  const bool smallest = (which == 's' || which == 'S');
  const R eps = lamch_e<R>();
  const R eps34 = std::pow(eps, R(0.75));
  const R epsn = R(std::max(m, n)) * eps / R(2);
  dim = std::min(dim, std::min(n + 1, m + 1));
  int k = dim - p;
  const R tol = std::min(one, std::max(R(16) * eps, tolin));
  R anorm = zero, rnorm;
  std::vector<R> wbnd(dim + 2, zero), al(dim, zero), be(dim, zero), al1(dim, zero), be1(dim, zero), shift(dim, zero);
  std::vector<R> P((size_t)(dim + 1) * (dim + 1)), Q((size_t)dim * dim), wrk(4 * (size_t)dim + 16);
  std::vector<T> swork(size_t(m) + n);
  int ierr = 0, lapinfo = 0;
  R dummy[1];

  rnorm = pnrm2(m, U);
  if (rnorm == zero)
    getu0<T>('n', m, n, 0, 1, U, rnorm, U, ldu, aprod, ctx, ierr, ioption[0], anorm, swork.data());
  int iter = 0, nconv = 0, kold = 0;
  info = 0;
  while (nconv < neig && iter < maxiter) {
    int dimio = dim;
    lanbpro<T>(m, n, kold, dimio, aprod, ctx, U, ldu, V, ldv, al.data(), be.data(), rnorm, doption, ioption, ierr);  // :213
    dim = dimio;  // dlanbpro may shrink K (Fortran passes dim by reference)
    kold = k;
    al1 = al; be1 = be;
    std::fill(wbnd.begin(), wbnd.begin() + dim + 1, zero);
    bdqr(dim == std::min(m, n), false, dim, al1.data(), be1.data(), wbnd[dim - 1], wbnd[dim], (R*)nullptr, dim + 1);  // :224
    bdsqr("u", dim, 0, 1, 0, al1.data(), be1.data(), dummy, 1, wbnd.data(), 1, dummy, 1, wrk.data(), &lapinfo);     // :228
    stats().nbsvd += 1;
    if (dim > 5) anorm = al1[0]; else anorm = std::max(anorm, al1[0]);
    for (int i = 0; i < dim; ++i) wbnd[i] = std::fabs(rnorm * wbnd[i]);
    if (smallest) refinebounds(std::min(m, n), dim, al1.data(), wbnd.data(), epsn * anorm, eps34);
    else refinebounds(std::min(m, n), std::min(dim, neig), al1.data(), wbnd.data(), epsn * anorm, eps34);
    // count converged (:262-290), 1-based i
    if (smallest) {
      int i = dim - neig + 1; nconv = 0;
      while (i <= dim) {
        if (wbnd[i - 1] <= tol * al1[0]) { sigma[nconv] = al1[i - 1]; bnd[nconv] = wbnd[i - 1]; nconv += 1; }
        i += 1;
      }
    } else {
      int i = 1; nconv = 0;
      while (i <= std::min(dim, neig)) {
        if (wbnd[i - 1] <= tol * al1[i - 1]) { sigma[nconv] = al1[i - 1]; bnd[nconv] = wbnd[i - 1]; nconv += 1; i += 1; }
        else i = k + 1;
      }
    }
    if (ierr < 0) { if (dim < k) info = dim; break; }  // :295-302
    if (nconv < neig) {
      // shifts (:318-344)
      std::fill(shift.begin(), shift.begin() + (dim - k), zero);
      int nshft = 0;
      if (smallest) {
        for (int i = 1; i <= k; ++i) {
          R relgap = (al1[i - 1] - wbnd[i - 1] - al1[dim - neig - 1]);
          if (relgap > doption[3] * al1[dim - neig - 1]) shift[nshft] = al1[i - 1]; else shift[nshft] = al1[0];
          nshft += 1;
        }
      } else {
        for (int i = dim; i >= k + 1; --i) {
          R relgap = al1[k - 1] - (al1[i - 1] + wbnd[i - 1]);
          if (relgap > doption[3] * al1[k - 1]) shift[nshft] = al1[i - 1]; else shift[nshft] = zero;
          nshft += 1;
        }
      }
      // accumulate rotations (:350-363)
      std::fill(P.begin(), P.end(), zero); std::fill(Q.begin(), Q.end(), zero);
      for (int i = 0; i < dim + 1; ++i) P[(size_t)i * (dim + 2)] = one;
      for (int i = 0; i < dim; ++i) Q[(size_t)i * (dim + 1)] = one;
      for (int i = dim; i >= k + 1; --i) {
        R sh = shift[dim - i];
        bsvdstep(true, true, dim + 1, dim, i, sh, al.data(), be.data(), P.data(), dim + 1, Q.data(), dim);
      }
      // U(:,1:k+1) = U(:,1:dim+1) P(:,1:k+1);  V(:,1:k) = V(:,1:dim) Q(:,1:k)   (:387-395)
      gemm_ovwr_left<T>('n', m, k + 1, dim + 1, U, ldu, P.data(), dim + 1);
      gemm_ovwr_left<T>('n', n, k, dim, V, ldv, Q.data(), dim);
      rnorm = be[k - 1];
      stats().nrestart += 1;
    }
    iter += 1;
  }
  if ((nconv >= neig || info > 0) && (jobu || jobv)) {  // :405-416
    al1 = al; be1 = be;
    ritzvec<T>(which, jobu, jobv, m, n, nconv, dim, al1.data(), be1.data(), U, ldu, V, ldv);
  }
  neig = nconv;
  stats().nlandim = dim;
}

}  // namespace oracle
