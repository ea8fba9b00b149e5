#include "host_lapack.hpp"

#include <dlfcn.h>
#include <glob.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace pb {
namespace host {

namespace {
extern "C" {
typedef void (*dbdsqr_t)(const char*, const int*, const int*, const int*, const int*, double*, double*, double*,
                         const int*, double*, const int*, double*, const int*, double*, int*, size_t);
typedef void (*sbdsqr_t)(const char*, const int*, const int*, const int*, const int*, float*, float*, float*,
                         const int*, float*, const int*, float*, const int*, float*, int*, size_t);
typedef void (*dbdsdc_t)(const char*, const char*, const int*, double*, double*, double*, const int*, double*,
                         const int*, double*, int*, double*, int*, int*, size_t, size_t);
typedef void (*sbdsdc_t)(const char*, const char*, const int*, float*, float*, float*, const int*, float*,
                         const int*, float*, int*, float*, int*, int*, size_t, size_t);
typedef void (*dlasq1_t)(const int*, double*, double*, double*, int*);
typedef void (*dstein_t)(const int*, const double*, const double*, const int*, const double*, const int*, const int*, double*,
                         const int*, double*, int*, int*, int*);
}
dlasq1_t p_dlasq1 = nullptr;   // optional (ritz_leading)
dstein_t p_dstein = nullptr;
dbdsqr_t p_dbdsqr = nullptr;
sbdsqr_t p_sbdsqr = nullptr;
dbdsdc_t p_dbdsdc = nullptr;
sbdsdc_t p_sbdsdc = nullptr;

bool try_bind(const std::string& path) {
  void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) return false;
  auto sym = [&](const char* base) -> void* {
    for (const char* pre : {"", "scipy_"}) {
      std::string s = std::string(pre) + base + "_";
      if (void* p = dlsym(h, s.c_str())) return p;
    }
    return nullptr;
  };
  dbdsqr_t a = (dbdsqr_t)sym("dbdsqr");
  sbdsqr_t b = (sbdsqr_t)sym("sbdsqr");
  dbdsdc_t c = (dbdsdc_t)sym("dbdsdc");
  sbdsdc_t d = (sbdsdc_t)sym("sbdsdc");
  if (!(a && b && c && d)) { dlclose(h); return false; }
  p_dbdsqr = a; p_sbdsqr = b; p_dbdsdc = c; p_sbdsdc = d;
  p_dlasq1 = (dlasq1_t)sym("dlasq1");
  p_dstein = (dstein_t)sym("dstein");
  return true;
}
}  // namespace

bool lapack_bound() { return p_dbdsqr != nullptr; }

void bind_lapack(const char* path) {
  if (lapack_bound() && !path) return;
  std::vector<std::string> cands;
  if (path && *path) cands.push_back(path);
  if (const char* env = std::getenv("PROPACK_B200_LAPACK")) cands.push_back(env);
  cands.push_back("liblapack.so.3");
  cands.push_back("libopenblas.so.0");
  for (const char* pat : {"/opt/*/.venv/lib/python3*/site-packages/scipy.libs/libscipy_openblas*.so",
                          "/usr/lib/python3*/site-packages/scipy.libs/libscipy_openblas*.so",
                          "/usr/local/lib/python3*/site-packages/scipy.libs/libscipy_openblas*.so"}) {
    glob_t g;
    if (glob(pat, 0, nullptr, &g) == 0) {
      for (size_t i = 0; i < g.gl_pathc; ++i) cands.push_back(g.gl_pathv[i]);
    }
    globfree(&g);
  }
  for (const auto& c : cands)
    if (try_bind(c)) return;
  throw std::runtime_error(
      "propack_b200: no LAPACK with {d,s}bdsqr/{d,s}bdsdc found; set PROPACK_B200_LAPACK to a LAPACK shared object");
}

void bdsqr_row(int n, double* d, double* e, double* urow, int* info) {
  bind_lapack();
  const int zero = 0, one = 1;
  std::vector<double> work(4 * (size_t)n + 8);
  double dum[1] = {0};
  p_dbdsqr("U", &n, &zero, &one, &zero, d, e, dum, &one, urow, &one, dum, &one, work.data(), info, 1);
}
void bdsqr_row(int n, float* d, float* e, float* urow, int* info) {
  bind_lapack();
  const int zero = 0, one = 1;
  std::vector<float> work(4 * (size_t)n + 8);
  float dum[1] = {0};
  p_sbdsqr("U", &n, &zero, &one, &zero, d, e, dum, &one, urow, &one, dum, &one, work.data(), info, 1);
}
void bdsdc_full(int n, double* d, double* e, double* U, int ldu, double* VT, int ldvt, int* info) {
  bind_lapack();
  std::vector<double> work(3 * (size_t)n * n + 4 * (size_t)n + 16);
  std::vector<int> iwork(8 * (size_t)n + 8);
  double q[1]; int iq[1];
  p_dbdsdc("U", "I", &n, d, e, U, &ldu, VT, &ldvt, q, iq, work.data(), iwork.data(), info, 1, 1);
}
void bdsdc_full(int n, float* d, float* e, float* U, int ldu, float* VT, int ldvt, int* info) {
  bind_lapack();
  std::vector<float> work(3 * (size_t)n * n + 4 * (size_t)n + 16);
  std::vector<int> iwork(8 * (size_t)n + 8);
  float q[1]; int iq[1];
  p_sbdsdc("U", "I", &n, d, e, U, &ldu, VT, &ldvt, q, iq, work.data(), iwork.data(), info, 1, 1);
}

namespace {
// sigma (all j, descending) by dqds and the Golub-Kahan eigenvectors Z (N x K, column c <-> sigma_{K-1-c}) of the K leading ones
bool leading_triplets(int j, const double* alpha, const double* beta, int K, std::vector<double>& sig, std::vector<double>& Z) {
  bind_lapack();
  if (!p_dlasq1 || !p_dstein || j < 2 || K < 1 || K > j) return false;
  double amax = 0;
  for (int i = 0; i < j; ++i) amax = std::max(amax, std::max(std::fabs(alpha[i]), std::fabs(beta[i])));
  for (int i = 0; i < j; ++i)
    if (!(std::fabs(alpha[i]) > 1e-13 * amax) || !(std::fabs(beta[i]) > 1e-13 * amax)) return false;  // (nearly) reducible
  // 1. singular values: QR-reduce the (j+1) x j lower bidiagonal to j x j upper (dbdqr's rotations, dbsvd.F:128-147), dqds
  std::vector<double> e(beta, beta + j), work(4 * (size_t)j + 8);
  sig.assign(alpha, alpha + j);
  std::vector<double>& d = sig;
  for (int i = 0; i < j - 1; ++i) {
    const double r = std::hypot(d[i], e[i]), cs = d[i] / r, sn = e[i] / r;
    d[i] = r; e[i] = sn * d[i + 1]; d[i + 1] = cs * d[i + 1];
  }
  d[j - 1] = std::hypot(d[j - 1], e[j - 1]); e[j - 1] = 0;
  int info = 0;
  p_dlasq1(&j, d.data(), e.data(), work.data(), &info);
  if (info != 0) return false;
  for (int i = 0; i + 1 < K; ++i)
    if (!(d[i] > d[i + 1])) return false;               // repeated leading values: leave it to the reference route
  // 2. eigenvectors of the Golub-Kahan tridiagonal (order u_1 v_1 u_2 v_2 ... u_j v_j u_{j+1}; zero diagonal,
  //    off-diagonals alpha_1 beta_1 alpha_2 beta_2 ... alpha_j beta_j) for the eigenvalues +sigma_K <= ... <= +sigma_1
  const int N = 2 * j + 1;
  std::vector<double> dz((size_t)N, 0.0), off((size_t)N), w((size_t)K), wk(5 * (size_t)N);
  std::vector<int> iblock((size_t)N, 1), isplit((size_t)N, 0), iwk((size_t)N), ifail((size_t)K, 0);
  Z.assign((size_t)N * K, 0.0);
  for (int i = 0; i < j; ++i) { off[2 * i] = alpha[i]; off[2 * i + 1] = beta[i]; }
  for (int i = 0; i < K; ++i) w[i] = d[K - 1 - i];
  isplit[0] = N;
  p_dstein(&N, dz.data(), off.data(), &K, w.data(), iblock.data(), isplit.data(), Z.data(), &N, wk.data(), iwk.data(), ifail.data(), &info);
  return info == 0;
}
}  // namespace

bool ritz_leading(int j, const double* alpha, const double* beta, int K, double* theta, double* last) {
  std::vector<double> sig, Z;
  if (!leading_triplets(j, alpha, beta, K, sig, Z)) return false;
  const int N = 2 * j + 1;
  for (int i = 0; i < K; ++i) {
    const double* z = Z.data() + (size_t)(K - 1 - i) * N;   // column of sigma_i (descending order out)
    double un = 0;
    for (int t = 0; t < N; t += 2) un += z[t] * z[t];
    if (!(un > 0.25 && un < 0.75)) return false;             // the u-part of a unit eigenvector has norm^2 1/2
    theta[i] = sig[i];
    last[i] = std::fabs(z[N - 1]) / std::sqrt(un);
  }
  return true;
}
bool ritz_vectors_leading(int dim, const double* alpha, const double* beta, int k, std::vector<double>& WU, std::vector<double>& WV) {
  std::vector<double> sig, Z;
  if (!leading_triplets(dim, alpha, beta, k, sig, Z)) return false;
  const int N = 2 * dim + 1;
  WU.assign((size_t)(dim + 1) * k, 0.0);
  WV.assign((size_t)dim * k, 0.0);
  for (int i = 0; i < k; ++i) {
    const double* z = Z.data() + (size_t)(k - 1 - i) * N;
    double un = 0, vn = 0;
    for (int t = 0; t <= dim; ++t) un += z[2 * t] * z[2 * t];
    for (int t = 0; t < dim; ++t) vn += z[2 * t + 1] * z[2 * t + 1];
    if (!(un > 0.25 && un < 0.75 && vn > 0.25 && vn < 0.75)) return false;
    const double su = 1.0 / std::sqrt(un), sv = 1.0 / std::sqrt(vn);
    for (int t = 0; t <= dim; ++t) WU[(size_t)i * (dim + 1) + t] = su * z[2 * t];
    for (int t = 0; t < dim; ++t) WV[(size_t)i * dim + t] = sv * z[2 * t + 1];
  }
  return true;
}
bool ritz_vectors_leading(int, const float*, const float*, int, std::vector<float>&, std::vector<float>&) { return false; }
bool ritz_leading(int, const float*, const float*, int, float*, float*) { return false; }

}  // namespace host
}  // namespace pb
