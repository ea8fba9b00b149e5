#include "host_lapack.hpp"

#include <dlfcn.h>
#include <glob.h>

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
}
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

}  // namespace host
}  // namespace pb
