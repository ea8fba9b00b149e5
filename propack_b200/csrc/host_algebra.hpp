// propack_b200 -- the O(k) / O(k^2) host-side algebra of the Lanczos drivers.
//
// The north star keeps the small bidiagonal problem on the host "exactly as in the reference
// algorithm": Larsen's omega-recurrences and interval selection (double/dlanbpro.F:555-723), the
// bidiagonal QR / implicit-shift sweeps / bound refinement of double/dbsvd.F, and LAPACK
// xBDSQR / xBDSDC (host_lapack.hpp).  Written 0-based over std::vector; the decision logic must
// reproduce the reference's choices, so every formula cites the line it restates.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <exception>
#include <thread>
#include <vector>

namespace pb {
namespace host {

template <class R> struct Machine;  // dlamch('e') = eps with rounding, dlamch('s') = safe minimum
template <> struct Machine<double> {
  static constexpr double eps = 1.1102230246251565e-16;
  static constexpr double sfmin = 2.2250738585072014e-308;
};
template <> struct Machine<float> {
  static constexpr float eps = 5.9604644775390625e-8f;
  static constexpr float sfmin = 1.17549435e-38f;
};

template <class R> inline R hypot2(R x, R y) {  // LAPACK dlapy2
  const R ax = std::fabs(x), ay = std::fabs(y);
  const R big = ax > ay ? ax : ay, small = ax > ay ? ay : ax;
  if (small == R(0)) return big;
  const R q = small / big;
  return big * std::sqrt(R(1) + q * q);
}

// Plane rotation with the LAPACK 3.0 dlartg conventions the reference's vendored copy uses
// (Lapack_Util/dlartg.f): r = +-sqrt(f^2+g^2), cs > 0 whenever |f| > |g|.
template <class R> struct Givens { R c, s, r; };
template <class R> inline Givens<R> make_givens(R f, R g) {
  Givens<R> G;
  if (g == R(0)) { G.c = 1; G.s = 0; G.r = f; return G; }
  if (f == R(0)) { G.c = 0; G.s = 1; G.r = g; return G; }
  static const R safmn2 = std::pow(R(2), R(int(std::log(Machine<R>::sfmin / Machine<R>::eps) / std::log(R(2)) / R(2))));
  static const R safmx2 = R(1) / safmn2;
  R f1 = f, g1 = g;
  R scale = std::max(std::fabs(f1), std::fabs(g1));
  int up = 0, down = 0;
  while (scale >= safmx2) { f1 *= safmn2; g1 *= safmn2; scale = std::max(std::fabs(f1), std::fabs(g1)); ++down; }
  if (!down) while (scale <= safmn2) { f1 *= safmx2; g1 *= safmx2; scale = std::max(std::fabs(f1), std::fabs(g1)); ++up; }
  G.r = std::sqrt(f1 * f1 + g1 * g1);
  G.c = f1 / G.r; G.s = g1 / G.r;
  for (; down > 0; --down) G.r *= safmx2;
  for (; up > 0; --up) G.r *= safmn2;
  if (std::fabs(f) > std::fabs(g) && G.c < R(0)) { G.c = -G.c; G.s = -G.s; G.r = -G.r; }
  return G;
}

// ------------------------------------------------------------------------------------------------
// Reorthogonalisation interval list: 1-based inclusive [s1,e1,s2,e2,...,T], consumers stop at the
// first start that is > k or <= 0 (dreorth.F:13-16,159).  Kept in the reference's
// encoding because the U side re-uses and patches the V side's list (dlanbpro.F:480-485).
// ------------------------------------------------------------------------------------------------
struct IntervalList {
  std::vector<int> v;  // v[0..] = index(1..)
  explicit IntervalList(int cap = 0) : v(cap, 0) {}
  void reset(int cap) { v.assign(cap, 0); }
  void set_single(int s, int e, int term) { v[0] = s; v[1] = e; v[2] = term; }
  template <class F> void for_each(int k, F f) const {  // f(p,q) over valid intervals
    for (size_t i = 0; i + 1 < v.size() && v[i] <= k && v[i] > 0; i += 2) f(v[i], v[i + 1]);
  }
};

// ------------------------------------------------------------------------------------------------
// omega-recurrences: mu (loss of orthogonality of u_{j+1} against U_j) and nu (v_j against V_{j-1}).
// ------------------------------------------------------------------------------------------------
template <class R> struct OmegaRecurrence {
  std::vector<R> mu, nu;  // mu[0..k], nu[0..k]  (reference: work(imu..), work(inu..))
  void reset(int k) { mu.assign(k + 2, R(0)); nu.assign(k + 2, R(0)); }

  // dupdate_nu (dlanbpro.F:684-723).  a = B(:,1), b = B(:,2), 0-based; j is the 1-based step.
  R update_nu(int j, const R* a, const R* b, R anorm, R eps1) {
    R numax = 0;
    if (j > 1) {
      for (int k = 0; k < j - 1; ++k) {
        R t = b[k] * mu[k + 1] + a[k] * mu[k] - b[j - 2] * nu[k];
        const R d = eps1 * (hypot2(a[k], b[k]) + hypot2(a[j - 1], b[j - 2])) + eps1 * anorm;
        t = (t + std::copysign(d, t)) / a[j - 1];
        nu[k] = t;
        numax = std::max(numax, std::fabs(t));
      }
      nu[j - 1] = 1;
    }
    return numax;
  }
  // dupdate_mu (dlanbpro.F:628-680)
  R update_mu(int j, const R* a, const R* b, R anorm, R eps1) {
    R mumax;
    const R aj = a[j - 1], bj = b[j - 1];
    if (j == 1) {
      mu[0] = eps1 / b[0];
      mumax = std::fabs(mu[0]);
    } else {
      R t = a[0] * nu[0] - aj * mu[0];
      R d = eps1 * (hypot2(aj, bj) + a[0]) + eps1 * anorm;
      mu[0] = (t + std::copysign(d, t)) / bj;
      mumax = std::fabs(mu[0]);
      for (int k = 1; k < j - 1; ++k) {
        t = a[k] * nu[k] + b[k - 1] * nu[k - 1] - aj * mu[k];
        d = eps1 * (hypot2(aj, bj) + hypot2(a[k], b[k - 1])) + eps1 * anorm;
        mu[k] = (t + std::copysign(d, t)) / bj;
        mumax = std::max(mumax, std::fabs(mu[k]));
      }
      t = b[j - 2] * nu[j - 2];
      d = eps1 * (hypot2(aj, bj) + hypot2(aj, b[j - 2])) + eps1 * anorm;
      mu[j - 1] = (t + std::copysign(d, t)) / bj;
      mumax = std::max(mumax, std::fabs(mu[j - 1]));
    }
    mu[j] = 1;
    return mumax;
  }
};

// dset_mu (dlanbpro.F:555-577)
template <class R> inline void fill_intervals(int k, std::vector<R>& w, const IntervalList& idx, R val) {
  idx.for_each(k, [&](int p, int q) { for (int t = p; t <= q; ++t) w[t - 1] = val; });
}

// dcompute_int (dlanbpro.F:581-624): maximal runs of |w| >= eta that contain at least
// one |w| > delta, over w[0..j).  Leaves `idx` untouched when delta < eta, as the reference does.
template <class R> inline void select_intervals(const std::vector<R>& w, int j, R delta, R eta, IntervalList& idx) {
  if (delta < eta) { std::fprintf(stderr, "propack_b200: warning delta<eta in compute_int\n"); return; }
  int out = 0;
  idx.v[0] = 0;
  int i = 0;  // 1-based scan position: everything <= i is already classified
  while (i < j) {
    int k = i + 1;
    while (k <= j && !(std::fabs(w[k - 1]) > delta)) ++k;
    if (k > j) break;
    int s = k;
    const int lo = std::max(i, 1);
    while (s >= lo && !(std::fabs(w[s - 1]) < eta)) --s;
    idx.v[out++] = s + 1;
    i = s + 1;
    while (i <= j && !(std::fabs(w[i - 1]) < eta)) ++i;
    idx.v[out++] = i - 1;
  }
  idx.v[out++] = j + 1;
}

// ------------------------------------------------------------------------------------------------
// dbsvd.F
// ------------------------------------------------------------------------------------------------
template <class R> inline void rotate_cols(int rows, R* x, R* y, R c, R s) {  // BLAS drot
  for (int i = 0; i < rows; ++i) { const R t = c * x[i] + s * y[i]; y[i] = c * y[i] - s * x[i]; x[i] = t; }
}

// dbdqr (dbsvd.F:87-157): QR of the (n+1) x n lower bidiagonal (d,e) by Givens rotations; returns
// the last two entries (c1,c2) of Q^T e_{n+1}; optionally accumulates Q^T (ldq x (n+1)).
template <class R> inline void bidiag_qr(bool ignore_last, bool want_q, int n, R* d, R* e, R& c1, R& c2, R* Qt, int ldq) {
  if (n < 1) return;
  auto Q = [&](int i, int j) -> R& { return Qt[(size_t)j * ldq + i]; };
  if (want_q)
    for (int j = 0; j <= n; ++j) { for (int i = 0; i <= n; ++i) Q(i, j) = 0; Q(j, j) = 1; }
  auto fold = [&](int i, R cs, R sn) {  // apply rotation i to rows i, i+1 of Q^T
    for (int j = 0; j <= i; ++j) { Q(i + 1, j) = -sn * Q(i, j); Q(i, j) = cs * Q(i, j); }
    Q(i, i + 1) = sn; Q(i + 1, i + 1) = cs;
  };
  for (int i = 0; i < n - 1; ++i) {
    const Givens<R> G = make_givens(d[i], e[i]);
    d[i] = G.r; e[i] = G.s * d[i + 1]; d[i + 1] = G.c * d[i + 1];
    if (want_q) fold(i, G.c, G.s);
  }
  if (!ignore_last) {
    const Givens<R> G = make_givens(d[n - 1], e[n - 1]);
    d[n - 1] = G.r; e[n - 1] = 0; c1 = G.s; c2 = G.c;
    if (want_q) fold(n - 1, G.c, G.s);
  }
}

// dbsvdstep (dbsvd.F:5-82): one implicit-shift QR sweep on the k leading entries of the lower
// bidiagonal (D,E).  The scalar recurrence is separated from the accumulation of the rotations: `left(i, c, s)` /
// `right(i, c, s)` receive the rotation of columns (i, i+1) of the left / right accumulator.
template <class R, class FL, class FR>
inline void bidiag_shift_sweep_core(int k, R shift, R* D, R* E, FL left, FR right) {
  if (k <= 1) return;
  R x = D[0] * D[0] - shift * shift;
  R y = E[0] * D[0];
  for (int i = 0; i < k - 1; ++i) {
    Givens<R> G = make_givens(x, y);
    if (i > 0) E[i - 1] = G.r;
    x = G.c * D[i] + G.s * E[i];
    E[i] = -G.s * D[i] + G.c * E[i];
    D[i] = x;
    y = G.s * D[i + 1];
    D[i + 1] = G.c * D[i + 1];
    left(i, G.c, G.s);
    G = make_givens(x, y);
    D[i] = G.r;
    x = G.c * E[i] + G.s * D[i + 1];
    D[i + 1] = -G.s * E[i] + G.c * D[i + 1];
    E[i] = x;
    y = G.s * E[i + 1];
    E[i + 1] = G.c * E[i + 1];
    right(i, G.c, G.s);
  }
  const Givens<R> G = make_givens(x, y);
  E[k - 2] = G.r;
  x = G.c * D[k - 1] + G.s * E[k - 1];
  E[k - 1] = -G.s * D[k - 1] + G.c * E[k - 1];
  D[k - 1] = x;
  left(k - 1, G.c, G.s);
}
// ... accumulating the left rotations in U (m rows) and the right ones in V (n rows) as they are generated
template <class R>
inline void bidiag_shift_sweep(int m, int n, int k, R shift, R* D, R* E, R* U, int ldu, R* V, int ldv) {
  bidiag_shift_sweep_core<R>(
      k, shift, D, E,
      [&](int i, R c, R s) { if (U && m > 0) rotate_cols(m, U + (size_t)i * ldu, U + (size_t)(i + 1) * ldu, c, s); },
      [&](int i, R c, R s) { if (V && n > 0) rotate_cols(n, V + (size_t)i * ldv, V + (size_t)(i + 1) * ldv, c, s); });
}

// The p shifted sweeps of an implicit restart (dlansvd_irl.F:350-363) accumulate ~2 p dim rotations of dim-long columns:
// O(p dim^2) host flops between two Lanczos blocks, replicated on every rank of a multi-GPU run.  A rotation of columns
// (i, i+1) treats every row independently, so the accumulation is done row block by row block: the scalar recurrences
// run first and record the rotations; then, for 8 rows at a time, an L1-resident working copy [column][8 rows]
// (19 KB at dim = 300) takes the whole rotation sequence -- 8-wide vectorisable, no striding through the matrix -- and is
// scattered into the result.  Blocks are independent (host threads take them round-robin) and every element sees
// exactly the operations of the sequential accumulation in the same order: bit-identical.
template <class R> struct RotRec { int col; R c, s; };
template <class R>
inline void apply_rotations_blocked(int rows, int cols, R* M, int ld, const std::vector<RotRec<R>>& rots, int t, int nthreads) {
  constexpr int W = 8;
  std::vector<R> buf((size_t)cols * W);
  const int nblocks = (rows + W - 1) / W;
  for (int blk = t; blk < nblocks; blk += nthreads) {
    const int r0 = blk * W, nr = std::min(W, rows - r0);
    for (int c = 0; c < cols; ++c)          // M is the identity on entry; gather what is there so the routine stays general
      for (int l = 0; l < W; ++l) buf[(size_t)c * W + l] = l < nr ? M[(size_t)c * ld + r0 + l] : R(0);
    for (const RotRec<R>& g : rots) {
      R* x = buf.data() + (size_t)g.col * W;
      R* y = x + W;
      const R cs = g.c, sn = g.s;
      for (int l = 0; l < W; ++l) { const R tt = cs * x[l] + sn * y[l]; y[l] = cs * y[l] - sn * x[l]; x[l] = tt; }
    }
    for (int c = 0; c < cols; ++c)
      for (int l = 0; l < nr; ++l) M[(size_t)c * ld + r0 + l] = buf[(size_t)c * W + l];
  }
}
// P ((dim+1) x (dim+1), identity on entry) and Q (dim x dim): B+ = P^T B Q for the sweeps i = dim .. k+1 with shift[dim-i]
template <class R>
inline void restart_sweeps(int dim, int k, const R* shift, R* a, R* b, R* P, R* Q, int nthreads) {
  std::vector<RotRec<R>> rl, rr;
  rl.reserve((size_t)(dim - k) * (dim + 1)); rr.reserve((size_t)(dim - k) * dim);
  for (int i = dim; i >= k + 1; --i)
    bidiag_shift_sweep_core<R>(i, shift[dim - i], a, b, [&](int col, R c, R s) { rl.push_back(RotRec<R>{col, c, s}); },
                               [&](int col, R c, R s) { rr.push_back(RotRec<R>{col, c, s}); });
  nthreads = std::max(1, std::min(nthreads, 64));
  auto work = [&](int t) {
    apply_rotations_blocked<R>(dim + 1, dim + 1, P, dim + 1, rl, t, nthreads);
    apply_rotations_blocked<R>(dim, dim, Q, dim, rr, t, nthreads);
  };
  std::vector<std::thread> pool;
  int spawned = 1;
  try {
    for (; spawned < nthreads; ++spawned) pool.emplace_back(work, spawned);
  } catch (const std::exception&) {   // no more threads to be had: the caller takes the remaining shares itself
  }
  work(0);
  for (int t = spawned; t < nthreads; ++t) work(t);
  for (auto& th : pool) th.join();
}

// drefinebounds (dbsvd.F:162-231): merge bounds of clustered Ritz values, then gap theorem.
template <class R> inline void refine_bounds(int n, int k, const R* theta, R* bound, R tol, R eps34) {
  if (k <= 1) return;
  for (int i = 0; i < k; ++i)
    for (int l = -1; l <= 1; l += 2) {
      const int o = i + l;
      if (o < 0 || o >= k) continue;
      if (std::fabs(theta[i] - theta[o]) < eps34 * theta[i] && bound[i] > tol && bound[o] > tol) {
        bound[o] = hypot2(bound[i], bound[o]);
        bound[i] = 0;
      }
    }
  for (int i = 0; i < k; ++i) {
    if (!(i < k - 1 || k == n)) continue;
    R gap;
    if (i == 0) gap = std::fabs(theta[0] - theta[1]) - std::max(bound[0], bound[1]);
    else if (i == n - 1) gap = std::fabs(theta[i - 1] - theta[i]) - std::max(bound[i - 1], bound[i]);
    else {
      gap = std::fabs(theta[i] - theta[i + 1]) - std::max(bound[i], bound[i + 1]);
      gap = std::min(gap, std::fabs(theta[i - 1] - theta[i]) - std::max(bound[i - 1], bound[i]));
    }
    if (gap > bound[i]) bound[i] = bound[i] * (bound[i] / gap);
  }
}

// ||A|| estimates from the bidiagonal (dlanbpro.F:318-334 after alpha_j, :451-458 after beta_j).
// a,b 0-based; j 1-based; FUDGE only on the alpha-side formula.
template <class R> inline R anorm_after_alpha(int j, const R* a, const R* b, R amax) {
  const R FUDGE = R(1.01);
  if (j == 2) {
    const R a1 = b[0] / amax;
    return FUDGE * amax * std::sqrt((a[0] / amax) * (a[0] / amax) + a1 * a1 + a[1] / amax * a1);
  }
  const R a1 = a[j - 2] / amax, b1 = b[j - 2] / amax;
  return FUDGE * amax * std::sqrt(a1 * a1 + b1 * b1 + a1 * b[j - 3] / amax + a[j - 1] / amax * b1);
}
template <class R> inline R anorm_after_beta(int j, const R* a, const R* b, R amax) {
  if (j <= 1) return hypot2(a[0], b[0]);
  const R a1 = a[j - 1] / amax;
  return amax * std::sqrt(a1 * a1 + (b[j - 1] / amax) * (b[j - 1] / amax) + a1 * b[j - 2] / amax);
}

}  // namespace host
}  // namespace pb
