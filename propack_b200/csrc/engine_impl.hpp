// propack_b200 -- Engine<T> member definitions (drivers + Lanczos loop).  See engine.hpp.
#pragma once

namespace pb {

// ================================================================================================
// dlanbpro: k0 -> k steps of Lanczos bidiagonalisation with partial reorthogonalisation.
// Per step (no reorth, no ELR) the device executes exactly four kernels:
//   v_j  <- A^H u_j - beta_{j-1} v_{j-1}, ||.||      (one fused SpMV launch; dlanbpro.F:288-296)
//   v_j  <- v_j / alpha_j                             (dsafescal :413)
//   u_j+1<- A v_j - alpha_j u_j, ||.||                (one fused SpMV launch; :420-424)
//   u_j+1<- u_j+1 / beta_j                            (:543)
// and the host runs the O(j) omega-recurrences between them (:340-343, :463-466).
// ================================================================================================
template <class T> int Engine<T>::lanbpro(int k0, int& k, R* a, R* b, R& rnorm, R* doption, const int* ioption) {
  const R zero = 0, one = 1, FUDGE = R(1.01), kappa = R(0.717);
  const R eps = host::Machine<R>::eps;
  const R eps34 = std::pow(eps, R(0.75));
  DistScope ds(c, dist);
  op->invalidate_staged();  // vectors may have been modified since the last call (restart, re-entry rescaling)
  const R epsn = R(std::max(mg, ng)) * eps;
  const R epsn2 = std::sqrt(R(std::max(mg, ng))) * eps;
  const bool elr = ioption[1] > 0;
  const int cgs = ioption[0];
  int ierr = 0;

  // defaults (:155-180)
  const R delta = doption[0] < zero ? std::sqrt(eps / R(k)) : doption[0];
  const R eta = doption[1] < zero ? eps34 / std::sqrt(R(k)) : doption[1];
  bool full_reorth = (delta <= eta || delta == zero);
  bool force_reorth = false;
  R anorm = zero, anormest = zero;
  if (doption[2] > zero) anorm = doption[2];
  else if (k0 > 0) {
    anorm = host::hypot2(a[0], b[0]);
    if (anorm <= zero) { doption[2] = anorm; return -1; }
  }

  if (rnorm == zero) {  // :186-191
    getu0(false, k0, 3, ucol(k0 + 1), rnorm, U, ldu, ierr, cgs, anormest);
    anorm = std::max(anorm, anormest);
  }

  host::OmegaRecurrence<R> om;
  om.reset(k);
  host::IntervalList idx(2 * k + 4);

  R alpha, beta, amax;
  int j0;
  if (k0 == 0) {  // :206-230
    amax = zero; alpha = zero; beta = rnorm;
    // ||A r|| / ||r|| probe so that ||A|| is not grossly underestimated early on (:212-224).
    // The probe vector itself is discarded: it lands in column 1 of the side that is still empty.
    R sn;
    int ierr2 = 0;
    // (u_2 resp. v_1 are fully overwritten by the first Lanczos step below.)
    if (n > m) getu0(false, 0, 1, ucol(2), sn, U, ldu, ierr2, cgs, anormest);
    else getu0(true, 0, 1, vcol(1), sn, V, ldv, ierr2, cgs, anormest);
    ierr = ierr2;
    anorm = std::max(anorm, FUDGE * anormest);
    j0 = 1;
    if (beta != zero) safescal(m, beta, ucol(1));
    om.mu[0] = one; om.nu[0] = one;
  } else {  // :231-275  extend an existing factorisation
    force_reorth = true;
    alpha = a[k0 - 1]; beta = rnorm;
    if (k0 < k && beta * delta < anorm * eps) { full_reorth = true; ierr = k0; }
    idx.set_single(1, k0, k0 + 1);
    { Context::PhaseScope ps(c, PH_LEVEL1); k_scal<T>(c, m, ucol(k0 + 1), rnorm); }
    reorth(m, k0, U, ldu, ucol(k0 + 1), rnorm, idx, kappa, cgs);
    safescal(m, rnorm, ucol(k0 + 1));
    host::fill_intervals(k0, om.mu, idx, epsn2);
    host::fill_intervals(k0, om.nu, idx, epsn2);
    beta = rnorm;
    b[k0 - 1] = beta;
    amax = zero;
    for (int j = 1; j <= k0; ++j) {  // re-estimate ||A|| from B (:255-273)
      amax = std::max(amax, std::max(a[j - 1], b[j - 1]));
      if (j == 1) anorm = std::max(anorm, FUDGE * alpha);
      else anorm = std::max(anorm, host::anorm_after_alpha(j, a, b, amax));
    }
    j0 = k0 + 1;
  }
  R numax = zero, mumax = zero;

  for (int j = j0; j <= k; ++j) {
    c.ctr.nsteps += 1;
    // ---- alpha_j v_j = A^H u_j - beta_j v_{j-1} ------------------------------------------------
    {
      Pending p;
      {
        Context::PhaseScope ps(c, PH_APROD);
        op->apply(c, true, ucol(j), vcol(j), -beta, j > 1 ? vcol(j - 1) : nullptr, &p);
      }
      c.ctr.nopx += 1;
      alpha = (R)c.wait(p);
    }
    if (j == 1) {
      anorm = std::max(anorm, FUDGE * alpha);
    } else {
      if (elr && alpha < kappa * beta) {  // extended local reorthogonalisation (:301-316)
        Context::PhaseScope ps(c, PH_LEVEL1);
        R nrm = alpha;
        for (int i = 0; i < ioption[1]; ++i) {
          Pending pd, pn;
          k_dotc<T>(c, n, vcol(j - 1), vcol(j), &pd);
          double si = 0;
          const double sr = c.wait(pd, &si);
          T s; set_scalar(s, sr, si);
          k_axpy_nrm<T>(c, n, neg(s), vcol(j - 1), vcol(j), &pn);
          if (!scalar_traits<T>::is_complex && beta != zero) { beta = beta + (R)sr; b[j - 2] = beta; }  // real only (zlanbpro.F:318-331)
          nrm = (R)c.wait(pn);
          if (nrm >= kappa * alpha) break;
          alpha = nrm;
        }
        om.nu[j - 2] = eps;
        alpha = nrm;
      }
      a[j - 1] = alpha;
      amax = std::max(amax, alpha);
      anorm = std::max(anorm, host::anorm_after_alpha(j, a, b, amax));
    }
    if (!full_reorth && alpha != zero) numax = (j > 1) ? om.update_nu(j, a, b, anorm, epsn2) : numax;  // :340-343

    // ---- reorthogonalise v_j (:348-367) -------------------------------------------------------------
    if ((full_reorth || numax > delta || force_reorth) && alpha != zero) {
      if (full_reorth || eta == zero) idx.set_single(1, j - 1, j);
      else if (!force_reorth) host::select_intervals(om.nu, j - 1, delta, eta, idx);
      reorth(n, j - 1, V, ldv, vcol(j), alpha, idx, kappa, cgs);
      host::fill_intervals(j - 1, om.nu, idx, eps);
      numax = eta;
      force_reorth = !force_reorth;
    }
    // ---- invariant subspace? (:372-408) ----------------------------------------------------------------
    if (alpha < anorm * epsn && j < k) {
      rnorm = alpha; alpha = zero;
      getu0(true, j - 1, 3, vcol(j), alpha, V, ldv, ierr, cgs, anormest);
      if (alpha == zero) { k = j - 1; ierr = -j; doption[2] = anorm; return ierr; }
      safescal(n, alpha, vcol(j));
      alpha = zero; force_reorth = true;
      if (delta > zero) full_reorth = false;
    } else if (j > 1 && !full_reorth && j < k && (delta * alpha < anorm * eps)) {
      ierr = j;
    }
    a[j - 1] = alpha;
    if (alpha != zero) normalize_for_apply(false, n, alpha, vcol(j));

    // ---- beta_{j+1} u_{j+1} = A v_j - alpha_j u_j -------------------------------------------------------
    {
      Pending p;
      {
        Context::PhaseScope ps(c, PH_APROD);
        op->apply(c, false, vcol(j), ucol(j + 1), -alpha, ucol(j), &p);
      }
      c.ctr.nopx += 1;
      beta = (R)c.wait(p);
    }
    if (elr && beta < kappa * alpha) {  // (:429-443)
      Context::PhaseScope ps(c, PH_LEVEL1);
      R nrm = beta;
      for (int i = 0; i < ioption[1]; ++i) {
        Pending pd, pn;
        k_dotc<T>(c, m, ucol(j), ucol(j + 1), &pd);
        double si = 0;
        const double sr = c.wait(pd, &si);
        T s; set_scalar(s, sr, si);
        k_axpy_nrm<T>(c, m, neg(s), ucol(j), ucol(j + 1), &pn);
        if (!scalar_traits<T>::is_complex && alpha != zero) { alpha = alpha + (R)sr; a[j - 1] = alpha; }
        nrm = (R)c.wait(pn);
        if (nrm >= kappa * beta) break;
        beta = nrm;
      }
      om.mu[j - 1] = eps;
      beta = nrm;
    }
    b[j - 1] = beta;
    amax = std::max(amax, beta);
    anorm = std::max(anorm, host::anorm_after_beta(j, a, b, amax));
    if (!full_reorth && beta != zero) mumax = om.update_mu(j, a, b, anorm, epsn2);  // :463-466

    // ---- reorthogonalise u_{j+1} (:471-498) -----------------------------------------------------------
    if ((full_reorth || mumax > delta || force_reorth) && beta != zero) {
      if (full_reorth || eta == zero) idx.set_single(1, j, j + 1);
      else if (!force_reorth) host::select_intervals(om.mu, j, delta, eta, idx);
      else {
        // forced: re-use the V-side list; its terminator (== j) must now exceed k = j (:480-485)
        for (int i = 0; i < 2 * j + 1 && i < (int)idx.v.size(); ++i)
          if (idx.v[i] == j) { idx.v[i] = j + 1; break; }
      }
      reorth(m, j, U, ldu, ucol(j + 1), beta, idx, kappa, cgs);
      host::fill_intervals(j, om.mu, idx, eps);
      mumax = eta;
      force_reorth = !force_reorth;
    }
    // ---- invariant subspace? (:503-539) ----------------------------------------------------------------
    if (beta < anorm * epsn && j < k) {
      rnorm = beta; beta = zero;
      getu0(false, j, 3, ucol(j + 1), beta, U, ldu, ierr, cgs, anormest);
      if (beta == zero) { k = j; ierr = -j; doption[2] = anorm; return ierr; }
      safescal(m, beta, ucol(j + 1));
      beta = zero; force_reorth = true;
      if (delta > zero) full_reorth = false;
    } else if (!full_reorth && j < k && (delta * beta < anorm * eps)) {
      ierr = j;
    }
    b[j - 1] = beta;
    if (beta != zero && beta != one) normalize_for_apply(true, m, beta, ucol(j + 1));
    rnorm = beta;
  }
  doption[2] = anorm;  // :547 (in/out)
  return ierr;
}

// ================================================================================================
// dritzvec: host bidiagonal SVD (dbdqr + dbdsdc), then the two tall in-place GEMMs on device.
// ================================================================================================
// The reference's route to the two small matrices of dritzvec (dritzvec.F:116-193): WU ((dim+1) x k, ld dim+1) and
// WV (dim x k, ld dim) such that U(:,1:k) <- U(:,1:dim+1) WU and V(:,1:k) <- V(:,1:dim) WV.  D, E are overwritten.
template <class R>
void ritz_w_reference(bool ignorelast, bool smallest, bool want_u, bool want_v, int k, int dim, R* D, R* E, std::vector<R>& WU,
                      std::vector<R>& WV) {
  std::vector<R> Mt((size_t)(dim + 1) * (dim + 1), R(0)), Qt((size_t)dim * dim, R(0)), P((size_t)dim * dim, R(0));
  R c1 = 0, c2 = 0;
  int info = 0;
  host::bidiag_qr(ignorelast, want_u, dim, D, E, c1, c2, Mt.data(), dim + 1);  // dritzvec.F:116
  host::bdsdc_full(dim, D, E, P.data(), dim, Qt.data(), dim, &info);         // :123
  const int mstart = smallest ? dim - k : 0;  // 0-based first wanted row of X / Q^T
  if (want_u) {
    // X = P^T M^T(1:dim,:) (:130), and the wanted product is U(:,1:dim+1) * X(mstart:mstart+k,:)^T (:160):
    // W (K=dim+1 x N=k) with W(l,jn) = X(mstart+jn, l)
    WU.assign((size_t)(dim + 1) * k, R(0));
    for (int jn = 0; jn < k; ++jn)
      for (int l = 0; l < dim + 1; ++l) {
        R s = 0;
        const R* pc = P.data() + (size_t)(mstart + jn) * dim;  // column mstart+jn of P
        const R* mc = Mt.data() + (size_t)l * (dim + 1);       // column l of M^T (first dim rows)
        for (int t = 0; t < dim; ++t) s += pc[t] * mc[t];
        WU[(size_t)jn * (dim + 1) + l] = s;
      }
  }
  if (want_v) {
    // V(:,1:k) = V(:,1:dim) * Qt(mstart:mstart+k,:)^T (:193): W(l,jn) = Qt(mstart+jn, l)
    WV.assign((size_t)dim * k, R(0));
    for (int jn = 0; jn < k; ++jn)
      for (int l = 0; l < dim; ++l) WV[(size_t)jn * dim + l] = Qt[(size_t)l * dim + mstart + jn];
  }
}

template <class T> void Engine<T>::ritzvec(bool smallest, bool jobu, bool jobv, int k, int dim, R* D, R* E, bool reference_route) {
  Context::PhaseScope ps(c, PH_RITZ);
  DistScope ds(c, dist);
  std::vector<R> WU, WV;
  // Large Krylov dimension, largest triplets: the k leading singular vector pairs of B directly (dqds + inverse iteration
  // on the Golub-Kahan tridiagonal) instead of a divide & conquer SVD of all dim pairs; the reference route is the fallback.
  bool fast = false;
  if (!reference_route && fast_ritz_bounds() && !smallest && dim >= 128 && dim != std::min(mg, ng) && k < dim)
    fast = host::ritz_vectors_leading(dim, D, E, k, WU, WV);
  if (!fast) ritz_w_reference<R>(dim == std::min(mg, ng), smallest, jobu, jobv, k, dim, D, E, WU, WV);
  if (jobu) k_gemm_tall<T>(c, m, k, dim + 1, U, ldu, WU.data());
  if (jobv) k_gemm_tall<T>(c, n, k, dim, V, ldv, WV.data());
  c.sync();
}

// ================================================================================================
// dlansvd
// ================================================================================================
template <class T>
int Engine<T>::lansvd(bool jobu, bool jobv, int& k, int kmax, R* sigma, R* bnd, R tolin, R* doption, const int* ioption) {
  const R one = 1, zero = 0;
  const R eps = host::Machine<R>::eps;
  const R eps34 = std::pow(eps, R(0.75));
  DistScope ds(c, dist);
  const R epsn = R(std::max(mg, ng)) * eps / R(2);
  const int lanmax = std::min(std::min(ng + 1, mg + 1), kmax);
  const R tol = std::min(one, std::max(R(16) * eps, tolin));
  if (lanmax + 1 > ucols || lanmax > vcols) throw std::runtime_error("propack_b200: basis buffers smaller than kmax");
  std::vector<R> a(lanmax + 1, zero), b(lanmax + 1, zero), th(lanmax + 1), ee(lanmax + 1), wb(lanmax + 2, zero);
  R anorm = zero, rnorm;
  int ierr = 0, info = 0;

  rnorm = nrm2(m, ucol(1));
  if (rnorm == zero) getu0(false, 0, 1, ucol(1), rnorm, U, ldu, ierr, ioption[0], anorm);  // dlansvd.F:165-169
  c.ctr.nsing = k;
  int neig = 0, jold = 0;
  int j = std::min(k + std::max(8, k) + 1, lanmax);
  while (neig < k) {
    ierr = lanbpro(jold, j, a.data(), b.data(), rnorm, doption, ioption);  // :185
    jold = j;
    {
      Context::PhaseScope ps(c, PH_HOST_BSVD);
      // Ritz values and bounds from the (j+1) x j bidiagonal (:193-215).  Only the k leading values and bounds are
      // used below, and drefinebounds couples a bound to its two neighbours only, so for large j the k+4 leading
      // values and last-row components are computed directly (host::ritz_leading: dqds + inverse iteration, ~3x
      // cheaper than the QR sweep over all j values) and the reference's xBDSQR route is the fallback.
      int nb = j;
      bool fast = false;
      if (fast_ritz_bounds() && j >= 128 && j != std::min(mg, ng) && j > k + 8) {
        nb = k + 4;
        fast = host::ritz_leading(j, a.data(), b.data(), nb, th.data(), wb.data());
        if (!fast) nb = j;
      }
      if (!fast) {
        std::copy(a.begin(), a.begin() + j, th.begin());
        std::copy(b.begin(), b.begin() + j, ee.begin());
        std::fill(wb.begin(), wb.begin() + j + 1, zero);
        int lapinfo = 0;
        host::bidiag_qr<R>(j == std::min(mg, ng), false, j, th.data(), ee.data(), wb[j - 1], wb[j], nullptr, 0);
        host::bdsqr_row(j, th.data(), ee.data(), wb.data(), &lapinfo);
      }
      c.ctr.nbsvd += 1;
      anorm = (j > 5) ? th[0] : std::max(anorm, th[0]);
      for (int i = 0; i < nb; ++i) wb[i] = std::fabs(rnorm * wb[i]);
      host::refine_bounds(std::min(mg, ng), nb, th.data(), wb.data(), epsn * anorm, eps34);
      for (int i = 0; i < std::min(j, k); ++i) bnd[i] = wb[i];
      neig = 0;  // leading converged values only (:222-236)
      for (int i = 0; i < std::min(j, k); ++i) {
        if (wb[i] <= tol * th[i]) sigma[neig++] = th[i];
        else break;
      }
    }
    if (ierr < 0) { if (j < k) info = j; break; }       // invariant subspace (:242-249)
    if (j >= lanmax) { if (neig < k) info = -1; break; }  // Krylov space exhausted (:250-259)
    int dj;  // grow the Krylov dimension (:268-275)
    if (neig > 1) { dj = std::min(j / 2, ((k - neig) * (j - 6)) / (2 * neig + 1)); dj = std::min(100, std::max(2, dj)); }
    else { dj = j / 2; dj = std::min(100, std::max(10, dj)); }
    j = std::min(j + dj, lanmax);
  }
  if ((neig >= k || info > 0) && (jobu || jobv)) {  // :278-288
    std::vector<R> D(a.begin(), a.begin() + jold), E(b.begin(), b.begin() + jold);
    ritzvec(false, jobu, jobv, neig, jold, D.data(), E.data());
  }
  k = neig;
  c.ctr.nlandim = j;
  return info;
}

// ================================================================================================
// dlansvd_irl
// ================================================================================================
template <class T>
int Engine<T>::lansvd_irl(bool smallest, bool jobu, bool jobv, int& dim, int p, int& neig, int maxiter, R* sigma, R* bnd,
                          R tolin, R* doption, const int* ioption) {
  const R one = 1, zero = 0;
  const R eps = host::Machine<R>::eps;
  const R eps34 = std::pow(eps, R(0.75));
  DistScope ds(c, dist);
  const R epsn = R(std::max(mg, ng)) * eps / R(2);
  dim = std::min(dim, std::min(ng + 1, mg + 1));  // dlansvd_irl.F:170
  const int k = dim - p;
  const R tol = std::min(one, std::max(R(16) * eps, tolin));
  if (dim + 1 > ucols || dim > vcols) throw std::runtime_error("propack_b200: basis buffers smaller than dim");
  std::vector<R> a(dim + 1, zero), b(dim + 1, zero), th(dim + 1), ee(dim + 1), wb(dim + 2, zero), shift(dim + 1, zero);
  std::vector<R> P((size_t)(dim + 1) * (dim + 1)), Q((size_t)dim * dim);
  R anorm = zero, rnorm;
  int ierr = 0, info = 0;

  rnorm = nrm2(m, ucol(1));
  if (rnorm == zero) getu0(false, 0, 1, ucol(1), rnorm, U, ldu, ierr, ioption[0], anorm);
  int iter = 0, nconv = 0, kold = 0;
  while (nconv < neig && iter < maxiter) {
    ierr = lanbpro(kold, dim, a.data(), b.data(), rnorm, doption, ioption);  // :213
    kold = k;
    {
      Context::PhaseScope ps(c, PH_HOST_BSVD);
      std::copy(a.begin(), a.begin() + dim, th.begin());
      std::copy(b.begin(), b.begin() + dim, ee.begin());
      std::fill(wb.begin(), wb.begin() + dim + 1, zero);
      int lapinfo = 0;
      host::bidiag_qr<R>(dim == std::min(mg, ng), false, dim, th.data(), ee.data(), wb[dim - 1], wb[dim], nullptr, 0);
      host::bdsqr_row(dim, th.data(), ee.data(), wb.data(), &lapinfo);
      c.ctr.nbsvd += 1;
      anorm = (dim > 5) ? th[0] : std::max(anorm, th[0]);
      for (int i = 0; i < dim; ++i) wb[i] = std::fabs(rnorm * wb[i]);
      host::refine_bounds(std::min(mg, ng), smallest ? dim : std::min(dim, neig), th.data(), wb.data(), epsn * anorm, eps34);
      nconv = 0;  // :262-290
      if (smallest) {
        for (int i = dim - neig; i < dim; ++i)
          if (wb[i] <= tol * th[0]) { sigma[nconv] = th[i]; bnd[nconv] = wb[i]; ++nconv; }
      } else {
        for (int i = 0; i < std::min(dim, neig); ++i) {
          if (wb[i] <= tol * th[i]) { sigma[nconv] = th[i]; bnd[nconv] = wb[i]; ++nconv; }
          else break;
        }
      }
    }
    if (ierr < 0) { if (dim < k) info = dim; break; }  // :295-302
    if (nconv < neig) {
      Context::PhaseScope ps(c, PH_RESTART);
      // exact shifts with the relative-gap guard doption(4) (:318-344)
      int nshft = 0;
      if (smallest) {
        // The reference fills k = dim-p shifts here (dlansvd_irl.F:320-332) although the p sweeps below consume p of them: for
        // p > dim/2 the last p-k sweeps run with the zero the workspace was cleared to, which is the documented
        // "incorrect results if WHICH='S' and P>DIM/2" (Changelog:56).  Filling all p shifts -- the p largest Ritz values,
        // the ones a smallest-triplet restart must purge -- is identical for p <= dim/2 and correct beyond.
        for (int i = 0; i < p; ++i) {
          const R ref = th[dim - neig - 1];
          const R relgap = th[i] - wb[i] - ref;
          shift[nshft++] = (relgap > doption[3] * ref) ? th[i] : th[0];
        }
      } else {
        for (int i = dim - 1; i >= k; --i) {
          const R relgap = th[k - 1] - (th[i] + wb[i]);
          shift[nshft++] = (relgap > doption[3] * th[k - 1]) ? th[i] : zero;
        }
      }
      // accumulate the rotations of the p shifted QR sweeps: B+ = P^T B Q (:350-363)
      std::fill(P.begin(), P.end(), zero);
      std::fill(Q.begin(), Q.end(), zero);
      for (int i = 0; i <= dim; ++i) P[(size_t)i * (dim + 2)] = one;
      for (int i = 0; i < dim; ++i) Q[(size_t)i * (dim + 1)] = one;
      // (scalar recurrences first, then the ~2 p dim recorded rotations applied row-parallel by a few host threads:
      // host_algebra.hpp::restart_sweeps -- bit-identical to the sequential accumulation)
      host::restart_sweeps<R>(dim, k, shift.data(), a.data(), b.data(), P.data(), Q.data(), host_threads());
      // U(:,1:k+1) <- U(:,1:dim+1) P(:,1:k+1);  V(:,1:k) <- V(:,1:dim) Q(:,1:k)  (:387-395)
      k_gemm_tall<T>(c, m, k + 1, dim + 1, U, ldu, P.data());   // P is (dim+1)x(dim+1), ld = dim+1 = K
      k_gemm_tall<T>(c, n, k, dim, V, ldv, Q.data());           // Q is dim x dim, ld = dim = K
      c.sync();
      rnorm = b[k - 1];
      c.ctr.nrestart += 1;
    }
    iter += 1;
  }
  if ((nconv >= neig || info > 0) && (jobu || jobv)) {  // :405-416
    std::vector<R> D(a.begin(), a.begin() + dim), E(b.begin(), b.begin() + dim);
    ritzvec(smallest, jobu, jobv, nconv, dim, D.data(), E.data());
  }
  neig = nconv;
  c.ctr.nlandim = dim;
  return info;
}

}  // namespace pb
