// propack_b200 -- device-resident Lanczos bidiagonalisation engine (host control flow).
//
// Mirrors the reference's call tree  xLANSVD / xLANSVD_IRL -> xLANBPRO -> {APROD, xREORTH,
// xGETU0, xSAFESCAL} and xRITZVEC -> xGEMM_OVWR_LEFT  (SURVEY.md section 3), but the Lanczos bases U, V and
// the operator live in HBM for the whole solve: the host only sees the scalars the algorithm
// branches on (alpha, beta, dot products, norms) and the O(k) recurrences / O(k^2) bidiagonal
// problem.  There is no CPU fallback: every vector operation below is a CUDA kernel launch.
#pragma once
#include <cmath>
#include <cstdlib>
#include <functional>
#include <memory>
#include <stdexcept>
#include <thread>
#include <vector>

#include "comm.hpp"
#include "context.hpp"
#include "host_algebra.hpp"
#include "host_lapack.hpp"
#include "kernels.cuh"

namespace pb {

// Fortran-ABI user APROD (dlansvd.F:23-31): hidden CHARACTER length appended by value.
template <class T>
using aprod_f77_t = void (*)(const char* transa, const int* m, const int* n, const T* x, T* y, void* parm, int* iparm,
                             size_t transa_len);

// PROPACK_B200_FAST_RITZ_BOUNDS=0 forces the reference's xBDSQR route for the per-iteration Ritz bounds (cross-check)
inline bool fast_ritz_bounds() {
  static const bool on = [] { const char* e = std::getenv("PROPACK_B200_FAST_RITZ_BOUNDS"); return !(e && e[0] == '0'); }();
  return on;
}

// host threads for the O(p dim^2) restart accumulation (PROPACK_B200_HOST_THREADS; default: the cores this rank can
// claim, at most 8 -- one process per GPU shares the host with the other ranks)
inline int host_threads() {
  static const int n = [] {
    if (const char* e = std::getenv("PROPACK_B200_HOST_THREADS")) return std::max(1, std::atoi(e));
    const unsigned hc = std::thread::hardware_concurrency();
    const int world = std::max(1, Comm::get().world);
    return (int)std::max(1u, std::min(8u, (hc ? hc : 1u) / (unsigned)world));
  }();
  return n;
}

// measurement only (set_option("bench_skip_gather", 1) + propack_b200_bench_spmv on a row-sharded operator): time the local
// panel launches alone, on whatever the gather buffer holds, without the all-gather of the input vector
inline bool& bench_skip_gather() { static bool v = false; return v; }

inline void set_scalar(float& s, double re, double) { s = (float)re; }
inline void set_scalar(double& s, double re, double) { s = re; }
template <class R> inline void set_scalar(cplx<R>& s, double re, double im) { s = cplx<R>((R)re, (R)im); }
inline float neg(float a) { return -a; }
inline double neg(double a) { return -a; }
template <class R> inline cplx<R> neg(cplx<R> a) { return cplx<R>(-a.x, -a.y); }

// ------------------------------------------------------------------------------------------------
// Linear operators.  apply(): y = op(A) x + coef*prev (prev may be null), optional ||y|| publication.
// ------------------------------------------------------------------------------------------------
template <class T> struct LinOp {
  using R = real_t<T>;
  int m = 0, n = 0;            // rows / columns held by THIS process (= the global sizes on one GPU)
  // row-sharded operators (one process per GPU): global sizes, this rank's first global row / column,
  // and the common slice lengths that the engine must use as leading dimensions of U and V
  int mg = 0, ng = 0;
  long m_off = 0, n_off = 0, ld_m = 0, ld_n = 0;
  bool sharded = false;
  int global_m() const { return sharded ? mg : m; }
  int global_n() const { return sharded ? ng : n; }
  virtual ~LinOp() {}
  virtual void apply(Context& c, bool adjoint, const T* x, T* y, R coef, const T* prev, Pending* nrm) = 0;
  virtual double algorithmic_bytes(bool adjoint) const = 0;  // per apply, SURVEY 8d byte model
  // Row-sharded operators may fuse "x <- scale*x" with the all-gather the next apply(adjoint, x, ...) needs.
  // Returns false when not supported (the caller scales x itself and apply() gathers).
  virtual bool stage_scaled(Context&, bool /*adjoint*/, R /*scale*/, T* /*x*/) { return false; }
  virtual void invalidate_staged() {}
};

// One sparse operand in its two device forms: the CSR arrays (kept: the long-row kernel and the transpose export read
// them; PROPACK_B200_SPMV=csr also runs the whole product from them) and the sliced jagged-ELL copy the default kernel streams.
template <class T> struct SparseOperand {
  CsrDevice<T> csr;
  SellStorage<T> sell;
  DeviceBuffer<int> long_rows;
};

inline bool spmv_use_sell() {
  static const bool on = [] { const char* e = std::getenv("PROPACK_B200_SPMV"); return !(e && (e[0] == 'c' || e[0] == 'C')); }();
  return on;
}

// CSR arrays already on the device (0-based, validated) -> long-row list, lanes-per-row of the CSR kernel, SELL copy
template <class T>
void finish_operand(Context& c, SparseOperand<T>& S, int rows, int cols, long nnz, const int* rp, const int* ci, const T* va,
                    int ctas_per_sm = kSellCtasPerSm) {
  CsrDevice<T>& D = S.csr;
  D.rows = rows; D.cols = cols; D.nnz = nnz; D.rp = rp; D.ci = ci; D.va = va;
  const bool sell = spmv_use_sell();
  D.n_long = k_csr_long_rows(c, rows, rp, sell ? kSellLong : spmv_group_nnz<T>() / 2, S.long_rows);
  D.long_rows = D.n_long ? S.long_rows.p : nullptr;
  D.lpr_log2 = csr_lanes_per_row_log2(nnz, rows, spmv_group_nnz<T>());
  if (const char* e = std::getenv("PROPACK_B200_SPMV_LPR_LOG2")) D.lpr_log2 = std::min(5, std::max(0, std::atoi(e)));
  if (sell) sell_build<T>(c, rows, cols, nnz, rp, ci, va, S.sell, ctas_per_sm);
}

// A sparse operand split by column range into panels ("phases"): y = sum_g A_g x.  One panel = the whole operand.
// Single GPU: the panels are L2-sized column blocks (sell.cu, "Column blocking"); row-sharded: the source ranks of the
// gathered vector, in the order their slices arrive.
template <class T> struct PanelSet {
  using R = real_t<T>;
  struct Panel {
    DeviceBuffer<int> rp, ci;
    DeviceBuffer<T> va;
    SparseOperand<T> S;
    unsigned int src_mask = 0;   // row-sharded: ranks whose slices this panel reads (own rank excluded: no wait needed)
  };
  std::vector<std::unique_ptr<Panel>> panels;
  long nnz = 0;
  int rows = 0;
  // Split the CSR (device pointers) into G panels: panel of column c = ring distance of (c / ld) behind `rank`, / (P/G).
  // Returns the validation status of k_csr_split_phases (0 = ok).
  int build(Context& c, int rows_, long width, long nnz_, const int* rp, const int* ci, const T* va, long ld, int P, int rank, int G,
            int ctas_per_sm) {
    rows = rows_; nnz = nnz_;
    panels.clear();
    std::vector<int*> out_rp(G);
    std::vector<DeviceBuffer<int>*> out_ci(G);
    std::vector<DeviceBuffer<T>*> out_va(G);
    std::vector<long> nnz_g(G, 0);
    for (int g = 0; g < G; ++g) {
      panels.emplace_back(new Panel());
      panels[g]->rp.alloc((size_t)rows + 1);
      out_rp[g] = panels[g]->rp.p; out_ci[g] = &panels[g]->ci; out_va[g] = &panels[g]->va;
    }
    const int st = k_csr_split_phases<T>(c, rows, width, rp, ci, va, ld, P, rank, G, out_rp.data(), out_ci.data(), out_va.data(),
                                         nnz_g.data());
    if (st) return st;
    const int per = P / G;
    for (int g = 0; g < G; ++g) {
      Panel& pn = *panels[g];
      finish_operand<T>(c, pn.S, rows, (int)width, nnz_g[g], pn.rp.p, pn.ci.p, pn.va.p, ctas_per_sm);
      if (spmv_use_sell() && pn.S.csr.n_long == 0) {   // the jagged copy is all the kernels read: drop the panel's CSR entries
        pn.ci.free(); pn.va.free();
        pn.S.csr.ci = nullptr; pn.S.csr.va = nullptr;
      }
      pn.src_mask = 0;
      if (P > 1 && G <= P)
        for (int k = g * per; k < (g + 1) * per; ++k) {
          const int src = ((rank - k) % P + P) % P;
          if (src != rank) pn.src_mask |= 1u << src;
        }
    }
    return 0;
  }
  // y = sum_g A_g x + coef*prev, ||y||.  flags != null: panel g first waits for the arrival flags of its sources.
  void apply(Context& c, bool conj, const T* x, T* y, R coef, const T* prev, Pending* nrm, const unsigned long long* flags,
             unsigned long long epoch) {
    const size_t G = panels.size();
    for (size_t g = 0; g < G; ++g) {
      Panel& pn = *panels[g];
      const unsigned int mask = flags ? pn.src_mask : 0u;
      const bool last = g + 1 == G;
      if (spmv_use_sell()) {
        const int mode = (g ? kSellModeAcc : 0) | (last ? kSellModeFinal : 0);
        if (mask && pn.S.csr.n_long > 0) k_wait_flags(c, flags, mask, epoch);   // the long-row kernel has no in-kernel wait
        k_spmv_sell<T>(c, pn.S.sell.dev, &pn.S.csr, conj, x, y, coef, prev, last ? nrm : nullptr, mode, flags, mask, epoch);
      } else {
        // CSR kernel: the epilogue term goes in with the first panel, later panels add onto y
        if (mask) k_wait_flags(c, flags, mask, epoch);
        k_spmv<T>(c, pn.S.csr, conj, x, y, g == 0 ? coef : R(1), g == 0 ? prev : y, last ? nrm : nullptr);
      }
    }
  }
  double stream_bytes() const {   // matrix entries + row lengths / pointers + y traffic of the panel passes
    const double w = sizeof(T);
    return (double)nnz * (w + 4) + ((double)rows + 1) * 4 + (double)rows * w;
  }
};

// Panels of a single-GPU operand: as many column blocks as it takes to keep one block of the gathered vector L2-resident
// (PROPACK_B200_SPMV_COLBLOCK_MB, default 48 MB per block -- measured on config 5: two 40 MB blocks beat one 80 MB vector by 18 %
// and four 20 MB blocks lose to the per-panel overhead; PROPACK_B200_SPMV_COLBLOCKS forces the count), at most 8.
template <class T> inline int column_blocks(long cols) {
  if (const char* e = std::getenv("PROPACK_B200_SPMV_COLBLOCKS")) return std::min(8, std::max(1, std::atoi(e)));
  double mb = 48.0;
  if (const char* e = std::getenv("PROPACK_B200_SPMV_COLBLOCK_MB")) mb = std::max(1.0, std::atof(e));
  const double need = (double)cols * sizeof(T) / (mb * 1048576.0);
  int G = 1;
  while (G < 8 && G < need) G *= 2;
  return G;
}

template <class T> struct CsrOperator : LinOp<T> {
  using R = real_t<T>;
  DeviceBuffer<int> rp, ci, trp, tci;
  DeviceBuffer<T> va, tva;
  PanelSet<T> A, At;  // At = A^T (values not conjugated)
  void apply(Context& c, bool adjoint, const T* x, T* y, R coef, const T* prev, Pending* nrm) override {
    (adjoint ? At : A).apply(c, /*conj=*/adjoint, x, y, coef, prev, nrm, nullptr, 0ull);
  }
  double algorithmic_bytes(bool adjoint) const override {
    const PanelSet<T>& S = adjoint ? At : A;
    return S.stream_bytes() + (double)(adjoint ? this->m : this->n) * sizeof(T);
  }
};

// Row-sharded CSR operator: this rank holds its row block of A (m_loc x ng) and the transpose of its column block
// (n_loc x mg), so BOTH products are "all-gather the input vector, then a local gather-SpMV": no cross-rank summation
// inside a matvec.  Slices are contiguous and of equal padded length, so global column indices address the gathered
// vector directly.
//
// Each local operand is split by the SOURCE RANK of the gathered vector into G panels (G = min(world, 4) by default,
// PROPACK_B200_SPMV_PHASES): panel g holds the entries whose column belongs to the ranks at ring distance
// [g*P/G, (g+1)*P/G) behind this rank -- the order in which the staggered push (k_scal_push) delivers the slices.
// Panel g waits, inside its kernel, only for its own sources, so the SpMV over the slices that have landed overlaps the
// NVLink transfer of the ones still in flight.  The push runs on a few CTAs of a side stream; the SpMV kernels leave
// it room (one CTA slot per SM stays free).
template <class T> struct ShardedCsrOperator : LinOp<T> {
  using R = real_t<T>;
  PanelSet<T> sets[2];
  DeviceBuffer<T> xbuf[2];
  // gather buffers: [world*ld elements | kMaxRanks arrival flags]; peer windows when NVLink peer memory is mapped
  // (index 0: the n-vector gathered for A x, index 1: the m-vector gathered for A^H x)
  T* xfull[2] = {nullptr, nullptr};
  Comm::Window win[2];
  bool fused = false;
  const T* staged_ptr[2] = {nullptr, nullptr};
  bool staged_valid[2] = {false, false};
  unsigned long long epoch[2] = {0, 0};
  long ld_of(int d) const { return d ? this->ld_m : this->ld_n; }
  void alloc_gather_buffers() {
    Comm& cm = Comm::get();
    fused = cm.peer_ok;
    for (int d = 0; d < 2; ++d) {
      const size_t bytes = sizeof(T) * (size_t)ld_of(d) * cm.world + sizeof(unsigned long long) * Comm::kMaxRanks;
      if (fused) { win[d] = cm.alloc_window(bytes); xfull[d] = static_cast<T*>(win[d].base[cm.rank]); }
      else {
        xbuf[d].alloc(bytes / sizeof(T) + 1);
        PB_CUDA(cudaMemset(xbuf[d].p, 0, bytes));
        xfull[d] = xbuf[d].p;
      }
    }
  }
  ~ShardedCsrOperator() override {
    if (fused) { cudaDeviceSynchronize(); Comm& cm = Comm::get(); cm.free_window(win[0]); cm.free_window(win[1]); }
  }
  // ---- transport of the fused all-gather ------------------------------------------------------------------------
  // Default: the SM-store push kernel of k_scal_push (64 thin CTAs on the side stream).  PROPACK_B200_PUSH=ce moves the
  // slices with the COPY ENGINES instead: for every destination at ring distance k = 1..P-1 a peer cudaMemcpyAsync of
  // this rank's slice (own gather slot -> the same slot of the destination's window) followed by an 8-byte copy of the
  // epoch word into the destination's arrival flag, all on the side stream -- stream order makes the flag land after the
  // data and the slices arrive in ring order, as the phase-split SpMV expects.  No SM executes a store for the transfer
  // (the SpMV may then use every CTA slot).  Measured on 8 GPUs, config 5: 982-1000 ms against 937-954 ms for the push
  // kernel (a dependent chain of 10 MB peer copies pays ~6 us of engine start-up per copy; the same copies as a CUDA
  // graph with 1 / 2 / 3 parallel chains: 1017 / 1042 / 1102 ms, all at once: 1293 ms), so the push kernel stays the default.
  const bool ce_push = [] { const char* e = std::getenv("PROPACK_B200_PUSH"); return e && (e[0] == 'c' || e[0] == 'C'); }();   // per operator
  bool push_by_copy_engine() const { return ce_push; }
  DeviceBuffer<unsigned long long> epoch_dev;
  void push_by_copies(Context& c, int d, long len) {
    Comm& cm = Comm::get();
    const long ld = ld_of(d);
    T* self = xfull[d] + (size_t)cm.rank * ld;
    for (int k = 1; k < cm.world; ++k) {
      const int dest = (cm.rank + k) % cm.world;
      T* dst_base = static_cast<T*>(win[d].base[dest]);
      unsigned long long* dst_flags = reinterpret_cast<unsigned long long*>(dst_base + (size_t)cm.world * ld);
      if (len > 0)
        PB_CUDA(cudaMemcpyAsync(dst_base + (size_t)cm.rank * ld, self, sizeof(T) * (size_t)len, cudaMemcpyDeviceToDevice, c.stream2));
      PB_CUDA(cudaMemcpyAsync(dst_flags + cm.rank, epoch_dev.p + d, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c.stream2));
    }
  }
  const unsigned long long* flags(int d) const {
    return reinterpret_cast<const unsigned long long*>(xfull[d] + (size_t)ld_of(d) * Comm::get().world);
  }
  bool stage_scaled(Context& c, bool adjoint, R scale, T* x) override {
    if (!fused) return false;
    const int d = adjoint ? 1 : 0;
    Comm& cm = Comm::get();
    epoch[d] += 1;
    const long len = adjoint ? this->m : this->n;
    T* self = xfull[d] + (size_t)cm.rank * ld_of(d);
    if (push_by_copy_engine()) {
      if (!epoch_dev.p) { epoch_dev.alloc(2); PB_CUDA(cudaMemset(epoch_dev.p, 0, 2 * sizeof(unsigned long long))); }
      k_scal_local<T>(c, len, x, scale, self, epoch_dev.p + d, epoch[d]);
      PB_CUDA(cudaEventRecord(c.ev_fork, c.stream));
      PB_CUDA(cudaStreamWaitEvent(c.stream2, c.ev_fork, 0));
      push_by_copies(c, d, len);
    } else {
      k_scal_push<T>(c, len, ld_of(d), x, scale, win[d].table_dev, cm.rank, cm.world, epoch[d], self);
    }
    staged_ptr[d] = x; staged_valid[d] = true;
    return true;
  }
  void invalidate_staged() override { staged_valid[0] = staged_valid[1] = false; }
  void apply(Context& c, bool adjoint, const T* x, T* y, R coef, const T* prev, Pending* nrm) override {
    Comm& cm = Comm::get();
    const int d = adjoint ? 1 : 0;
    // staged: the slices are being pushed by the producers (stage_scaled); each panel waits only for its own sources.
    // (nrm != nullptr: the cross-rank norm reduction that follows is what makes reusing the buffer safe.)
    const bool staged = staged_valid[d] && staged_ptr[d] == x && nrm != nullptr;
    if (!staged && !bench_skip_gather()) cm.allgather(x, xfull[d], sizeof(T) * (size_t)ld_of(d), c.stream);
    staged_valid[d] = false;
    sets[d].apply(c, /*conj=*/adjoint, xfull[d], y, coef, prev, nrm, staged ? flags(d) : nullptr, epoch[d]);
  }
  double algorithmic_bytes(bool adjoint) const override {
    return sets[adjoint ? 1 : 0].stream_bytes() + (double)ld_of(adjoint ? 1 : 0) * Comm::get().world * sizeof(T);
  }
};

// Dense column-major operator (BASELINE config 3): both products are the reorthogonalisation
// GEMV kernels over A itself.
template <class T> struct DenseOperator : LinOp<T> {
  using R = real_t<T>;
  DeviceBuffer<T> store;
  const T* A = nullptr; long lda = 0;
  void apply(Context& c, bool adjoint, const T* x, T* y, R coef, const T* prev, Pending* nrm) override {
    if (adjoint) {
      k_gemv_t<T>(c, this->m, this->n, A, lda, x, y);
      if (prev) k_axpy_nrm<T>(c, this->n, T(coef), prev, y, nrm ? nrm : &scratch_p);
      else if (nrm) k_nrm2<T>(c, this->n, y, nrm);
    } else {
      k_gemv_n<T>(c, this->m, this->n, A, lda, x, coef, prev, +1, y, nrm);
    }
  }
  double algorithmic_bytes(bool) const override { return (double)this->m * this->n * sizeof(T); }
  Pending scratch_p{};
};

// Row-sharded dense operator (BASELINE config 3 on N GPUs): this rank holds the rows [m_off, m_off+m) of A
// (column-major, all ng columns).  V-vectors are block-sharded like everywhere else, so
//   A x   : all-gather the n-vector (ng*w bytes -- tiny next to the m_loc x ng block), then the local GEMV;
//   A^H u : local GEMV^T gives this rank's contribution to all ng coefficients, all-reduce of those ng values
//           (the north star's "allreduce of length-k coefficients" pattern), then every rank keeps its own slice.
// The products themselves are the reorthogonalisation GEMV kernels over A, as on one GPU (DenseOperator).
template <class T> struct ShardedDenseOperator : LinOp<T> {
  using R = real_t<T>;
  DeviceBuffer<T> store, xfull, tfull;
  const T* A = nullptr; long lda = 0;
  void alloc_buffers() {
    xfull.alloc((size_t)this->ld_n * Comm::get().world);
    tfull.alloc((size_t)this->ld_n * Comm::get().world);
    PB_CUDA(cudaMemset(xfull.p, 0, sizeof(T) * xfull.n));
    PB_CUDA(cudaMemset(tfull.p, 0, sizeof(T) * tfull.n));
  }
  void apply(Context& c, bool adjoint, const T* x, T* y, R coef, const T* prev, Pending* nrm) override {
    Comm& cm = Comm::get();
    if (!adjoint) {
      // slices are ld_n long and contiguous in the global index, so the gathered buffer is the global vector
      cm.allgather(x, xfull.p, sizeof(T) * (size_t)this->ld_n, c.stream);
      k_gemv_n<T>(c, this->m, this->ng, A, lda, xfull.p, coef, prev, +1, y, nrm);
    } else {
      {  // local contribution only: the cross-rank sum is the explicit all-reduce below, whatever the engine's mode
        const bool old = c.dist_reduce; c.dist_reduce = false;
        k_gemv_t<T>(c, this->m, this->ng, A, lda, x, tfull.p);
        c.dist_reduce = old;
      }
      cm.allreduce_sum(reinterpret_cast<R*>(tfull.p), (size_t)this->ng * (scalar_traits<T>::is_complex ? 2 : 1), c.stream);
      if (this->n > 0)
        PB_CUDA(cudaMemcpyAsync(y, tfull.p + this->n_off, sizeof(T) * (size_t)this->n, cudaMemcpyDeviceToDevice, c.stream));
      if (prev) k_axpy_nrm<T>(c, this->n, T(coef), prev, y, nrm ? nrm : &scratch_p);
      else if (nrm) k_nrm2<T>(c, this->n, y, nrm);
    }
  }
  double algorithmic_bytes(bool) const override { return (double)this->m * this->ng * sizeof(T); }
  Pending scratch_p{};
};

// Generic user callback: host-staged (D2H x, call, H2D y).  Correct, PCIe-bound; the fast path is a
// built-in operator handle.
template <class T> struct CallbackOperator : LinOp<T> {
  using R = real_t<T>;
  aprod_f77_t<T> fn = nullptr; void* parm = nullptr; int* iparm = nullptr;
  std::vector<T> hx, hy;
  void apply(Context& c, bool adjoint, const T* x, T* y, R coef, const T* prev, Pending* nrm) override {
    const int nx = adjoint ? this->m : this->n, ny = adjoint ? this->n : this->m;
    hx.resize(nx); hy.resize(ny);
    PB_CUDA(cudaMemcpyAsync(hx.data(), x, sizeof(T) * nx, cudaMemcpyDeviceToHost, c.stream));
    PB_CUDA(cudaStreamSynchronize(c.stream));
    const char t = adjoint ? (scalar_traits<T>::is_complex ? 'c' : 't') : 'n';
    fn(&t, &this->m, &this->n, hx.data(), hy.data(), parm, iparm, 1);
    PB_CUDA(cudaMemcpyAsync(y, hy.data(), sizeof(T) * ny, cudaMemcpyHostToDevice, c.stream));
    PB_CUDA(cudaStreamSynchronize(c.stream));
    if (prev) k_axpy_nrm<T>(c, ny, T(coef), prev, y, nrm ? nrm : &scratch_p);
    else if (nrm) k_nrm2<T>(c, ny, y, nrm);
  }
  double algorithmic_bytes(bool) const override { return 0.0; }
  Pending scratch_p{};
};

// ------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------
template <class T> class Engine {
 public:
  using R = real_t<T>;
  Context& c;
  LinOp<T>* op;
  int m, n;           // local vector lengths (this rank's rows of U / V)
  int mg, ng;         // global problem size (what the reference's formulas see)
  bool dist;          // row-sharded run: reductions are completed across ranks
  long ldu, ldv;      // padded leading dimensions (multiples of 32 elements => 256-byte aligned columns)
  int ucols, vcols;   // allocated columns
  DeviceBuffer<T> wrk, hbuf;
  T* U; T* V;
  Context::BasisLayout ulay, vlay;

  static long pad_ld(long rows) { return (rows + 31) / 32 * 32; }

  Engine(Context& ctx, LinOp<T>* op_, int ucols_, int vcols_)
      : c(ctx), op(op_), m(op_->m), n(op_->n), mg(op_->global_m()), ng(op_->global_n()),
        dist(op_->sharded && Comm::get().active()), ldu(op_->ld_m > 0 ? op_->ld_m : pad_ld(op_->m)),
        ldv(op_->ld_n > 0 ? op_->ld_n : pad_ld(op_->n)), ucols(ucols_), vcols(vcols_) {
    wrk.alloc((size_t)std::max(ldu, ldv));
    hbuf.alloc((size_t)std::max(ucols, vcols) + 8);
    // padding rows must be (and stay) zero: kernels read whole 128-bit packs past the last row.  Recycled basis buffers
    // of the same layout already satisfy that (Context::basis_acquire).
    ulay.ld = ldu; ulay.rows = m; ulay.cols = ucols; ulay.elem = (int)sizeof(T);
    vlay.ld = ldv; vlay.rows = n; vlay.cols = vcols; vlay.elem = (int)sizeof(T);
    bool uz = false, vz = false;
    U = static_cast<T*>(c.basis_acquire(ulay, &uz));
    V = static_cast<T*>(c.basis_acquire(vlay, &vz));
    if (!uz) PB_CUDA(cudaMemsetAsync(U, 0, sizeof(T) * (size_t)ldu * ucols, c.stream));
    if (!vz) PB_CUDA(cudaMemsetAsync(V, 0, sizeof(T) * (size_t)ldv * vcols, c.stream));
    PB_CUDA(cudaMemsetAsync(wrk.p, 0, sizeof(T) * wrk.n, c.stream));
  }
  ~Engine() {
    cudaStreamSynchronize(c.stream);
    c.basis_release(U, ulay);
    c.basis_release(V, vlay);
  }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  // While alive, kernels launched through this engine publish cross-rank reductions (no-op on one GPU).
  struct DistScope {
    Context& c; bool old;
    DistScope(Context& c_, bool on) : c(c_), old(c_.dist_reduce) { c.dist_reduce = on; }
    ~DistScope() { c.dist_reduce = old; }
  };
  T* ucol(int j) { return U + (size_t)(j - 1) * ldu; }  // 1-based column
  T* vcol(int j) { return V + (size_t)(j - 1) * ldv; }

  // --- dsafescal (dsafescal.F:4-55) ---------------------------------------------------------------
  void safescal(long len, R alpha, T* x) {
    Context::PhaseScope ps(c, PH_LEVEL1);
    const R sfmin = host::Machine<R>::sfmin;
    if (std::fabs(alpha) >= sfmin) { k_scal<T>(c, len, x, R(1) / alpha); return; }
    // dlascl('General',..,cfrom=alpha,cto=1): multiply in safe steps (Lapack_Util/dlascl.f:113-137)
    const R smlnum = sfmin, bignum = R(1) / smlnum;
    R cfromc = alpha, ctoc = 1;
    for (bool done = false; !done;) {
      const R cfrom1 = cfromc * smlnum, cto1 = ctoc / bignum;
      R mul;
      if (std::fabs(cfrom1) > std::fabs(ctoc) && ctoc != R(0)) { mul = smlnum; cfromc = cfrom1; }
      else if (std::fabs(cto1) > std::fabs(cfromc)) { mul = bignum; ctoc = cto1; }
      else { mul = ctoc / cfromc; done = true; }
      k_scal<T>(c, len, x, mul);
    }
  }

  // x <- x/alpha where x is the vector the next apply(adjoint_next, x, ...) consumes: on a row-sharded operator with
  // mapped peers the scaling kernel also pushes the slice to every rank (fused all-gather); otherwise plain dsafescal.
  void normalize_for_apply(bool adjoint_next, long len, R alpha, T* x) {
    if (dist && std::fabs(alpha) >= host::Machine<R>::sfmin) {
      Context::PhaseScope ps(c, PH_LEVEL1);
      if (op->stage_scaled(c, adjoint_next, R(1) / alpha, x)) return;
    }
    safescal(len, alpha, x);
  }

  R nrm2(long len, const T* x) {
    Pending p; k_nrm2<T>(c, len, x, &p);
    return (R)c.wait(p);
  }

  // --- dreorth (dreorth.F:5-101) ----------------------------------------------------------------------
  // Iterated Gram-Schmidt of vnew against basis(:, intervals), DGKS test with factor `kappa`, at most
  // NTRY = 5 passes, else the vector is declared in span(basis) and zeroed.  Each interval is one GEMV
  // pair (dcgs, dreorth.F:106-210); intervals are swept in order, as the reference's block loop does.
  // iflag (CGS=1 / MGS=0) selects the same kernels: on this hardware column-sequential MGS
  // (dmgs.risc.F:58-79) would be l dependent grid-wide reductions, so it is realised as the blocked
  // GEMV pair with DGKS re-iteration, which meets the same orthogonality test.
  void reorth(long len, int k, const T* basis, long ld, T* vnew, R& normvnew, const host::IntervalList& idx, R kappa,
              int iflag) {
    (void)iflag;
    // (a rank that owns no rows of a sharded vector still takes part in every cross-rank reduction below)
    if (k <= 0 || (len <= 0 && !dist)) return;
    Context::PhaseScope ps(c, PH_REORTH);
    const int NTRY = 5;
    for (int itry = 0; itry < NTRY; ++itry) {
      const R norm0 = normvnew;
      Pending p;
      bool published = false;
      int count = 0;
      idx.for_each(k, [&](int, int) { ++count; });
      int seen = 0;
      idx.for_each(k, [&](int pcol, int qcol) {
        ++seen;
        const int l = qcol - pcol + 1;
        c.ctr.ndot += l;
        if (l <= 0) return;
        const T* blk = basis + (size_t)(pcol - 1) * ld;
        k_gemv_t<T>(c, len, l, blk, ld, vnew, hbuf.p);
        const bool last = (seen == count);
        k_gemv_n<T>(c, len, l, blk, ld, hbuf.p, R(1), vnew, -1, vnew, last ? &p : nullptr);
        if (last) published = true;
        c.ctr.reorth_cols += l;
        c.ctr.reorth_elems += (long long)l * len;
        c.ctr.reorth_vec_elems += len;
      });
      c.ctr.reorth_passes += 1;
      c.ctr.ndot += k;
      if (!published) k_nrm2<T>(c, len, vnew, &p);
      normvnew = (R)c.wait(p);
      if (normvnew > kappa * norm0) { c.ctr.nreorth += 1; return; }
    }
    normvnew = 0;
    k_zero<T>(c, len, vnew);
    c.ctr.nreorth += 1;
  }
  // (row-sharded run: the all-reduce of the l coefficients h = V_local^H q_local -- the OpenMP build's CRITICAL sum over
  // threads, dreorth.F:177-197 -- happens inside k_gemv_t: fused into its finalize kernel over NVLink peer memory, or NCCL.)

  // --- dgetu0 (dgetu0.F:11-89) ---------------------------------------------------------------------------
  // u0 <- op(A) r, r ~ LAPACK uniform(-1,1) stream from iseed (1,3,5,7) (reset on every call), then
  // orthogonalised against basis(:,1:j).  adjoint=false: u0 is an m-vector.
  void getu0(bool adjoint, int j, int ntry, T* u0, R& u0norm, const T* basis, long ld, int& ierr, int icgs, R& anormest) {
    Context::PhaseScope ps(c, PH_GETU0);
    const R kappa = R(0.717f);  // single-precision literal in dgetu0.F:28-29
    const long rsize = adjoint ? m : n, usize = adjoint ? n : m;          // local lengths
    const long rglobal = adjoint ? mg : ng, roff = adjoint ? op->m_off : op->n_off;
    int iseed[4] = {1, 3, 5, 7};
    ierr = 0;
    for (int itry = 0; itry < ntry; ++itry) {
      // stream position: itry-th block of rsize values (dlarnv keeps advancing iseed between tries)
      Pending pr, pu;
      advance_and_draw(iseed, rsize, rglobal, roff, wrk.p, &pr);
      const R nrm = (R)c.wait(pr);
      op->apply(c, adjoint, wrk.p, u0, R(0), nullptr, &pu);
      c.ctr.nopx += 1;
      u0norm = (R)c.wait(pu);
      anormest = u0norm / nrm;
      if (j >= 1) {
        host::IntervalList idx(4);
        idx.set_single(1, j, j + 1);
        reorth(usize, j, basis, ld, u0, u0norm, idx, kappa, icgs);
      }
      if (u0norm > 0) return;
    }
    ierr = -1;
  }
  // draw this rank's `len` values (global positions off .. off+len-1) of the stream that starts at the current seed
  // and advance the seed past all `glen` values, like xLARNV
  void advance_and_draw(int iseed[4], long len, long glen, long off, T* x, Pending* p) {
    k_larnv_nrm<T>(c, len, x, iseed, p, off);
    const unsigned long long M = (1ull << 48) - 1;
    unsigned long long s = ((unsigned long long)iseed[0] << 36) | ((unsigned long long)iseed[1] << 24) |
                           ((unsigned long long)iseed[2] << 12) | (unsigned long long)iseed[3];
    unsigned long long e = (unsigned long long)glen * (scalar_traits<T>::is_complex ? 2 : 1), base = 33952834046453ull, r = 1;
    while (e) { if (e & 1) r = (r * base) & M; base = (base * base) & M; e >>= 1; }
    s = (s * r) & M;
    iseed[0] = int((s >> 36) & 4095); iseed[1] = int((s >> 24) & 4095); iseed[2] = int((s >> 12) & 4095); iseed[3] = int(s & 4095);
  }

  // --- dlanbpro (dlanbpro.F:1-549) ----------------------------------------------------------------------
  // a = B(:,1) (alpha), b = B(:,2) (beta): host arrays of length >= k.  Returns ierr.
  int lanbpro(int k0, int& k, R* a, R* b, R& rnorm, R* doption, const int* ioption);

  // --- dritzvec (dritzvec.F:1-199) -------------------------------------------------------------------------
  void ritzvec(bool smallest, bool jobu, bool jobv, int k, int dim, R* D, R* E, bool reference_route = false);

  // --- drivers ------------------------------------------------------------------------------------------
  // dlansvd (dlansvd.F:1-291); U(:,1) on device holds the start vector (zero => random).  Returns info.
  int lansvd(bool jobu, bool jobv, int& k, int kmax, R* sigma, R* bnd, R tolin, R* doption, const int* ioption);
  // dlansvd_irl (dlansvd_irl.F:1-419)
  int lansvd_irl(bool smallest, bool jobu, bool jobv, int& dim, int p, int& neig, int maxiter, R* sigma, R* bnd, R tolin,
                 R* doption, const int* ioption);
};

}  // namespace pb

#include "engine_impl.hpp"
