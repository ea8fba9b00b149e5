// propack_b200 -- the C-ABI (include/propack_b200.h): Fortran-ABI drivers of the reference library,
// device-resident operator handles, solver sessions, counters.  Host buffers are staged here; the
// numerical work is in engine.hpp / the .cu kernels.  Nothing in this file computes on the CPU
// except the O(k^2) bidiagonal algebra the reference also keeps on the host.
#include <cstdarg>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>

#include "../../include/propack_b200.h"
#include "engine.hpp"

using namespace pb;

extern "C" {
struct pb200_timing_common timing_;
}

namespace {

thread_local std::string g_last_error;   // per calling thread, like errno
void set_error(const std::string& s) { g_last_error = s; }

int error_code(const std::exception& e) {
  set_error(e.what());
  if (auto* ce = dynamic_cast<const CudaError*>(&e)) return -100 - (int)ce->code;
  return -99;
}
#define PB_API_TRY try {
#define PB_API_CATCH(ret_stmt)             \
  }                                        \
  catch (const std::exception& e) {        \
    int code__ = error_code(e);            \
    (void)code__;                          \
    ret_stmt;                              \
  }

template <class T> struct abi;  // maps the ABI struct types onto the kernel scalar types
template <> struct abi<float> { using type = float; static constexpr char tag = 's'; };
template <> struct abi<double> { using type = double; static constexpr char tag = 'd'; };
template <> struct abi<cplx<float>> { using type = pb200_complex8; static constexpr char tag = 'c'; };
template <> struct abi<cplx<double>> { using type = pb200_complex16; static constexpr char tag = 'z'; };

struct OpEntry {
  char tag = 0;
  int kind = 0;  // 0 csr, 1 dense, 2 row-sharded csr, 3 row-sharded dense
  std::shared_ptr<void> op;
};
struct SolverEntry {
  char tag = 0;
  int op_handle = 0;
  std::shared_ptr<void> op;       // keeps the operator alive while the session exists (op_destroy only drops the handle)
  std::shared_ptr<void> engine;
};
// Handle tables.  The library is not re-entrant (like the reference: COMMON /timing/, SAVEd variables), but handle
// creation / destruction may come from different host threads, so the maps themselves are guarded.
std::mutex g_mu;
std::map<int, OpEntry> g_ops;
std::map<int, SolverEntry> g_solvers;
int g_next_op = 1, g_next_solver = 1;

int register_op(const OpEntry& e) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int h = g_next_op++;
  g_ops[h] = e;
  return h;
}
OpEntry find_op(int handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_ops.find(handle);
  if (it == g_ops.end()) throw std::runtime_error("propack_b200: unknown operator handle " + std::to_string(handle));
  return it->second;
}
template <class T> std::shared_ptr<void> lookup_op_shared(int handle) {
  OpEntry e = find_op(handle);
  if (e.tag != abi<T>::tag) throw std::runtime_error("propack_b200: operator handle has a different precision");
  return e.op;
}
template <class T> LinOp<T>* lookup_op(int handle) { return static_cast<LinOp<T>*>(lookup_op_shared<T>(handle).get()); }

bool is_yes(const char* c) { return c && (*c == 'y' || *c == 'Y'); }

// ---------------------------------------------------------------------------------------------------
// CSR registration: upload, device-independent host analysis of row lengths, transpose.
// ---------------------------------------------------------------------------------------------------
template <class T> int csr_create(int m, int n, const int* rowptr, const int* colind, const void* values_, int base) {
  PB_API_TRY
  Context& c = Context::get();
  const T* values = static_cast<const T*>(values_);
  if (m <= 0 || n <= 0 || !rowptr || !colind || !values) throw std::runtime_error("propack_b200: bad CSR arguments");
  // the three arrays may live on the host or on this device (e.g. a torch CSR tensor): cudaMemcpyDefault copies either
  int rp_ends[2] = {0, 0};
  PB_CUDA(cudaMemcpy(&rp_ends[0], rowptr, sizeof(int), cudaMemcpyDefault));
  PB_CUDA(cudaMemcpy(&rp_ends[1], rowptr + m, sizeof(int), cudaMemcpyDefault));
  const long nnz = (long)rp_ends[1] - base;
  if (rp_ends[0] != base || nnz < 0) throw std::runtime_error("propack_b200: CSR row pointers do not start at the index base");
  auto op = std::make_shared<CsrOperator<T>>();
  op->m = m; op->n = n;
  op->rp.alloc(m + 1); op->ci.alloc(std::max<long>(nnz, 1)); op->va.alloc(std::max<long>(nnz, 1));
  op->trp.alloc(n + 1); op->tci.alloc(std::max<long>(nnz, 1)); op->tva.alloc(std::max<long>(nnz, 1));
  // the three host arrays go up as they are; re-basing, validation (row pointers, column range, sorted rows), the
  // canonical transpose and the SELL copies are all built on the device (csr_build.cu, sell.cu)
  PB_CUDA(cudaMemcpyAsync(op->rp.p, rowptr, sizeof(int) * (m + 1), cudaMemcpyDefault, c.stream));
  if (nnz) {
    PB_CUDA(cudaMemcpyAsync(op->ci.p, colind, sizeof(int) * nnz, cudaMemcpyDefault, c.stream));
    PB_CUDA(cudaMemcpyAsync(op->va.p, values, sizeof(T) * nnz, cudaMemcpyDefault, c.stream));
  }
  k_rebase(c, m + 1, op->rp.p, base);
  k_rebase(c, nnz, op->ci.p, base);
  if (k_csr_check_rowptr(c, m, nnz, op->rp.p)) throw std::runtime_error("propack_b200: CSR row pointers must be non-decreasing");
  const int st = k_csr_transpose<T>(c, m, n, nnz, op->rp.p, op->ci.p, op->va.p, op->trp.p, op->tci.p, op->tva.p);
  if (st & 2) throw std::runtime_error("propack_b200: CSR column index out of range");
  if (st & 1) throw std::runtime_error("propack_b200: CSR column indices must be sorted within each row");
  // column panels whose block of the gathered vector stays L2-resident (one panel unless the vector is tens of MB)
  const int Ga = column_blocks<T>(n), Gt = column_blocks<T>(m);
  const long lda_ = shard_slice(n, Ga), ldt_ = shard_slice(m, Gt);
  op->A.build(c, m, n, nnz, op->rp.p, op->ci.p, op->va.p, lda_, Ga, 0, Ga, kSellCtasPerSm);
  op->At.build(c, n, m, nnz, op->trp.p, op->tci.p, op->tva.p, ldt_, Gt, 0, Gt, kSellCtasPerSm);
  OpEntry e; e.tag = abi<T>::tag; e.kind = 0; e.op = op;
  return register_op(e);
  PB_API_CATCH(return code__)
}

// Row-sharded CSR: `row_*` = CSR of this rank's row block A[r0:r1, :] (global column ids), `colt_*` = CSR
// of the transpose of this rank's column block, (A[:, c0:c1])^T, i.e. (c1-c0) rows x mg columns (global row ids) --
// scipy: A[r0:r1].tocsr() and A[:, c0:c1].tocsc().  Bounds follow shard_bounds(); values are not conjugated.
template <class T>
int csr_create_sharded(int mg, int ng, const int* row_rp, const int* row_ci, const void* row_va_, const int* colt_rp,
                       const int* colt_ci, const void* colt_va_, int base) {
  PB_API_TRY
  Context& c = Context::get();
  Comm& cm = Comm::get();
  const T* row_va = static_cast<const T*>(row_va_);
  const T* colt_va = static_cast<const T*>(colt_va_);
  if (mg <= 0 || ng <= 0 || !row_rp || !colt_rp) throw std::runtime_error("propack_b200: bad sharded CSR arguments");
  long r0, r1, c0, c1;
  shard_bounds(mg, cm.world, cm.rank, r0, r1);
  shard_bounds(ng, cm.world, cm.rank, c0, c1);
  const int ml = (int)(r1 - r0), nl = (int)(c1 - c0);
  auto op = std::make_shared<ShardedCsrOperator<T>>();
  op->m = ml; op->n = nl; op->mg = mg; op->ng = ng; op->m_off = r0; op->n_off = c0;
  op->ld_m = shard_slice(mg, cm.world); op->ld_n = shard_slice(ng, cm.world);
  op->sharded = true;
  // phases (sub-products gated on the arrival of their source slices) per product: must divide world
  int G = std::min(cm.world, 4);
  if (const char* e = std::getenv("PROPACK_B200_SPMV_PHASES")) G = std::max(1, std::atoi(e));
  G = std::min(std::min(G, cm.world), 8);
  while (cm.world % G) --G;
  // The shard goes up as it is; validation, the split into phases and the SELL copies are built on the device.
  auto upload = [&](int d, int rows, long width, long ld, const int* rp_in, const int* ci_in, const T* va_in) {
    if (rp_in[0] != base) throw std::runtime_error("propack_b200: sharded CSR row pointers must start at the index base");
    const long nnz = (long)rp_in[rows] - base;
    if (nnz < 0) throw std::runtime_error("propack_b200: bad sharded CSR row pointers");
    DeviceBuffer<int> rp((size_t)rows + 1), ci((size_t)std::max<long>(nnz, 1));
    DeviceBuffer<T> va((size_t)std::max<long>(nnz, 1));
    PB_CUDA(cudaMemcpyAsync(rp.p, rp_in, sizeof(int) * ((size_t)rows + 1), cudaMemcpyHostToDevice, c.stream));
    if (nnz) {
      PB_CUDA(cudaMemcpyAsync(ci.p, ci_in, sizeof(int) * nnz, cudaMemcpyHostToDevice, c.stream));
      PB_CUDA(cudaMemcpyAsync(va.p, va_in, sizeof(T) * nnz, cudaMemcpyHostToDevice, c.stream));
    }
    k_rebase(c, (long)rows + 1, rp.p, base);
    k_rebase(c, nnz, ci.p, base);
    if (k_csr_check_rowptr(c, rows, nnz, rp.p)) throw std::runtime_error("propack_b200: sharded CSR row pointers must be non-decreasing");
    // one CTA slot per SM stays free, so that the NVLink push kernel of the side stream can always become resident while
    // the SpMV CTAs spin on arrival flags
    // (the copy-engine transport needs no SM: the SpMV may then take every slot)
    const bool reserve = !op->push_by_copy_engine();
    const int st = op->sets[d].build(c, rows, ld * cm.world, nnz, rp.p, ci.p, va.p, ld, cm.world, cm.rank, G,
                                     reserve ? /*all resident slots but one*/ -1 : /*all resident slots*/ 0);
    if (st & 2) throw std::runtime_error("propack_b200: sharded CSR index out of range");
    if (st & 1) throw std::runtime_error("propack_b200: CSR indices must be sorted within each row");
    (void)width;
  };
  upload(0, ml, ng, op->ld_n, row_rp, row_ci, row_va);
  upload(1, nl, mg, op->ld_m, colt_rp, colt_ci, colt_va);
  op->alloc_gather_buffers();
  OpEntry e; e.tag = abi<T>::tag; e.kind = 2; e.op = op;
  return register_op(e);
  PB_API_CATCH(return code__)
}

// Row-sharded dense operator (BASELINE config 3 on N GPUs): `A_rows` = this rank's row block A[r0:r1, :]
// (column-major, leading dimension lda >= r1-r0; host or device memory), bounds from shard_bounds().
template <class T> std::shared_ptr<ShardedDenseOperator<T>> sharded_dense_shell(int mg, int ng) {
  Comm& cm = Comm::get();
  if (mg <= 0 || ng <= 0) throw std::runtime_error("propack_b200: bad sharded dense arguments");
  long r0, r1, c0, c1;
  shard_bounds(mg, cm.world, cm.rank, r0, r1);
  shard_bounds(ng, cm.world, cm.rank, c0, c1);
  auto op = std::make_shared<ShardedDenseOperator<T>>();
  op->m = (int)(r1 - r0); op->n = (int)(c1 - c0); op->mg = mg; op->ng = ng; op->m_off = r0; op->n_off = c0;
  op->ld_m = shard_slice(mg, cm.world); op->ld_n = shard_slice(ng, cm.world);
  op->sharded = true;
  const long ld = Engine<T>::pad_ld(std::max(op->m, 1));
  op->store.alloc((size_t)ld * ng);
  PB_CUDA(cudaMemsetAsync(op->store.p, 0, sizeof(T) * (size_t)ld * ng, Context::get().stream));   // zero padding rows
  op->A = op->store.p; op->lda = ld;
  op->alloc_buffers();
  return op;
}
template <class T> int dense_create_sharded(int mg, int ng, const void* A_rows, long lda) {
  PB_API_TRY
  Context& c = Context::get();
  auto op = sharded_dense_shell<T>(mg, ng);
  if (op->m > 0) {
    if (!A_rows || lda < op->m) throw std::runtime_error("propack_b200: bad sharded dense row block");
    PB_CUDA(cudaMemcpy2DAsync(op->store.p, sizeof(T) * op->lda, A_rows, sizeof(T) * lda, sizeof(T) * op->m, ng, cudaMemcpyDefault, c.stream));
  }
  c.sync();
  OpEntry e; e.tag = abi<T>::tag; e.kind = 3; e.op = op;
  return register_op(e);
  PB_API_CATCH(return code__)
}

template <class T> int dense_create(int m, int n, const void* A_, long lda, bool adopt_device) {
  PB_API_TRY
  Context& c = Context::get();
  auto op = std::make_shared<DenseOperator<T>>();
  op->m = m; op->n = n;
  if (m <= 0 || n <= 0 || !A_ || lda < m) throw std::runtime_error("propack_b200: bad dense operator arguments");
  if (adopt_device) {
    // the GEMV kernels read whole 128-bit packs: columns must be 16-byte aligned and the rows m..lda-1 of every column
    // (the padding up to the pack boundary) must be zero -- the caller owns the buffer, so only the layout is checked
    constexpr int VEC = 16 / (int)sizeof(T);
    if (lda % VEC != 0 || (reinterpret_cast<uintptr_t>(A_) & 15u) != 0 || (m % VEC != 0 && lda < (m + VEC - 1) / VEC * VEC))
      throw std::runtime_error("propack_b200: adopted dense operator needs a 16-byte aligned base, lda a multiple of the 128-bit pack "
                               "and zeroed padding rows up to the pack boundary");
    op->A = static_cast<const T*>(A_); op->lda = lda;
  } else {
    const long ld = Engine<T>::pad_ld(m);
    op->store.alloc((size_t)ld * n);
    PB_CUDA(cudaMemsetAsync(op->store.p, 0, sizeof(T) * (size_t)ld * n, c.stream));
    PB_CUDA(cudaMemcpy2DAsync(op->store.p, sizeof(T) * ld, A_, sizeof(T) * lda, sizeof(T) * m, n, cudaMemcpyHostToDevice, c.stream));
    c.sync();
    op->A = op->store.p; op->lda = ld;
  }
  OpEntry e; e.tag = abi<T>::tag; e.kind = 1; e.op = op;
  return register_op(e);
  PB_API_CATCH(return code__)
}

// ---------------------------------------------------------------------------------------------------
// operator resolution for the Fortran-ABI entry points
// ---------------------------------------------------------------------------------------------------
template <class T> void* builtin_aprod();
template <> void* builtin_aprod<float>() { return (void*)&propack_b200_aprod_s_; }
template <> void* builtin_aprod<double>() { return (void*)&propack_b200_aprod_d_; }
template <> void* builtin_aprod<cplx<float>>() { return (void*)&propack_b200_aprod_c_; }
template <> void* builtin_aprod<cplx<double>>() { return (void*)&propack_b200_aprod_z_; }

template <class T> struct ResolvedOp {
  LinOp<T>* op = nullptr;
  std::shared_ptr<void> keep;
  CallbackOperator<T> cb;
  ResolvedOp(void* aprod, int m, int n, void* parm, int* iparm) {
    if (aprod == builtin_aprod<T>()) {
      if (!iparm) throw std::runtime_error("propack_b200: built-in APROD needs the operator handle in iparm(1)");
      keep = lookup_op_shared<T>(iparm[0]);
      op = static_cast<LinOp<T>*>(keep.get());
      if (op->m != m || op->n != n) throw std::runtime_error("propack_b200: m,n do not match the registered operator");
    } else {
      if (!aprod) throw std::runtime_error("propack_b200: APROD is null");
      cb.m = m; cb.n = n; cb.fn = (aprod_f77_t<T>)aprod; cb.parm = parm; cb.iparm = iparm;
      op = &cb;
    }
  }
};

void publish_timing(const Context& c) {  // keep the reference's COMMON /timing/ counters live (stat.h:7-15)
  timing_.nopx = (int)c.ctr.nopx; timing_.nreorth = (int)c.ctr.nreorth; timing_.ndot = (int)c.ctr.ndot;
  timing_.nitref = (int)c.ctr.nitref; timing_.nrestart = (int)c.ctr.nrestart; timing_.nbsvd = (int)c.ctr.nbsvd;
  timing_.nlandim = (int)c.ctr.nlandim; timing_.nsing = (int)c.ctr.nsing;
  timing_.tmvopx = (float)(c.ctr.phase_ms[PH_APROD] * 1e-3); timing_.treorth = (float)(c.ctr.phase_ms[PH_REORTH] * 1e-3);
  timing_.tgetu0 = (float)(c.ctr.phase_ms[PH_GETU0] * 1e-3); timing_.tritzvec = (float)(c.ctr.phase_ms[PH_RITZ] * 1e-3);
  timing_.trestart = (float)(c.ctr.phase_ms[PH_RESTART] * 1e-3); timing_.tbsvd = (float)(c.ctr.phase_ms[PH_HOST_BSVD] * 1e-3);
}

template <class T> void upload_cols(Context& c, T* dst, long ldd, const void* src, long lds, long rows, int cols) {
  if (cols <= 0 || rows <= 0) return;
  PB_CUDA(cudaMemcpy2DAsync(dst, sizeof(T) * ldd, src, sizeof(T) * lds, sizeof(T) * rows, cols, cudaMemcpyHostToDevice, c.stream));
}
template <class T> void download_cols(Context& c, void* dst, long ldd, const T* src, long lds, long rows, int cols) {
  if (cols <= 0 || rows <= 0) return;
  PB_CUDA(cudaMemcpy2DAsync(dst, sizeof(T) * ldd, src, sizeof(T) * lds, sizeof(T) * rows, cols, cudaMemcpyDeviceToHost, c.stream));
}

// ---------------------------------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------------------------------
template <class T>
void lansvd_entry(const char* jobu, const char* jobv, int m, int n, int* k, int kmax, void* aprod, void* U, int ldu,
                  real_t<T>* sigma, real_t<T>* bnd, void* V, int ldv, real_t<T> tolin, real_t<T>* option, int* ioption, int* info,
                  void* parm, int* iparm) {
  PB_API_TRY
  Context& c = Context::get();
  ResolvedOp<T> ro(aprod, m, n, parm, iparm);
  const int lanmax = std::min(std::min(n + 1, m + 1), kmax);
  Engine<T> eng(c, ro.op, lanmax + 1, std::max(lanmax, 1));
  upload_cols<T>(c, eng.U, eng.ldu, U, ldu, m, 1);
  const bool ju = is_yes(jobu), jv = is_yes(jobv);
  *info = eng.lansvd(ju, jv, *k, kmax, sigma, bnd, tolin, option, ioption);
  if (ju) download_cols<T>(c, U, ldu, eng.U, eng.ldu, m, *k);
  if (jv) download_cols<T>(c, V, ldv, eng.V, eng.ldv, n, *k);
  c.sync();
  publish_timing(c);
  PB_API_CATCH(*info = code__)
}

template <class T>
void lansvd_irl_entry(const char* which, const char* jobu, const char* jobv, int m, int n, int* dim, int p, int* neig, int maxiter,
                      void* aprod, void* U, int ldu, real_t<T>* sigma, real_t<T>* bnd, void* V, int ldv, real_t<T> tolin,
                      real_t<T>* option, int* ioption, int* info, void* parm, int* iparm) {
  PB_API_TRY
  Context& c = Context::get();
  ResolvedOp<T> ro(aprod, m, n, parm, iparm);
  const int d = std::min(*dim, std::min(n + 1, m + 1));
  Engine<T> eng(c, ro.op, d + 1, std::max(d, 1));
  upload_cols<T>(c, eng.U, eng.ldu, U, ldu, m, 1);
  const bool ju = is_yes(jobu), jv = is_yes(jobv);
  const bool smallest = which && (*which == 's' || *which == 'S');
  *info = eng.lansvd_irl(smallest, ju, jv, *dim, p, *neig, maxiter, sigma, bnd, tolin, option, ioption);
  if (ju) download_cols<T>(c, U, ldu, eng.U, eng.ldu, m, *neig);
  if (jv) download_cols<T>(c, V, ldv, eng.V, eng.ldv, n, *neig);
  c.sync();
  publish_timing(c);
  PB_API_CATCH(*info = code__)
}

template <class T>
void lanbpro_entry(int m, int n, int k0, int* k, void* aprod, void* U, int ldu, void* V, int ldv, real_t<T>* B, int ldb,
                   real_t<T>* rnorm, real_t<T>* option, int* ioption, void* parm, int* iparm, int* ierr) {
  PB_API_TRY
  Context& c = Context::get();
  ResolvedOp<T> ro(aprod, m, n, parm, iparm);
  const int kin = *k;
  Engine<T> eng(c, ro.op, kin + 1, std::max(kin, 1));
  upload_cols<T>(c, eng.U, eng.ldu, U, ldu, m, k0 + 1);
  upload_cols<T>(c, eng.V, eng.ldv, V, ldv, n, k0);
  *ierr = eng.lanbpro(k0, *k, B, B + ldb, *rnorm, option, ioption);
  download_cols<T>(c, U, ldu, eng.U, eng.ldu, m, kin + 1);
  download_cols<T>(c, V, ldv, eng.V, eng.ldv, n, kin);
  c.sync();
  publish_timing(c);
  PB_API_CATCH(*ierr = code__)
}

template <class T> struct NullOp : LinOp<T> {
  void apply(Context&, bool, const T*, T*, real_t<T>, const T*, Pending*) override {
    throw std::runtime_error("propack_b200: internal: NullOp applied");
  }
  double algorithmic_bytes(bool) const override { return 0; }
};

template <class T>
void reorth_entry(int n, int k, const void* V, int ldv, void* vnew, real_t<T>* normvnew, const int* index, real_t<T> alpha, int iflag) {
  PB_API_TRY
  if (k <= 0 || n <= 0) return;
  Context& c = Context::get();
  NullOp<T> nop; nop.m = n; nop.n = 1;
  Engine<T> eng(c, &nop, k + 1, 1);
  upload_cols<T>(c, eng.U, eng.ldu, V, ldv, n, k);
  upload_cols<T>(c, eng.ucol(k + 1), eng.ldu, vnew, n, n, 1);
  host::IntervalList idx(2 * k + 4);
  for (int i = 0; i < 2 * k + 3; ++i) {
    idx.v[i] = index[i];
    if ((i % 2 == 0) && (index[i] > k || index[i] <= 0)) break;
  }
  eng.reorth(n, k, eng.U, eng.ldu, eng.ucol(k + 1), *normvnew, idx, alpha, iflag);
  download_cols<T>(c, vnew, n, eng.ucol(k + 1), eng.ldu, n, 1);
  c.sync();
  publish_timing(c);
  PB_API_CATCH(*normvnew = real_t<T>(-1))
}

template <class T>
void getu0_entry(const char* transa, int m, int n, int j, int ntry, void* u0, real_t<T>* u0norm, const void* Ub, int ldu, void* aprod,
                 void* parm, int* iparm, int* ierr, int icgs, real_t<T>* anormest) {
  PB_API_TRY
  Context& c = Context::get();
  ResolvedOp<T> ro(aprod, m, n, parm, iparm);
  const bool adjoint = !(transa && (*transa == 'n' || *transa == 'N'));
  Engine<T> eng(c, ro.op, adjoint ? 1 : j + 1, adjoint ? j + 1 : 1);
  T* basis = adjoint ? eng.V : eng.U;
  const long ld = adjoint ? eng.ldv : eng.ldu;
  const long rows = adjoint ? n : m;
  upload_cols<T>(c, basis, ld, Ub, ldu, rows, j);
  T* out = basis + (size_t)j * ld;
  eng.getu0(adjoint, j, ntry, out, *u0norm, basis, ld, *ierr, icgs, *anormest);
  download_cols<T>(c, u0, rows, out, ld, rows, 1);
  c.sync();
  publish_timing(c);
  PB_API_CATCH(*ierr = code__)
}

template <class T> void safescal_entry(int n, real_t<T> alpha, void* x) {
  PB_API_TRY
  if (n <= 0) return;
  Context& c = Context::get();
  NullOp<T> nop; nop.m = n; nop.n = 1;
  Engine<T> eng(c, &nop, 1, 1);
  upload_cols<T>(c, eng.U, eng.ldu, x, n, n, 1);
  eng.safescal(n, alpha, eng.U);
  download_cols<T>(c, x, n, eng.U, eng.ldu, n, 1);
  c.sync();
  PB_API_CATCH(return )
}

// A(m x k) <- alpha * A * op(B): result m x n (dgemm_ovwr.F:56-87)
template <class T>
void gemm_ovwr_left_entry(const char* transb, int m, int n, int k, real_t<T> alpha, void* A, int lda, const real_t<T>* B, int ldb) {
  using R = real_t<T>;
  PB_API_TRY
  if (m <= 0 || n <= 0 || k <= 0) return;
  Context& c = Context::get();
  const bool tr = transb && (*transb == 't' || *transb == 'T');
  std::vector<R> W((size_t)k * n);  // W(l,jn) = alpha * op(B)(l,jn)
  for (int jn = 0; jn < n; ++jn)
    for (int l = 0; l < k; ++l) W[(size_t)jn * k + l] = alpha * (tr ? B[(size_t)l * ldb + jn] : B[(size_t)jn * ldb + l]);
  NullOp<T> nop; nop.m = m; nop.n = 1;
  Engine<T> eng(c, &nop, std::max(k, n), 1);
  upload_cols<T>(c, eng.U, eng.ldu, A, lda, m, k);
  k_gemm_tall<T>(c, m, n, k, eng.U, eng.ldu, W.data());
  download_cols<T>(c, A, lda, eng.U, eng.ldu, m, n);
  c.sync();
  PB_API_CATCH(return )
}

// dritzvec (dritzvec.F:1-199; complex: zritzvec.F:1-2): U(:,1:k) <- U(:,1:dim+1) X^T, V(:,1:k) <- V(:,1:dim) Q with the
// small factors from dbdqr + dbdsdc of the bidiagonal (D, E) -- the reference route, so D returns the singular values
// of B in descending order and E is destroyed, as in the Fortran.
template <class T>
void ritzvec_entry(const char* which, const char* jobu, const char* jobv, int m, int n, int k, int dim, real_t<T>* D, real_t<T>* E,
                   real_t<T>* S, void* U, int ldu, void* V, int ldv) {
  PB_API_TRY
  if (k <= 0 || dim <= 0 || m <= 0 || n <= 0) return;
  Context& c = Context::get();
  NullOp<T> nop; nop.m = m; nop.n = n;
  Engine<T> eng(c, &nop, dim + 1, dim);
  const bool ju = is_yes(jobu), jv = is_yes(jobv);
  const bool smallest = which && (*which == 's' || *which == 'S');
  if (ju) upload_cols<T>(c, eng.U, eng.ldu, U, ldu, m, dim + 1);
  if (jv) upload_cols<T>(c, eng.V, eng.ldv, V, ldv, n, dim);
  eng.ritzvec(smallest, ju, jv, k, dim, D, E, /*reference_route=*/true);
  if (ju) download_cols<T>(c, U, ldu, eng.U, eng.ldu, m, k);
  if (jv) download_cols<T>(c, V, ldv, eng.V, eng.ldv, n, k);
  c.sync();
  (void)S;   // documented as output in dritzvec.F:26-28 but never written by the reference either; D carries the values
  publish_timing(c);
  PB_API_CATCH(fprintf(stderr, "%s\n", g_last_error.c_str()))
}

// dgemm_ovwr (dgemm_ovwr.F:5-53): B(m x n) <- alpha*op(A)*B + beta*B, op(A) m x k, B k x n on entry (ldb >= max(m,k)).
// The product runs through the tall-GEMM kernel on the transposed problem  B^T (n x k) * op(A)^T (k x m).
template <class R>
void gemm_ovwr_entry(const char* transa, int m, int n, int k, R alpha, const R* A, int lda, R beta, R* B, int ldb) {
  PB_API_TRY
  if (m <= 0 || n <= 0 || k <= 0) return;
  if (m > ldb) throw std::runtime_error("propack_b200: m>ldb in xGEMM_OVWR");
  Context& c = Context::get();
  const bool tr = transa && (*transa == 't' || *transa == 'T');
  std::vector<R> W((size_t)k * m);        // W(l, i) = alpha * op(A)(i, l)
  for (int i = 0; i < m; ++i)
    for (int l = 0; l < k; ++l) W[(size_t)i * k + l] = alpha * (tr ? A[(size_t)i * lda + l] : A[(size_t)l * lda + i]);
  std::vector<R> Bt((size_t)n * k);       // B^T, n x k column-major
  for (int l = 0; l < k; ++l)
    for (int j = 0; j < n; ++j) Bt[(size_t)l * n + j] = B[(size_t)j * ldb + l];
  NullOp<R> nop; nop.m = n; nop.n = 1;
  Engine<R> eng(c, &nop, std::max(k, m), 1);
  upload_cols<R>(c, eng.U, eng.ldu, Bt.data(), n, n, k);
  k_gemm_tall<R>(c, n, m, k, eng.U, eng.ldu, W.data());
  std::vector<R> Ct((size_t)n * m);
  download_cols<R>(c, Ct.data(), n, eng.U, eng.ldu, n, m);
  c.sync();
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      const R v = Ct[(size_t)i * n + j];
      B[(size_t)j * ldb + i] = beta == R(0) ? v : v + beta * B[(size_t)j * ldb + i];
    }
  PB_API_CATCH(fprintf(stderr, "%s\n", g_last_error.c_str()))
}

// ---- blasext level-1 (dblasext.F:6-255, zblasext.F): host vectors staged through the device kernels -----------------
template <class T> struct HostVec {   // strided host vector -> padded device vector (and back)
  DeviceBuffer<T> d;
  std::vector<T> packed;
  long n;
  HostVec(Context& c, long n_, const void* x, int incx) : n(n_) {
    d.alloc((size_t)Engine<T>::pad_ld(std::max<long>(n, 1)));
    PB_CUDA(cudaMemsetAsync(d.p, 0, sizeof(T) * d.n, c.stream));
    if (n <= 0 || !x) return;
    const T* xs = static_cast<const T*>(x);
    if (incx == 1) { PB_CUDA(cudaMemcpyAsync(d.p, xs, sizeof(T) * n, cudaMemcpyHostToDevice, c.stream)); return; }
    packed.resize(n);
    const long start = incx < 0 ? (long)(1 - n) * incx : 0;   // BLAS convention for negative increments
    for (long i = 0; i < n; ++i) packed[i] = xs[start + i * incx];
    PB_CUDA(cudaMemcpyAsync(d.p, packed.data(), sizeof(T) * n, cudaMemcpyHostToDevice, c.stream));
  }
  void store(Context& c, void* x, int incx) {
    if (n <= 0) return;
    T* xs = static_cast<T*>(x);
    if (incx == 1) { PB_CUDA(cudaMemcpyAsync(xs, d.p, sizeof(T) * n, cudaMemcpyDeviceToHost, c.stream)); c.sync(); return; }
    packed.resize(n);
    PB_CUDA(cudaMemcpyAsync(packed.data(), d.p, sizeof(T) * n, cudaMemcpyDeviceToHost, c.stream));
    c.sync();
    const long start = incx < 0 ? (long)(1 - n) * incx : 0;
    for (long i = 0; i < n; ++i) xs[start + i * incx] = packed[i];
  }
};
template <class T> real_t<T> l1_nrm2(int n, const void* x, int incx) {
  PB_API_TRY
  if (n <= 0) return 0;
  Context& c = Context::get();
  HostVec<T> hx(c, n, x, incx);
  Pending p; k_nrm2<T>(c, n, hx.d.p, &p);
  return (real_t<T>)c.wait(p);
  PB_API_CATCH(return real_t<T>(-1))
}
template <class T> void l1_dot(int n, const void* x, int incx, const void* y, int incy, bool conj, double* re, double* im) {
  *re = *im = 0;
  PB_API_TRY
  if (n <= 0) return;
  Context& c = Context::get();
  HostVec<T> hx(c, n, x, incx), hy(c, n, y, incy);
  Pending p; k_dotc<T>(c, n, hx.d.p, hy.d.p, &p);
  *re = c.wait(p, im);
  if (!conj && scalar_traits<T>::is_complex) {   // x.y = conj(conj(x).conj(y)): unconjugated product via the conjugated vector
    std::vector<T> xc(n);
    const T* xs = static_cast<const T*>(x);
    const long start = incx < 0 ? (long)(1 - n) * incx : 0;
    for (long i = 0; i < n; ++i) xc[i] = conj_(xs[start + i * incx]);
    HostVec<T> hc(c, n, xc.data(), 1);
    Pending q; k_dotc<T>(c, n, hc.d.p, hy.d.p, &q);
    *re = c.wait(q, im);
  }
  PB_API_CATCH(return )
}
template <class T> void l1_axpy(int n, T alpha, const void* x, int incx, void* y, int incy) {
  PB_API_TRY
  if (n <= 0) return;
  Context& c = Context::get();
  HostVec<T> hx(c, n, x, incx), hy(c, n, y, incy);
  Pending p; k_axpy_nrm<T>(c, n, alpha, hx.d.p, hy.d.p, &p);
  c.wait(p);
  hy.store(c, y, incy);
  PB_API_CATCH(return )
}
template <class T> void l1_scal(int n, real_t<T> alpha, void* x, int incx) {
  PB_API_TRY
  if (n <= 0) return;
  Context& c = Context::get();
  HostVec<T> hx(c, n, x, incx);
  k_scal<T>(c, n, hx.d.p, alpha);
  hx.store(c, x, incx);
  PB_API_CATCH(return )
}
template <class T> void l1_zero(int n, void* x, int incx) {
  PB_API_TRY
  if (n <= 0) return;
  Context& c = Context::get();
  HostVec<T> hx(c, n, x, incx);
  k_zero<T>(c, n, hx.d.p);
  hx.store(c, x, incx);
  PB_API_CATCH(return )
}

template <class T> void aprod_entry(const char* transa, int m, int n, const void* x, void* y, int* iparm) {
  PB_API_TRY
  Context& c = Context::get();
  std::shared_ptr<void> keep = lookup_op_shared<T>(iparm[0]);
  LinOp<T>* op = static_cast<LinOp<T>*>(keep.get());
  const bool adjoint = !(transa && (*transa == 'n' || *transa == 'N'));
  const long nx = adjoint ? m : n, ny = adjoint ? n : m;
  DeviceBuffer<T> dx(Engine<T>::pad_ld(nx)), dy(Engine<T>::pad_ld(ny));
  PB_CUDA(cudaMemsetAsync(dx.p, 0, sizeof(T) * dx.n, c.stream));
  PB_CUDA(cudaMemsetAsync(dy.p, 0, sizeof(T) * dy.n, c.stream));
  PB_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(T) * nx, cudaMemcpyHostToDevice, c.stream));
  op->apply(c, adjoint, dx.p, dy.p, real_t<T>(0), nullptr, nullptr);
  PB_CUDA(cudaMemcpyAsync(y, dy.p, sizeof(T) * ny, cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  PB_API_CATCH(fprintf(stderr, "%s\n", g_last_error.c_str()))
}

// ---------------------------------------------------------------------------------------------------
// solver sessions
// ---------------------------------------------------------------------------------------------------
template <class T> int solver_create_t(int op_handle, int ucols, int vcols) {
  Context& c = Context::get();
  std::shared_ptr<void> keep = lookup_op_shared<T>(op_handle);
  auto eng = std::make_shared<Engine<T>>(c, static_cast<LinOp<T>*>(keep.get()), ucols, vcols);
  SolverEntry e; e.tag = abi<T>::tag; e.op_handle = op_handle; e.op = keep; e.engine = eng;
  c.sync();
  std::lock_guard<std::mutex> lk(g_mu);
  const int id = g_next_solver++;
  g_solvers[id] = e;
  return id;
}
SolverEntry find_solver(int id) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_solvers.find(id);
  if (it == g_solvers.end()) throw std::runtime_error("propack_b200: unknown solver id");
  return it->second;
}
template <class F> auto dispatch(char tag, F f) {
  switch (tag) {
    case 's': return f((float*)nullptr);
    case 'd': return f((double*)nullptr);
    case 'c': return f((cplx<float>*)nullptr);
    default: return f((cplx<double>*)nullptr);
  }
}

}  // namespace

// =====================================================================================================
// extern "C"
// =====================================================================================================
extern "C" {

#define PB_DRIVERS_REAL(P, T, R, APT)                                                                                        \
  void P##lansvd_(const char* jobu, const char* jobv, const int* m, const int* n, int* k, const int* kmax, APT aprod, R* U,    \
                  const int* ldu, R* sigma, R* bnd, R* V, const int* ldv, const R* tolin, R* work, const int* lwork, int* iwork, \
                  const int* liwork, R* option, int* ioption, int* info, R* parm, int* iparm, size_t, size_t) {                  \
    (void)work; (void)lwork; (void)iwork; (void)liwork;                                                                          \
    lansvd_entry<T>(jobu, jobv, *m, *n, k, *kmax, (void*)aprod, U, *ldu, sigma, bnd, V, *ldv, *tolin, option, ioption, info,    \
                    parm, iparm);                                                                                                \
  }                                                                                                                              \
  void P##lansvd_irl_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, int* dim, const int* p, \
                      int* neig, const int* maxiter, APT aprod, R* U, const int* ldu, R* sigma, R* bnd, R* V, const int* ldv,    \
                      const R* tolin, R* work, const int* lwork, int* iwork, const int* liwork, R* option, int* ioption,         \
                      int* info, R* parm, int* iparm, size_t, size_t, size_t) {                                                  \
    (void)work; (void)lwork; (void)iwork; (void)liwork;                                                                          \
    lansvd_irl_entry<T>(which, jobu, jobv, *m, *n, dim, *p, neig, *maxiter, (void*)aprod, U, *ldu, sigma, bnd, V, *ldv, *tolin,  \
                        option, ioption, info, parm, iparm);                                                                     \
  }                                                                                                                              \
  void P##lanbpro_(const int* m, const int* n, const int* k0, int* k, APT aprod, R* U, const int* ldu, R* V, const int* ldv,     \
                   R* B, const int* ldb, R* rnorm, R* option, int* ioption, R* work, int* iwork, R* parm, int* iparm,            \
                   int* ierr) {                                                                                                  \
    (void)work; (void)iwork;                                                                                                     \
    lanbpro_entry<T>(*m, *n, *k0, k, (void*)aprod, U, *ldu, V, *ldv, B, *ldb, rnorm, option, ioption, parm, iparm, ierr);        \
  }                                                                                                                              \
  void P##gemm_ovwr_left_(const char* transb, const int* m, const int* n, const int* k, const R* alpha, R* A, const int* lda,    \
                          const R* beta, const R* B, const int* ldb, R* work, const int* lwork, size_t) {                        \
    (void)beta; (void)work; (void)lwork;                                                                                         \
    gemm_ovwr_left_entry<T>(transb, *m, *n, *k, *alpha, A, *lda, B, *ldb);                                                       \
  }

#define PB_DRIVERS_CPLX(P, T, R, CT, APT, GEMMNAME)                                                                             \
  void P##lansvd_(const char* jobu, const char* jobv, const int* m, const int* n, int* k, const int* kmax, APT aprod, CT* U,     \
                  const int* ldu, R* sigma, R* bnd, CT* V, const int* ldv, const R* tolin, R* work, const int* lwork, CT* zwork, \
                  const int* lzwrk, int* iwork, const int* liwork, R* option, int* ioption, int* info, CT* parm, int* iparm,     \
                  size_t, size_t) {                                                                                              \
    (void)work; (void)lwork; (void)zwork; (void)lzwrk; (void)iwork; (void)liwork;                                                \
    lansvd_entry<T>(jobu, jobv, *m, *n, k, *kmax, (void*)aprod, U, *ldu, sigma, bnd, V, *ldv, *tolin, option, ioption, info,    \
                    parm, iparm);                                                                                                \
  }                                                                                                                              \
  void P##lansvd_irl_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, int* dim, const int* p, \
                      int* neig, const int* maxiter, APT aprod, CT* U, const int* ldu, R* sigma, R* bnd, CT* V, const int* ldv,  \
                      const R* tolin, R* work, const int* lwork, CT* zwork, const int* lzwrk, int* iwork, const int* liwork,     \
                      R* option, int* ioption, int* info, CT* parm, int* iparm, size_t, size_t, size_t) {                        \
    (void)work; (void)lwork; (void)zwork; (void)lzwrk; (void)iwork; (void)liwork;                                                \
    lansvd_irl_entry<T>(which, jobu, jobv, *m, *n, dim, *p, neig, *maxiter, (void*)aprod, U, *ldu, sigma, bnd, V, *ldv, *tolin,  \
                        option, ioption, info, parm, iparm);                                                                     \
  }                                                                                                                              \
  void P##lanbpro_(const int* m, const int* n, const int* k0, int* k, APT aprod, CT* U, const int* ldu, CT* V, const int* ldv,   \
                   R* B, const int* ldb, R* rnorm, R* option, int* ioption, R* dwork, CT* zwork, int* iwork, CT* parm,           \
                   int* iparm, int* ierr) {                                                                                      \
    (void)dwork; (void)zwork; (void)iwork;                                                                                       \
    lanbpro_entry<T>(*m, *n, *k0, k, (void*)aprod, U, *ldu, V, *ldv, B, *ldb, rnorm, option, ioption, parm, iparm, ierr);        \
  }                                                                                                                              \
  void GEMMNAME(const char* transb, const int* m, const int* n, const int* k, CT* A, const int* lda, const R* B, const int* ldb, \
                CT* zwork, const int* lzwork, size_t) {                                                                          \
    (void)zwork; (void)lzwork;                                                                                                   \
    gemm_ovwr_left_entry<T>(transb, *m, *n, *k, R(1), A, *lda, B, *ldb);                                                         \
  }

#define PB_COMMON(P, T, R, CT, APT)                                                                                             \
  void P##reorth_(const int* n, const int* k, const CT* V, const int* ldv, CT* vnew, R* normvnew, const int* index,              \
                  const R* alpha, CT* work, const int* iflag) {                                                                  \
    (void)work;                                                                                                                  \
    reorth_entry<T>(*n, *k, V, *ldv, vnew, normvnew, index, *alpha, *iflag);                                                     \
  }                                                                                                                              \
  void P##getu0_(const char* transa, const int* m, const int* n, const int* j, const int* ntry, CT* u0, R* u0norm, const CT* U,  \
                 const int* ldu, APT aprod, CT* parm, int* iparm, int* ierr, const int* icgs, R* anormest, CT* work, size_t) {   \
    (void)work;                                                                                                                  \
    getu0_entry<T>(transa, *m, *n, *j, *ntry, u0, u0norm, U, *ldu, (void*)aprod, parm, iparm, ierr, *icgs, anormest);            \
  }                                                                                                                              \
  void P##safescal_(const int* n, const R* alpha, CT* x) { safescal_entry<T>(*n, *alpha, x); }                                   \
  void propack_b200_aprod_##P##_(const char* transa, const int* m, const int* n, const CT* x, CT* y, CT* parm, int* iparm,       \
                                 size_t) {                                                                                       \
    (void)parm;                                                                                                                  \
    aprod_entry<T>(transa, *m, *n, x, y, iparm);                                                                                 \
  }                                                                                                                              \
  int propack_b200_csr_create_##P(int m, int n, const int* rowptr, const int* colind, const CT* values, int index_base) {        \
    return csr_create<T>(m, n, rowptr, colind, values, index_base);                                                              \
  }                                                                                                                              \
  int propack_b200_dense_create_##P(int m, int n, const CT* A, long lda) { return dense_create<T>(m, n, A, lda, false); } \
  int propack_b200_csr_create_sharded_##P(int m_global, int n_global, const int* row_rowptr, const int* row_colind,             \
                                          const CT* row_values, const int* colt_rowptr, const int* colt_rowind,                  \
                                          const CT* colt_values, int index_base) {                                               \
    return csr_create_sharded<T>(m_global, n_global, row_rowptr, row_colind, row_values, colt_rowptr, colt_rowind, colt_values,  \
                                 index_base);                                                                                    \
  }

PB_DRIVERS_REAL(s, float, float, pb200_aprod_s_t)
PB_DRIVERS_REAL(d, double, double, pb200_aprod_d_t)
PB_DRIVERS_CPLX(c, cplx<float>, float, pb200_complex8, pb200_aprod_c_t, csgemm_ovwr_left_)
PB_DRIVERS_CPLX(z, cplx<double>, double, pb200_complex16, pb200_aprod_z_t, zdgemm_ovwr_left_)
PB_COMMON(s, float, float, float, pb200_aprod_s_t)
PB_COMMON(d, double, double, double, pb200_aprod_d_t)
PB_COMMON(c, cplx<float>, float, pb200_complex8, pb200_aprod_c_t)
PB_COMMON(z, cplx<double>, double, pb200_complex16, pb200_aprod_z_t)

// ---- xRITZVEC (dritzvec.F:1-2; zritzvec.F:1-2 adds zwork, lzwrk) and xGEMM_OVWR (dgemm_ovwr.F:5) ---------------------------
#define PB_RITZ_REAL(P, T, R)                                                                                                   \
  void P##ritzvec_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, const int* k,            \
                   const int* dim, R* D, R* E, R* S, R* U, const int* ldu, R* V, const int* ldv, R* work, const int* in_lwrk,  \
                   int* iwork, size_t, size_t, size_t) {                                                                        \
    (void)work; (void)in_lwrk; (void)iwork;                                                                                     \
    ritzvec_entry<T>(which, jobu, jobv, *m, *n, *k, *dim, D, E, S, U, *ldu, V, *ldv);                                           \
  }                                                                                                                             \
  void P##gemm_ovwr_(const char* transa, const int* m, const int* n, const int* k, const R* alpha, const R* A, const int* lda, \
                     const R* beta, R* B, const int* ldb, R* dwork, const int* ldwork, size_t) {                                \
    (void)dwork; (void)ldwork;                                                                                                  \
    gemm_ovwr_entry<R>(transa, *m, *n, *k, *alpha, A, *lda, *beta, B, *ldb);                                                    \
  }
#define PB_RITZ_CPLX(P, T, R, CT)                                                                                               \
  void P##ritzvec_(const char* which, const char* jobu, const char* jobv, const int* m, const int* n, const int* k,            \
                   const int* dim, R* D, R* E, R* S, CT* U, const int* ldu, CT* V, const int* ldv, R* work, const int* in_lwrk, \
                   CT* zwork, const int* lzwrk, int* iwork, size_t, size_t, size_t) {                                           \
    (void)work; (void)in_lwrk; (void)zwork; (void)lzwrk; (void)iwork;                                                           \
    ritzvec_entry<T>(which, jobu, jobv, *m, *n, *k, *dim, D, E, S, U, *ldu, V, *ldv);                                           \
  }
PB_RITZ_REAL(s, float, float)
PB_RITZ_REAL(d, double, double)
PB_RITZ_CPLX(c, cplx<float>, float, pb200_complex8)
PB_RITZ_CPLX(z, cplx<double>, double, pb200_complex16)

// ---- blasext level-1 (dblasext.F:6,38,92,121,202; zblasext.F:6,60,113,167,344; s/c likewise) --------------------------------
float psnrm2_(const int* n, const float* x, const int* incx) { return l1_nrm2<float>(*n, x, *incx); }
double pdnrm2_(const int* n, const double* x, const int* incx) { return l1_nrm2<double>(*n, x, *incx); }
float pscnrm2_(const int* n, const pb200_complex8* x, const int* incx) { return l1_nrm2<cplx<float>>(*n, x, *incx); }
double pdznrm2_(const int* n, const pb200_complex16* x, const int* incx) { return l1_nrm2<cplx<double>>(*n, x, *incx); }
float psdot_(const int* n, const float* x, const int* incx, const float* y, const int* incy) {
  double re, im; l1_dot<float>(*n, x, *incx, y, *incy, true, &re, &im); return (float)re;
}
double pddot_(const int* n, const double* x, const int* incx, const double* y, const int* incy) {
  double re, im; l1_dot<double>(*n, x, *incx, y, *incy, true, &re, &im); return re;
}
pb200_complex8 pcdotc_(const int* n, const pb200_complex8* x, const int* incx, const pb200_complex8* y, const int* incy) {
  double re, im; l1_dot<cplx<float>>(*n, x, *incx, y, *incy, true, &re, &im); return pb200_complex8{(float)re, (float)im};
}
pb200_complex16 pzdotc_(const int* n, const pb200_complex16* x, const int* incx, const pb200_complex16* y, const int* incy) {
  double re, im; l1_dot<cplx<double>>(*n, x, *incx, y, *incy, true, &re, &im); return pb200_complex16{re, im};
}
pb200_complex8 pcdotu_(const int* n, const pb200_complex8* x, const int* incx, const pb200_complex8* y, const int* incy) {
  double re, im; l1_dot<cplx<float>>(*n, x, *incx, y, *incy, false, &re, &im); return pb200_complex8{(float)re, (float)im};
}
pb200_complex16 pzdotu_(const int* n, const pb200_complex16* x, const int* incx, const pb200_complex16* y, const int* incy) {
  double re, im; l1_dot<cplx<double>>(*n, x, *incx, y, *incy, false, &re, &im); return pb200_complex16{re, im};
}
void psaxpy_(const int* n, const float* a, const float* x, const int* incx, float* y, const int* incy) { l1_axpy<float>(*n, *a, x, *incx, y, *incy); }
void pdaxpy_(const int* n, const double* a, const double* x, const int* incx, double* y, const int* incy) { l1_axpy<double>(*n, *a, x, *incx, y, *incy); }
void pcaxpy_(const int* n, const pb200_complex8* a, const pb200_complex8* x, const int* incx, pb200_complex8* y, const int* incy) {
  l1_axpy<cplx<float>>(*n, cplx<float>(a->re, a->im), x, *incx, y, *incy);
}
void pzaxpy_(const int* n, const pb200_complex16* a, const pb200_complex16* x, const int* incx, pb200_complex16* y, const int* incy) {
  l1_axpy<cplx<double>>(*n, cplx<double>(a->re, a->im), x, *incx, y, *incy);
}
void pcsaxpy_(const int* n, const float* a, const pb200_complex8* x, const int* incx, pb200_complex8* y, const int* incy) {
  l1_axpy<cplx<float>>(*n, cplx<float>(*a, 0.f), x, *incx, y, *incy);
}
void pzdaxpy_(const int* n, const double* a, const pb200_complex16* x, const int* incx, pb200_complex16* y, const int* incy) {
  l1_axpy<cplx<double>>(*n, cplx<double>(*a, 0.0), x, *incx, y, *incy);
}
void psscal_(const int* n, const float* a, float* x, const int* incx) { l1_scal<float>(*n, *a, x, *incx); }
void pdscal_(const int* n, const double* a, double* x, const int* incx) { l1_scal<double>(*n, *a, x, *incx); }
void pcsscal_(const int* n, const float* a, pb200_complex8* x, const int* incx) { l1_scal<cplx<float>>(*n, *a, x, *incx); }
void pzdscal_(const int* n, const double* a, pb200_complex16* x, const int* incx) { l1_scal<cplx<double>>(*n, *a, x, *incx); }
void pszero_(const int* n, float* x, const int* incx) { l1_zero<float>(*n, x, *incx); }
void pdzero_(const int* n, double* x, const int* incx) { l1_zero<double>(*n, x, *incx); }
void pczero_(const int* n, pb200_complex8* x, const int* incx) { l1_zero<cplx<float>>(*n, x, *incx); }
void pzzero_(const int* n, pb200_complex16* x, const int* incx) { l1_zero<cplx<double>>(*n, x, *incx); }

// ---- host bidiagonal algebra (real only, as in the reference) -------------------------------------------
#define PB_HOST_ALG(P, R)                                                                                                       \
  void P##bsvdstep_(const char* jobu, const char* jobv, const int* m, const int* n, const int* k, const R* sigma, R* D, R* E,    \
                    R* U, const int* ldu, R* V, const int* ldv, size_t, size_t) {                                                \
    host::bidiag_shift_sweep<R>(*m, *n, *k, *sigma, D, E, is_yes(jobu) ? U : nullptr, *ldu, is_yes(jobv) ? V : nullptr, *ldv);   \
  }                                                                                                                              \
  void P##bdqr_(const int* ignorelast, const char* jobq, const int* n, R* D, R* E, R* c1, R* c2, R* Qt, const int* ldq,          \
                size_t) {                                                                                                        \
    host::bidiag_qr<R>(*ignorelast != 0, is_yes(jobq), *n, D, E, *c1, *c2, Qt, *ldq);                                            \
  }                                                                                                                              \
  void P##refinebounds_(const int* n, const int* k, const R* theta, R* bound, const R* tol, const R* eps34) {                    \
    host::refine_bounds<R>(*n, *k, theta, bound, *tol, *eps34);                                                                  \
  }                                                                                                                              \
  void P##set_mu_(const int* k, R* mu, const int* index, const R* val) {                                                         \
    for (int i = 0; index[i] <= *k && index[i] > 0; i += 2)                                                                      \
      for (int t = index[i]; t <= index[i + 1]; ++t) mu[t - 1] = *val;                                                           \
  }                                                                                                                              \
  void P##compute_int_(const R* mu, const int* j, const R* delta, const R* eta, int* index) {                                    \
    std::vector<R> w(mu, mu + *j);                                                                                               \
    host::IntervalList idx(2 * *j + 4);                                                                                          \
    if (*delta < *eta) { host::select_intervals<R>(w, *j, *delta, *eta, idx); return; }                                          \
    host::select_intervals<R>(w, *j, *delta, *eta, idx);                                                                         \
    int i = 0;                                                                                                                   \
    for (;; i += 2) { index[i] = idx.v[i]; if (idx.v[i] > *j || idx.v[i] <= 0) break; index[i + 1] = idx.v[i + 1]; }             \
  }                                                                                                                              \
  void P##update_mu_(R* mumax, R* mu, const R* nu, const int* j, const R* alpha, const R* beta, const R* anorm,                  \
                     const R* eps1) {                                                                                            \
    host::OmegaRecurrence<R> om;                                                                                                 \
    om.mu.assign(mu, mu + *j + 1); om.nu.assign(nu, nu + *j + 1);                                                                \
    *mumax = om.update_mu(*j, alpha, beta, *anorm, *eps1);                                                                       \
    for (int i = 0; i <= *j; ++i) mu[i] = om.mu[i];                                                                              \
  }                                                                                                                              \
  void P##update_nu_(R* numax, const R* mu, R* nu, const int* j, const R* alpha, const R* beta, const R* anorm,                  \
                     const R* eps1) {                                                                                            \
    if (*j <= 1) return;                                                                                                         \
    host::OmegaRecurrence<R> om;                                                                                                 \
    om.mu.assign(mu, mu + *j + 1); om.nu.assign(nu, nu + *j + 1);                                                                \
    *numax = om.update_nu(*j, alpha, beta, *anorm, *eps1);                                                                       \
    for (int i = 0; i < *j; ++i) nu[i] = om.nu[i];                                                                               \
  }
PB_HOST_ALG(s, float)
PB_HOST_ALG(d, double)

void clearstat_(void) {
  std::memset(&timing_, 0, sizeof timing_);
  try { Context::get().ctr = Counters(); } catch (...) {}
}
void printstat_(void) {  // layout follows double/printstat.F:35-75
  printf(" +-----------------------------------------------------------+\n");
  printf(" Dimension of Lanczos basis                  = %12d\n", timing_.nlandim);
  printf(" Number of singular values requested         = %12d\n", timing_.nsing);
  printf(" Number of restarts                          = %12d\n", timing_.nrestart);
  printf(" Number of matrix-vector multiplications     = %12d\n", timing_.nopx);
  printf(" Number of reorthogonalizations              = %12d\n", timing_.nreorth);
  printf(" Number of inner products in reorth.         = %12d\n", timing_.ndot);
  printf(" Number of bidiagonal SVDs calculated        = %12d\n", timing_.nbsvd);
  printf("\n");
  printf("  Time spent doing matrix-vector multiply    = %12.4e\n", timing_.tmvopx);
  printf("  Time spent generating starting vectors     = %12.4e\n", timing_.tgetu0);
  printf("    Time spent reorthogonalizing             = %12.4e\n", timing_.treorth);
  printf("  Time spent calculating bidiagonal SVDs     = %12.4e\n", timing_.tbsvd);
  printf("  Time spent on restarts                     = %12.4e\n", timing_.trestart);
  printf("  Time spent calculating Ritz vectors        = %12.4e\n", timing_.tritzvec);
  printf(" +-----------------------------------------------------------+\n");
}

// ---- operators ------------------------------------------------------------------------------------------
int propack_b200_dense_adopt_device_s(int m, int n, const float* A_device, long lda) { return dense_create<float>(m, n, A_device, lda, true); }
int propack_b200_dense_adopt_device_d(int m, int n, const double* A_device, long lda) { return dense_create<double>(m, n, A_device, lda, true); }
int propack_b200_dense_adopt_device_c(int m, int n, const pb200_complex8* A_device, long lda) {
  return dense_create<cplx<float>>(m, n, A_device, lda, true);
}
int propack_b200_dense_adopt_device_z(int m, int n, const pb200_complex16* A_device, long lda) {
  return dense_create<cplx<double>>(m, n, A_device, lda, true);
}
int propack_b200_dense_create_synthetic_d(int m, int n, unsigned long long seed, const double* table16x256) {
  PB_API_TRY
  Context& c = Context::get();
  if (m <= 0 || n <= 0 || !table16x256) throw std::runtime_error("propack_b200: bad synthetic dense arguments");
  auto op = std::make_shared<DenseOperator<double>>();
  op->m = m; op->n = n;
  const long ld = Engine<double>::pad_ld(m);
  op->store.alloc((size_t)ld * n);
  if (ld != m) PB_CUDA(cudaMemsetAsync(op->store.p, 0, sizeof(double) * (size_t)ld * n, c.stream));  // zero padding rows
  k_dense_synth(c, m, n, ld, seed, table16x256, op->store.p);
  op->A = op->store.p; op->lda = ld;
  OpEntry e; e.tag = 'd'; e.kind = 1; e.op = op;
  return register_op(e);
  PB_API_CATCH(return code__)
}
int propack_b200_dense_create_sharded_s(int m_global, int n_global, const float* A_rows, long lda) {
  return dense_create_sharded<float>(m_global, n_global, A_rows, lda);
}
int propack_b200_dense_create_sharded_d(int m_global, int n_global, const double* A_rows, long lda) {
  return dense_create_sharded<double>(m_global, n_global, A_rows, lda);
}
int propack_b200_dense_create_sharded_c(int m_global, int n_global, const pb200_complex8* A_rows, long lda) {
  return dense_create_sharded<cplx<float>>(m_global, n_global, A_rows, lda);
}
int propack_b200_dense_create_sharded_z(int m_global, int n_global, const pb200_complex16* A_rows, long lda) {
  return dense_create_sharded<cplx<double>>(m_global, n_global, A_rows, lda);
}
// this rank's rows of the synthetic config-3 matrix, evaluated on the device (the same matrix as the 1-GPU generator)
int propack_b200_dense_create_synthetic_sharded_d(int m_global, int n_global, unsigned long long seed, const double* table16x256) {
  PB_API_TRY
  Context& c = Context::get();
  if (!table16x256) throw std::runtime_error("propack_b200: bad synthetic dense arguments");
  auto op = sharded_dense_shell<double>(m_global, n_global);
  k_dense_synth(c, op->m, n_global, op->lda, seed, table16x256, op->store.p, op->m_off);
  OpEntry e; e.tag = 'd'; e.kind = 3; e.op = op;
  return register_op(e);
  PB_API_CATCH(return code__)
}
// Drops the handle.  Solver sessions created on the operator keep it alive until they are destroyed themselves.
int propack_b200_op_destroy(int handle) {
  std::shared_ptr<void> last;   // the operator (device memory, peer windows) is released outside the lock
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ops.find(handle);
    if (it == g_ops.end()) return -1;
    last = it->second.op;
    g_ops.erase(it);
  }
  return 0;
}
double propack_b200_op_bytes(int handle, int adjoint) {
  try {
    OpEntry e = find_op(handle);
    return dispatch(e.tag, [&](auto* t) {
      using T = std::remove_pointer_t<decltype(t)>;
      return static_cast<LinOp<T>*>(e.op.get())->algorithmic_bytes(adjoint != 0);
    });
  } catch (...) { return -1.0; }
}
int propack_b200_csr_get_transpose(int handle, int* t_rowptr, int* t_colind, void* t_values) {
  PB_API_TRY
  OpEntry e = find_op(handle);
  if (e.kind != 0) throw std::runtime_error("propack_b200: not a CSR operator handle");
  return dispatch(e.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    auto* op = static_cast<CsrOperator<T>*>(e.op.get());
    PB_CUDA(cudaMemcpy(t_rowptr, op->trp.p, sizeof(int) * (op->n + 1), cudaMemcpyDeviceToHost));
    if (op->A.nnz) {
      PB_CUDA(cudaMemcpy(t_colind, op->tci.p, sizeof(int) * op->A.nnz, cudaMemcpyDeviceToHost));
      PB_CUDA(cudaMemcpy(t_values, op->tva.p, sizeof(T) * op->A.nnz, cudaMemcpyDeviceToHost));
    }
    return 0;
  });
  PB_API_CATCH(return code__)
}

// Sliced jagged-ELL copy of a registered CSR operator (adjoint = 1: of A^T; panel = column block): sizes, then the arrays
// (integer work, bit-exact against the numpy restatement in tests/sell_ref.py).
// info[0..3] = slices, stored entries, number of panels, long-row threshold.
int propack_b200_csr_sell_info(int handle, int adjoint, int panel, long long* info4) {
  PB_API_TRY
  OpEntry e = find_op(handle);
  if (e.kind != 0) throw std::runtime_error("propack_b200: not a CSR operator handle");
  return dispatch(e.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    auto* op = static_cast<CsrOperator<T>*>(e.op.get());
    PanelSet<T>& P = adjoint ? op->At : op->A;
    if (panel < 0 || panel >= (int)P.panels.size()) throw std::runtime_error("propack_b200: no such panel");
    const SellDevice<T>& S = P.panels[panel]->S.sell.dev;
    info4[0] = S.nslices; info4[1] = S.stored; info4[2] = (long long)P.panels.size(); info4[3] = kSellLong;
    return 0;
  });
  PB_API_CATCH(return code__)
}
int propack_b200_csr_get_sell(int handle, int adjoint, int panel, long long* slice_offsets, unsigned char* row_len, int* colind,
                              void* values) {
  PB_API_TRY
  OpEntry e = find_op(handle);
  if (e.kind != 0) throw std::runtime_error("propack_b200: not a CSR operator handle");
  return dispatch(e.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    auto* op = static_cast<CsrOperator<T>*>(e.op.get());
    PanelSet<T>& P = adjoint ? op->At : op->A;
    if (panel < 0 || panel >= (int)P.panels.size()) throw std::runtime_error("propack_b200: no such panel");
    const SellDevice<T>& S = P.panels[panel]->S.sell.dev;
    if (S.joff == nullptr) throw std::runtime_error("propack_b200: the operator has no jagged-ELL copy (PROPACK_B200_SPMV=csr)");
    PB_CUDA(cudaMemcpy(slice_offsets, S.joff, sizeof(long long) * (S.nslices + 1), cudaMemcpyDeviceToHost));
    if (S.rows) PB_CUDA(cudaMemcpy(row_len, S.len8, (size_t)S.rows, cudaMemcpyDeviceToHost));
    if (S.stored) {
      PB_CUDA(cudaMemcpy(colind, S.ci, sizeof(int) * S.stored, cudaMemcpyDeviceToHost));
      PB_CUDA(cudaMemcpy(values, S.va, sizeof(T) * S.stored, cudaMemcpyDeviceToHost));
    }
    return 0;
  });
  PB_API_CATCH(return code__)
}

// ---- solver sessions ----------------------------------------------------------------------------------------
int propack_b200_solver_create(int op_handle, int ucols, int vcols) {
  PB_API_TRY
  OpEntry e = find_op(op_handle);
  return dispatch(e.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    return solver_create_t<T>(op_handle, ucols, vcols);
  });
  PB_API_CATCH(return code__)
}
int propack_b200_solver_destroy(int solver) {
  SolverEntry last;   // engine buffers are freed outside the lock
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_solvers.find(solver);
    if (it == g_solvers.end()) return -1;
    last = it->second;
    g_solvers.erase(it);
  }
  return 0;
}
int propack_b200_solver_set_start(int solver, const void* u0_host) {
  PB_API_TRY
  SolverEntry s = find_solver(solver);
  return dispatch(s.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    auto* e = static_cast<Engine<T>*>(s.engine.get());
    if (u0_host) PB_CUDA(cudaMemcpyAsync(e->U, u0_host, sizeof(T) * e->m, cudaMemcpyHostToDevice, e->c.stream));
    else PB_CUDA(cudaMemsetAsync(e->U, 0, sizeof(T) * e->ldu, e->c.stream));
    e->c.sync();
    return 0;
  });
  PB_API_CATCH(return code__)
}
int propack_b200_solver_lansvd(int solver, int jobu, int jobv, int* k, int kmax, void* sigma, void* bnd, double tolin, void* option3,
                               int* ioption, int* info) {
  PB_API_TRY
  SolverEntry s = find_solver(solver);
  return dispatch(s.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    using R = real_t<T>;
    auto* e = static_cast<Engine<T>*>(s.engine.get());
    *info = e->lansvd(jobu != 0, jobv != 0, *k, kmax, (R*)sigma, (R*)bnd, (R)tolin, (R*)option3, ioption);
    e->c.sync();
    publish_timing(e->c);
    return 0;
  });
  PB_API_CATCH(*info = code__; return code__)
}
int propack_b200_solver_lansvd_irl(int solver, int which_smallest, int jobu, int jobv, int* dim, int p, int* neig, int maxiter,
                                   void* sigma, void* bnd, double tolin, void* option4, int* ioption, int* info) {
  PB_API_TRY
  SolverEntry s = find_solver(solver);
  return dispatch(s.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    using R = real_t<T>;
    auto* e = static_cast<Engine<T>*>(s.engine.get());
    *info = e->lansvd_irl(which_smallest != 0, jobu != 0, jobv != 0, *dim, p, *neig, maxiter, (R*)sigma, (R*)bnd, (R)tolin,
                          (R*)option4, ioption);
    e->c.sync();
    publish_timing(e->c);
    return 0;
  });
  PB_API_CATCH(*info = code__; return code__)
}
int propack_b200_solver_get_u(int solver, int ncols, void* U_host, long ldu) {
  PB_API_TRY
  SolverEntry s = find_solver(solver);
  return dispatch(s.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    auto* e = static_cast<Engine<T>*>(s.engine.get());
    download_cols<T>(e->c, U_host, ldu, e->U, e->ldu, e->m, std::min(ncols, e->ucols));
    e->c.sync();
    return 0;
  });
  PB_API_CATCH(return code__)
}
int propack_b200_solver_get_v(int solver, int ncols, void* V_host, long ldv) {
  PB_API_TRY
  SolverEntry s = find_solver(solver);
  return dispatch(s.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    auto* e = static_cast<Engine<T>*>(s.engine.get());
    download_cols<T>(e->c, V_host, ldv, e->V, e->ldv, e->n, std::min(ncols, e->vcols));
    e->c.sync();
    return 0;
  });
  PB_API_CATCH(return code__)
}

// ---- multi-GPU (one process per GPU) -------------------------------------------------------------------------------
int propack_b200_comm_unique_id(void* id128_out) {
  PB_API_TRY
  Comm::unique_id(id128_out);
  return 0;
  PB_API_CATCH(return code__)
}
int propack_b200_comm_init(int rank, int world, const void* id128) {
  PB_API_TRY
  Context& c = Context::get();  // binds the device first
  Comm& cm = Comm::get();
  cm.init(rank, world, id128);
  c.peer_table = cm.peer_ok ? cm.slots.table_dev : nullptr;
  c.coef_table = cm.peer_ok ? cm.coef.table_dev : nullptr;
  c.peer_rank = cm.rank; c.peer_world = cm.world;
  c.reset_peer_counters();   // collective call: every rank restarts the (slot, sequence) matching of the fused reductions
  return 0;
  PB_API_CATCH(return code__)
}
int propack_b200_comm_finalize(void) {
  PB_API_TRY
  try { Context& c = Context::get(); c.sync(); c.peer_table = nullptr; c.coef_table = nullptr; c.peer_rank = 0; c.peer_world = 1; } catch (...) {}
  Comm::get().finalize();
  return 0;
  PB_API_CATCH(return code__)
}
int propack_b200_comm_rank(void) { return Comm::get().rank; }
int propack_b200_comm_world(void) { return Comm::get().world; }
long propack_b200_shard_slice(long dim, int world) { return shard_slice(dim, world); }
void propack_b200_shard_bounds(long dim, int world, int rank, long* lo, long* hi) { shard_bounds(dim, world, rank, *lo, *hi); }
void propack_b200_comm_stats(long long* n_allreduce, long long* n_allgather, double* allgather_bytes) {
  Comm& cm = Comm::get();
  if (n_allreduce) *n_allreduce = cm.n_allreduce;
  if (n_allgather) *n_allgather = cm.n_allgather;
  if (allgather_bytes) *allgather_bytes = cm.allgather_bytes;
}
int propack_b200_solver_local_rows(int solver, int* m_local, int* n_local, long* ldu, long* ldv) {
  PB_API_TRY
  SolverEntry s = find_solver(solver);
  return dispatch(s.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    auto* e = static_cast<Engine<T>*>(s.engine.get());
    if (m_local) *m_local = e->m;
    if (n_local) *n_local = e->n;
    if (ldu) *ldu = e->ldu;
    if (ldv) *ldv = e->ldv;
    return 0;
  });
  PB_API_CATCH(return code__)
}

// Host-only test hook (no device needed): leading Ritz values / last-row components of the (j+1) x j lower bidiagonal by
// method 0 = the reference route (dbdqr + dbdsqr, dlansvd.F:193-199) or 1 = host::ritz_leading.  Returns 0, or 1 when
// method 1 declined (the driver then falls back to method 0).
int propack_b200_host_ritz_bounds_d(int j, const double* alpha, const double* beta, int K, int method, double* theta, double* last) {
  PB_API_TRY
  if (method == 1) return host::ritz_leading(j, alpha, beta, K, theta, last) ? 0 : 1;
  std::vector<double> th(alpha, alpha + j), ee(beta, beta + j), wb((size_t)j + 2, 0.0);
  int info = 0;
  host::bidiag_qr<double>(false, false, j, th.data(), ee.data(), wb[j - 1], wb[j], nullptr, 0);
  host::bdsqr_row(j, th.data(), ee.data(), wb.data(), &info);
  for (int i = 0; i < K; ++i) { theta[i] = th[i]; last[i] = std::fabs(wb[i]); }
  return info == 0 ? 0 : -1;
  PB_API_CATCH(return code__)
}

// Host-only test hook: the small matrices of dritzvec, WU ((dim+1) x k) and WV (dim x k), by the reference route
// (method 0: dbdqr + dbdsdc) or the fast route (method 1: host::ritz_vectors_leading).  Returns 1 when method 1 declines.
int propack_b200_host_ritz_vectors_d(int dim, const double* alpha, const double* beta, int k, int method, double* WU, double* WV) {
  PB_API_TRY
  std::vector<double> wu, wv;
  if (method == 1) {
    if (!host::ritz_vectors_leading(dim, alpha, beta, k, wu, wv)) return 1;
  } else {
    std::vector<double> D(alpha, alpha + dim), E(beta, beta + dim);
    ritz_w_reference<double>(false, false, true, true, k, dim, D.data(), E.data(), wu, wv);
  }
  std::copy(wu.begin(), wu.end(), WU);
  std::copy(wv.begin(), wv.end(), WV);
  return 0;
  PB_API_CATCH(return code__)
}

// Host-only test hook: the p = dim-k shifted QR sweeps of an implicit restart (dlansvd_irl.F:350-363) on the bidiagonal
// (alpha, beta), accumulating P ((dim+1)^2) and Q (dim^2) from the identity.  nthreads = 0: the reference's sequential
// dbsvdstep accumulation (rotation by rotation); nthreads >= 1: the recorded-rotation, row-parallel route the driver uses.
int propack_b200_host_restart_sweeps_d(int dim, int k, const double* shift, double* alpha, double* beta, double* P, double* Q, int nthreads) {
  PB_API_TRY
  std::fill(P, P + (size_t)(dim + 1) * (dim + 1), 0.0);
  std::fill(Q, Q + (size_t)dim * dim, 0.0);
  for (int i = 0; i <= dim; ++i) P[(size_t)i * (dim + 2)] = 1.0;
  for (int i = 0; i < dim; ++i) Q[(size_t)i * (dim + 1)] = 1.0;
  if (nthreads <= 0) {
    for (int i = dim; i >= k + 1; --i) host::bidiag_shift_sweep<double>(dim + 1, dim, i, shift[dim - i], alpha, beta, P, dim + 1, Q, dim);
  } else {
    host::restart_sweeps<double>(dim, k, shift, alpha, beta, P, Q, nthreads);
  }
  return 0;
  PB_API_CATCH(return code__)
}

// ---- runtime --------------------------------------------------------------------------------------------------
int propack_b200_init(void) {
  PB_API_TRY
  Context::get();
  host::bind_lapack();
  return 0;
  PB_API_CATCH(return code__)
}
int propack_b200_set_stream(void* s) {
  PB_API_TRY
  Context::get().set_stream((cudaStream_t)s);
  return 0;
  PB_API_CATCH(return code__)
}
int propack_b200_set_lapack(const char* path) {
  PB_API_TRY
  host::bind_lapack(path);
  return 0;
  PB_API_CATCH(return code__)
}
int propack_b200_set_option(const char* name, int value) {
  PB_API_TRY
  Context& c = Context::get();
  const std::string n = name ? name : "";
  if (n == "peer_timeout_s") { c.set_peer_timeout(value); return 0; }
  if (n == "bench_skip_gather") { bench_skip_gather() = value != 0; return 0; }   // measurement only, see engine.hpp
  throw std::runtime_error("propack_b200: unknown option '" + n + "'");
  PB_API_CATCH(return code__)
}
void propack_b200_release_cache(void) { try { Context& c = Context::get(); c.sync(); c.basis_cache_clear(); } catch (...) {} }
void propack_b200_set_profile(int on) { try { Context::get().profile = on != 0; } catch (...) {} }
void propack_b200_reset_counters(void) { try { Context::get().ctr = Counters(); } catch (...) {} }
void propack_b200_get_counters(long long* out) {
  std::memset(out, 0, sizeof(long long) * 16);
  try {
    const Counters& c = Context::get().ctr;
    long long v[16] = {c.nopx, c.nreorth, c.ndot, c.nitref, c.nrestart, c.nbsvd, c.nlandim, c.nsing, c.nsteps, c.reorth_passes,
                       c.reorth_cols, c.reorth_elems, c.reorth_vec_elems, c.launches, c.host_syncs, 0};
    std::memcpy(out, v, sizeof v);
  } catch (...) {}
}
void propack_b200_get_phase_ms(double* out_ms, long long* out_launches) {
  try {
    const Counters& c = Context::get().ctr;
    for (int i = 0; i < PH_COUNT; ++i) { if (out_ms) out_ms[i] = c.phase_ms[i]; if (out_launches) out_launches[i] = c.phase_launches[i]; }
  } catch (...) {}
}
const char* propack_b200_last_error(void) { return g_last_error.c_str(); }
int propack_b200_device_sms(void) { try { return Context::get().num_sms; } catch (...) { return -1; } }

}  // extern "C"

// ---- micro-benchmark hooks ------------------------------------------------------------------------------------
namespace {
struct L2Flusher {
  DeviceBuffer<char> buf;
  void flush(Context& c) {
    if (!buf.p) buf.alloc(256u << 20);
    PB_CUDA(cudaMemsetAsync(buf.p, 1, buf.n, c.stream));
  }
};
template <class F> double time_launches(Context& c, int reps, bool flush, F f) {
  static L2Flusher fl;
  cudaEvent_t e0, e1;
  PB_CUDA(cudaEventCreate(&e0)); PB_CUDA(cudaEventCreate(&e1));
  double total = 0;
  for (int r = -2; r < reps; ++r) {  // 2 warm-up launches
    if (flush) fl.flush(c);
    PB_CUDA(cudaEventRecord(e0, c.stream));
    f();
    PB_CUDA(cudaEventRecord(e1, c.stream));
    PB_CUDA(cudaEventSynchronize(e1));
    float ms = 0; PB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 0) total += ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return total / std::max(reps, 1);
}
}  // namespace

extern "C" {

double propack_b200_bench_reorth_d(long L, int l, int reps, int flush_l2) {
  PB_API_TRY
  Context& c = Context::get();
  const long ld = Engine<double>::pad_ld(L);
  DeviceBuffer<double> V((size_t)ld * l), q(ld), h(l + 8);
  PB_CUDA(cudaMemsetAsync(V.p, 0, sizeof(double) * V.n, c.stream));
  PB_CUDA(cudaMemsetAsync(q.p, 0, sizeof(double) * q.n, c.stream));
  int iseed[4] = {1, 3, 5, 7};
  Pending p;
  for (int j = 0; j < l; ++j) { iseed[3] = 2 * j + 1; k_larnv_nrm<double>(c, L, V.p + (size_t)j * ld, iseed, &p); }
  iseed[2] = 77; k_larnv_nrm<double>(c, L, q.p, iseed, &p);
  c.wait(p);
  k_scal<double>(c, (long)ld * l, V.p, 1.0 / std::sqrt((double)L));  // keep q bounded over repeated passes
  return time_launches(c, reps, flush_l2 != 0, [&] {
    Pending pn;
    k_gemv_t<double>(c, L, l, V.p, ld, q.p, h.p);
    k_gemv_n<double>(c, L, l, V.p, ld, h.p, 1.0, q.p, -1, q.p, &pn);
  });
  PB_API_CATCH(return (double)code__)
}
double propack_b200_bench_spmv(int op_handle, int adjoint, int reps, int flush_l2) {
  PB_API_TRY
  Context& c = Context::get();
  OpEntry e = find_op(op_handle);
  return dispatch(e.tag, [&](auto* t) {
    using T = std::remove_pointer_t<decltype(t)>;
    LinOp<T>* op = static_cast<LinOp<T>*>(e.op.get());
    const long nx = adjoint ? op->m : op->n, ny = adjoint ? op->n : op->m;
    DeviceBuffer<T> x(Engine<T>::pad_ld(nx)), y(Engine<T>::pad_ld(ny)), prev(Engine<T>::pad_ld(ny));
    PB_CUDA(cudaMemsetAsync(x.p, 0, sizeof(T) * x.n, c.stream));
    PB_CUDA(cudaMemsetAsync(y.p, 0, sizeof(T) * y.n, c.stream));
    PB_CUDA(cudaMemsetAsync(prev.p, 0, sizeof(T) * prev.n, c.stream));
    int iseed[4] = {1, 3, 5, 7};
    Pending p;
    k_larnv_nrm<T>(c, nx, x.p, iseed, &p);
    iseed[0] = 9; k_larnv_nrm<T>(c, ny, prev.p, iseed, &p);
    c.wait(p);
    return time_launches(c, reps, flush_l2 != 0, [&] {
      Pending pn;
      op->apply(c, adjoint != 0, x.p, y.p, real_t<T>(-0.5), prev.p, &pn);
    });
  });
  PB_API_CATCH(return (double)code__)
}
double propack_b200_bench_gemm_d(long M, int N, int K, int reps) {
  PB_API_TRY
  Context& c = Context::get();
  const long ld = Engine<double>::pad_ld(M);
  DeviceBuffer<double> A((size_t)ld * K);
  PB_CUDA(cudaMemsetAsync(A.p, 0, sizeof(double) * A.n, c.stream));
  int iseed[4] = {1, 3, 5, 7};
  Pending p;
  for (int j = 0; j < K; ++j) { iseed[3] = 2 * j + 1; k_larnv_nrm<double>(c, M, A.p + (size_t)j * ld, iseed, &p); }
  c.wait(p);
  std::vector<double> W((size_t)K * N);
  for (size_t i = 0; i < W.size(); ++i) W[i] = ((i * 2654435761u) % 1000) / 1000.0 / K;
  return time_launches(c, reps, false, [&] { k_gemm_tall<double>(c, M, N, K, A.p, ld, W.data()); });
  PB_API_CATCH(return (double)code__)
}

}  // extern "C"
