// propack_b200 -- per-process device context: stream, scalar slots, reduction workspace,
// phase timers.  One context per process (one process per GPU).
#pragma once
#include <chrono>
#include <vector>

#include "common.cuh"

namespace pb {

// Phases timed with CUDA events on the library stream (the reference's dead `tmvopx treorth
// tritzvec ...` timers of stat.h:9-15, made live).  Only filled when profiling is enabled.
enum Phase { PH_APROD = 0, PH_REORTH, PH_LEVEL1, PH_GETU0, PH_RITZ, PH_RESTART, PH_HOST_BSVD, PH_COUNT };

struct Counters {  // COMMON /timing/ integer counters (stat.h:7-8) + byte model inputs
  long long nopx = 0, nreorth = 0, ndot = 0, nitref = 0, nrestart = 0, nbsvd = 0, nlandim = 0, nsing = 0;
  long long nsteps = 0;        // Lanczos steps (dlanbpro.F:283 loop iterations)
  long long reorth_passes = 0; // GS passes executed
  long long reorth_cols = 0;   // sum over passes of columns swept
  long long reorth_elems = 0;  // sum over passes of L * l  (for the byte model  w*L*(2l+3))
  long long reorth_vec_elems = 0;  // sum over passes of L
  long long launches = 0;      // kernels launched by this library
  long long host_syncs = 0;    // scalar read-backs the host waited on
  double phase_ms[PH_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  long long phase_launches[PH_COUNT] = {0, 0, 0, 0, 0, 0, 0};
};

class Context {
 public:
  static Context& get();       // lazily created on the current device
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;       // side stream: the NVLink push of a row-sharded SpMV input runs here
  cudaEvent_t ev_fork = nullptr;        // main stream -> side stream dependency
  bool owns_stream = true;
  int device = 0;
  int num_sms = 148;
  Counters ctr;
  bool profile = false;        // per-phase CUDA-event timing (adds syncs; off for benchmarks)

  // --- scalar slots -------------------------------------------------------------------------
  static constexpr int kSlots = 64;   // == kPeerSlotsPerRank
  ScalarSlot* host_slots = nullptr;      // pinned + mapped
  ScalarSlot* host_slots_dev = nullptr;  // device view of the same memory
  ScalarSlot* dev_slots = nullptr;
  double* partials = nullptr;
  unsigned int* ticket = nullptr;
  unsigned int* tickets8 = nullptr;     // per-destination arrival counters of the staggered all-gather push
  unsigned long long seq = 0;
  int next_slot = 0;

  struct Pending { int slot; unsigned long long seq; int local_only = 0; };
  // Row-sharded runs (one process per GPU): while set, every reduction a kernel publishes is this rank's partial
  // and is completed across ranks (Comm all-reduce) before the host sees it.  Set by Engine for sharded operators.
  bool dist_reduce = false;
  // The fused cross-rank reductions match peers by (slot, sequence number), so the counters they use must advance in
  // lock step on every rank: they are separate from the single-GPU counters (a rank may run extra local work, e.g. a
  // validation solve on rank 0 only), use their own half of the slot table, and are reset by comm_init on all ranks.
  static constexpr int kLocalSlots = kSlots / 2;
  unsigned long long peer_seq = 0;
  int peer_next_slot = 0;
  long long peer_timeout_cycles = 60000000000LL;   // ~30 s at 1.9 GHz; PROPACK_B200_PEER_TIMEOUT_S / set_option("peer_timeout_s")
  void set_peer_timeout(int seconds) { peer_timeout_cycles = (long long)(seconds > 0 ? seconds : 1) * 2000000000LL; }
  void reset_peer_counters() { peer_seq = 0; peer_next_slot = 0; coef_seq = 0; }
  // Allocate a result slot for the next reducing kernel.
  ReduceWs new_reduce(Pending* p) {
    ReduceWs ws;
    ws.partials = partials; ws.ticket = ticket;
    ws.peer_table = nullptr; ws.rank = 0; ws.world = 1; ws.host_err = host_err_dev;
    ws.timeout_cycles = peer_timeout_cycles;
    ws.local_only = dist_reduce ? 1 : 0;
    int s;
    if (dist_reduce && peer_table) {   // fused path: the kernel itself completes the cross-rank reduction
      s = kLocalSlots + peer_next_slot; peer_next_slot = (peer_next_slot + 1) % (kSlots - kLocalSlots);
      ws.seq = ++peer_seq;
      ws.peer_table = peer_table; ws.rank = peer_rank; ws.world = peer_world; ws.local_only = 0;
    } else {
      s = next_slot; next_slot = (next_slot + 1) % kLocalSlots;
      ws.seq = ++seq;
    }
    ws.dev_slot = dev_slots + s; ws.host_slot = host_slots_dev + s; ws.slot = s;
    p->slot = s; p->seq = ws.seq; p->local_only = ws.local_only;
    return ws;
  }
  // set by Comm-aware code (c_api: comm_init) when peer windows are mapped
  void** peer_table = nullptr; int peer_rank = 0, peer_world = 1;
  void** coef_table = nullptr;           // Comm::coef window bases (device array), for the fused coefficient all-reduce
  unsigned long long coef_seq = 0;
  unsigned int* host_err = nullptr;      // pinned + mapped error word
  unsigned int* host_err_dev = nullptr;
  // Call right after launching the kernel that owns `p`.  kind: 0 = sum (dot products), 1 = sqrt(sum) (norms).
  // Single-GPU: nothing to do (the kernel's last CTA already published).
  void complete_reduce(const Pending& p, int kind);
  const ScalarSlot* dev_slot(const Pending& p) const { return dev_slots + p.slot; }
  // Spin until the kernel that owns `p` has published; returns the real part (imag via out param).
  double wait(const Pending& p, double* imag = nullptr);

  void set_stream(cudaStream_t s);
  void sync() { PB_CUDA(cudaStreamSynchronize(stream)); PB_CUDA(cudaStreamSynchronize(stream2)); }
  int grid_for(long work_items, int per_cta, int ctas_per_sm) const {
    long need = (work_items + per_cta - 1) / per_cta;
    long cap = (long)num_sms * ctas_per_sm;
    if (cap > kMaxCtas) cap = kMaxCtas;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
  }

  // --- phase timing ---------------------------------------------------------------------------
  struct PhaseScope {
    Context& c; Phase ph; cudaEvent_t e0 = nullptr, e1 = nullptr; long long l0;
    PhaseScope(Context& c_, Phase p);
    ~PhaseScope();
  };

  // scratch device buffer that grows on demand (coefficients, partials of gemv_t, ...)
  void* scratch(size_t bytes);

  // Basis-buffer cache.  The Fortran-ABI drivers own no state between calls, but allocating and zeroing the Krylov
  // bases (2 x 24 GB on BASELINE config 5) on every call is pure overhead for a caller that solves repeatedly: a
  // released basis buffer is parked here and handed back, WITHOUT a memset, to the next request with the same layout
  // (leading dimension, valid rows, column count, element size) -- every kernel keeps the padding rows zero and the drivers overwrite
  // each column before reading it, so a recycled buffer is as good as a zeroed one.  At most two buffers are parked (U
  // and V); a request with another layout frees them.  propack_b200_release_cache() empties it.
  struct BasisLayout {
    long ld = 0, rows = 0; int cols = 0; int elem = 0;   // rows: the valid rows -- everything below them must be zero
    bool operator==(const BasisLayout& o) const { return ld == o.ld && rows == o.rows && cols == o.cols && elem == o.elem; }
  };
  void* basis_acquire(const BasisLayout& lay, bool* zeroed);   // *zeroed = false: fresh allocation, the caller must clear it
  void basis_release(void* p, const BasisLayout& lay);
  void basis_cache_clear();

 private:
  Context();
  void* scratch_ = nullptr;
  size_t scratch_bytes_ = 0;
  struct Parked { void* p; BasisLayout lay; };
  std::vector<Parked> parked_;
};

template <class T> struct DeviceBuffer {
  T* p = nullptr; size_t n = 0;
  DeviceBuffer() {}
  explicit DeviceBuffer(size_t n_) { alloc(n_); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  void alloc(size_t n_) {
    free();
    n = n_;
    if (n) PB_CUDA(cudaMalloc((void**)&p, n * sizeof(T)));
  }
  void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DeviceBuffer() { free(); }
};

}  // namespace pb
