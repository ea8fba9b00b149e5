// propack_b200 -- inter-GPU communication (one process per GPU).
//
// The reference has no distributed layer (its parallelism is OpenMP row-chunking: dreorth.F:147-208,
// dritzvec.F:145-196); this is that SPMD scheme lifted across GPUs.  NCCL (over NVLink 5 / NVSwitch) is bound
// at run time with dlopen, so single-GPU users need no NCCL at all and the library has no link dependency.
// Collectives used on the hot path:
//   all-gather   the input vector of a sharded SpMV (n*w or m*w bytes per product)
//   all-reduce   the l reorthogonalisation coefficients after the local GEMV^T, and the scalar partials of
//                every norm / dot (2 doubles)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace pb {

class Comm {
 public:
  static Comm& get();
  int rank = 0, world = 1;
  bool active() const { return world > 1 && comm_ != nullptr; }
  static void unique_id(void* out128);                         // rank 0: ncclGetUniqueId
  void init(int rank, int world, const void* id128);            // ncclCommInitRank (collective)
  void finalize();
  void allreduce_sum(double* buf, size_t count, cudaStream_t s);
  void allreduce_sum(float* buf, size_t count, cudaStream_t s);
  void allgather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
  long long n_allreduce = 0, n_allgather = 0;
  double allgather_bytes = 0;

  // ---- NVLink peer memory (CUDA IPC): the transport of the fused collectives ---------------------------------
  // A window is a cudaMalloc'd buffer of the same size on every rank, mapped into every other rank's address space;
  // base[r] is rank r's copy as seen from this process (base[rank] = the local one).  Collective call.
  static constexpr int kMaxRanks = 8;
  struct Window { void* base[kMaxRanks] = {nullptr}; void** table_dev = nullptr; size_t bytes = 0; };
  Window alloc_window(size_t bytes);
  void free_window(Window& w);
  bool peer_ok = false;          // all peers mapped (set by init when PROPACK_B200_FUSED_COLLECTIVES != 0)
  Window slots;                  // PeerSlot[kMaxRanks][Context::kSlots]: scalar partials of every rank
  // reorthogonalisation coefficients: 2 buffers x [kMaxRanks][kCoefMax] 16-byte elements, then 2 x kMaxRanks flags
  static constexpr int kCoefMax = 2048;
  static constexpr size_t kCoefBufBytes = (size_t)kMaxRanks * kCoefMax * 16;
  Window coef;

 private:
  void* comm_ = nullptr;  // ncclComm_t
};

// One rank's partial of one reduction, written by that rank into every peer's `slots` window.
struct PeerSlot {
  double re, im;
  volatile unsigned long long seq;
  unsigned long long pad;
};

// Block partition of a dimension over `world` ranks: every rank owns `shard_slice(dim, world)` consecutive
// indices (a multiple of 32 elements so that slices stay 256-byte aligned), the last owners fewer or none.
inline long shard_slice(long dim, int world) {
  const long per = (dim + world - 1) / world;
  return (per + 31) / 32 * 32;
}
inline void shard_bounds(long dim, int world, int rank, long& lo, long& hi) {
  const long s = shard_slice(dim, world);
  lo = s * rank < dim ? s * rank : dim;
  hi = s * (rank + 1) < dim ? s * (rank + 1) : dim;
}

}  // namespace pb
