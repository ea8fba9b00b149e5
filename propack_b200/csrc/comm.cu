#include "comm.hpp"

#include <dlfcn.h>
#include <glob.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"

namespace pb {

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool bind_from(void* h) {
  if (!h) return false;
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
  a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!(a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.GetErrorString)) return false;
  g_nccl = a;
  return true;
}

void bind_nccl() {
  if (g_nccl.handle) return;
  // 1. a copy already loaded into the process (e.g. torch's bundled NCCL) -- keeps one NCCL per process
  if (bind_from(dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL))) return;
  std::vector<std::string> cands;
  if (const char* e = std::getenv("PROPACK_B200_NCCL")) cands.push_back(e);
  for (const char* pat : {"/opt/*/.venv/lib/python3*/site-packages/nvidia/nccl/lib/libnccl.so.2",
                          "/usr/lib/python3*/site-packages/nvidia/nccl/lib/libnccl.so.2"}) {
    glob_t g;
    if (glob(pat, 0, nullptr, &g) == 0)
      for (size_t i = 0; i < g.gl_pathc; ++i) cands.push_back(g.gl_pathv[i]);
    globfree(&g);
  }
  cands.push_back("libnccl.so.2");
  for (const auto& c : cands)
    if (bind_from(dlopen(c.c_str(), RTLD_NOW | RTLD_GLOBAL))) return;
  throw std::runtime_error("propack_b200: libnccl.so.2 not found (set PROPACK_B200_NCCL); multi-GPU needs NCCL");
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess)
    throw std::runtime_error(std::string("propack_b200: NCCL error in ") + what + ": " +
                             (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
}
}  // namespace

Comm& Comm::get() {
  static Comm* c = new Comm();
  return *c;
}

void Comm::unique_id(void* out128) {
  bind_nccl();
  ncclUniqueId id;
  nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, &id, sizeof id);
}

void Comm::init(int rank_, int world_, const void* id128) {
  if (world_ <= 1) { rank = 0; world = 1; return; }
  bind_nccl();
  if (comm_) finalize();
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  ncclComm_t c;
  nccl_check(g_nccl.CommInitRank(&c, world_, id, rank_), "ncclCommInitRank");
  comm_ = c; rank = rank_; world = world_;
  // Fused collectives over NVLink peer memory (default on; PROPACK_B200_FUSED_COLLECTIVES=0 keeps every reduction on NCCL)
  const char* e = std::getenv("PROPACK_B200_FUSED_COLLECTIVES");
  if (world <= kMaxRanks && !(e && e[0] == '0')) {
    slots = alloc_window(sizeof(PeerSlot) * kMaxRanks * 64);
    coef = alloc_window(2 * kCoefBufBytes + 2 * kMaxRanks * sizeof(unsigned long long));
    peer_ok = true;
  }
}

void Comm::finalize() {
  if (peer_ok) { free_window(slots); free_window(coef); peer_ok = false; }
  if (comm_) { g_nccl.CommDestroy((ncclComm_t)comm_); comm_ = nullptr; }
  rank = 0; world = 1;
}

Comm::Window Comm::alloc_window(size_t bytes) {
  Window w;
  w.bytes = bytes;
  void* local = nullptr;
  PB_CUDA(cudaMalloc(&local, bytes));
  PB_CUDA(cudaMemset(local, 0, bytes));
  cudaIpcMemHandle_t mine;
  PB_CUDA(cudaIpcGetMemHandle(&mine, local));
  // exchange the handles with an NCCL all-gather (the only out-of-band channel the library owns)
  cudaIpcMemHandle_t* dev = nullptr;
  PB_CUDA(cudaMalloc((void**)&dev, sizeof(cudaIpcMemHandle_t) * world));
  PB_CUDA(cudaMemcpy(dev + rank, &mine, sizeof mine, cudaMemcpyHostToDevice));
  nccl_check(g_nccl.AllGather(dev + rank, dev, sizeof mine, ncclInt8, (ncclComm_t)comm_, nullptr), "ncclAllGather(ipc handles)");
  PB_CUDA(cudaDeviceSynchronize());
  std::vector<cudaIpcMemHandle_t> all(world);
  PB_CUDA(cudaMemcpy(all.data(), dev, sizeof(cudaIpcMemHandle_t) * world, cudaMemcpyDeviceToHost));
  cudaFree(dev);
  for (int r = 0; r < world; ++r) {
    if (r == rank) { w.base[r] = local; continue; }
    PB_CUDA(cudaIpcOpenMemHandle(&w.base[r], all[r], cudaIpcMemLazyEnablePeerAccess));
  }
  PB_CUDA(cudaMalloc((void**)&w.table_dev, sizeof(void*) * kMaxRanks));
  PB_CUDA(cudaMemcpy(w.table_dev, w.base, sizeof(void*) * kMaxRanks, cudaMemcpyHostToDevice));
  return w;
}

void Comm::free_window(Window& w) {
  for (int r = 0; r < world; ++r) {
    if (!w.base[r]) continue;
    if (r == rank) cudaFree(w.base[r]);
    else cudaIpcCloseMemHandle(w.base[r]);
    w.base[r] = nullptr;
  }
  if (w.table_dev) cudaFree(w.table_dev);
  w.table_dev = nullptr;
}

void Comm::allreduce_sum(double* buf, size_t count, cudaStream_t s) {
  if (!active() || count == 0) return;
  nccl_check(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)comm_, s), "ncclAllReduce(f64)");
  n_allreduce += 1;
}
void Comm::allreduce_sum(float* buf, size_t count, cudaStream_t s) {
  if (!active() || count == 0) return;
  nccl_check(g_nccl.AllReduce(buf, buf, count, ncclFloat, ncclSum, (ncclComm_t)comm_, s), "ncclAllReduce(f32)");
  n_allreduce += 1;
}
void Comm::allgather(const void* send, void* recv, size_t bytes, cudaStream_t s) {
  if (!active()) {
    if (send != recv) PB_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s));
    return;
  }
  nccl_check(g_nccl.AllGather(send, recv, bytes, ncclInt8, (ncclComm_t)comm_, s), "ncclAllGather");
  n_allgather += 1;
  allgather_bytes += (double)bytes * world;
}

}  // namespace pb
