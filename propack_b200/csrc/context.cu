#include "context.hpp"

#include "comm.hpp"

#include <cstdlib>
#include <cstring>

namespace pb {

Context& Context::get() {
  static Context* ctx = new Context();  // leaked on purpose: CUDA teardown order at exit is undefined
  return *ctx;
}

Context::Context() {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw CudaError(e == cudaSuccess ? cudaErrorNoDevice : e,
                    "propack_b200: no CUDA device available -- this library has no CPU fallback");
  PB_CUDA(cudaGetDevice(&device));
  cudaDeviceProp prop;
  PB_CUDA(cudaGetDeviceProperties(&prop, device));
  num_sms = prop.multiProcessorCount;
  if (prop.major < 10 && !std::getenv("PROPACK_B200_ALLOW_ANY_ARCH")) {
    char buf[256];
    snprintf(buf, sizeof buf, "propack_b200: device '%s' is sm_%d%d; this build targets sm_100a (B200) only", prop.name,
             prop.major, prop.minor);
    throw CudaError(cudaErrorInvalidDevice, buf);
  }
  PB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  PB_CUDA(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
  PB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  PB_CUDA(cudaHostAlloc((void**)&host_slots, sizeof(ScalarSlot) * kSlots, cudaHostAllocMapped));
  std::memset((void*)host_slots, 0, sizeof(ScalarSlot) * kSlots);
  PB_CUDA(cudaHostGetDevicePointer((void**)&host_slots_dev, (void*)host_slots, 0));
  PB_CUDA(cudaHostAlloc((void**)&host_err, sizeof(unsigned int), cudaHostAllocMapped));
  *host_err = 0u;
  PB_CUDA(cudaHostGetDevicePointer((void**)&host_err_dev, (void*)host_err, 0));
  PB_CUDA(cudaMalloc((void**)&dev_slots, sizeof(ScalarSlot) * kSlots));
  PB_CUDA(cudaMemset(dev_slots, 0, sizeof(ScalarSlot) * kSlots));
  PB_CUDA(cudaMalloc((void**)&partials, sizeof(double) * 2 * kMaxCtas));
  PB_CUDA(cudaMalloc((void**)&ticket, sizeof(unsigned int)));
  PB_CUDA(cudaMemset(ticket, 0, sizeof(unsigned int)));
  PB_CUDA(cudaMalloc((void**)&tickets8, sizeof(unsigned int) * 8));
  PB_CUDA(cudaMemset(tickets8, 0, sizeof(unsigned int) * 8));
  PB_CUDA(cudaDeviceSynchronize());
  profile = std::getenv("PROPACK_B200_PROFILE") != nullptr;
  if (const char* e = std::getenv("PROPACK_B200_PEER_TIMEOUT_S")) set_peer_timeout(std::atoi(e));
}

void Context::set_stream(cudaStream_t s) {
  PB_CUDA(cudaStreamSynchronize(stream));
  if (owns_stream && stream) cudaStreamDestroy(stream);
  stream = s;
  owns_stream = false;
}

double Context::wait(const Pending& p, double* imag) {
  volatile ScalarSlot* s = host_slots + p.slot;
  ctr.host_syncs += 1;
  unsigned long spins = 0;
  while (s->seq != p.seq) {
    if ((++spins & 0xfffu) == 0) {
      cudaError_t e = cudaStreamQuery(stream);
      if (e == cudaSuccess) {
        // stream drained: the value must be there now (the kernel's system-scope fence precedes completion)
        if (s->seq == p.seq) break;
        throw CudaError(cudaErrorUnknown, "propack_b200: reduction result was never published (internal error)");
      }
      if (e != cudaErrorNotReady) cuda_check(e, "cudaStreamQuery while waiting for a scalar", __FILE__, __LINE__);
    }
  }
  if (*(volatile unsigned int*)host_err) {
    *host_err = 0u;
    throw CudaError(cudaErrorUnknown, "propack_b200: a peer rank never delivered its partial of a cross-GPU reduction (timeout)");
  }
  if (imag) *imag = s->im;
  return s->re;
}

// Second half of a cross-rank reduction: after the all-reduce of (re, im) in dev_slot, apply the norm's sqrt and
// publish to the host-mapped slot.
__global__ void publish_slot_kernel(ScalarSlot* dev_slot, ScalarSlot* host_slot, unsigned long long seq, int kind) {
  double sr = dev_slot->re, si = dev_slot->im;
  if (kind == 1) { sr = sqrt(sr); dev_slot->re = sr; }
  host_slot->re = sr; host_slot->im = si;
  __threadfence_system();
  host_slot->seq = seq;
}

void Context::complete_reduce(const Pending& p, int kind) {
  if (!p.local_only) return;
  ScalarSlot* d = dev_slots + p.slot;
  Comm::get().allreduce_sum(&d->re, 2, stream);   // (re, im) are adjacent doubles
  publish_slot_kernel<<<1, 1, 0, stream>>>(d, host_slots_dev + p.slot, p.seq, kind);
  PB_LAUNCH_CHECK();
  ctr.launches += 1;
}

void* Context::scratch(size_t bytes) {
  if (bytes > scratch_bytes_) {
    if (scratch_) { PB_CUDA(cudaStreamSynchronize(stream)); cudaFree(scratch_); }
    size_t want = bytes + bytes / 2 + 4096;
    PB_CUDA(cudaMalloc(&scratch_, want));
    scratch_bytes_ = want;
  }
  return scratch_;
}

void* Context::basis_acquire(const BasisLayout& lay, bool* zeroed) {
  for (size_t i = 0; i < parked_.size(); ++i)
    if (parked_[i].lay == lay) {
      void* p = parked_[i].p;
      parked_.erase(parked_.begin() + i);
      *zeroed = true;
      return p;
    }
  // another layout: the parked buffers would only pin memory the new solve needs
  basis_cache_clear();
  void* p = nullptr;
  const size_t bytes = (size_t)lay.ld * (size_t)lay.cols * (size_t)lay.elem;
  PB_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
  *zeroed = false;
  return p;
}
void Context::basis_release(void* p, const BasisLayout& lay) {
  if (!p) return;
  if (parked_.size() >= 2 || std::getenv("PROPACK_B200_NO_BASIS_CACHE")) { cudaFree(p); return; }
  parked_.push_back(Parked{p, lay});
}
void Context::basis_cache_clear() {
  for (auto& b : parked_) cudaFree(b.p);
  parked_.clear();
}

Context::PhaseScope::PhaseScope(Context& c_, Phase p) : c(c_), ph(p), l0(c_.ctr.launches) {
  if (c.profile) {
    PB_CUDA(cudaEventCreate(&e0));
    PB_CUDA(cudaEventCreate(&e1));
    PB_CUDA(cudaEventRecord(e0, c.stream));
  }
}
Context::PhaseScope::~PhaseScope() {
  c.ctr.phase_launches[ph] += c.ctr.launches - l0;
  if (c.profile) {
    cudaEventRecord(e1, c.stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    c.ctr.phase_ms[ph] += ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
}

}  // namespace pb
