// Library-wide state of the C ABI: last error string, version, launch counter.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstring>

namespace eavit {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// dropout epoch word, one per device (see DropCfg in common.cuh)
static uint32_t* g_epoch[64] = {nullptr};
const uint32_t* drop_epoch_ptr() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (g_epoch[dev] == nullptr) {
    uint32_t* p = nullptr;
    if (cudaMalloc(&p, sizeof(uint32_t)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(uint32_t));
    g_epoch[dev] = p;
  }
  return g_epoch[dev];
}
__global__ void epoch_kernel(uint32_t* p, uint32_t add, uint32_t set_to, int set) { *p = set ? set_to : *p + add; }
}  // namespace eavit

extern "C" {
const char* eavit_last_error(void) { return eavit::g_err; }
int eavit_version(void) { return 100; }
long long eavit_launch_count(void) { return eavit::g_launches.load(std::memory_order_relaxed); }

int eavit_dropout_epoch_bump(void* stream) {
  uint32_t* p = const_cast<uint32_t*>(eavit::drop_epoch_ptr());
  if (p == nullptr) { eavit::set_error("dropout epoch: allocation failed"); return EAVIT_ECUDA; }
  eavit::epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p, 1u, 0u, 0);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
int eavit_dropout_epoch_set(int value, void* stream) {
  uint32_t* p = const_cast<uint32_t*>(eavit::drop_epoch_ptr());
  if (p == nullptr) { eavit::set_error("dropout epoch: allocation failed"); return EAVIT_ECUDA; }
  eavit::epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p, 0u, (uint32_t)value, 1);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
}
