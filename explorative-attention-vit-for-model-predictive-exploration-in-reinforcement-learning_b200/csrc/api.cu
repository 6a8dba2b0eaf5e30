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
}  // namespace eavit

extern "C" {
const char* eavit_last_error(void) { return eavit::g_err; }
int eavit_version(void) { return 100; }
long long eavit_launch_count(void) { return eavit::g_launches.load(std::memory_order_relaxed); }
}
