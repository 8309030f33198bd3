// Error reporting and version of the C-ABI (include/ignnition_b200.h).
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

static thread_local char g_last_error[512] = "";

void ign_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

extern "C" int ign_version(void) { return 100; }   // 0.1.0

extern "C" int ign_last_error(char* buf, size_t n) {
  const size_t len = strlen(g_last_error);
  if (buf && n > 0) {
    const size_t c = len < n - 1 ? len : n - 1;
    memcpy(buf, g_last_error, c);
    buf[c] = '\0';
  }
  return (int)len;
}

static std::atomic<long long> g_launches{0};
void ign_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t ign_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

// process-wide switch between the tcgen05 (3xTF32) kernels and their fp32 CUDA-core twins
static std::atomic<int> g_tensor_cores{1};
bool ign_tensor_cores_enabled() { return g_tensor_cores.load(std::memory_order_relaxed) != 0; }
extern "C" int ign_set_tensor_cores(int enable) {
  return g_tensor_cores.exchange(enable ? 1 : 0, std::memory_order_relaxed);
}
