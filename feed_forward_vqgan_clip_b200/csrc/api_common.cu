// Error reporting / launch accounting for libffvc_sm100.so.
#include <cuda_runtime.h>
#include <atomic>
#include <cstring>
#include "ffvc_internal.h"

namespace ffvc {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace ffvc

extern "C" const char* ffvc_last_error(void) { return ffvc::g_err; }
extern "C" int ffvc_arch(void) { return 100; }
extern "C" long long ffvc_launch_count(void) { return ffvc::g_launches.load(); }
extern "C" void ffvc_reset_launch_count(void) { ffvc::g_launches.store(0); }

// Struct sizes, so the ctypes mirror in _lib.py can be checked without a GPU (tests/test_abi.py).
extern "C" int ffvc_sizeof(const char* name) {
  if (!name) return -1;
  if (!strcmp(name, "ffvc_gemm_params")) return (int)sizeof(ffvc_gemm_params);
  return -1;
}
