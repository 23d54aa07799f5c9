// Error reporting / launch accounting for libffvc_sm100.so.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
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

// ---- kernel-selection switches.  Defaults are the variants measured faster on B200 (profiles/); FFVC_OPTS overrides.
static const char* const kOptNames[OPT_COUNT] = {"ln_fwd_v2", "ln_bwd_v2", "pool_v2", "gn_ring", "halo_epi16", "sm_limit", "gemm_quad"};
static std::atomic<int> g_opts[OPT_COUNT];
static std::once_flag g_opts_once;
static int opt_index(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < OPT_COUNT; ++i)
    if (!strcmp(name, kOptNames[i])) return i;
  return -1;
}
static void opts_init() {
  static const int defaults[OPT_COUNT] = {2, 1, 1, 1, 0, 0, 0}   /* measured: profiles/r01_ab_kernels.md */;
  for (int i = 0; i < OPT_COUNT; ++i) g_opts[i].store(defaults[i]);
  const char* env = getenv("FFVC_OPTS");
  if (!env) return;
  char buf[256];
  strncpy(buf, env, sizeof(buf) - 1);
  buf[sizeof(buf) - 1] = 0;
  for (char* tok = strtok(buf, ","); tok; tok = strtok(nullptr, ",")) {
    char* eq = strchr(tok, '=');
    if (!eq) continue;
    *eq = 0;
    const int i = opt_index(tok);
    if (i >= 0) g_opts[i].store(atoi(eq + 1));
  }
}
int option(int id) {
  std::call_once(g_opts_once, opts_init);
  return g_opts[id].load(std::memory_order_relaxed);
}
}  // namespace ffvc

extern "C" const char* ffvc_last_error(void) { return ffvc::g_err; }
extern "C" int ffvc_arch(void) { return 100; }
extern "C" long long ffvc_launch_count(void) { return ffvc::g_launches.load(); }
extern "C" void ffvc_reset_launch_count(void) { ffvc::g_launches.store(0); }

extern "C" int ffvc_set_option(const char* name, int value) {
  std::call_once(ffvc::g_opts_once, ffvc::opts_init);
  const int i = ffvc::opt_index(name);
  if (i < 0) return -1;
  return ffvc::g_opts[i].exchange(value);
}
extern "C" int ffvc_get_option(const char* name) {
  std::call_once(ffvc::g_opts_once, ffvc::opts_init);
  const int i = ffvc::opt_index(name);
  return i < 0 ? -1 : ffvc::g_opts[i].load();
}

// Struct sizes, so the ctypes mirror in _lib.py can be checked without a GPU (tests/test_abi.py).
extern "C" int ffvc_sizeof(const char* name) {
  if (!name) return -1;
  if (!strcmp(name, "ffvc_gemm_params")) return (int)sizeof(ffvc_gemm_params);
  return -1;
}
