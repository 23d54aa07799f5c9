// Internal helpers shared by the translation units of libffvc_sm100.so.
#pragma once
#include "../../include/ffvc.h"

namespace ffvc {
int set_error(int code, const char* msg);
void count_launch(int n = 1);
int option(int id);   // kernel-selection switch (ffvc_set_option / FFVC_OPTS)
enum { OPT_LN_FWD_V2 = 0, OPT_LN_BWD_V2 = 1, OPT_POOL_V2 = 2, OPT_GN_RING = 3, OPT_HALO_EPI16 = 4, OPT_SM_LIMIT = 5, OPT_GEMM_QUAD = 6, OPT_COUNT = 7 };
// SMs the persistent tcgen05 kernels may occupy: min(num_sms, option sm_limit) (0 = all).  The data-parallel step lowers it while a
// gradient all-reduce is in flight: a persistent grid of 148 CTAs on a GPU where NCCL holds k SMs runs its last k CTAs in a second wave.
inline int sm_budget(int num_sms) {
  const int lim = option(OPT_SM_LIMIT);
  return (lim > 0 && lim < num_sms) ? lim : num_sms;
}
}  // namespace ffvc

#define FFVC_CHECK_LAUNCH()                                                        \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) return ffvc::set_error(FFVC_ERR_CUDA, cudaGetErrorString(e__)); \
    ffvc::count_launch();                                                          \
  } while (0)
