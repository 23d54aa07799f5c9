// Internal helpers shared by the translation units of libffvc_sm100.so.
#pragma once
#include "../../include/ffvc.h"

namespace ffvc {
int set_error(int code, const char* msg);
void count_launch(int n = 1);
enum { OPT_LN_FWD_V2 = 0, OPT_LN_BWD_V2 = 1, OPT_POOL_V2 = 2, OPT_GN_RING = 3, OPT_HALO_EPI16 = 4, OPT_COUNT = 5 };
int option(int id);   // kernel-selection switch (ffvc_set_option / FFVC_OPTS)
}  // namespace ffvc

#define FFVC_CHECK_LAUNCH()                                                        \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) return ffvc::set_error(FFVC_ERR_CUDA, cudaGetErrorString(e__)); \
    ffvc::count_launch();                                                          \
  } while (0)
