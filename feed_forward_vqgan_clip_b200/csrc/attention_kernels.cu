// Small-sequence multi-head attention for the CLIP ViT (T = 50 tokens, 12 heads x 64), forward and backward.
// Reference: nn.MultiheadAttention inside ResidualAttentionBlock.attention, cloob.py:188,198-200.
// One CTA per (sequence, head): Q, K, V (and dO) live in shared memory as fp32, scores never touch HBM.
// qkv layout: [N][T][3*W] bf16 (the fused in_proj output: q | k | v along the last dim), head h owns columns
// h*64..h*64+63 of each third.  out / dout: [N][T][W] bf16.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cfloat>

#include "ffvc_internal.h"
#include "ptx.cuh"

namespace ffvc {

static constexpr int kDh = 64;
static constexpr int kTmax = 64;

__device__ __forceinline__ void load_head(const __nv_bfloat16* __restrict__ base, long long row_stride, int T, float* dst) {
  // dst[t][kDh+1] fp32, coalesced 16B loads: 8 vectors per row
  for (int i = threadIdx.x; i < T * 8; i += blockDim.x) {
    const int t = i >> 3, v = i & 7;
    const uint4 pk = *reinterpret_cast<const uint4*>(base + (long long)t * row_stride + v * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      dst[t * (kDh + 1) + v * 8 + 2 * j] = f.x;
      dst[t * (kDh + 1) + v * 8 + 2 * j + 1] = f.y;
    }
  }
}

__global__ void __launch_bounds__(256) mha_small_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ out, int T, int heads, float scale) {
  extern __shared__ float sm[];
  float* q = sm;
  float* k = q + kTmax * (kDh + 1);
  float* v = k + kTmax * (kDh + 1);
  float* s = v + kTmax * (kDh + 1);  // [T][kTmax+1]
  const int n = blockIdx.x / heads, h = blockIdx.x % heads;
  const int W = heads * kDh;
  const __nv_bfloat16* base = qkv + (long long)n * T * 3 * W + h * kDh;
  load_head(base, 3 * W, T, q);
  load_head(base + W, 3 * W, T, k);
  load_head(base + 2 * W, 3 * W, T, v);
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int a = i / T, b = i % T;
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < kDh; ++d) acc = fmaf(q[a * (kDh + 1) + d], k[b * (kDh + 1) + d], acc);
    s[a * (kTmax + 1) + b] = acc * scale;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int a = warp; a < T; a += (blockDim.x >> 5)) {
    float mx = -FLT_MAX;
    for (int b = lane; b < T; b += 32) mx = fmaxf(mx, s[a * (kTmax + 1) + b]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int b = lane; b < T; b += 32) {
      const float e = __expf(s[a * (kTmax + 1) + b] - mx);
      s[a * (kTmax + 1) + b] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int b = lane; b < T; b += 32) s[a * (kTmax + 1) + b] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * (kDh / 2); i += blockDim.x) {
    const int a = i / (kDh / 2), d2 = i % (kDh / 2);
    float o0 = 0.f, o1 = 0.f;
    for (int b = 0; b < T; ++b) {
      const float p = s[a * (kTmax + 1) + b];
      o0 = fmaf(p, v[b * (kDh + 1) + 2 * d2], o0);
      o1 = fmaf(p, v[b * (kDh + 1) + 2 * d2 + 1], o1);
    }
    *reinterpret_cast<__nv_bfloat162*>(out + ((long long)n * T + a) * W + h * kDh + 2 * d2) = __floats2bfloat162_rn(o0, o1);
  }
}

// dqkv: [N][T][3W] bf16 gradient of the fused projection output.  Probabilities are recomputed.
__global__ void __launch_bounds__(256) mha_small_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            const __nv_bfloat16* __restrict__ dout,
                                                            __nv_bfloat16* __restrict__ dqkv, int T, int heads, float scale) {
  extern __shared__ float sm[];
  float* q = sm;
  float* k = q + kTmax * (kDh + 1);
  float* v = k + kTmax * (kDh + 1);
  float* dO = v + kTmax * (kDh + 1);
  float* p = dO + kTmax * (kDh + 1);     // [T][kTmax+1] probabilities
  float* ds = p + kTmax * (kTmax + 1);   // [T][kTmax+1] dS
  const int n = blockIdx.x / heads, h = blockIdx.x % heads;
  const int W = heads * kDh;
  const __nv_bfloat16* base = qkv + (long long)n * T * 3 * W + h * kDh;
  load_head(base, 3 * W, T, q);
  load_head(base + W, 3 * W, T, k);
  load_head(base + 2 * W, 3 * W, T, v);
  load_head(dout + (long long)n * T * W + h * kDh, W, T, dO);
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int a = i / T, b = i % T;
    float acc = 0.f, acc2 = 0.f;
#pragma unroll 16
    for (int d = 0; d < kDh; ++d) {
      acc = fmaf(q[a * (kDh + 1) + d], k[b * (kDh + 1) + d], acc);
      acc2 = fmaf(dO[a * (kDh + 1) + d], v[b * (kDh + 1) + d], acc2);
    }
    p[a * (kTmax + 1) + b] = acc * scale;
    ds[a * (kTmax + 1) + b] = acc2;  // dP
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int a = warp; a < T; a += (blockDim.x >> 5)) {
    float mx = -FLT_MAX;
    for (int b = lane; b < T; b += 32) mx = fmaxf(mx, p[a * (kTmax + 1) + b]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int b = lane; b < T; b += 32) {
      const float e = __expf(p[a * (kTmax + 1) + b] - mx);
      p[a * (kTmax + 1) + b] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float dot = 0.f;
    for (int b = lane; b < T; b += 32) {
      const float pv = p[a * (kTmax + 1) + b] * inv;
      p[a * (kTmax + 1) + b] = pv;
      dot += pv * ds[a * (kTmax + 1) + b];
    }
    dot = warp_sum(dot);
    for (int b = lane; b < T; b += 32)
      ds[a * (kTmax + 1) + b] = p[a * (kTmax + 1) + b] * (ds[a * (kTmax + 1) + b] - dot) * scale;
  }
  __syncthreads();
  __nv_bfloat16* ob = dqkv + (long long)n * T * 3 * W + h * kDh;
  for (int i = threadIdx.x; i < T * (kDh / 2); i += blockDim.x) {
    const int a = i / (kDh / 2), d2 = i % (kDh / 2);
    float dq0 = 0.f, dq1 = 0.f, dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
    for (int b = 0; b < T; ++b) {
      const float dsab = ds[a * (kTmax + 1) + b];  // dS[a][b]
      const float dsba = ds[b * (kTmax + 1) + a];  // dS[b][a]
      const float pba = p[b * (kTmax + 1) + a];    // P[b][a]
      dq0 = fmaf(dsab, k[b * (kDh + 1) + 2 * d2], dq0);
      dq1 = fmaf(dsab, k[b * (kDh + 1) + 2 * d2 + 1], dq1);
      dk0 = fmaf(dsba, q[b * (kDh + 1) + 2 * d2], dk0);
      dk1 = fmaf(dsba, q[b * (kDh + 1) + 2 * d2 + 1], dk1);
      dv0 = fmaf(pba, dO[b * (kDh + 1) + 2 * d2], dv0);
      dv1 = fmaf(pba, dO[b * (kDh + 1) + 2 * d2 + 1], dv1);
    }
    __nv_bfloat16* r = ob + (long long)a * 3 * W + 2 * d2;
    *reinterpret_cast<__nv_bfloat162*>(r) = __floats2bfloat162_rn(dq0, dq1);
    *reinterpret_cast<__nv_bfloat162*>(r + W) = __floats2bfloat162_rn(dk0, dk1);
    *reinterpret_cast<__nv_bfloat162*>(r + 2 * W) = __floats2bfloat162_rn(dv0, dv1);
  }
}

}  // namespace ffvc

using namespace ffvc;

extern "C" int ffvc_mha_small_fwd(const void* qkv, void* out, int N, int T, int heads, int head_dim, float scale,
                                  void* stream) {
  if (head_dim != kDh || T > kTmax || T < 1) return set_error(FFVC_ERR_UNSUPPORTED, "mha_small: head_dim must be 64 and T <= 64");
  const size_t smem = (size_t)(3 * kTmax * (kDh + 1) + kTmax * (kTmax + 1)) * sizeof(float);
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(mha_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(mha_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    done = true;
  }
  mha_small_fwd_kernel<<<N * heads, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), T, heads, scale);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_mha_small_bwd(const void* qkv, const void* dout, void* dqkv, int N, int T, int heads, int head_dim,
                                  float scale, void* stream) {
  if (head_dim != kDh || T > kTmax || T < 1) return set_error(FFVC_ERR_UNSUPPORTED, "mha_small: head_dim must be 64 and T <= 64");
  const size_t smem = (size_t)(4 * kTmax * (kDh + 1) + 2 * kTmax * (kTmax + 1)) * sizeof(float);
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(mha_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(mha_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    done = true;
  }
  mha_small_bwd_kernel<<<N * heads, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout),
      reinterpret_cast<__nv_bfloat16*>(dqkv), T, heads, scale);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
