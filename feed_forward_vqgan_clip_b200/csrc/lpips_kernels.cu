// Pieces of the LPIPS-VGG16 diversity term of the train step (main.py:532-537,776-791; taming lpips.vgg16 slices):
// 2x2 max-pool on NHWC bf16 (forward / backward), CLIP-style image normalisation, and the fused
// normalize_tensor + pairwise squared-difference loss over `repeat` samples of the same prompt (forward + backward).
// The VGG convolutions themselves run on the tcgen05 implicit-GEMM kernels with a ReLU epilogue.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cfloat>

#include "ffvc_internal.h"
#include "ptx.cuh"

namespace ffvc {

static inline unsigned grid_for_l(long long n, int threads, int cap = 148 * 16) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  return (unsigned)(g < cap ? g : cap);
}

__device__ __forceinline__ void unpack8l(const uint4& pk, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8l(const float (&f)[8]) {
  uint4 pk;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
  pk.x = *reinterpret_cast<uint32_t*>(&h0);
  pk.y = *reinterpret_cast<uint32_t*>(&h1);
  pk.z = *reinterpret_cast<uint32_t*>(&h2);
  pk.w = *reinterpret_cast<uint32_t*>(&h3);
  return pk;
}

// y[n][oy][ox][c] = max over the 2x2 window; C % 8 == 0
__global__ void maxpool2x2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H, int W, int C) {
  const int CV = C >> 3, OH = H >> 1, OW = W >> 1;
  const long long total = (long long)N * OH * OW * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV);
    long long p = i / CV;
    const int ox = (int)(p % OW);
    p /= OW;
    const int oy = (int)(p % OH);
    const int n = (int)(p / OH);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -FLT_MAX;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float v[8];
        unpack8l(*reinterpret_cast<const uint4*>(x + ((((long long)n * H + 2 * oy + dy) * W + 2 * ox + dx) * C + c * 8)), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8l(m);
  }
}
// dx[window] = dy routed to the first maximum of the window (row-major scan), 0 elsewhere
__global__ void maxpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                      __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C) {
  const int CV = C >> 3, OH = H >> 1, OW = W >> 1;
  const long long total = (long long)N * OH * OW * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV);
    long long p = i / CV;
    const int ox = (int)(p % OW);
    p /= OW;
    const int oy = (int)(p % OH);
    const int n = (int)(p / OH);
    float v[4][8], g[8];
    unpack8l(*reinterpret_cast<const uint4*>(dy + i * 8), g);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      unpack8l(*reinterpret_cast<const uint4*>(x + ((((long long)n * H + 2 * oy + (t >> 1)) * W + 2 * ox + (t & 1)) * C + c * 8)), v[t]);
    float o[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int best = 0;
      float m = v[0][j];
#pragma unroll
      for (int t = 1; t < 4; ++t)
        if (v[t][j] > m) {
          m = v[t][j];
          best = t;
        }
#pragma unroll
      for (int t = 0; t < 4; ++t) o[t][j] = (t == best) ? g[j] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
      *reinterpret_cast<uint4*>(dx + ((((long long)n * H + 2 * oy + (t >> 1)) * W + 2 * ox + (t & 1)) * C + c * 8)) = pack8l(o[t]);
  }
}

// out = g where post > 0 else 0 (ReLU backward with the post-activation tensor as the mask)
__global__ void relu_mask_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ post,
                                 __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(post[i]) > 0.f ? g[i] : __float2bfloat16(0.f);
}

// y = (x - mean[c]) * istd[c] on NHWC fp32 with 3 channels; bwd: dx += dy * istd[c]
__global__ void normalize3_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float m0, float m1, float m2,
                                  float s0, float s1, float s2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), s = c == 0 ? s0 : (c == 1 ? s1 : s2);
    y[i] = (x[i] - m) * s;
  }
}
__global__ void normalize3_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long n, float s0, float s1, float s2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    dx[i] += dy[i] * (c == 0 ? s0 : (c == 1 ? s1 : s2));
  }
}

// Diversity term of one VGG tap (main.py:779-782, mode 'between_same_prompts'):
//   a_r = f_r / (||f_r||_2 + 1e-10) over channels;  div = mean_{r1,r2,b,y,x} sum_c (a_r1 - a_r2)^2
// feats: [R*B][HW][C] bf16 (sample index r*B + b).  One warp per (b, pixel); C <= 512, R <= 4.
// loss_accum += scale * div_contribution;  dfeat (same layout) = scale * d(div)/d(feats)  (overwritten)
template <int kMaxR>
__global__ void __launch_bounds__(256) diversity_kernel(const __nv_bfloat16* __restrict__ feats, float* __restrict__ loss_accum,
                                                        __nv_bfloat16* __restrict__ dfeat, int R, int B, int HW, int C, float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + warp;   // over B * HW
  if (item >= (long long)B * HW) return;
  const int b = (int)(item / HW);
  const int pix = (int)(item % HW);
  constexpr int kMaxPerLane = 16;   // C <= 512
  float a[kMaxR][kMaxPerLane], inv[kMaxR], nrm[kMaxR];
  const int per = (C + 31) / 32;
  for (int r = 0; r < R; ++r) {
    const __nv_bfloat16* f = feats + (((long long)(r * B + b)) * HW + pix) * C;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int c = lane + 32 * k;
      const float v = (k < per && c < C) ? __bfloat162float(f[c]) : 0.f;
      a[r][k] = v;
      ss += v * v;
    }
    ss = warp_sum(ss);
    nrm[r] = sqrtf(ss);
    inv[r] = 1.0f / (nrm[r] + 1e-10f);
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) a[r][k] *= inv[r];
  }
  const float w = scale / ((float)R * R * B * HW);
  float lsum = 0.f;
  for (int r = 0; r < R; ++r) {
    // g = 4 * sum_r2 (a_r - a_r2)  (each unordered pair appears twice in the double sum)
    float g[kMaxPerLane];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      float acc = 0.f;
      for (int r2 = 0; r2 < R; ++r2) {
        const float d = a[r][k] - a[r2][k];
        acc += d;
        lsum += d * d;
      }
      g[k] = 4.0f * w * acc;
      dot += g[k] * a[r][k];
    }
    dot = warp_sum(dot);
    if (dfeat) {
      // through a = x / (n + eps): dx = g / (n + eps) - x (x.g) / (n (n + eps)^2) = inv * (g - a * (a.g) * (n + eps) / n)
      const float corr = nrm[r] > 0.f ? (nrm[r] + 1e-10f) / nrm[r] : 0.f;
      __nv_bfloat16* o = dfeat + (((long long)(r * B + b)) * HW + pix) * C;
#pragma unroll
      for (int k = 0; k < kMaxPerLane; ++k) {
        const int c = lane + 32 * k;
        if (k < per && c < C) o[c] = __float2bfloat16(inv[r] * (g[k] - a[r][k] * dot * corr));
      }
    }
  }
  lsum = warp_sum(lsum);
  if (lane == 0) atomicAdd(loss_accum, w * lsum);
}

// Vector form of the kernel above for R in {2, 3, 4} and C in {64, 128, 256, 512} (every VGG16 tap): a lane owns V 8-channel
// vectors (16-byte loads / stores) of one pixel, G = C / (8 V) <= 32 neighbouring lanes share a pixel, 32 / G pixels per warp;
// everything is unrolled at compile time (no indexed register arrays -> no local memory); CTAs walk the pixels grid-stride and
// issue ONE loss atomic each.  Round 1's form used 2-byte loads, one warp per pixel whatever C, and one same-address atomic per
// pixel: 6.2 ms for the 64-channel tap of 16 images at 512 x 512 (2.1 M atomics on one float) against 0.17 ms of HBM time.
template <int R, int V>
__global__ void __launch_bounds__(256) diversity_vec_kernel(const __nv_bfloat16* __restrict__ feats, float* __restrict__ loss_accum,
                                                            __nv_bfloat16* __restrict__ dfeat, int B, int HW, int C, float scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = C / (8 * V);                       // lanes per pixel (power of two, <= 32)
  const int ppw = 32 / G;                          // pixels per warp
  const int gl = lane % G, pw = lane / G;
  const long long items = (long long)B * HW;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const float w = scale / ((float)R * R * B * HW);
  float lsum = 0.f;
  for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * ppw; base < items; base += warps_total * ppw) {
    const long long item = base + pw;
    const bool ok = item < items;
    const int b = ok ? (int)(item / HW) : 0;
    const int pix = ok ? (int)(item % HW) : 0;
    float a[R][V][8], inv[R], nrm[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const __nv_bfloat16* f = feats + (((long long)(r * B + b)) * HW + pix) * C;
      float ss = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (ok) pk = *reinterpret_cast<const uint4*>(f + (v * G + gl) * 8);
        unpack8l(pk, a[r][v]);
#pragma unroll
        for (int k = 0; k < 8; ++k) ss = fmaf(a[r][v][k], a[r][v][k], ss);
      }
      for (int o = 1; o < G; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      nrm[r] = sqrtf(ss);
      inv[r] = 1.0f / (nrm[r] + 1e-10f);
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[r][v][k] *= inv[r];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float g[V][8], dot = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float acc = 0.f;
#pragma unroll
          for (int r2 = 0; r2 < R; ++r2) {
            const float d = a[r][v][k] - a[r2][v][k];
            acc += d;
            if (ok) lsum = fmaf(d, d, lsum);
          }
          g[v][k] = 4.0f * w * acc;
          dot = fmaf(g[v][k], a[r][v][k], dot);
        }
      for (int o = 1; o < G; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (dfeat && ok) {
        const float corr = nrm[r] > 0.f ? (nrm[r] + 1e-10f) / nrm[r] : 0.f;
        __nv_bfloat16* o = dfeat + (((long long)(r * B + b)) * HW + pix) * C;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float out[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) out[k] = inv[r] * (g[v][k] - a[r][v][k] * dot * corr);
          *reinterpret_cast<uint4*>(o + (v * G + gl) * 8) = pack8l(out);
        }
      }
    }
  }
  lsum = warp_sum(lsum);
  __shared__ float part[8];
  if (lane == 0) part[warp] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    atomicAdd(loss_accum, w * t);
  }
}

// Same term for any number of repeats R (mode 'all', main.py:783-787, is the same expression with R = batch, B = 1; also
// 'between_same_prompts' with repeat > 4).  With S = sum_r a_r:  sum_{r1,r2} |a_r1 - a_r2|^2 = 2 R sum_r |a_r|^2 - 2 |S|^2 and
// d/da_r = 4 (R a_r - S): two streaming passes over the R feature vectors of a pixel (the second one hits L2), nothing but S
// in registers.  One warp per (b, pixel); C <= 512.
__global__ void __launch_bounds__(256) diversity_anyr_kernel(const __nv_bfloat16* __restrict__ feats, float* __restrict__ loss_accum,
                                                             __nv_bfloat16* __restrict__ dfeat, int R, int B, int HW, int C,
                                                             float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + warp;   // over B * HW
  if (item >= (long long)B * HW) return;
  const int b = (int)(item / HW);
  const int pix = (int)(item % HW);
  constexpr int kMaxPerLane = 16;   // C <= 512
  const int per = (C + 31) / 32;
  float S[kMaxPerLane];
#pragma unroll
  for (int k = 0; k < kMaxPerLane; ++k) S[k] = 0.f;
  float q = 0.f;                     // sum_r |a_r|^2
  for (int r = 0; r < R; ++r) {
    const __nv_bfloat16* f = feats + (((long long)r * B + b) * HW + pix) * C;
    float v[kMaxPerLane], ss = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int c = lane + 32 * k;
      v[k] = (k < per && c < C) ? __bfloat162float(f[c]) : 0.f;
      ss += v[k] * v[k];
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / (sqrtf(ss) + 1e-10f);
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) S[k] += v[k] * inv;
    q += ss * inv * inv;
  }
  float s2 = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxPerLane; ++k) s2 += S[k] * S[k];
  s2 = warp_sum(s2);
  const float w = scale / ((float)R * R * B * HW);
  if (lane == 0) atomicAdd(loss_accum, w * (2.0f * R * q - 2.0f * s2));
  if (!dfeat) return;
  for (int r = 0; r < R; ++r) {
    const long long off = (((long long)r * B + b) * HW + pix) * C;
    const __nv_bfloat16* f = feats + off;
    float a[kMaxPerLane], ss = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int c = lane + 32 * k;
      a[k] = (k < per && c < C) ? __bfloat162float(f[c]) : 0.f;
      ss += a[k] * a[k];
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss), inv = 1.0f / (nrm + 1e-10f);
    float g[kMaxPerLane], dot = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      a[k] *= inv;
      g[k] = 4.0f * w * (R * a[k] - S[k]);
      dot += g[k] * a[k];
    }
    dot = warp_sum(dot);
    const float corr = nrm > 0.f ? (nrm + 1e-10f) / nrm : 0.f;
    __nv_bfloat16* o = dfeat + off;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int c = lane + 32 * k;
      if (k < per && c < C) o[c] = __float2bfloat16(inv * (g[k] - a[k] * dot * corr));
    }
  }
}

}  // namespace ffvc

using namespace ffvc;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int ffvc_maxpool2x2_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  if (C % 8 || H % 2 || W % 2) return set_error(FFVC_ERR_ARG, "maxpool2x2: C % 8, H % 2, W % 2 must be 0");
  maxpool2x2_fwd_kernel<<<grid_for_l((long long)N * (H / 2) * (W / 2) * (C / 8), 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), N, H, W, C);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, void* stream) {
  if (C % 8 || H % 2 || W % 2) return set_error(FFVC_ERR_ARG, "maxpool2x2: C % 8, H % 2, W % 2 must be 0");
  maxpool2x2_bwd_kernel<<<grid_for_l((long long)N * (H / 2) * (W / 2) * (C / 8), 256), 256, 0, ST(stream)>>>(CBF(x), CBF(dy), BF(dx), N, H,
                                                                                                           W, C);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
// mean / std: HOST float[3]
extern "C" int ffvc_normalize3_fwd(const float* x, float* y, long long n, const float* mean, const float* std_, void* stream) {
  normalize3_kernel<<<grid_for_l(n, 256), 256, 0, ST(stream)>>>(x, y, n, mean[0], mean[1], mean[2], 1.f / std_[0], 1.f / std_[1],
                                                               1.f / std_[2]);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_normalize3_bwd(const float* dy, float* dx_accum, long long n, const float* std_, void* stream) {
  normalize3_bwd_kernel<<<grid_for_l(n, 256), 256, 0, ST(stream)>>>(dy, dx_accum, n, 1.f / std_[0], 1.f / std_[1], 1.f / std_[2]);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_diversity_tap(const void* feats, float* loss_accum, void* dfeat, int R, int B, int HW, int C, float scale,
                                  void* stream) {
  if (R < 1 || C > 512) return set_error(FFVC_ERR_UNSUPPORTED, "diversity_tap: C <= 512");
  const long long items = (long long)B * HW;
  const bool aligned = ((reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(dfeat)) & 15) == 0;
  if (R >= 2 && R <= 4 && aligned && (C == 64 || C == 128 || C == 256 || C == 512)) {
    const int V = C == 512 ? 2 : 1;
    const int ppw = 32 / (C / (8 * V));
    long long ctas = (items + 8LL * ppw - 1) / (8LL * ppw);
    if (ctas > 148 * 8) ctas = 148 * 8;
    const unsigned gr = (unsigned)ctas;
#define FFVC_DIV(RR, VV) diversity_vec_kernel<RR, VV><<<gr, 256, 0, ST(stream)>>>(CBF(feats), loss_accum, BF(dfeat), B, HW, C, scale)
    if (R == 2 && V == 1) FFVC_DIV(2, 1);
    else if (R == 2) FFVC_DIV(2, 2);
    else if (R == 3 && V == 1) FFVC_DIV(3, 1);
    else if (R == 3) FFVC_DIV(3, 2);
    else if (V == 1) FFVC_DIV(4, 1);
    else FFVC_DIV(4, 2);
#undef FFVC_DIV
  } else if (R <= 4)
    diversity_kernel<4><<<(unsigned)((items + 7) / 8), 256, 0, ST(stream)>>>(CBF(feats), loss_accum, BF(dfeat), R, B, HW, C, scale);
  else
    diversity_anyr_kernel<<<(unsigned)((items + 7) / 8), 256, 0, ST(stream)>>>(CBF(feats), loss_accum, BF(dfeat), R, B, HW, C, scale);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_relu_mask(const void* g, const void* post, void* out, long long n, void* stream) {
  relu_mask_kernel<<<grid_for_l(n, 256), 256, 0, ST(stream)>>>(CBF(g), CBF(post), BF(out), n);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
