// Spherical-distance loss between the normalised image embedding and the normalised target embedding,
// forward and backward fused in one pass (main.py:801-811):
//   H = normalize(out_feats.repeat(cutn,1)); e = normalize(embed); loss = coef * mean(2 * asin(|H - e| / 2)^2)
// One warp per embedding row; produces d(loss)/d(embed) directly.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "ffvc_internal.h"
#include "ptx.cuh"

namespace ffvc {

// embed: [N][D] fp32, target: [B][D] fp32 (row n uses target[n % B]).  loss_out: one float (accumulated, zeroed by caller).
// target2 / coef2 (optional): the reference's `input_loss` term (main.py:812-824) — the same distance to a second target
// (the source embeddings) added with its own coefficient; both terms share the normalised image embedding.
__global__ void __launch_bounds__(256) spherical_loss_kernel(const float* __restrict__ embed, const float* __restrict__ target,
                                                             const float* __restrict__ target2, float* __restrict__ loss_out,
                                                             float* __restrict__ dembed, __nv_bfloat16* __restrict__ dembed_bf16,
                                                             int N, int B, int D, float coef, float coef2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + warp;
  if (n >= N) return;
  const float* e = embed + (long long)n * D;
  float se = 0.f;
  for (int i = lane; i < D; i += 32) se += e[i] * e[i];
  se = warp_sum(se);
  const float ie = 1.0f / fmaxf(sqrtf(se), 1e-12f);   // F.normalize eps
  // per target k: it[k] = 1 / |t|, s[k] = coef_k / N * dl/dd / d  (dl/dd = 2 * a / sqrt(1 - d^2/4), a = asin(d / 2))
  float its[2] = {0.f, 0.f}, ss[2] = {0.f, 0.f}, lsum = 0.f;
  const float* ts[2] = {target + (long long)(n % B) * D, target2 ? target2 + (long long)(n % B) * D : nullptr};
  const float cs[2] = {coef, coef2};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float* t = ts[k];
    if (!t) continue;
    float st = 0.f;
    for (int i = lane; i < D; i += 32) st += t[i] * t[i];
    st = warp_sum(st);
    const float it = 1.0f / fmaxf(sqrtf(st), 1e-12f);
    float d2 = 0.f;
    for (int i = lane; i < D; i += 32) {
      const float diff = t[i] * it - e[i] * ie;
      d2 += diff * diff;
    }
    d2 = warp_sum(d2);
    const float d = sqrtf(d2);
    const float half = fminf(0.5f * d, 1.0f);
    const float a = asinf(half);
    lsum += cs[k] * 2.0f * a * a;
    const float dl_dd = (d > 0.f) ? 2.0f * a * rsqrtf(fmaxf(1.0f - half * half, 1e-12f)) : 0.f;
    its[k] = it;
    ss[k] = (d > 0.f) ? (cs[k] / N) * dl_dd / d : 0.f;
  }
  if (lane == 0) atomicAdd(loss_out, lsum / N);
  // g = sum_k s_k * (ehat - Hhat_k);  de = (g - ehat * <g, ehat>) / |e|
  float dot_eh_g = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float eh = e[i] * ie;
    float g = ss[0] * (eh - ts[0][i] * its[0]);
    if (ts[1]) g += ss[1] * (eh - ts[1][i] * its[1]);
    dot_eh_g += g * eh;
  }
  dot_eh_g = warp_sum(dot_eh_g);
  for (int i = lane; i < D; i += 32) {
    const float eh = e[i] * ie;
    float g = ss[0] * (eh - ts[0][i] * its[0]);
    if (ts[1]) g += ss[1] * (eh - ts[1][i] * its[1]);
    const float de = (g - eh * dot_eh_g) * ie;
    if (dembed) dembed[(long long)n * D + i] = de;
    if (dembed_bf16) dembed_bf16[(long long)n * D + i] = __float2bfloat16(de);
  }
}

// y[r] = x[r] / max(|x[r]|, 1e-12): F.normalize(inp_feats, dim=1) of `normalize_input` (main.py:734-735); one warp per row
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int D) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += x[(long long)r * D + i] * x[(long long)r * D + i];
  s = warp_sum(s);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int i = lane; i < D; i += 32) y[(long long)r * D + i] = x[(long long)r * D + i] * inv;
}

// Total-variation loss of main.py:423-428 on an NHWC fp32 image, forward + backward fused:
//   tv = 0.5 * (mean |Y[y+1] - Y[y]| + mean |Y[x+1] - Y[x]|);  dimg += coef * d(tv)/d(img);  loss += coef * tv
__global__ void __launch_bounds__(256) tv_loss_kernel(const float* __restrict__ img, float* __restrict__ loss, float* __restrict__ dimg,
                                                      int B, int H, int W, int C, float coef) {
  const long long total = (long long)B * H * W * C;
  const float wy = 0.5f * coef / ((float)B * C * (H - 1) * W), wx = 0.5f * coef / ((float)B * C * H * (W - 1));
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const float v = img[i];
    float g = 0.f;
    auto sgn = [](float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); };
    if (y > 0) g += wy * sgn(v - img[i - (long long)W * C]);
    if (y < H - 1) {
      const float d = img[i + (long long)W * C] - v;
      g -= wy * sgn(d);
      acc += wy * fabsf(d);
    }
    if (x > 0) g += wx * sgn(v - img[i - C]);
    if (x < W - 1) {
      const float d = img[i + C] - v;
      g -= wx * sgn(d);
      acc += wx * fabsf(d);
    }
    (void)c;
    if (dimg) dimg[i] += g;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && loss) atomicAdd(loss, acc);
}
// y[i] += a * x[i]  (the gradient of l2_coef * mean(z^2), main.py:758-762: a = 2 * l2_coef / numel)
__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, float a, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}

}  // namespace ffvc

using namespace ffvc;

extern "C" int ffvc_spherical_loss(const float* embed, const float* target, float* loss_out, float* dembed,
                                   void* dembed_bf16, int N, int B, int D, float coef, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(loss_out, 0, sizeof(float), st);
  spherical_loss_kernel<<<(N + 7) / 8, 256, 0, st>>>(embed, target, nullptr, loss_out, dembed,
                                                    reinterpret_cast<__nv_bfloat16*>(dembed_bf16), N, B, D, coef, 0.f);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_spherical_loss2(const float* embed, const float* target, const float* target2, float* loss_out, float* dembed,
                                    void* dembed_bf16, int N, int B, int D, float coef, float coef2, void* stream) {
  if (!target2) return set_error(FFVC_ERR_ARG, "spherical_loss2: null second target");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(loss_out, 0, sizeof(float), st);
  spherical_loss_kernel<<<(N + 7) / 8, 256, 0, st>>>(embed, target, target2, loss_out, dembed,
                                                    reinterpret_cast<__nv_bfloat16*>(dembed_bf16), N, B, D, coef, coef2);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_normalize_rows(const float* x, float* y, int rows, int D, void* stream) {
  normalize_rows_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, rows, D);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_tv_loss(const float* img, float* loss_accum, float* dimg_accum, int B, int H, int W, int C, float coef,
                            void* stream) {
  if (H < 2 || W < 2) return set_error(FFVC_ERR_ARG, "tv_loss: image must be at least 2x2");
  const long long total = (long long)B * H * W * C;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  tv_loss_kernel<<<(unsigned)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(img, loss_accum, dimg_accum, B, H, W, C, coef);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_axpy_f32(const float* x, float* y, float a, long long n, void* stream) {
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  axpy_kernel<<<(unsigned)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, a, n);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
