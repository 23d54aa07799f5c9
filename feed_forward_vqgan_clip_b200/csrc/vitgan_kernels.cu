// Kernels specific to the VitGAN mapper (vitgan.py): self-modulated LayerNorm (SLN), the mapper's small multi-head
// attention with the reference's interleaved '(d k h)' projection layout, and a pitched fp32->bf16 cast for the
// 1020-wide projection (TMA needs 16-byte row strides).  The token count is tiny (T = 16), so these are plain
// coalesced SIMT kernels; all GEMMs of the mapper run on ffvc_gemm.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cfloat>

#include "ffvc_internal.h"
#include "ptx.cuh"

namespace ffvc {

static inline unsigned grid_for_v(long long n, int threads, int cap = 148 * 16) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  return (unsigned)(g < cap ? g : cap);
}

// SLN(hl, w) = gamma * w * n + beta * w with n = LayerNorm(hl) (vitgan.py:20-21): s = w * (gamma * n + beta)
__global__ void sln_mod_fwd_kernel(const __nv_bfloat16* __restrict__ n, const __nv_bfloat16* __restrict__ w,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   __nv_bfloat16* __restrict__ s, long long total) {
  const float g = gamma[0], b = beta[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    s[i] = __float2bfloat16(__bfloat162float(w[i]) * (g * __bfloat162float(n[i]) + b));
}
// ds -> dn = ds * w * gamma (bf16);  dw_acc += ds * (gamma * n + beta) (fp32, accumulated over every SLN that uses w);
// dgamma += sum ds * w * n;  dbeta += sum ds * w
__global__ void __launch_bounds__(256) sln_mod_bwd_kernel(const __nv_bfloat16* __restrict__ ds, const __nv_bfloat16* __restrict__ n,
                                                          const __nv_bfloat16* __restrict__ w, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, __nv_bfloat16* __restrict__ dn,
                                                          float* __restrict__ dw_acc, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, long long total) {
  const float g = gamma[0], b = beta[0];
  float ag = 0.f, ab = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float d = __bfloat162float(ds[i]), nv = __bfloat162float(n[i]), wv = __bfloat162float(w[i]);
    dn[i] = __float2bfloat16(d * wv * g);
    dw_acc[i] += d * (g * nv + b);
    ag += d * wv * nv;
    ab += d * wv;
  }
  ag = warp_sum(ag);
  ab = warp_sum(ab);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(dgamma, ag);
    atomicAdd(dbeta, ab);
  }
}

// Attention of vitgan.py:81-97.  qkv: [B][T][ld_qkv] bf16 with column index d*3H + k*H + h (k = 0 q, 1 k, 2 v);
// out: [B][T][ld_out] bf16 with column h*dh + d.  One CTA per (b, h); T <= 32, dh <= 256.
__global__ void __launch_bounds__(128) vitgan_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                              float* __restrict__ probs, int T, int H, int dh, int ld_qkv,
                                                              int ld_out, float scale) {
  extern __shared__ float sm[];
  float* q = sm;                 // [T][dh]
  float* k = q + T * dh;
  float* v = k + T * dh;
  float* s = v + T * dh;         // [T][T]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld_qkv;
  for (int i = threadIdx.x; i < T * dh; i += blockDim.x) {
    const int t = i / dh, d = i % dh;
    const __nv_bfloat16* r = base + (long long)t * ld_qkv + d * 3 * H + h;
    q[i] = __bfloat162float(r[0]);
    k[i] = __bfloat162float(r[H]);
    v[i] = __bfloat162float(r[2 * H]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int a = i / T, c = i % T;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(q[a * dh + d], k[c * dh + d], acc);
    s[i] = acc * scale;
  }
  __syncthreads();
  if (threadIdx.x < T) {
    const int a = threadIdx.x;
    float mx = -FLT_MAX;
    for (int c = 0; c < T; ++c) mx = fmaxf(mx, s[a * T + c]);
    float sum = 0.f;
    for (int c = 0; c < T; ++c) {
      const float e = __expf(s[a * T + c] - mx);
      s[a * T + c] = e;
      sum += e;
    }
    const float inv = 1.0f / sum;
    for (int c = 0; c < T; ++c) {
      s[a * T + c] *= inv;
      probs[((long long)blockIdx.x * T + a) * T + c] = s[a * T + c];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * dh; i += blockDim.x) {
    const int a = i / dh, d = i % dh;
    float acc = 0.f;
    for (int c = 0; c < T; ++c) acc = fmaf(s[a * T + c], v[c * dh + d], acc);
    out[((long long)b * T + a) * ld_out + h * dh + d] = __float2bfloat16(acc);
  }
}

__global__ void __launch_bounds__(128) vitgan_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ probs,
                                                              const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqkv,
                                                              int T, int H, int dh, int ld_qkv, int ld_out, float scale) {
  extern __shared__ float sm[];
  float* q = sm;
  float* k = q + T * dh;
  float* v = k + T * dh;
  float* dO = v + T * dh;
  float* p = dO + T * dh;        // [T][T]
  float* dsm = p + T * T;        // [T][T]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const __nv_bfloat16* base = qkv + (long long)b * T * ld_qkv;
  for (int i = threadIdx.x; i < T * dh; i += blockDim.x) {
    const int t = i / dh, d = i % dh;
    const __nv_bfloat16* r = base + (long long)t * ld_qkv + d * 3 * H + h;
    q[i] = __bfloat162float(r[0]);
    k[i] = __bfloat162float(r[H]);
    v[i] = __bfloat162float(r[2 * H]);
    dO[i] = __bfloat162float(dout[((long long)b * T + t) * ld_out + h * dh + d]);
  }
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) p[i] = probs[(long long)blockIdx.x * T * T + i];
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int a = i / T, c = i % T;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(dO[a * dh + d], v[c * dh + d], acc);
    dsm[i] = acc;   // dP
  }
  __syncthreads();
  if (threadIdx.x < T) {
    const int a = threadIdx.x;
    float dot = 0.f;
    for (int c = 0; c < T; ++c) dot += p[a * T + c] * dsm[a * T + c];
    for (int c = 0; c < T; ++c) dsm[a * T + c] = p[a * T + c] * (dsm[a * T + c] - dot) * scale;
  }
  __syncthreads();
  __nv_bfloat16* ob = dqkv + (long long)b * T * ld_qkv;
  for (int i = threadIdx.x; i < T * dh; i += blockDim.x) {
    const int a = i / dh, d = i % dh;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int c = 0; c < T; ++c) {
      dq = fmaf(dsm[a * T + c], k[c * dh + d], dq);
      dk = fmaf(dsm[c * T + a], q[c * dh + d], dk);
      dv = fmaf(p[c * T + a], dO[c * dh + d], dv);
    }
    __nv_bfloat16* r = ob + (long long)a * ld_qkv + d * 3 * H + h;
    r[0] = __float2bfloat16(dq);
    r[H] = __float2bfloat16(dk);
    r[2 * H] = __float2bfloat16(dv);
  }
}

// dst[r][c] = bf16(src[r][c]) for c < cols, 0 for cols <= c < ld  (row pitch ld, a multiple of 8)
__global__ void cast_pitched_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols, int ld) {
  const long long total = (long long)rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    const long long r = i / ld;
    dst[i] = __float2bfloat16(c < cols ? src[r * cols + c] : 0.f);
  }
}
// y[b][i] = x[i] for every b (pos_emb1D has no batch dimension, vitgan.py:256)
__global__ void broadcast_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, long long n) {
  const long long total = (long long)B * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i % n]);
}


// ---- SimpleGenerator (vitgan.py:262-305): the same attention at T = size*size tokens runs as batched tcgen05 GEMMs, which need
// head-contiguous, 16-byte aligned q | k | v.  The projection WEIGHTS are re-packed (cheap, once per forward) instead of the
// activations: row (k, h, d) of the packed to_qkv weight is row d*3H + k*H + h of the reference layout ('(d k h)', vitgan.py:82),
// with the head dimension padded from dh to dhp (zero rows => zero q/k/v columns); w_out's columns are padded the same way.
__global__ void pack_qkv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int H, int dh, int dhp, int D) {
  const long long total = 3LL * H * dhp * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % D);
    const long long r = i / D;
    const int d = (int)(r % dhp), kh = (int)(r / dhp);
    wp[i] = __float2bfloat16(d < dh ? w[((long long)d * 3 * H + kh) * D + c] : 0.f);
  }
}
__global__ void unpack_qkv_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int H, int dh, int dhp, int D) {
  const long long total = 3LL * H * dh * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % D);
    const long long r = i / D;                       // reference row d*3H + kh
    const int kh = (int)(r % (3 * H)), d = (int)(r / (3 * H));
    dw[i] += dwp[((long long)kh * dhp + d) * D + c];
  }
}
__global__ void pack_out_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int H, int dh, int dhp, int D) {
  const long long total = (long long)D * H * dhp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (H * dhp));
    const long long r = i / (H * dhp);
    const int d = col % dhp, h = col / dhp;
    wp[i] = __float2bfloat16(d < dh ? w[r * (H * dh) + h * dh + d] : 0.f);
  }
}
__global__ void unpack_out_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int H, int dh, int dhp, int D) {
  const long long total = (long long)D * H * dh;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (H * dh));
    const long long r = i / (H * dh);
    const int d = col % dh, h = col / dh;
    dw[i] += dwp[r * (H * dhp) + h * dhp + d];
  }
}

}  // namespace ffvc

using namespace ffvc;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int ffvc_sln_mod_fwd(const void* n, const void* w, const float* gamma, const float* beta, void* s, long long total,
                                void* stream) {
  sln_mod_fwd_kernel<<<grid_for_v(total, 256), 256, 0, ST(stream)>>>(CBF(n), CBF(w), gamma, beta, BF(s), total);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_sln_mod_bwd(const void* ds, const void* n, const void* w, const float* gamma, const float* beta, void* dn,
                                float* dw_acc, float* dgamma, float* dbeta, long long total, void* stream) {
  sln_mod_bwd_kernel<<<grid_for_v(total, 256, 148 * 4), 256, 0, ST(stream)>>>(CBF(ds), CBF(n), CBF(w), gamma, beta, BF(dn), dw_acc,
                                                                              dgamma, dbeta, total);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
static int attn_check(int T, int dh) {
  if (T < 1 || T > 32 || dh < 1 || dh > 256) return set_error(FFVC_ERR_UNSUPPORTED, "vitgan_attn: T <= 32 and head dim <= 256");
  return FFVC_OK;
}
extern "C" int ffvc_vitgan_attn_fwd(const void* qkv, void* out, float* probs, int B, int T, int H, int dh, int ld_qkv, int ld_out,
                                    float scale, void* stream) {
  int rc = attn_check(T, dh);
  if (rc) return rc;
  const size_t smem = (size_t)(3 * T * dh + T * T) * sizeof(float);
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(vitgan_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(vitgan_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    done = true;
  }
  vitgan_attn_fwd_kernel<<<B * H, 128, smem, ST(stream)>>>(CBF(qkv), BF(out), probs, T, H, dh, ld_qkv, ld_out, scale);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_vitgan_attn_bwd(const void* qkv, const float* probs, const void* dout, void* dqkv, int B, int T, int H, int dh,
                                    int ld_qkv, int ld_out, float scale, void* stream) {
  int rc = attn_check(T, dh);
  if (rc) return rc;
  const size_t smem = (size_t)(4 * T * dh + 2 * T * T) * sizeof(float);
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(vitgan_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(vitgan_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    done = true;
  }
  vitgan_attn_bwd_kernel<<<B * H, 128, smem, ST(stream)>>>(CBF(qkv), probs, CBF(dout), BF(dqkv), T, H, dh, ld_qkv, ld_out, scale);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_cast_f32_bf16_pitched(const float* src, void* dst, int rows, int cols, int ld, void* stream) {
  if (ld < cols) return set_error(FFVC_ERR_ARG, "cast_pitched: ld < cols");
  cast_pitched_kernel<<<grid_for_v((long long)rows * ld, 256), 256, 0, ST(stream)>>>(src, BF(dst), rows, cols, ld);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_broadcast_rows(const float* x, void* y, int B, long long n, void* stream) {
  broadcast_rows_kernel<<<grid_for_v((long long)B * n, 256), 256, 0, ST(stream)>>>(x, BF(y), B, n);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

static int pack_check(int H, int dh, int dhp, int D) {
  if (H < 1 || dh < 1 || dhp < dh || D < 1) return set_error(FFVC_ERR_ARG, "vitgan pack: need H >= 1, 1 <= dh <= dhp, D >= 1");
  return FFVC_OK;
}
extern "C" int ffvc_vitgan_pack_qkv_weight(const float* w, void* wp, int H, int dh, int dhp, int D, void* stream) {
  int rc = pack_check(H, dh, dhp, D);
  if (rc) return rc;
  pack_qkv_weight_kernel<<<grid_for_v(3LL * H * dhp * D, 256), 256, 0, ST(stream)>>>(w, BF(wp), H, dh, dhp, D);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_vitgan_unpack_qkv_wgrad(const float* dwp, float* dw, int H, int dh, int dhp, int D, void* stream) {
  int rc = pack_check(H, dh, dhp, D);
  if (rc) return rc;
  unpack_qkv_wgrad_kernel<<<grid_for_v(3LL * H * dh * D, 256), 256, 0, ST(stream)>>>(dwp, dw, H, dh, dhp, D);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_vitgan_pack_out_weight(const float* w, void* wp, int H, int dh, int dhp, int D, void* stream) {
  int rc = pack_check(H, dh, dhp, D);
  if (rc) return rc;
  pack_out_weight_kernel<<<grid_for_v((long long)D * H * dhp, 256), 256, 0, ST(stream)>>>(w, BF(wp), H, dh, dhp, D);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_vitgan_unpack_out_wgrad(const float* dwp, float* dw, int H, int dh, int dhp, int D, void* stream) {
  int rc = pack_check(H, dh, dhp, D);
  if (rc) return rc;
  unpack_out_wgrad_kernel<<<grid_for_v((long long)D * H * dh, 256), 256, 0, ST(stream)>>>(dwp, dw, H, dh, dhp, D);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
