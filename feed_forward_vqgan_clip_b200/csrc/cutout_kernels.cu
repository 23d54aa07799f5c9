// MakeCutouts (main.py:154-229) + CLIP normalisation (main.py:797) as gather/scatter HBM kernels.
//   pool      : (AdaptiveAvgPool2d + AdaptiveMaxPool2d) / 2                     main.py:218
//   warp      : bilinear homography resample, border or zero padding            kornia RandomAffine / RandomPerspective
//   final     : perspective warp + hue/saturation jitter + erase + noise + (x-mean)/std, written PATCH-MAJOR in bf16
//               so that the ViT patch embedding (cloob.py:224,237) is a plain GEMM      main.py:171-190,223-225,797
// Images are NHWC fp32 with 3 channels.  All randomness arrives as explicit tensors (host-side sampling policy).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cfloat>

#include "ffvc_internal.h"

namespace ffvc {

#define TWO_PI_F 6.283185307179586f

static inline unsigned grid_for_c(long long n, int threads, int cap = 148 * 16) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  return (unsigned)(g < cap ? g : cap);
}

// fixed-point accumulators of the backward's scatter stages (see scatter3 below)
#define FFVC_FX_SCALE 1099511627776.0f            /* 2^40 */
#define FFVC_FX_INV 9.094947017729282e-13f        /* 2^-40 */
__device__ __forceinline__ void fx_add(long long* p, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__float2ll_rn(v * FFVC_FX_SCALE));
}
__device__ __forceinline__ float fx_get(const long long* p) { return (float)(*p) * FFVC_FX_INV; }

// ------------------------------------------------------------------------------ adaptive (avg+max)/2 pooling
__device__ __forceinline__ int win_start(int i, int in, int out) { return (int)(((long long)i * in) / out); }
__device__ __forceinline__ int win_end(int i, int in, int out) { return (int)((((long long)(i + 1)) * in + out - 1) / out); }

// Index math: every kernel of this file maps blockIdx.y to an output row and blockIdx.z to an image, threads run along x — the
// flat 64-bit index of round 1 cost ~10 emulated 64-bit div / mod per element and made these kernels issue-bound (69 - 82 % issue
// active at 0.4 - 2 TB/s, profiles/r02_ncu_hbm_kernels.md).
__global__ void pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int P) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= P) return;
  const int oy = blockIdx.y, b = blockIdx.z;
  const int y0 = (oy * H) / P, y1 = ((oy + 1) * H + P - 1) / P, x0 = (ox * W) / P, x1 = ((ox + 1) * W + P - 1) / P;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, m0 = -FLT_MAX, m1 = -FLT_MAX, m2 = -FLT_MAX;
  for (int yy = y0; yy < y1; ++yy) {
    const float* row = x + (((long long)b * H + yy) * W + x0) * 3;
    for (int xx = 0; xx < x1 - x0; ++xx) {
      const float v0 = row[3 * xx], v1 = row[3 * xx + 1], v2 = row[3 * xx + 2];
      s0 += v0;
      s1 += v1;
      s2 += v2;
      m0 = fmaxf(m0, v0);
      m1 = fmaxf(m1, v1);
      m2 = fmaxf(m2, v2);
    }
  }
  const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
  float* o = y + (((long long)b * P + oy) * P + ox) * 3;
  o[0] = 0.5f * (s0 * inv + m0);
  o[1] = 0.5f * (s1 * inv + m1);
  o[2] = 0.5f * (s2 * inv + m2);
}
// gather form: each input pixel collects from every pooled output whose window contains it
template <typename TG>
__device__ __forceinline__ float grad_get(const TG* p);
template <>
__device__ __forceinline__ float grad_get<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float grad_get<long long>(const long long* p) { return fx_get(p); }

template <typename TG>
__global__ void pool_bwd_kernel(const float* __restrict__ x, const TG* __restrict__ dy, float* __restrict__ dx, int B, int H,
                                int W, int P) {
  const long long total = (long long)B * H * W * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    long long p = i / 3;
    const int ix = (int)(p % W);
    p /= W;
    const int iy = (int)(p % H);
    const int b = (int)(p / H);
    float acc = 0.f;
    const int oy_lo = max(0, (int)(((long long)iy * P) / H) - 1), oy_hi = min(P - 1, (int)((((long long)(iy + 1)) * P + H - 1) / H));
    const int ox_lo = max(0, (int)(((long long)ix * P) / W) - 1), ox_hi = min(P - 1, (int)((((long long)(ix + 1)) * P + W - 1) / W));
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const int y0 = win_start(oy, H, P), y1 = win_end(oy, H, P);
      if (iy < y0 || iy >= y1) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const int x0 = win_start(ox, W, P), x1 = win_end(ox, W, P);
        if (ix < x0 || ix >= x1) continue;
        const float g = grad_get(dy + (((long long)b * P + oy) * P + ox) * 3 + c);
        acc += 0.5f * g / ((y1 - y0) * (x1 - x0));
        // arg max of the window (first maximum in row-major order, like ATen's adaptive_max_pool2d)
        float m = -FLT_MAX;
        int my = y0, mx = x0;
        for (int yy = y0; yy < y1; ++yy)
          for (int xx = x0; xx < x1; ++xx) {
            const float v = x[(((long long)b * H + yy) * W + xx) * 3 + c];
            if (v > m) {
              m = v;
              my = yy;
              mx = xx;
            }
          }
        if (my == iy && mx == ix) acc += 0.5f * g;
      }
    }
    dx[i] = acc;
  }
}

// Same gather, laid out so that no 64-bit division is needed and the window geometry is shared by the 3 channels:
// blockIdx.y = (image, input row), threads run along x.  (The flat form above spends most of its time in emulated
// 64-bit div/mod: ~10 of them per element.)  Identical arithmetic per element.
template <typename TG>
__global__ void __launch_bounds__(128) pool_bwd_rows_kernel(const float* __restrict__ x, const TG* __restrict__ dy,
                                                            float* __restrict__ dx, int H, int W, int P) {
  const int ix = blockIdx.x * 128 + threadIdx.x;
  if (ix >= W) return;
  const int b = blockIdx.y / H, iy = blockIdx.y - b * H;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  const int oy_lo = max(0, (iy * P) / H - 1), oy_hi = min(P - 1, ((iy + 1) * P + H - 1) / H);
  const int ox_lo = max(0, (ix * P) / W - 1), ox_hi = min(P - 1, ((ix + 1) * P + W - 1) / W);
  const float* xb = x + (long long)b * H * W * 3;
  const int me = iy * W + ix;
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    const int y0 = (oy * H) / P, y1 = ((oy + 1) * H + P - 1) / P;
    if (iy < y0 || iy >= y1) continue;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const int x0 = (ox * W) / P, x1 = ((ox + 1) * W + P - 1) / P;
      if (ix < x0 || ix >= x1) continue;
      const TG* gp = dy + (((long long)b * P + oy) * P + ox) * 3;
      const float g0 = grad_get(gp), g1 = grad_get(gp + 1), g2 = grad_get(gp + 2);
      const int area = (y1 - y0) * (x1 - x0);
      acc0 += 0.5f * g0 / area;
      acc1 += 0.5f * g1 / area;
      acc2 += 0.5f * g2 / area;
      // arg max of the window per channel (first maximum in row-major order, like ATen's adaptive_max_pool2d)
      float m0 = -FLT_MAX, m1 = -FLT_MAX, m2 = -FLT_MAX;
      int p0 = y0 * W + x0, p1 = p0, p2 = p0;
      for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) {
          const int pos = yy * W + xx;
          const float* xp = xb + (long long)pos * 3;
          const float v0 = xp[0], v1 = xp[1], v2 = xp[2];
          if (v0 > m0) { m0 = v0; p0 = pos; }
          if (v1 > m1) { m1 = v1; p1 = pos; }
          if (v2 > m2) { m2 = v2; p2 = pos; }
        }
      if (p0 == me) acc0 += 0.5f * g0;
      if (p1 == me) acc1 += 0.5f * g1;
      if (p2 == me) acc2 += 0.5f * g2;
    }
  }
  float* o = dx + ((long long)b * H * W + me) * 3;
  o[0] = acc0;
  o[1] = acc1;
  o[2] = acc2;
}

// ------------------------------------------------------------------------------ bilinear homography sampling
struct Taps {
  int x0, y0, x1, y1;
  float w00, w01, w10, w11;  // weights for (y0,x0), (y0,x1), (y1,x0), (y1,x1); zero if tap is outside (zeros mode)
};
__device__ __forceinline__ Taps make_taps(const float* __restrict__ hm, int ox, int oy, int P, int border) {
  const float fx = (float)ox, fy = (float)oy;
  const float den = hm[6] * fx + hm[7] * fy + hm[8];
  float sx = (hm[0] * fx + hm[1] * fy + hm[2]) / den;
  float sy = (hm[3] * fx + hm[4] * fy + hm[5]) / den;
  if (border) {
    sx = fminf(fmaxf(sx, 0.f), (float)(P - 1));
    sy = fminf(fmaxf(sy, 0.f), (float)(P - 1));
  }
  const float flx = floorf(sx), fly = floorf(sy);
  const float ax = sx - flx, ay = sy - fly;
  Taps t;
  // keep the integer conversion safe for wild coordinates
  const float cl = 4.0f * P;
  t.x0 = (int)fminf(fmaxf(flx, -cl), cl);
  t.y0 = (int)fminf(fmaxf(fly, -cl), cl);
  t.x1 = t.x0 + 1;
  t.y1 = t.y0 + 1;
  t.w00 = (1.f - ax) * (1.f - ay);
  t.w01 = ax * (1.f - ay);
  t.w10 = (1.f - ax) * ay;
  t.w11 = ax * ay;
  const bool vx0 = t.x0 >= 0 && t.x0 < P, vx1 = t.x1 >= 0 && t.x1 < P;
  const bool vy0 = t.y0 >= 0 && t.y0 < P, vy1 = t.y1 >= 0 && t.y1 < P;
  if (!(vx0 && vy0)) t.w00 = 0.f;
  if (!(vx1 && vy0)) t.w01 = 0.f;
  if (!(vx0 && vy1)) t.w10 = 0.f;
  if (!(vx1 && vy1)) t.w11 = 0.f;
  t.x0 = min(max(t.x0, 0), P - 1);
  t.x1 = min(max(t.x1, 0), P - 1);
  t.y0 = min(max(t.y0, 0), P - 1);
  t.y1 = min(max(t.y1, 0), P - 1);
  return t;
}
__device__ __forceinline__ void sample3(const float* __restrict__ img, const Taps& t, int P, float (&o)[3]) {
  const float* a = img + ((long long)t.y0 * P + t.x0) * 3;
  const float* b = img + ((long long)t.y0 * P + t.x1) * 3;
  const float* c = img + ((long long)t.y1 * P + t.x0) * 3;
  const float* d = img + ((long long)t.y1 * P + t.x1) * 3;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) o[ch] = t.w00 * a[ch] + t.w01 * b[ch] + t.w10 * c[ch] + t.w11 * d[ch];
}
// The two bilinear stages scatter in the backward (several output pixels sample one source pixel).  Float atomics would make
// d(image) depend on the order of arrival — and one flipped bf16 rounding downstream grows to the bf16 noise floor within a few
// layers (profiles/r02_parity_fullsize.md).  The scatter targets are therefore 64-bit FIXED-POINT accumulators (value x 2^40:
// integer addition commutes, so the sums are bit-identical from run to run and for any sharding of the batch; resolution 9e-13,
// range +-8e6, against gradient elements of 1e-8 .. 1e-2); the consumer of each buffer converts back while reading.
__device__ __forceinline__ void scatter3(long long* __restrict__ img, const Taps& t, int P, const float (&g)[3]) {
  long long* a = img + ((long long)t.y0 * P + t.x0) * 3;
  long long* b = img + ((long long)t.y0 * P + t.x1) * 3;
  long long* c = img + ((long long)t.y1 * P + t.x0) * 3;
  long long* d = img + ((long long)t.y1 * P + t.x1) * 3;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (t.w00 != 0.f) fx_add(a + ch, t.w00 * g[ch]);
    if (t.w01 != 0.f) fx_add(b + ch, t.w01 * g[ch]);
    if (t.w10 != 0.f) fx_add(c + ch, t.w10 * g[ch]);
    if (t.w11 != 0.f) fx_add(d + ch, t.w11 * g[ch]);
  }
}

// out[n] = warp(in[n % n_src], hinv[n])
__global__ void warp_fwd_kernel(const float* __restrict__ in, const float* __restrict__ hinv, float* __restrict__ out, int N,
                                int n_src, int P, int border) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= P) return;
  const int oy = blockIdx.y, n = blockIdx.z;
  const long long i = ((long long)n * P + oy) * P + ox;
  const Taps t = make_taps(hinv + n * 9, ox, oy, P, border);
  float o[3];
  sample3(in + (long long)(n % n_src) * P * P * 3, t, P, o);
  out[i * 3 + 0] = o[0];
  out[i * 3 + 1] = o[1];
  out[i * 3 + 2] = o[2];
}
// din[n % n_src] += warp^T(dout[n])   (both fixed-point, din zeroed by the caller)
__global__ void warp_bwd_kernel(const long long* __restrict__ dout, const float* __restrict__ hinv, long long* __restrict__ din, int N,
                                int n_src, int P, int border) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= P) return;
  const int oy = blockIdx.y, n = blockIdx.z;
  const long long i = ((long long)n * P + oy) * P + ox;
  const Taps t = make_taps(hinv + n * 9, ox, oy, P, border);
  const float g[3] = {fx_get(dout + i * 3), fx_get(dout + i * 3 + 1), fx_get(dout + i * 3 + 2)};
  scatter3(din + (long long)(n % n_src) * P * P * 3, t, P, g);
}

// ------------------------------------------------------------------------------ hue / saturation jitter with forward-mode
// derivatives (value + d/d(r,g,b)) so the backward kernel gets the exact 3x3 Jacobian of the HSV round trip.
struct D3 {
  float v, d0, d1, d2;
};
__device__ __forceinline__ D3 dconst(float c) { return {c, 0.f, 0.f, 0.f}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.v + b.v, a.d0 + b.d0, a.d1 + b.d1, a.d2 + b.d2}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.v - b.v, a.d0 - b.d0, a.d1 - b.d1, a.d2 - b.d2}; }
__device__ __forceinline__ D3 operator*(D3 a, D3 b) {
  return {a.v * b.v, a.d0 * b.v + a.v * b.d0, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2};
}
__device__ __forceinline__ D3 operator/(D3 a, D3 b) {
  const float inv = 1.0f / b.v;
  const float q = a.v * inv;
  return {q, (a.d0 - q * b.d0) * inv, (a.d1 - q * b.d1) * inv, (a.d2 - q * b.d2) * inv};
}
__device__ __forceinline__ D3 operator*(float s, D3 a) { return {s * a.v, s * a.d0, s * a.d1, s * a.d2}; }
__device__ __forceinline__ D3 shift(D3 a, float c) { return {a.v + c, a.d0, a.d1, a.d2}; }

__device__ __forceinline__ void jitter(D3 r, D3 g, D3 b, float sat, float hue, D3& ro, D3& go, D3& bo) {
  // rgb -> hsv (kornia.color.rgb_to_hsv)
  int am = 0;
  D3 maxc = r;
  if (g.v > maxc.v) {
    maxc = g;
    am = 1;
  }
  if (b.v > maxc.v) {
    maxc = b;
    am = 2;
  }
  D3 minc = r;
  if (g.v < minc.v) minc = g;
  if (b.v < minc.v) minc = b;
  const D3 v = maxc;
  D3 deltac = maxc - minc;
  D3 s = deltac / shift(v, 1e-6f);
  if (deltac.v == 0.f) deltac = dconst(1.f);
  const D3 rc = maxc - r, gc = maxc - g, bc = maxc - b;
  D3 h;
  if (am == 0) h = bc - gc;
  else if (am == 1) h = 2.f * deltac + rc - bc;
  else h = 4.f * deltac + gc - rc;
  h = h / deltac;
  h = (1.0f / 6.0f) * h;
  h = shift(h, -floorf(h.v));  // python-style % 1.0
  h = TWO_PI_F * h;
  // jitter
  s = sat * s;
  if (s.v < 0.f) s = dconst(0.f);
  else if (s.v > 1.f) s = dconst(1.f);
  h = shift(h, hue + TWO_PI_F);
  h = shift(h, -TWO_PI_F * truncf(h.v / TWO_PI_F));  // fmod(h, 2pi), h >= 0
  // hsv -> rgb (kornia.color.hsv_to_rgb)
  D3 h6 = (6.0f / TWO_PI_F) * h;
  const float fl = floorf(h6.v);
  int hi = ((int)fl) % 6;
  if (hi < 0) hi += 6;
  const float m6 = h6.v - 6.0f * floorf(h6.v / 6.0f);  // h6 % 6
  D3 f = shift(h6, (m6 - (float)hi) - h6.v);
  const D3 one = dconst(1.f);
  const D3 p = v * (one - s);
  const D3 q = v * (one - f * s);
  const D3 t = v * (one - (one - f) * s);
  switch (hi) {
    case 0: ro = v; go = t; bo = p; break;
    case 1: ro = q; go = v; bo = p; break;
    case 2: ro = p; go = v; bo = t; break;
    case 3: ro = p; go = q; bo = v; break;
    case 4: ro = t; go = p; bo = v; break;
    default: ro = v; go = p; bo = q; break;
  }
}

struct FinalParams {
  const float* cut1;    // [N][P][P][3] image after the affine warp
  const float* hinv;    // [N][9] perspective inverse homographies
  const float* sat;     // [N]
  const float* hue;     // [N]
  const float* noise;   // [N][3][P][P]  (standard normal, NCHW like randn_like(batch))
  const float* facs;    // [N]           (U(0, noise_fac))
  const int* erase;     // DEVICE int[4]: x0, y0, x1, y1 (empty if x1 <= x0)
  float mean[3], istd[3];
  int N, P, patch, grid;   // grid = P / patch
};

// forward: -> patches [N][grid*grid][3*patch*patch] bf16   (k = c*patch^2 + py*patch + px)
__global__ void cutout_final_fwd_kernel(FinalParams fp, __nv_bfloat16* __restrict__ patches, float* __restrict__ img_out) {
  const int P = fp.P;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= P) return;
  {
    const int oy = blockIdx.y, n = blockIdx.z;
    const Taps t = make_taps(fp.hinv + n * 9, ox, oy, P, 0);
    float c[3];
    sample3(fp.cut1 + (long long)n * P * P * 3, t, P, c);
    D3 ro, go, bo;
    jitter({c[0], 1, 0, 0}, {c[1], 0, 1, 0}, {c[2], 0, 0, 1}, fp.sat[n], fp.hue[n], ro, go, bo);
    float o[3] = {ro.v, go.v, bo.v};
    const bool erased = ox >= fp.erase[0] && ox < fp.erase[2] && oy >= fp.erase[1] && oy < fp.erase[3];
    const float fac = fp.facs[n];
    const int pp = fp.patch * fp.patch;
    const int pidx = (oy / fp.patch) * fp.grid + ox / fp.patch;
    const int kin = (oy % fp.patch) * fp.patch + ox % fp.patch;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float val = erased ? 0.f : o[ch];
      val += fac * fp.noise[(((long long)n * 3 + ch) * P + oy) * P + ox];
      val = (val - fp.mean[ch]) * fp.istd[ch];
      patches[((long long)n * fp.grid * fp.grid + pidx) * (3 * pp) + ch * pp + kin] = __float2bfloat16(val);
      if (img_out) img_out[(((long long)n * 3 + ch) * P + oy) * P + ox] = val;
    }
  }
}
// backward: dpatches (bf16, same layout) -> dcut1 (fixed-point accumulators, zeroed by caller)
__global__ void cutout_final_bwd_kernel(FinalParams fp, const __nv_bfloat16* __restrict__ dpatches, long long* __restrict__ dcut1) {
  const int P = fp.P;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= P) return;
  {
    const int oy = blockIdx.y, n = blockIdx.z;
    const bool erased = ox >= fp.erase[0] && ox < fp.erase[2] && oy >= fp.erase[1] && oy < fp.erase[3];
    if (erased) return;
    const Taps t = make_taps(fp.hinv + n * 9, ox, oy, P, 0);
    float c[3];
    sample3(fp.cut1 + (long long)n * P * P * 3, t, P, c);
    D3 ro, go, bo;
    jitter({c[0], 1, 0, 0}, {c[1], 0, 1, 0}, {c[2], 0, 0, 1}, fp.sat[n], fp.hue[n], ro, go, bo);
    const int pp = fp.patch * fp.patch;
    const int pidx = (oy / fp.patch) * fp.grid + ox / fp.patch;
    const int kin = (oy % fp.patch) * fp.patch + ox % fp.patch;
    float go_[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      go_[ch] = __bfloat162float(dpatches[((long long)n * fp.grid * fp.grid + pidx) * (3 * pp) + ch * pp + kin]) * fp.istd[ch];
    float gi[3];
    gi[0] = go_[0] * ro.d0 + go_[1] * go.d0 + go_[2] * bo.d0;
    gi[1] = go_[0] * ro.d1 + go_[1] * go.d1 + go_[2] * bo.d1;
    gi[2] = go_[0] * ro.d2 + go_[1] * go.d2 + go_[2] * bo.d2;
    scatter3(dcut1 + (long long)n * P * P * 3, t, P, gi);
  }
}

}  // namespace ffvc

using namespace ffvc;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
// one block (of whole warps, at most 256 threads) per stretch of an output row; grid = (stretches, rows, images)
static inline unsigned row_block(int P) { return (unsigned)(P >= 256 ? 256 : (P + 31) / 32 * 32); }
static inline dim3 row_grid(int P, int n) { return dim3((unsigned)((P + (int)row_block(P) - 1) / (int)row_block(P)), (unsigned)P, (unsigned)n); }

extern "C" int ffvc_cutout_pool_fwd(const float* x, float* y, int B, int H, int W, int P, void* stream) {
  if (B > 65535 || P > 65535) return set_error(FFVC_ERR_ARG, "cutout_pool_fwd: more than 65535 images or rows");
  pool_fwd_kernel<<<row_grid(P, B), row_block(P), 0, ST(stream)>>>(x, y, B, H, W, P);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
template <typename TG>
static void pool_bwd_launch(const float* x, const TG* dy, float* dx, int B, int H, int W, int P, cudaStream_t st) {
  if (option(OPT_POOL_V2) && (long long)B * H <= 65535 && (long long)(H > W ? H : W) * (P + 1) < (1LL << 30))
    pool_bwd_rows_kernel<TG><<<dim3((unsigned)((W + 127) / 128), (unsigned)(B * H)), 128, 0, st>>>(x, dy, dx, H, W, P);
  else
    pool_bwd_kernel<TG><<<grid_for_c((long long)B * H * W * 3, 256), 256, 0, st>>>(x, dy, dx, B, H, W, P);
}
// dy_fixed != 0: dy is the fixed-point (x 2^40, int64) buffer ffvc_cutout_warp_bwd leaves; 0: plain fp32
extern "C" int ffvc_cutout_pool_bwd(const float* x, const void* dy, float* dx, int B, int H, int W, int P, int dy_fixed, void* stream) {
  if (dy_fixed) pool_bwd_launch(x, reinterpret_cast<const long long*>(dy), dx, B, H, W, P, ST(stream));
  else pool_bwd_launch(x, reinterpret_cast<const float*>(dy), dx, B, H, W, P, ST(stream));
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_cutout_warp_fwd(const float* in, const float* hinv, float* out, int N, int n_src, int P, int border,
                                    void* stream) {
  if (N > 65535 || P > 65535) return set_error(FFVC_ERR_ARG, "cutout_warp: more than 65535 cutouts or rows");
  warp_fwd_kernel<<<row_grid(P, N), row_block(P), 0, ST(stream)>>>(in, hinv, out, N, n_src, P, border);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_cutout_warp_bwd(const long long* dout, const float* hinv, long long* din, int N, int n_src, int P, int border,
                                    void* stream) {
  cudaMemsetAsync(din, 0, sizeof(long long) * (size_t)n_src * P * P * 3, ST(stream));
  if (N > 65535 || P > 65535) return set_error(FFVC_ERR_ARG, "cutout_warp: more than 65535 cutouts or rows");
  warp_bwd_kernel<<<row_grid(P, N), row_block(P), 0, ST(stream)>>>(dout, hinv, din, N, n_src, P, border);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

static int fill_final(FinalParams& fp, const float* cut1, const float* hinv, const float* sat, const float* hue,
                      const float* noise, const float* facs, const int* erase, const float* mean, const float* std_, int N,
                      int P, int patch) {
  if (P % patch) return set_error(FFVC_ERR_ARG, "cutout_final: cut size must be a multiple of the patch size");
  fp.cut1 = cut1;
  fp.hinv = hinv;
  fp.sat = sat;
  fp.hue = hue;
  fp.noise = noise;
  fp.facs = facs;
  fp.erase = erase;
  for (int i = 0; i < 3; ++i) {
    fp.mean[i] = mean[i];
    fp.istd[i] = 1.0f / std_[i];
  }
  fp.N = N;
  fp.P = P;
  fp.patch = patch;
  fp.grid = P / patch;
  return FFVC_OK;
}
// erase is a DEVICE int[4]; mean / std are HOST pointers (3 floats each).  img_out (optional, [N][3][P][P] fp32) receives the
// normalised cutouts in the reference's NCHW layout (for parity tests / callers that want the tensor).
extern "C" int ffvc_cutout_final_fwd(const float* cut1, const float* hinv, const float* sat, const float* hue,
                                     const float* noise, const float* facs, const int* erase, const float* mean,
                                     const float* std_, void* patches, float* img_out, int N, int P, int patch, void* stream) {
  FinalParams fp;
  int rc = fill_final(fp, cut1, hinv, sat, hue, noise, facs, erase, mean, std_, N, P, patch);
  if (rc) return rc;
  if (N > 65535 || P > 65535) return set_error(FFVC_ERR_ARG, "cutout_final: more than 65535 cutouts or rows");
  cutout_final_fwd_kernel<<<row_grid(P, N), row_block(P), 0, ST(stream)>>>(fp, reinterpret_cast<__nv_bfloat16*>(patches), img_out);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_cutout_final_bwd(const float* cut1, const float* hinv, const float* sat, const float* hue,
                                     const int* erase, const float* mean, const float* std_, const void* dpatches,
                                     long long* dcut1, int N, int P, int patch, void* stream) {
  FinalParams fp;
  int rc = fill_final(fp, cut1, hinv, sat, hue, nullptr, nullptr, erase, mean, std_, N, P, patch);
  if (rc) return rc;
  cudaMemsetAsync(dcut1, 0, sizeof(long long) * (size_t)N * P * P * 3, ST(stream));
  if (N > 65535 || P > 65535) return set_error(FFVC_ERR_ARG, "cutout_final: more than 65535 cutouts or rows");
  cutout_final_bwd_kernel<<<row_grid(P, N), row_block(P), 0, ST(stream)>>>(fp, reinterpret_cast<const __nv_bfloat16*>(dpatches), dcut1);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
