// Coalesced / vectorised HBM-bound kernels of the train step that are not normalisations:
// nearest-2x upsample, batched transpose, row softmax, bias-gradient reductions, casts, ClampWithGrad,
// VQ nearest-code search, image post-processing, the tiny-Cin 3x3 conv (dgrad of conv_out) and fused Adam.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstring>
#include <cfloat>

#include "ffvc_internal.h"
#include "ptx.cuh"

namespace ffvc {

static inline unsigned grid_for(long long n, int threads, int cap = 148 * 16) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  return (unsigned)(g < cap ? g : cap);
}

// ---------------------------------------------------------------- nearest 2x upsample (taming Upsample)
__global__ void upsample2x_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int CV) {
  const long long total = (long long)N * 2 * H * 2 * W * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV);
    long long p = i / CV;
    const int ox = (int)(p % (2 * W));
    p /= (2 * W);
    const int oy = (int)(p % (2 * H));
    const int n = (int)(p / (2 * H));
    y[i] = x[(((long long)n * H + (oy >> 1)) * W + (ox >> 1)) * CV + c];
  }
}
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int H,
                                      int W, int C) {
  const int CV = C >> 3;
  const long long total = (long long)N * H * W * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV);
    long long p = i / CV;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const int n = (int)(p / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy_ = 0; dy_ < 2; ++dy_)
#pragma unroll
      for (int dx_ = 0; dx_ < 2; ++dx_) {
        const uint4 pk = *reinterpret_cast<const uint4*>(
            dy + ((((long long)n * 2 * H + 2 * y + dy_) * 2 * W + 2 * x + dx_) * C + c * 8));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    uint4 o;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0], acc[1]), h1 = __floats2bfloat162_rn(acc[2], acc[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[4], acc[5]), h3 = __floats2bfloat162_rn(acc[6], acc[7]);
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    o.z = *reinterpret_cast<uint32_t*>(&h2);
    o.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(dx + i * 8) = o;
  }
}

// ---------------------------------------------------------------- batched 2-D transpose  in[b][R][Cc] -> out[b][Cc][R]
template <typename Tin, typename Tout>
__global__ void transpose_kernel(const Tin* __restrict__ in, Tout* __restrict__ out, int R, int Cc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const Tin* ib = in + (long long)b * R * Cc;
  Tout* ob = out + (long long)b * R * Cc;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[j][threadIdx.x] = (float)ib[(long long)r * Cc + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < Cc) ob[(long long)c * R + r] = (Tout)tile[threadIdx.x][j];
  }
}

// ---------------------------------------------------------------- row softmax (attention: fp32 scores -> bf16 probs)
// rows are `ld` elements apart, the first n are valid; the padding [n, ld) of the output is zero-filled.
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p,
                                                          long long rows, int n, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float* sr = s + row * ld;
  float mx = -FLT_MAX;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, sr[i]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < n; i += 32) sum += __expf(sr[i] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int i = lane; i < ld; i += 32) p[row * ld + i] = __float2bfloat16(i < n ? __expf(sr[i] - mx) * inv : 0.f);
}
// ds = p * (dp - sum(p*dp)) * scale   (dp fp32 from the dO.V^T GEMM; ds bf16 feeds the dQ/dK GEMMs)
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const __nv_bfloat16* __restrict__ p, const float* __restrict__ dp,
                                                          __nv_bfloat16* __restrict__ ds, long long rows, int n, int ld,
                                                          float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  float dot = 0.f;
  for (int i = lane; i < n; i += 32) dot += __bfloat162float(p[row * ld + i]) * dp[row * ld + i];
  dot = warp_sum(dot);
  for (int i = lane; i < ld; i += 32)
    ds[row * ld + i] = __float2bfloat16(i < n ? __bfloat162float(p[row * ld + i]) * (dp[row * ld + i] - dot) * scale : 0.f);
}

// causal variant (CLIP text transformer, cloob.py:304-310): row r of sequence position t = r % T attends to columns 0..t
__global__ void __launch_bounds__(256) softmax_causal_fwd_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p,
                                                                 long long rows, int T, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const int n = (int)(row % T) + 1;
  const float* sr = s + row * ld;
  float mx = -FLT_MAX;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, sr[i]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < n; i += 32) sum += __expf(sr[i] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int i = lane; i < ld; i += 32) p[row * ld + i] = __float2bfloat16(i < n ? __expf(sr[i] - mx) * inv : 0.f);
}
// token embedding gather + positional embedding (cloob.py:526-528): x[b][t][:] = bf16(emb[tok[b][t]][:] + pos[t][:])
__global__ void embed_tokens_kernel(const long long* __restrict__ tok, const float* __restrict__ emb, const float* __restrict__ pos,
                                    __nv_bfloat16* __restrict__ x, long long rows, int T, int W) {
  const long long total = rows * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / W;
    const int w = (int)(i % W);
    x[i] = __float2bfloat16(emb[tok[r] * W + w] + pos[(r % T) * W + w]);
  }
}
// dst[b][:] = src[b*T + idx[b]][:]   (features at the EOT token, cloob.py:536)
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const long long* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ dst, int B, int T, int W) {
  const long long total = (long long)B * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / W), w = (int)(i % W);
    dst[i] = src[((long long)b * T + idx[b]) * W + w];
  }
}

// ---------------------------------------------------------------- bias gradients
// db[n] += sum_rows dy[row][n], n % 8 == 0.  A CTA (512 threads) owns a slab of <= 32 eight-column vectors (512 B of every
// row) x a chunk of rows: thread = (row lane, column vector), 8 rows in flight per thread (16-byte streaming loads), row
// lanes folded through shared memory, then ONE vector reduction (red.global.add.v4.f32) per 4 columns per CTA.  Splitting
// by column slabs keeps the number of atomics that hit one address at (row chunks) ~ SMs / slabs: the earlier forms (one
// scalar atomic per column from 512 small CTAs; then full-width rows per CTA) were bound by same-line atomic serialisation.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__global__ void __launch_bounds__(512, 2) colsum_vec_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db,
                                                         long long rows, int n, int cvw, int rows_per_cta) {
  extern __shared__ float cs[];                       // [lanes][cvw * 8] partials
  const int ncv = n >> 3;                             // column vectors in a row
  const int lanes = (int)blockDim.x / cvw;            // row lanes
  const int cvl = threadIdx.x % cvw, rl = threadIdx.x / cvw;
  const int cv = blockIdx.x * cvw + cvl;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  const bool active = cv < ncv && rl < lanes;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    const __nv_bfloat16* base = dy + (long long)cv * 8;
    for (long long r = r0 + rl; r < r1; r += (long long)lanes * 8) {
      uint4 pk[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long rr = r + (long long)u * lanes;
        pk[u] = rr < r1 ? __ldcs(reinterpret_cast<const uint4*>(base + rr * n)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[(rl * cvw + cvl) * 8 + j] = acc[j];
  }
  __syncthreads();
  // fold the row lanes: thread (l, cvl) with l < 8 sums column j = l of vector cvl over all lanes
  if (cv < ncv && rl < 8) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += cs[(l * cvw + cvl) * 8 + rl];
    cs[cvl * 8 + rl] = t;   // lane 0's slot; every reader of slot (0, cvl, rl) is this thread
  }
  __syncthreads();
  if (cv < ncv && rl < 2) {
    const float* q = cs + cvl * 8 + rl * 4;
    red_add_v4(db + cv * 8 + rl * 4, q[0], q[1], q[2], q[3]);
  }
}
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db,
                                                     long long rows, int n, int rows_per_cta) {
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += __bfloat162float(dy[r * n + c]);
  atomicAdd(&db[c], acc);
}
// db[j] += sum_{b,d} dy[b][j][d]   (row-bias of the token-mixing Conv1d); one warp per row, 16-byte loads when D % 8 == 0
__global__ void __launch_bounds__(256) rowsum_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db, int B,
                                                     int J, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;  // over B*J
  if (row >= (long long)B * J) return;
  const __nv_bfloat16* r = dy + row * D;
  float acc = 0.f;
  if ((D & 7) == 0) {
    const uint4* rv = reinterpret_cast<const uint4*>(r);
    const int nv = D >> 3;
    for (int i = lane; i < nv; i += 128) {
      uint4 pk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) pk[u] = (i + 32 * u < nv) ? __ldcs(rv + i + 32 * u) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          acc += f.x + f.y;
        }
      }
    }
  } else {
    for (int i = lane; i < D; i += 32) acc += __bfloat162float(r[i]);
  }
  acc = warp_sum(acc);
  if (lane == 0) atomicAdd(&db[row % J], acc);
}

// ---------------------------------------------------------------- casts / adds
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}
__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}

// ---------------------------------------------------------------- VQ: clamp -> nearest code -> gather (main.py:763,134-138)
// fp32 SIMT distance search (argmin is discontinuous: no bf16 here):
//   d(row, code) = |c|^2 - 2 x.c        (|x|^2 is constant per row and dropped)
// CTA = 64 latent vectors x all codes; the clamped x tile lives in smem transposed [C][64]; codes stream through
// smem in [32 k][64 codes] chunks read from a pre-transposed codebook codeT[C][ncodes] (frozen -> transposed once).
// Thread micro-tile 4 rows x 4 codes, float4 LDS.  Ties resolve to the lowest index like torch.argmin.
template <int C>
__global__ void __launch_bounds__(256) vq_nearest_kernel(const float* __restrict__ z, const float* __restrict__ codebook,
                                                         const float* __restrict__ codeT, const float* __restrict__ cnorm,
                                                         int* __restrict__ idx_out, __nv_bfloat16* __restrict__ zq_bf16,
                                                         float* __restrict__ zq_f32, float* __restrict__ zc_out,
                                                         long long P, int ncodes, float lo, float hi) {
  extern __shared__ float vq_smem[];
  float* xs = vq_smem;                 // [C][64]
  float* cs = vq_smem + C * 64;        // [32][64]
  float* red_d = cs + 32 * 64;         // [64][16]
  int* red_i = reinterpret_cast<int*>(red_d + 64 * 16);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long row0 = (long long)blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * C; i += 256) {
    const int r = i / C, c = i % C;
    float v = 0.f;
    if (row0 + r < P) {
      v = z[(row0 + r) * C + c];
      v = fminf(fmaxf(v, lo), hi);     // clamp_with_grad forward (main.py:763)
      if (zc_out) zc_out[(row0 + r) * C + c] = v;
    }
    xs[c * 64 + r] = v;
  }
  float best[4];
  int besti[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    best[i] = FLT_MAX;
    besti[i] = 0;
  }
  for (int c0 = 0; c0 < ncodes; c0 += 64) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < C; k0 += 32) {
      __syncthreads();
      for (int i = threadIdx.x; i < 32 * 16; i += 256) {
        const int k = i >> 4, c4 = i & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c4 * 4 + 3 < ncodes) v = *reinterpret_cast<const float4*>(codeT + (long long)(k0 + k) * ncodes + c0 + c4 * 4);
        else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < 4; ++j)
            if (c0 + c4 * 4 + j < ncodes) t[j] = codeT[(long long)(k0 + k) * ncodes + c0 + c4 * 4 + j];
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
        *reinterpret_cast<float4*>(cs + k * 64 + c4 * 4) = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + (k0 + k) * 64 + ty * 4);
        const float4 cv = *reinterpret_cast<const float4*>(cs + k * 64 + tx * 4);
        const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
        const float cr[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], cr[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int code = c0 + tx * 4 + j;
      if (code < ncodes) {
        const float cn = cnorm[code];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float d = cn - 2.0f * acc[i][j];
          if (d < best[i]) {
            best[i] = d;
            besti[i] = code;
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red_d[(ty * 4 + i) * 16 + tx] = best[i];
    red_i[(ty * 4 + i) * 16 + tx] = besti[i];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int r = threadIdx.x;
    float b = red_d[r * 16];
    int bi = red_i[r * 16];
    for (int t = 1; t < 16; ++t) {
      const float d = red_d[r * 16 + t];
      const int di = red_i[r * 16 + t];
      if (d < b || (d == b && di < bi)) {
        b = d;
        bi = di;
      }
    }
    red_i[r * 16] = bi;
    if (row0 + r < P) idx_out[row0 + r] = bi;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * C; i += 256) {
    const int r = i / C, c = i % C;
    if (row0 + r < P) {
      const float v = codebook[(long long)red_i[r * 16] * C + c];
      if (zq_bf16) zq_bf16[(row0 + r) * C + c] = __float2bfloat16(v);
      if (zq_f32) zq_f32[(row0 + r) * C + c] = v;
    }
  }
}
// ---- tensor-core nearest-code search (ffvc_vq_nearest_tc): operand preparation and the final gather.
// A 2-term bf16 split x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) carries 16 mantissa bits; the GEMM contracts
// [z_hi | z_hi | z_lo] with [c_hi | c_lo | c_hi] (K = 3C), i.e. z.c up to the lo.lo term (relative 2^-16 per product, random
// sign: ~1e-4 absolute on a 256-term dot product of O(1) values), accumulated in fp32 by the tensor core.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(x);
  lo = __float2bfloat16(x - __bfloat162float(hi));
}
__global__ void vq_split_codebook_kernel(const float* __restrict__ cb, __nv_bfloat16* __restrict__ cs, float* __restrict__ cnorm,
                                         int ncodes, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= ncodes) return;
  float s = 0.f;
  for (int i = lane; i < C; i += 32) {
    const float v = cb[(long long)row * C + i];
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    __nv_bfloat16* o = cs + (long long)row * 3 * C;
    o[i] = hi;
    o[C + i] = lo;
    o[2 * C + i] = hi;
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) cnorm[row] = s;
}
__global__ void vq_split_z_kernel(const float* __restrict__ z, __nv_bfloat16* __restrict__ zs, float* __restrict__ zc,
                                  long long P, int C, float lo_, float hi_) {
  const long long total = P * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i % C);
    const float v = fminf(fmaxf(z[i], lo_), hi_);       // clamp_with_grad forward (main.py:763)
    if (zc) zc[i] = v;
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    __nv_bfloat16* o = zs + r * 3 * C;
    o[c] = hi;
    o[C + c] = hi;
    o[2 * C + c] = lo;
  }
}
__global__ void vq_gather_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ codebook,
                                 int* __restrict__ idx_out, __nv_bfloat16* __restrict__ zq_bf16, float* __restrict__ zq_f32,
                                 long long P, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= P) return;
  const int code = (int)(keys[row] & 0xffffffffull);
  if (lane == 0) idx_out[row] = code;
  for (int c = lane; c < C; c += 32) {
    const float v = codebook[(long long)code * C + c];
    if (zq_bf16) zq_bf16[row * C + c] = __float2bfloat16(v);
    if (zq_f32) zq_f32[row * C + c] = v;
  }
}
__global__ void rownorm2_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  float s = 0.f;
  for (int i = lane; i < C; i += 32) s += x[(long long)row * C + i] * x[(long long)row * C + i];
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// ClampWithGrad.backward (main.py:126-129): pass g iff g * (x - clamp(x)) >= 0
__global__ void clamp_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ gx,
                                 long long n, float lo, float hi) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = x[i], gv = g[i];
    const float d = xv - fminf(fmaxf(xv, lo), hi);
    gx[i] = (gv * d >= 0.f) ? gv : 0.f;
  }
}

// ---------------------------------------------------------------- image post: xr = clamp_with_grad((d+1)/2, 0, 1)  (main.py:142)
__global__ void image_post_fwd_kernel(const float* __restrict__ d, float* __restrict__ xr, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    xr[i] = fminf(fmaxf((d[i] + 1.0f) * 0.5f, 0.f), 1.f);
}
__global__ void image_post_bwd_kernel(const float* __restrict__ g, const float* __restrict__ d, float* __restrict__ gd,
                                      long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u = (d[i] + 1.0f) * 0.5f;
    const float diff = u - fminf(fmaxf(u, 0.f), 1.f);
    const float gv = g[i];
    gd[i] = (gv * diff >= 0.f) ? 0.5f * gv : 0.f;
  }
}

// ---------------------------------------------------------------- 3x3 conv with tiny Cout as a 1x1 GEMM + tap gather (conv_out)
// The decoder's last conv has 3 output channels.  As an implicit GEMM it reads every pixel nine times through the tensor core
// for a 3-wide (padded to 32) N: the tcgen05 pipe was busy 60 % of 0.87 ms for 34 TFLOP/s (profiles/r02_ncu_conv_out.md) because
// an MMA costs its A rows whatever N is.  Here the tensor core sees every pixel ONCE: v[p][tap*COUT + co] = sum_c a[p][c] w[co][tap][c]
// (a plain GEMM, N = 9 * COUT <= 32, fp32 out), and this kernel adds the nine shifted taps:
//   y[p][co] = bias[co] + sum_tap v[p + off(tap)][tap*COUT + co]   (zero outside the image = padding 1)
// It also applies the image post-processing xr = clamp((y + 1) / 2, 0, 1) (main.py:142) when xr != nullptr.
// CTA = 8 rows x 32 pixels: the 10 x 34 rows of v it touches (one 128-byte line each) stay in L1 across the taps.
// CTA = 8 rows x 32 pixels.  The 10 x 34 source rows of v it needs (one 128-byte line each) are staged in shared memory with
// coalesced 16-byte loads (8 lanes per line; row pitch 33 floats so that the per-pixel reads below are bank-conflict free) —
// reading them straight from global memory, one 4-byte load per lane at a 128-byte stride, kept L1 at 97 % for 408 us against
// ~95 us of HBM time (profiles/r02_ncu_hbm_kernels.md).
template <int COUT>
__global__ void __launch_bounds__(256) conv_taps_gather_kernel(const float* __restrict__ v, const float* __restrict__ bias,
                                                               float* __restrict__ y, float* __restrict__ xr, int H, int W) {
  constexpr int TW = 34, TH = 10, PITCH = 33;
  __shared__ float sv[TH * TW * PITCH];                  // 44,880 B
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8, n = blockIdx.z;
  const float* vn = v + (long long)n * H * W * 32;
  for (int i = threadIdx.x; i < TH * TW * 8; i += 256) {
    const int piece = i & 7, line = i >> 3;
    const int ly = line / TW, lx = line - ly * TW;
    const int sy = y0 + ly - 1, sx = x0 + lx - 1;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);          // outside the image: zero (padding 1)
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) q = *reinterpret_cast<const float4*>(vn + ((long long)sy * W + sx) * 32 + piece * 4);
    float* d = sv + line * PITCH + piece * 4;
    d[0] = q.x;
    d[1] = q.y;
    d[2] = q.z;
    d[3] = q.w;
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int x = x0 + lx, yy = y0 + ly;
  if (x >= W || yy >= H) return;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = bias ? bias[c] : 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const float* src = sv + ((ly + tap / 3) * TW + lx + tap % 3) * PITCH + tap * COUT;
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] += src[c];
  }
  const long long o = (((long long)n * H + yy) * W + x) * COUT;
#pragma unroll
  for (int c = 0; c < COUT; ++c) {
    y[o + c] = acc[c];
    if (xr) xr[o + c] = fminf(fmaxf((acc[c] + 1.0f) * 0.5f, 0.f), 1.f);
  }
}

// ---------------------------------------------------------------- 3x3 conv with tiny Cin (dgrad of conv_out: 3 -> 128 ch)
// x: [N][H][W][CIN] fp32, w: [COUT][9][CIN] fp32 (already flipped/transposed for dgrad), y: NHWC bf16.  pad 1.
template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_smallcin_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               __nv_bfloat16* __restrict__ y, int N, int H, int W, int COUT) {
  extern __shared__ float ws[];  // [COUT][9*CIN]
  for (int i = threadIdx.x; i < COUT * 9 * CIN; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int cv = COUT >> 3;
  const long long total = (long long)N * H * W * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    long long p = i / cv;
    const int px = (int)(p % W);
    p /= W;
    const int py = (int)(p % H);
    const int n = (int)(p / H);
    float in[9 * CIN];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
      for (int c = 0; c < CIN; ++c) in[t * CIN + c] = ok ? x[(((long long)n * H + yy) * W + xx) * CIN + c] : 0.f;
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* wr = ws + (c8 * 8 + j) * 9 * CIN;
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 9 * CIN; ++k) a = fmaf(in[k], wr[k], a);
      acc[j] = a;
    }
    uint4 o;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0], acc[1]), h1 = __floats2bfloat162_rn(acc[2], acc[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[4], acc[5]), h3 = __floats2bfloat162_rn(acc[6], acc[7]);
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    o.z = *reinterpret_cast<uint32_t*>(&h2);
    o.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(y + i * 8) = o;
  }
}

// ---------------------------------------------------------------- im2col for a 3-channel 3x3 conv (dgrad of conv_out on tensor cores)
// x: [N][H][W][3] fp32 -> col: [N*H*W][32] bf16, col[p][tap*3 + c] = x[pixel p shifted by tap][c] (zero outside / for k >= 27)
__global__ void __launch_bounds__(256) im2col3x3_cin3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int N,
                                                             int H, int W) {
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % W);
    const int py = (int)((i / W) % H);
    const int n = (int)(i / ((long long)W * H));
    float v[32];
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const float* src = x + (((long long)n * H + yy) * W + xx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[t * 3 + c] = ok ? src[c] : 0.f;
    }
    uint4* dst = reinterpret_cast<uint4*>(col + i * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 o;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[q * 8 + 0], v[q * 8 + 1]), h1 = __floats2bfloat162_rn(v[q * 8 + 2], v[q * 8 + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[q * 8 + 4], v[q * 8 + 5]), h3 = __floats2bfloat162_rn(v[q * 8 + 6], v[q * 8 + 7]);
      o.x = *reinterpret_cast<uint32_t*>(&h0);
      o.y = *reinterpret_cast<uint32_t*>(&h1);
      o.z = *reinterpret_cast<uint32_t*>(&h2);
      o.w = *reinterpret_cast<uint32_t*>(&h3);
      dst[q] = o;
    }
  }
}

// ---------------------------------------------------------------- fused Adam (torch.optim.Adam defaults, main.py:591,835)
// p,g,m,v fp32 flat arenas; also refreshes the bf16 shadow the GEMMs read.  grad_scale folds the DP average.
// hyper (DEVICE float[16], so a captured CUDA graph sees per-step values):
//   [0] lr of this step   [1] beta1   [2] beta2   [3] eps   [4] 1-beta1^t   [5] sqrt(1-beta2^t)   [6] grad_scale   [7] weight_decay
//   [8] t (steps taken)   [9] clip_grad_norm max_norm (0 = off, main.py:693,833-834)   [10] sum g^2 of this step (ffvc_sumsq)
//   [11] clip coefficient of this step (adam_tick)   [12] base lr   [13] cosine T_max (0 = constant lr, main.py:702-705)
//   [14] cosine eta_min   [15] EMA decay (0 = off; torch_ema semantics, main.py:524-525,843-844)
struct AdamCoef {
  float b1, b2, eps, bc2_sqrt, gscale, wd, step_size, ema_omd;
};
__device__ __forceinline__ AdamCoef adam_coef(const float* __restrict__ hyper) {
  AdamCoef c;
  c.b1 = hyper[1];
  c.b2 = hyper[2];
  c.eps = hyper[3];
  c.bc2_sqrt = hyper[5];
  c.gscale = hyper[6] * (hyper[9] > 0.f ? hyper[11] : 1.f);
  c.wd = hyper[7];
  c.step_size = hyper[0] / hyper[4];
  // torch_ema: num_updates += 1; decay = min(decay, (1 + num_updates) / (10 + num_updates)); shadow -= (1 - decay) * (shadow - p)
  const float t = hyper[8];
  c.ema_omd = 1.f - fminf(hyper[15], (1.f + t) / (10.f + t));
  return c;
}
__device__ __forceinline__ float adam_one(const AdamCoef& c, float pv, float gv, float& mv, float& vv) {
  gv *= c.gscale;
  if (c.wd != 0.f) gv += c.wd * pv;
  mv = c.b1 * mv + (1.f - c.b1) * gv;
  vv = c.b2 * vv + (1.f - c.b2) * gv * gv;
  const float denom = sqrtf(vv) / c.bc2_sqrt + c.eps;
  return pv - c.step_size * (mv / denom);
}
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, __nv_bfloat16* __restrict__ shadow,
                                                   float* __restrict__ ema, long long n, const float* __restrict__ hyper) {
  const AdamCoef c = adam_coef(hyper);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float mv = m[i], vv = v[i];
    const float np = adam_one(c, p[i], g[i], mv, vv);
    m[i] = mv;
    v[i] = vv;
    p[i] = np;
    if (shadow) shadow[i] = __float2bfloat16(np);
    if (ema) {
      const float e = ema[i];
      ema[i] = e - c.ema_omd * (e - np);
    }
  }
}
// 16-byte form of the same update (n % 4 == 0, 16-byte aligned arenas): identical arithmetic per element
template <bool kEma>
__global__ void __launch_bounds__(256) adam_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                        float4* __restrict__ m, float4* __restrict__ v,
                                                        uint2* __restrict__ shadow, float4* __restrict__ ema, long long n4,
                                                        const float* __restrict__ hyper) {
  const AdamCoef c = adam_coef(hyper);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 pv = p[i], gv = g[i];
    float4 mv = m[i], vv = v[i];
    float4 np;
    np.x = adam_one(c, pv.x, gv.x, mv.x, vv.x);
    np.y = adam_one(c, pv.y, gv.y, mv.y, vv.y);
    np.z = adam_one(c, pv.z, gv.z, mv.z, vv.z);
    np.w = adam_one(c, pv.w, gv.w, mv.w, vv.w);
    m[i] = mv;
    v[i] = vv;
    p[i] = np;
    if (shadow) {
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(np.x, np.y), h1 = __floats2bfloat162_rn(np.z, np.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
      shadow[i] = o;
    }
    if (kEma) {
      float4 e = ema[i];
      e.x -= c.ema_omd * (e.x - np.x);
      e.y -= c.ema_omd * (e.y - np.y);
      e.z -= c.ema_omd * (e.z - np.z);
      e.w -= c.ema_omd * (e.w - np.w);
      ema[i] = e;
    }
  }
}

// advance Adam's step counter on the device and refresh the per-step scalars (graph-capturable, no host staging):
// bias corrections, the cosine-annealed learning rate (torch CosineAnnealingLR closed form; the reference steps the scheduler
// AFTER opt.step(), main.py:835-837, so update t uses epoch t-1) and the clip_grad_norm_ coefficient min(1, max/(norm+1e-6)).
__global__ void adam_tick_kernel(float* __restrict__ hyper) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float t = hyper[8] + 1.0f;
    hyper[8] = t;
    // one thread: double precision is free here and matches the Python-float arithmetic of torch.optim.Adam / CosineAnnealingLR
    // (--use_fast_math powf left 3e-5 of relative error in the second bias correction)
    hyper[4] = (float)(1.0 - pow((double)hyper[1], (double)t));
    hyper[5] = (float)sqrt(1.0 - pow((double)hyper[2], (double)t));
    if (hyper[13] > 0.f)
      hyper[0] = (float)((double)hyper[14] + ((double)hyper[12] - (double)hyper[14]) * 0.5 * (1.0 + cospi(((double)t - 1.0) / (double)hyper[13])));
    if (hyper[9] > 0.f) {
      const float norm = sqrtf(hyper[10]) * hyper[6];         // norm of the averaged gradient (grad_scale = 1 / world)
      hyper[11] = fminf(1.0f, hyper[9] / (norm + 1e-6f));
    }
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
  float acc = 0.f;
  const long long n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? (n >> 2) : 0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 t = x4[i];
    acc += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += x[i] * x[i];
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    atomicAdd(out, s);
  }
}


// ---------------------------------------------------------------- CLIP ViT token assembly (cloob.py:240-243)
// x[n][0] = cls + pos[0];  x[n][1+p] = pe[n][p] + pos[1+p]        (bf16 activations, fp32 cls/pos)
__global__ void clip_assemble_kernel(const __nv_bfloat16* __restrict__ pe, const float* __restrict__ cls,
                                     const float* __restrict__ pos, __nv_bfloat16* __restrict__ x, int N, int T, int W) {
  const long long total = (long long)N * T * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int t = (int)((i / W) % T);
    const long long n = i / ((long long)W * T);
    const float v = (t == 0) ? cls[w] : __bfloat162float(pe[(n * (T - 1) + (t - 1)) * W + w]);
    x[i] = __float2bfloat16(v + pos[t * W + w]);
  }
}
// generic strided row copy (bf16): dst[r*dst_stride + c] = src[r*src_stride + c]
__global__ void copy_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows, int D,
                                 long long src_stride, long long dst_stride) {
  const long long total = rows * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D;
    const int c = (int)(i % D);
    dst[r * dst_stride + c] = src[r * src_stride + c];
  }
}

}  // namespace ffvc

using namespace ffvc;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int ffvc_upsample2x_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  if (C % 8) return set_error(FFVC_ERR_ARG, "upsample: C % 8 != 0");
  const long long total = (long long)N * 4 * H * W * (C / 8);
  upsample2x_fwd_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const uint4*>(x),
                                                                     reinterpret_cast<uint4*>(y), N, H, W, C / 8);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_upsample2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, void* stream) {
  if (C % 8) return set_error(FFVC_ERR_ARG, "upsample: C % 8 != 0");
  const long long total = (long long)N * H * W * (C / 8);
  upsample2x_bwd_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(CBF(dy), BF(dx), N, H, W, C);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
// in[b][R][Cc] -> out[b][Cc][R]; dtype codes: 0 = bf16, 1 = fp32
extern "C" int ffvc_transpose(const void* in, void* out, int B, int R, int Cc, int in_fp32, int out_fp32, void* stream) {
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, B), block(32, 8);
  if (!in_fp32 && !out_fp32)
    transpose_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, ST(stream)>>>(CBF(in), BF(out), R, Cc);
  else if (in_fp32 && !out_fp32)
    transpose_kernel<float, __nv_bfloat16><<<grid, block, 0, ST(stream)>>>(reinterpret_cast<const float*>(in), BF(out), R, Cc);
  else if (!in_fp32 && out_fp32)
    transpose_kernel<__nv_bfloat16, float><<<grid, block, 0, ST(stream)>>>(CBF(in), reinterpret_cast<float*>(out), R, Cc);
  else
    transpose_kernel<float, float><<<grid, block, 0, ST(stream)>>>(reinterpret_cast<const float*>(in),
                                                                  reinterpret_cast<float*>(out), R, Cc);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_softmax_fwd(const float* s, void* p, long long rows, int n, int ld, void* stream) {
  if (ld < n) return set_error(FFVC_ERR_ARG, "softmax: ld < n");
  softmax_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST(stream)>>>(s, BF(p), rows, n, ld);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_softmax_bwd(const void* p, const float* dp, void* ds, long long rows, int n, int ld, float scale,
                                void* stream) {
  if (ld < n) return set_error(FFVC_ERR_ARG, "softmax: ld < n");
  softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST(stream)>>>(CBF(p), dp, BF(ds), rows, n, ld, scale);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_colsum(const void* dy, float* db, long long rows, int n, void* stream) {
  if (n % 8 == 0 && (reinterpret_cast<uintptr_t>(db) & 15) == 0) {
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 148;
    }
    const int ncv = n / 8;
    const int cvw = ncv < 32 ? ncv : 32;                           // column vectors per CTA slab
    const int lanes = 512 / cvw;                                   // >= 16 row lanes (needs >= 8 for the fold)
    const int slabs = (ncv + cvw - 1) / cvw;
    long long chunks = 2 * sms / slabs > 0 ? 2 * sms / slabs : 1;  // row chunks so that slabs * chunks ~ 2 CTAs per SM
    const long long q = (long long)lanes * 8;                      // whole unrolled row batches per CTA
    long long rows_per_cta = ((rows + chunks - 1) / chunks + q - 1) / q * q;
    chunks = (rows + rows_per_cta - 1) / rows_per_cta;
    if (slabs > 65535 || chunks > 65535) return set_error(FFVC_ERR_ARG, "colsum: tensor too large");
    colsum_vec_kernel<<<dim3((unsigned)slabs, (unsigned)chunks), 512, 512 * 8 * sizeof(float), ST(stream)>>>(
        CBF(dy), db, rows, n, cvw, (int)rows_per_cta);
  } else {
    const int rows_per_cta = 512;
    dim3 grid((n + 255) / 256, (unsigned)((rows + rows_per_cta - 1) / rows_per_cta));
    colsum_kernel<<<grid, 256, 0, ST(stream)>>>(CBF(dy), db, rows, n, rows_per_cta);
  }
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_rowsum(const void* dy, float* db, int B, int J, int D, void* stream) {
  const long long rows = (long long)B * J;
  rowsum_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST(stream)>>>(CBF(dy), db, B, J, D);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_cast_f32_bf16(const float* x, void* y, long long n, void* stream) {
  cast_f32_bf16_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(x, BF(y), n);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_cast_bf16_f32(const void* x, float* y, long long n, void* stream) {
  cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(CBF(x), y, n);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_add_bf16(const void* a, const void* b, void* y, long long n, void* stream) {
  add_bf16_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(CBF(a), CBF(b), BF(y), n);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_rownorm2(const float* x, float* out, int rows, int C, void* stream) {
  rownorm2_kernel<<<(rows + 7) / 8, 256, 0, ST(stream)>>>(x, out, rows, C);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
// z: [P][C] fp32 (token-major latent), clamp to [lo,hi], nearest code; outputs: idx int32 [P], zq (bf16 and/or fp32),
// zc (clamped z, optional).  codeT = codebook transposed [C][ncodes] (ffvc_transpose), cnorm = |code|^2 (ffvc_rownorm2).
extern "C" int ffvc_vq_nearest(const float* z, const float* codebook, const float* codeT, const float* cnorm, int* idx,
                               void* zq_bf16, float* zq_f32, float* zc, long long P, int C, int ncodes, float lo, float hi,
                               void* stream) {
  if (ncodes % 4) return set_error(FFVC_ERR_ARG, "vq_nearest: ncodes must be a multiple of 4");
  const unsigned grid = (unsigned)((P + 63) / 64);
  const size_t smem = (size_t)(C * 64 + 32 * 64 + 64 * 16) * sizeof(float) + 64 * 16 * sizeof(int);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(vq_nearest_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(vq_nearest_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_done = true;
  }
  if (C == 256)
    vq_nearest_kernel<256><<<grid, 256, smem, ST(stream)>>>(z, codebook, codeT, cnorm, idx, BF(zq_bf16), zq_f32, zc, P, ncodes, lo, hi);
  else if (C == 64)
    vq_nearest_kernel<64><<<grid, 256, smem, ST(stream)>>>(z, codebook, codeT, cnorm, idx, BF(zq_bf16), zq_f32, zc, P, ncodes, lo, hi);
  else
    return set_error(FFVC_ERR_UNSUPPORTED, "vq_nearest: embed dim must be 64 or 256");
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_vq_prepare_codebook(const float* codebook, void* csplit_bf16, float* cnorm, int ncodes, int C, void* stream) {
  if (!codebook || !csplit_bf16 || !cnorm) return set_error(FFVC_ERR_ARG, "vq_prepare_codebook: null pointer");
  vq_split_codebook_kernel<<<(ncodes + 7) / 8, 256, 0, ST(stream)>>>(codebook, BF(csplit_bf16), cnorm, ncodes, C);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_vq_nearest_tc(const float* z, const float* codebook, const void* csplit_bf16, const float* cnorm,
                                  void* zsplit_bf16, void* keys_u64, int* idx, void* zq_bf16, float* zq_f32, float* zc, long long P,
                                  int C, int ncodes, float lo, float hi, void* stream) {
  if (!z || !codebook || !csplit_bf16 || !cnorm || !zsplit_bf16 || !keys_u64 || !idx)
    return set_error(FFVC_ERR_ARG, "vq_nearest_tc: null pointer");
  if ((3 * C) % 8 != 0 || P <= 0 || P > 0x7fffffffLL) return set_error(FFVC_ERR_ARG, "vq_nearest_tc: 3*C must be a multiple of 8");
  vq_split_z_kernel<<<grid_for(P * C, 256), 256, 0, ST(stream)>>>(z, BF(zsplit_bf16), zc, P, C, lo, hi);
  FFVC_CHECK_LAUNCH();
  cudaMemsetAsync(keys_u64, 0xFF, sizeof(unsigned long long) * P, ST(stream));
  ffvc_gemm_params g;
  memset(&g, 0, sizeof(g));
  g.a = zsplit_bf16;
  g.b = csplit_bf16;
  g.a_ld = g.b_ld = 3 * C;
  g.M = (int)P;
  g.N = ncodes;
  g.K = 3 * C;
  g.batch = g.k_segs = g.splits = 1;
  g.bias = cnorm;                  // d = |c|^2 - 2 z.c   (|z|^2 is constant per row and dropped)
  g.bias_mode = 1;
  g.alpha = -2.0f;
  g.ldc = ncodes;
  g.argmin_out = keys_u64;
  int rc = ffvc_gemm(&g, stream);
  if (rc) return rc;
  vq_gather_kernel<<<(unsigned)((P + 7) / 8), 256, 0, ST(stream)>>>(reinterpret_cast<const unsigned long long*>(keys_u64), codebook,
                                                                     idx, BF(zq_bf16), zq_f32, P, C);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_clamp_bwd(const float* g, const float* x, float* gx, long long n, float lo, float hi, void* stream) {
  clamp_bwd_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(g, x, gx, n, lo, hi);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_image_post_fwd(const float* d, float* xr, long long n, void* stream) {
  image_post_fwd_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(d, xr, n);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_image_post_bwd(const float* g, const float* d, float* gd, long long n, void* stream) {
  image_post_bwd_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(g, d, gd, n);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_conv_taps_gather(const float* v, const float* bias, float* y, float* xr, int N, int H, int W, int COUT,
                                     void* stream) {
  if (!v || !y) return set_error(FFVC_ERR_ARG, "conv_taps_gather: null pointer");
  dim3 grid((W + 31) / 32, (H + 7) / 8, N);
  if (COUT == 3) conv_taps_gather_kernel<3><<<grid, 256, 0, ST(stream)>>>(v, bias, y, xr, H, W);
  else if (COUT == 1) conv_taps_gather_kernel<1><<<grid, 256, 0, ST(stream)>>>(v, bias, y, xr, H, W);
  else if (COUT == 2) conv_taps_gather_kernel<2><<<grid, 256, 0, ST(stream)>>>(v, bias, y, xr, H, W);
  else return set_error(FFVC_ERR_ARG, "conv_taps_gather: COUT must be 1, 2 or 3 (9 * COUT tap columns in a 32-wide row)");
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_conv3x3_cin3(const float* x, const float* w, void* y, int N, int H, int W, int COUT, void* stream) {
  if (COUT % 8 || COUT > 512) return set_error(FFVC_ERR_ARG, "conv3x3_cin3: COUT % 8 != 0 or too large");
  const long long total = (long long)N * H * W * (COUT / 8);
  conv3x3_smallcin_kernel<3><<<grid_for(total, 256), 256, COUT * 27 * sizeof(float), ST(stream)>>>(x, w, BF(y), N, H, W, COUT);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
static int adam_launch(float* p, const float* g, float* m, float* v, void* shadow_bf16, float* ema, long long n,
                       const float* hyper_dev, void* stream) {
  if (!hyper_dev) return set_error(FFVC_ERR_ARG, "adam: hyper-parameter block is null");
  if (n <= 0) return FFVC_OK;
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(ema) |
                       (reinterpret_cast<uintptr_t>(shadow_bf16) << 1);       // the bf16 shadow only needs 8-byte alignment
  if ((n & 3) == 0 && (al & 15) == 0) {
    const long long n4 = n >> 2;
    if (ema)
      adam_vec4_kernel<true><<<grid_for(n4, 256, 148 * 8), 256, 0, ST(stream)>>>(
          reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
          reinterpret_cast<float4*>(v), reinterpret_cast<uint2*>(shadow_bf16), reinterpret_cast<float4*>(ema), n4, hyper_dev);
    else
      adam_vec4_kernel<false><<<grid_for(n4, 256, 148 * 8), 256, 0, ST(stream)>>>(
          reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
          reinterpret_cast<float4*>(v), reinterpret_cast<uint2*>(shadow_bf16), nullptr, n4, hyper_dev);
  } else {
    adam_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, ST(stream)>>>(p, g, m, v, BF(shadow_bf16), ema, n, hyper_dev);
  }
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n,
                              const float* hyper_dev, void* stream) {
  return adam_launch(p, g, m, v, shadow_bf16, nullptr, n, hyper_dev, stream);
}
extern "C" int ffvc_adam_step_ema(float* p, const float* g, float* m, float* v, void* shadow_bf16, float* ema, long long n,
                                  const float* hyper_dev, void* stream) {
  if (!ema) return set_error(FFVC_ERR_ARG, "adam_step_ema: ema arena is null");
  return adam_launch(p, g, m, v, shadow_bf16, ema, n, hyper_dev, stream);
}
extern "C" int ffvc_im2col3x3_cin3(const float* x, void* col, int N, int H, int W, void* stream) {
  im2col3x3_cin3_kernel<<<grid_for((long long)N * H * W, 256), 256, 0, ST(stream)>>>(x, BF(col), N, H, W);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_adam_tick(float* hyper_dev, void* stream) {
  adam_tick_kernel<<<1, 32, 0, ST(stream)>>>(hyper_dev);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_sumsq(const float* x, float* out, long long n, void* stream) {
  cudaMemsetAsync(out, 0, sizeof(float), ST(stream));
  sumsq_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, ST(stream)>>>(x, out, n);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_clip_assemble(const void* pe, const float* cls, const float* pos, void* x, int N, int T, int W, void* stream) {
  clip_assemble_kernel<<<grid_for((long long)N * T * W, 256), 256, 0, ST(stream)>>>(CBF(pe), cls, pos, BF(x), N, T, W);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_copy_rows(const void* src, void* dst, long long rows, int D, long long src_stride, long long dst_stride,
                              void* stream) {
  copy_rows_kernel<<<grid_for(rows * D, 256), 256, 0, ST(stream)>>>(CBF(src), BF(dst), rows, D, src_stride, dst_stride);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_softmax_causal_fwd(const float* s, void* p, long long rows, int T, int ld, void* stream) {
  if (ld < T) return set_error(FFVC_ERR_ARG, "softmax_causal: ld < T");
  softmax_causal_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST(stream)>>>(s, BF(p), rows, T, ld);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_embed_tokens(const long long* tok, const float* emb, const float* pos, void* x, long long rows, int T, int W,
                                 void* stream) {
  embed_tokens_kernel<<<grid_for(rows * W, 256), 256, 0, ST(stream)>>>(tok, emb, pos, BF(x), rows, T, W);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
extern "C" int ffvc_gather_rows(const void* src, const long long* idx, void* dst, int B, int T, int W, void* stream) {
  gather_rows_kernel<<<grid_for((long long)B * W, 256), 256, 0, ST(stream)>>>(CBF(src), idx, BF(dst), B, T, W);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
