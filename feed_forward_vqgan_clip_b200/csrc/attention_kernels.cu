// Small-sequence multi-head attention for the CLIP ViT (T = 50 tokens, 12 heads x 64), forward and backward.
// Reference: nn.MultiheadAttention inside ResidualAttentionBlock.attention, cloob.py:188,198-200.
//
// One CTA (4 warps) per (sequence, head).  Q, K, V (and dO) are staged once in shared memory as bf16 (144-byte row pitch:
// conflict-free ldmatrix), every matmul runs on the warp-level tensor-core path (mma.sync m16n8k16, bf16 in, fp32
// accumulate) and scores / probabilities never touch HBM:
//   forward : warp w owns query rows 16w..16w+15:  S = Q K^T -> softmax in registers -> O = P V
//             (the S accumulators of two adjacent 8-column tiles ARE the A fragment of the next MMA: no shuffles)
//   backward: probabilities are recomputed (nothing but qkv is saved by the forward);
//             phase 1, per query-row tile: P, dP = dO V^T, dS = P (dP - rowsum(P dP)) scale, dQ = dS K; P and dS go to smem
//             phase 2, per key-row tile  : dV = P^T dO, dK = dS^T Q   (A operands read transposed with ldmatrix.trans)
// A 50 x 50 x 64 problem is far too small for a 128-row tcgen05 tile (the batched tcgen05 form it replaces padded 50
// rows to 128 and spent 100 us per launch in tile overheads); mma.sync at a few hundred MMAs per head is the right tool.
// qkv layout: [N][T][3*W] bf16 (the fused in_proj output: q | k | v along the last dim), head h owns columns
// h*64..h*64+63 of each third.  out / dout: [N][T][W] bf16.  T <= 64, head_dim == 64.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cfloat>

#include "ffvc_internal.h"
#include "ptx.cuh"

namespace ffvc {

static constexpr int kDh = 64;
static constexpr int kTmax = 64;
static constexpr int kPitch = 72;                       // bf16 elements per shared-memory row (144 B)
static constexpr int kTileElems = kTmax * kPitch;       // one [64][72] bf16 operand tile

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D (16x8, fp32) += A (16x16, bf16, row) * B (16x8, bf16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// stage T rows x 64 columns (bf16, global row stride `row_stride` elements) into a [64][kPitch] tile; rows >= T are zeroed
__device__ __forceinline__ void stage_tile(const __nv_bfloat16* __restrict__ base, long long row_stride, int T, __nv_bfloat16* dst) {
  for (int i = threadIdx.x; i < kTmax * 8; i += blockDim.x) {
    const int t = i >> 3, v = i & 7;
    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
    if (t < T) pk = *reinterpret_cast<const uint4*>(base + (long long)t * row_stride + v * 8);
    *reinterpret_cast<uint4*>(dst + t * kPitch + v * 8) = pk;
  }
}

// acc[nt] (16 x 8 tiles, nt = 0..7) = A[m0.., :64] * B^T where A, B are [row][k] tiles in shared memory (k contiguous):
// the Q K^T / dO V^T form (both operands read with plain ldmatrix)
__device__ __forceinline__ void mm_rows_x_rows(float (&acc)[8][4], uint32_t a_tile, uint32_t b_tile, int m0, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4(a_tile + (uint32_t)(((m0 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + kk * 16 + (lane >> 4) * 8) * 2), a0, a1, a2, a3);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(b_tile + (uint32_t)(((np * 16 + (lane & 7) + (lane >> 4) * 8) * kPitch + kk * 16 + ((lane >> 3) & 1) * 8) * 2), b0, b1, b2, b3);
      mma16816(acc[2 * np], a0, a1, a2, a3, b0, b1);
      mma16816(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}
// acc[nt] (16 x 8 tiles over the 64 output columns) += A * B with A given as register fragments af[kk][4] (k = 64 in four
// 16-wide steps) and B a [k][n] tile in shared memory (n contiguous): the P V / dS K form (B read with ldmatrix.trans)
__device__ __forceinline__ void mm_frag_x_cols(float (&acc)[8][4], const uint32_t (&af)[4][4], uint32_t b_tile, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(b_tile + (uint32_t)(((kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + np * 16 + (lane >> 4) * 8) * 2), b0, b1, b2, b3);
      mma16816(acc[2 * np], af[kk][0], af[kk][1], af[kk][2], af[kk][3], b0, b1);
      mma16816(acc[2 * np + 1], af[kk][0], af[kk][1], af[kk][2], af[kk][3], b2, b3);
    }
  }
}
// acc[nt] += A^T * B where A^T(m, k) is stored as at_tile[k][m] and B(k, n) as b_tile[k][n] (both read with ldmatrix.trans):
// the P^T dO / dS^T Q form
__device__ __forceinline__ void mm_colsT_x_cols(float (&acc)[8][4], uint32_t at_tile, uint32_t b_tile, int m0, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a0, a1, a2, a3;
    const int mi = lane >> 3;
    ldsm_x4_t(at_tile + (uint32_t)(((kk * 16 + (lane & 7) + (mi >> 1) * 8) * kPitch + m0 + (mi & 1) * 8) * 2), a0, a1, a2, a3);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(b_tile + (uint32_t)(((kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + np * 16 + (lane >> 4) * 8) * 2), b0, b1, b2, b3);
      mma16816(acc[2 * np], a0, a1, a2, a3, b0, b1);
      mma16816(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

// softmax over the 64 (masked to T) columns of the two rows (g, g + 8) a thread holds pieces of; s holds raw dot products.
// On return s holds the normalised probabilities (exactly 0 in the masked columns).
__device__ __forceinline__ void softmax_rows(float (&s)[8][4], int T, float scale, int lane) {
  const float k = scale * 1.4426950408889634f;
  const int t2 = (lane & 3) * 2;
  float mx0 = -FLT_MAX, mx1 = -FLT_MAX;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = nt * 8 + t2 + (e & 1);
      s[nt][e] = col < T ? s[nt][e] * k : -FLT_MAX;
    }
    mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = nt * 8 + t2 + (e & 1);
      const float p = col < T ? ex2_approx(s[nt][e] - ((e & 2) ? mx1 : mx0)) : 0.f;
      s[nt][e] = p;
      if (e & 2) sum1 += p; else sum0 += p;
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    s[nt][0] *= inv0;
    s[nt][1] *= inv0;
    s[nt][2] *= inv1;
    s[nt][3] *= inv1;
  }
}
// the fp32 accumulator tiles of a 16 x 64 matrix as the bf16 A fragments of the next MMA (k = 64 in four steps)
__device__ __forceinline__ void acc_to_afrag(const float (&s)[8][4], uint32_t (&af)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    af[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    af[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    af[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    af[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}
// store a 16 x 64 accumulator (rows m0 + g, m0 + g + 8) as bf16 to global rows of stride `ld` elements; rows >= T are skipped
__device__ __forceinline__ void store_rows(const float (&acc)[8][4], __nv_bfloat16* __restrict__ dst, long long ld, int m0, int T, int lane) {
  const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (m0 + g < T) *reinterpret_cast<uint32_t*>(dst + (long long)(m0 + g) * ld + nt * 8 + t2) = pack_bf16(acc[nt][0], acc[nt][1]);
    if (m0 + g + 8 < T) *reinterpret_cast<uint32_t*>(dst + (long long)(m0 + g + 8) * ld + nt * 8 + t2) = pack_bf16(acc[nt][2], acc[nt][3]);
  }
}
__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
}

__global__ void __launch_bounds__(128) mha_small_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ out, int T, int heads, float scale) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(sm_raw);
  __nv_bfloat16* ks = qs + kTileElems;
  __nv_bfloat16* vs = ks + kTileElems;
  const int n = blockIdx.x / heads, h = blockIdx.x % heads;
  const int W = heads * kDh;
  const __nv_bfloat16* base = qkv + (long long)n * T * 3 * W + h * kDh;
  stage_tile(base, 3 * W, T, qs);
  stage_tile(base + W, 3 * W, T, ks);
  stage_tile(base + 2 * W, 3 * W, T, vs);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = warp * 16;
  if (m0 >= T) return;
  float s[8][4];
  zero_acc(s);
  mm_rows_x_rows(s, smem_u32(qs), smem_u32(ks), m0, lane);
  softmax_rows(s, T, scale, lane);
  uint32_t pf[4][4];
  acc_to_afrag(s, pf);
  float o[8][4];
  zero_acc(o);
  mm_frag_x_cols(o, pf, smem_u32(vs), lane);
  store_rows(o, out + (long long)n * T * W + h * kDh, W, m0, T, lane);
}

// dqkv: [N][T][3W] bf16 gradient of the fused projection output.  Probabilities are recomputed.
__global__ void __launch_bounds__(128) mha_small_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            const __nv_bfloat16* __restrict__ dout,
                                                            __nv_bfloat16* __restrict__ dqkv, int T, int heads, float scale) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(sm_raw);
  __nv_bfloat16* ks = qs + kTileElems;
  __nv_bfloat16* vs = ks + kTileElems;
  __nv_bfloat16* dos = vs + kTileElems;
  __nv_bfloat16* ps = dos + kTileElems;     // P  [i][j]
  __nv_bfloat16* dss = ps + kTileElems;     // dS [i][j]
  const int n = blockIdx.x / heads, h = blockIdx.x % heads;
  const int W = heads * kDh;
  const __nv_bfloat16* base = qkv + (long long)n * T * 3 * W + h * kDh;
  stage_tile(base, 3 * W, T, qs);
  stage_tile(base + W, 3 * W, T, ks);
  stage_tile(base + 2 * W, 3 * W, T, vs);
  stage_tile(dout + (long long)n * T * W + h * kDh, W, T, dos);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = warp * 16;
  const int g = lane >> 2, t2 = (lane & 3) * 2;
  if (m0 >= T) {   // this warp's query rows do not exist: their P / dS rows must still be finite for phase 2
    for (int i = lane; i < 16 * 8; i += 32) {
      *reinterpret_cast<uint4*>(ps + (m0 + (i >> 3)) * kPitch + (i & 7) * 8) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(dss + (m0 + (i >> 3)) * kPitch + (i & 7) * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();
  __nv_bfloat16* ob = dqkv + (long long)n * T * 3 * W + h * kDh;
  if (m0 < T) {
    // ---- phase 1: rows i of this warp
    float p[8][4], dp[8][4];
    zero_acc(p);
    mm_rows_x_rows(p, smem_u32(qs), smem_u32(ks), m0, lane);
    softmax_rows(p, T, scale, lane);
    zero_acc(dp);
    mm_rows_x_rows(dp, smem_u32(dos), smem_u32(vs), m0, lane);
    float d0 = 0.f, d1 = 0.f;   // rowsum(P * dP) for rows g, g + 8
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      d0 = fmaf(p[nt][0], dp[nt][0], fmaf(p[nt][1], dp[nt][1], d0));
      d1 = fmaf(p[nt][2], dp[nt][2], fmaf(p[nt][3], dp[nt][3], d1));
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      // P to shared memory (bf16) for phase 2, then dS in place of dP
      *reinterpret_cast<uint32_t*>(ps + (m0 + g) * kPitch + nt * 8 + t2) = pack_bf16(p[nt][0], p[nt][1]);
      *reinterpret_cast<uint32_t*>(ps + (m0 + g + 8) * kPitch + nt * 8 + t2) = pack_bf16(p[nt][2], p[nt][3]);
      dp[nt][0] = p[nt][0] * (dp[nt][0] - d0) * scale;
      dp[nt][1] = p[nt][1] * (dp[nt][1] - d0) * scale;
      dp[nt][2] = p[nt][2] * (dp[nt][2] - d1) * scale;
      dp[nt][3] = p[nt][3] * (dp[nt][3] - d1) * scale;
      *reinterpret_cast<uint32_t*>(dss + (m0 + g) * kPitch + nt * 8 + t2) = pack_bf16(dp[nt][0], dp[nt][1]);
      *reinterpret_cast<uint32_t*>(dss + (m0 + g + 8) * kPitch + nt * 8 + t2) = pack_bf16(dp[nt][2], dp[nt][3]);
    }
    uint32_t dsf[4][4];
    acc_to_afrag(dp, dsf);
    float dq[8][4];
    zero_acc(dq);
    mm_frag_x_cols(dq, dsf, smem_u32(ks), lane);          // dQ = dS K
    store_rows(dq, ob, 3 * W, m0, T, lane);
  }
  __syncthreads();
  if (m0 < T) {
    // ---- phase 2: rows j (keys) of this warp
    float acc[8][4];
    zero_acc(acc);
    mm_colsT_x_cols(acc, smem_u32(dss), smem_u32(qs), m0, lane);   // dK = dS^T Q
    store_rows(acc, ob + W, 3 * W, m0, T, lane);
    zero_acc(acc);
    mm_colsT_x_cols(acc, smem_u32(ps), smem_u32(dos), m0, lane);   // dV = P^T dO
    store_rows(acc, ob + 2 * W, 3 * W, m0, T, lane);
  }
}

// ------------------------------------------------------------------------------------------------ long sequences (T > 64)
// Tiled ("flash") attention on the same 64 x 64 building blocks: the x-transformer mapper's causal attention at T = 1024
// (transformer.py:11-20 / x-transformers Attention, 6 heads x 64) and any ViT at more than 64 tokens.  Round 1 ran these as two
// batched tcgen05 GEMMs around a row softmax with the T x T scores (fp32) and probabilities (bf16) materialised in HBM: at
// config #4 that was 0.6 GB per layer and direction, 19 % of the step (profiles/r02_step_breakdown_c4.md).  Here scores and
// probabilities never leave the SM:
//   forward    CTA = (sequence, head, 64 query rows); key blocks of 64 stream through shared memory; online softmax in the log2
//              domain (running max m, running sum l), O rescaled per block; saves lse2 = m + log2(l) per row for the backward
//   backward   delta[i] = rowsum(dO_i * O_i) (one warp per row), then
//              dQ kernel : CTA = (sequence, head, query block), loop over key blocks: P = 2^(s2 - lse2), dP = dO V^T,
//                          dS = P (dP - delta) scale, dQ += dS K
//              dKV kernel: CTA = (sequence, head, key block), loop over query blocks: P and dS of the block go to shared memory,
//                          dV += P^T dO, dK += dS^T Q  (accumulators stay in registers over the loop)
// causal: key j attends query i only if j <= i (blocks above the diagonal are skipped, the diagonal block is masked).
__device__ __forceinline__ void stage_tile_rows(const __nv_bfloat16* __restrict__ base, long long row_stride, int row0, int T,
                                                __nv_bfloat16* dst) {
  for (int i = threadIdx.x; i < kTmax * 8; i += blockDim.x) {
    const int t = i >> 3, v = i & 7;
    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + t < T) pk = *reinterpret_cast<const uint4*>(base + (long long)(row0 + t) * row_stride + v * 8);
    *reinterpret_cast<uint4*>(dst + t * kPitch + v * 8) = pk;
  }
}
// scores of rows (q0 + m0 + g, + 8) against keys k0 .. k0 + 63 -> log2-domain logits, -FLT_MAX where masked
__device__ __forceinline__ void mask_scale(float (&s)[8][4], float k2, int qrow0, int k0, int T, int causal, int lane) {
  const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = k0 + nt * 8 + t2 + (e & 1);
      const int row = qrow0 + g + ((e & 2) ? 8 : 0);
      const bool ok = col < T && (!causal || col <= row);
      s[nt][e] = ok ? s[nt][e] * k2 : -FLT_MAX;
    }
}

__global__ void __launch_bounds__(128) mha_flash_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                            float* __restrict__ lse, int T, int heads, float scale, int causal) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(sm_raw);
  __nv_bfloat16* ks = qs + kTileElems;
  __nv_bfloat16* vs = ks + kTileElems;
  const int qb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int W = heads * kDh;
  const __nv_bfloat16* base = qkv + (long long)n * T * 3 * W + h * kDh;
  const int q0 = qb * 64;
  stage_tile_rows(base, 3 * W, q0, T, qs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = warp * 16, g = lane >> 2;
  const float k2 = scale * 1.4426950408889634f;
  float o[8][4];
  zero_acc(o);
  float mx0 = -FLT_MAX, mx1 = -FLT_MAX, l0 = 0.f, l1 = 0.f;
  const int nkb = causal ? (qb + 1) : (T + 63) / 64;
  for (int kb = 0; kb < nkb; ++kb) {
    __syncthreads();                                   // previous block's K / V fully consumed (and Q staged, first time)
    stage_tile_rows(base + W, 3 * W, kb * 64, T, ks);
    stage_tile_rows(base + 2 * W, 3 * W, kb * 64, T, vs);
    __syncthreads();
    float s[8][4];
    zero_acc(s);
    mm_rows_x_rows(s, smem_u32(qs), smem_u32(ks), m0, lane);
    mask_scale(s, k2, q0 + m0, kb * 64, T, causal, lane);
    float bm0 = -FLT_MAX, bm1 = -FLT_MAX;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
      bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float nm0 = fmaxf(mx0, bm0), nm1 = fmaxf(mx1, bm1);
    const float c0 = ex2_approx(mx0 - nm0), c1 = ex2_approx(mx1 - nm1);     // 0 on the first block (mx = -FLT_MAX)
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float m = (e & 2) ? nm1 : nm0;
        const float pv = s[nt][e] > -1e37f ? ex2_approx(s[nt][e] - m) : 0.f;
        s[nt][e] = pv;
        if (e & 2) rs1 += pv; else rs0 += pv;
      }
      o[nt][0] *= c0;
      o[nt][1] *= c0;
      o[nt][2] *= c1;
      o[nt][3] *= c1;
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
    mx0 = nm0;
    mx1 = nm1;
    uint32_t pf[4][4];
    acc_to_afrag(s, pf);
    mm_frag_x_cols(o, pf, smem_u32(vs), lane);
  }
  const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    o[nt][0] *= i0;
    o[nt][1] *= i0;
    o[nt][2] *= i1;
    o[nt][3] *= i1;
  }
  store_rows(o, out + (long long)n * T * W + h * kDh + (long long)q0 * W, W, m0, T - q0, lane);
  if ((lane & 3) == 0) {
    float* lr = lse + ((long long)n * heads + h) * T;
    if (q0 + m0 + g < T) lr[q0 + m0 + g] = mx0 + log2f(l0);
    if (q0 + m0 + g + 8 < T) lr[q0 + m0 + g + 8] = mx1 + log2f(l1);
  }
}

// delta[n][h][t] = sum_d dO[n][t][h*64 + d] * O[n][t][h*64 + d]; one warp per (n, t, h)
__global__ void __launch_bounds__(256) mha_flash_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                              float* __restrict__ delta, long long rows, int T, int heads) {
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // over N * T * heads
  if (item >= rows * heads) return;
  const int h = (int)(item % heads);
  const long long nt = item / heads;                 // n * T + t
  const long long off = nt * heads * kDh + h * kDh + lane * 2;
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(o + off));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dout + off));
  float d = warp_sum(a.x * b.x + a.y * b.y);
  if (lane == 0) delta[((nt / T) * heads + h) * T + nt % T] = d;
}

// P (recomputed from lse2) and dS for rows (qrow0 + g, + 8) of this warp against the staged key block; rows >= T give zeros
__device__ __forceinline__ void flash_p_ds(float (&p)[8][4], float (&dp)[8][4], uint32_t qs, uint32_t ks, uint32_t dos, uint32_t vs,
                                           const float* __restrict__ lse, const float* __restrict__ delta, int m0, int q0, int k0, int T,
                                           float scale, int causal, int lane) {
  const int g = lane >> 2;
  zero_acc(p);
  mm_rows_x_rows(p, qs, ks, m0, lane);
  mask_scale(p, scale * 1.4426950408889634f, q0 + m0, k0, T, causal, lane);
  zero_acc(dp);
  mm_rows_x_rows(dp, dos, vs, m0, lane);
  const int r0 = q0 + m0 + g, r1 = r0 + 8;
  const float ls0 = r0 < T ? lse[r0] : 0.f, ls1 = r1 < T ? lse[r1] : 0.f;
  const float d0 = r0 < T ? delta[r0] : 0.f, d1 = r1 < T ? delta[r1] : 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool hi = (e & 2) != 0;
      const bool ok = (hi ? r1 : r0) < T && p[nt][e] > -1e37f;
      const float pv = ok ? ex2_approx(p[nt][e] - (hi ? ls1 : ls0)) : 0.f;
      p[nt][e] = pv;
      dp[nt][e] = pv * (dp[nt][e] - (hi ? d1 : d0)) * scale;
    }
}

__global__ void __launch_bounds__(128) mha_flash_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                                                               const float* __restrict__ lse, const float* __restrict__ delta,
                                                               __nv_bfloat16* __restrict__ dqkv, int T, int heads, float scale, int causal) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(sm_raw);
  __nv_bfloat16* dos = qs + kTileElems;
  __nv_bfloat16* ks = dos + kTileElems;
  __nv_bfloat16* vs = ks + kTileElems;
  const int qb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int W = heads * kDh;
  const __nv_bfloat16* base = qkv + (long long)n * T * 3 * W + h * kDh;
  const int q0 = qb * 64;
  stage_tile_rows(base, 3 * W, q0, T, qs);
  stage_tile_rows(dout + (long long)n * T * W + h * kDh, W, q0, T, dos);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = warp * 16;
  const float* lr = lse + ((long long)n * heads + h) * T;
  const float* dr = delta + ((long long)n * heads + h) * T;
  float dq[8][4];
  zero_acc(dq);
  const int nkb = causal ? (qb + 1) : (T + 63) / 64;
  for (int kb = 0; kb < nkb; ++kb) {
    __syncthreads();
    stage_tile_rows(base + W, 3 * W, kb * 64, T, ks);
    stage_tile_rows(base + 2 * W, 3 * W, kb * 64, T, vs);
    __syncthreads();
    float p[8][4], ds[8][4];
    flash_p_ds(p, ds, smem_u32(qs), smem_u32(ks), smem_u32(dos), smem_u32(vs), lr, dr, m0, q0, kb * 64, T, scale, causal, lane);
    uint32_t dsf[4][4];
    acc_to_afrag(ds, dsf);
    mm_frag_x_cols(dq, dsf, smem_u32(ks), lane);          // dQ += dS K
  }
  store_rows(dq, dqkv + (long long)n * T * 3 * W + h * kDh + (long long)q0 * 3 * W, 3 * W, m0, T - q0, lane);
}

__global__ void __launch_bounds__(128) mha_flash_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                                                                const float* __restrict__ lse, const float* __restrict__ delta,
                                                                __nv_bfloat16* __restrict__ dqkv, int T, int heads, float scale, int causal) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(sm_raw);
  __nv_bfloat16* vs = ks + kTileElems;
  __nv_bfloat16* qs = vs + kTileElems;
  __nv_bfloat16* dos = qs + kTileElems;
  __nv_bfloat16* ps = dos + kTileElems;      // P  [i][j] of the current query block
  __nv_bfloat16* dss = ps + kTileElems;      // dS [i][j]
  const int kb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int W = heads * kDh;
  const __nv_bfloat16* base = qkv + (long long)n * T * 3 * W + h * kDh;
  const __nv_bfloat16* dob = dout + (long long)n * T * W + h * kDh;
  const int k0 = kb * 64;
  stage_tile_rows(base + W, 3 * W, k0, T, ks);
  stage_tile_rows(base + 2 * W, 3 * W, k0, T, vs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = warp * 16, g = lane >> 2, t2 = (lane & 3) * 2;
  const float* lr = lse + ((long long)n * heads + h) * T;
  const float* dr = delta + ((long long)n * heads + h) * T;
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  const int nqb = (T + 63) / 64;
  for (int qb = causal ? kb : 0; qb < nqb; ++qb) {
    __syncthreads();                          // the previous block's Q / dO / P / dS fully consumed
    stage_tile_rows(base, 3 * W, qb * 64, T, qs);
    stage_tile_rows(dob, W, qb * 64, T, dos);
    __syncthreads();
    float p[8][4], ds[8][4];
    flash_p_ds(p, ds, smem_u32(qs), smem_u32(ks), smem_u32(dos), smem_u32(vs), lr, dr, m0, qb * 64, k0, T, scale, causal, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t*>(ps + (m0 + g) * kPitch + nt * 8 + t2) = pack_bf16(p[nt][0], p[nt][1]);
      *reinterpret_cast<uint32_t*>(ps + (m0 + g + 8) * kPitch + nt * 8 + t2) = pack_bf16(p[nt][2], p[nt][3]);
      *reinterpret_cast<uint32_t*>(dss + (m0 + g) * kPitch + nt * 8 + t2) = pack_bf16(ds[nt][0], ds[nt][1]);
      *reinterpret_cast<uint32_t*>(dss + (m0 + g + 8) * kPitch + nt * 8 + t2) = pack_bf16(ds[nt][2], ds[nt][3]);
    }
    __syncthreads();
    mm_colsT_x_cols(dk, smem_u32(dss), smem_u32(qs), m0, lane);    // dK (keys m0..) += dS^T Q
    mm_colsT_x_cols(dv, smem_u32(ps), smem_u32(dos), m0, lane);    // dV += P^T dO
  }
  __nv_bfloat16* ob = dqkv + (long long)n * T * 3 * W + h * kDh + (long long)k0 * 3 * W;
  store_rows(dk, ob + W, 3 * W, m0, T - k0, lane);
  store_rows(dv, ob + 2 * W, 3 * W, m0, T - k0, lane);
}

}  // namespace ffvc

using namespace ffvc;

static int mha_setup() {
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(mha_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kTileElems * 2);
    if (e != cudaSuccess) return set_error(FFVC_ERR_CUDA, cudaGetErrorString(e));
    e = cudaFuncSetAttribute(mha_flash_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kTileElems * 2);
    if (e != cudaSuccess) return set_error(FFVC_ERR_CUDA, cudaGetErrorString(e));
    done = true;
  }
  return FFVC_OK;
}

extern "C" int ffvc_mha_small_fwd(const void* qkv, void* out, int N, int T, int heads, int head_dim, float scale,
                                  void* stream) {
  if (head_dim != kDh || T > kTmax || T < 1) return set_error(FFVC_ERR_UNSUPPORTED, "mha_small: head_dim must be 64 and T <= 64");
  if ((heads * kDh) % 8 != 0) return set_error(FFVC_ERR_ARG, "mha_small: width must be a multiple of 8");
  int rc = mha_setup();
  if (rc) return rc;
  mha_small_fwd_kernel<<<N * heads, 128, 3 * kTileElems * 2, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), T, heads, scale);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_mha_small_bwd(const void* qkv, const void* dout, void* dqkv, int N, int T, int heads, int head_dim,
                                  float scale, void* stream) {
  if (head_dim != kDh || T > kTmax || T < 1) return set_error(FFVC_ERR_UNSUPPORTED, "mha_small: head_dim must be 64 and T <= 64");
  int rc = mha_setup();
  if (rc) return rc;
  mha_small_bwd_kernel<<<N * heads, 128, 6 * kTileElems * 2, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout),
      reinterpret_cast<__nv_bfloat16*>(dqkv), T, heads, scale);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

// Tiled attention for any T (see above).  qkv [N][T][3*W] bf16 (q | k | v, head h = columns h*64..), out / dout [N][T][W] bf16,
// lse [N][heads][T] fp32 (log2-domain log-sum-exp of the scaled scores, written by the forward, read by the backward),
// delta_ws [N][heads][T] fp32 scratch.
extern "C" int ffvc_mha_flash_fwd(const void* qkv, void* out, float* lse, int N, int T, int heads, int head_dim, float scale,
                                  int causal, void* stream) {
  if (head_dim != kDh || T < 1) return set_error(FFVC_ERR_UNSUPPORTED, "mha_flash: head_dim must be 64");
  if (!qkv || !out || !lse) return set_error(FFVC_ERR_ARG, "mha_flash_fwd: null pointer");
  int rc = mha_setup();
  if (rc) return rc;
  dim3 grid((T + 63) / 64, heads, N);
  mha_flash_fwd_kernel<<<grid, 128, 3 * kTileElems * 2, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), lse, T, heads, scale, causal);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_mha_flash_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv,
                                  int N, int T, int heads, int head_dim, float scale, int causal, void* stream) {
  if (head_dim != kDh || T < 1) return set_error(FFVC_ERR_UNSUPPORTED, "mha_flash: head_dim must be 64");
  if (!qkv || !out || !dout || !lse || !delta_ws || !dqkv) return set_error(FFVC_ERR_ARG, "mha_flash_bwd: null pointer");
  int rc = mha_setup();
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)N * T;
  mha_flash_delta_kernel<<<(unsigned)((rows * heads + 7) / 8), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(out),
                                                                            reinterpret_cast<const __nv_bfloat16*>(dout), delta_ws, rows, T, heads);
  FFVC_CHECK_LAUNCH();
  dim3 grid((T + 63) / 64, heads, N);
  mha_flash_bwd_dq_kernel<<<grid, 128, 4 * kTileElems * 2, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                                reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta_ws,
                                                                reinterpret_cast<__nv_bfloat16*>(dqkv), T, heads, scale, causal);
  FFVC_CHECK_LAUNCH();
  mha_flash_bwd_dkv_kernel<<<grid, 128, 6 * kTileElems * 2, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                                 reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta_ws,
                                                                 reinterpret_cast<__nv_bfloat16*>(dqkv), T, heads, scale, causal);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
