// Shared device-side pieces of the tcgen05 GEMM family (gemm_tcgen05.cu, conv_halo_tcgen05.cu): the kernel parameter
// block and the fused register epilogue (bias / activation / activation-gradient / residual / store).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

#include "ffvc_internal.h"
#include "ptx.cuh"

namespace ffvc {

struct GemmDev {
  int M, N;
  int batch;            // output batches (outer * inner)
  int batch_inner;      // inner batch count (>= 1)
  int tile_m;           // 128 or 256
  int block_n;          // 32 / 64 / 128 / 256 (<= 128 when tile_m == 256)
  int a_mode, b_mode;   // FFVC_OP_*
  int kb_per_seg;       // ceil(K / 64)
  int k_segs;           // contraction additionally runs over this many "segments" (dim3 of the maps)
  int a_role, b_role;   // FFVC_ROLE_*
  int splits;           // split-K factor (atomic fp32 accumulation)
  int conv_h, conv_w, conv_cblocks;
  // epilogue
  void* out;
  void* pre_out;
  const __nv_bfloat16* aux;
  const __nv_bfloat16* res;
  const float* bias;
  long long ldc;
  long long out_bs;        // outer batch stride
  long long out_bs_inner;  // inner batch stride
  int out_fp32;
  int atomic;
  int bias_mode;  // 0 none, 1 per column, 2 per row
  int act;        // FFVC_ACT_*
  int mul_mode;   // multiply by act'(aux): FFVC_ACT_*
  float alpha;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == FFVC_ACT_GELU) return gelu_f(v);
  if (act == FFVC_ACT_QUICKGELU) return quick_gelu_f(v);
  if (act == FFVC_ACT_SWISH) return swish_f(v);
  if (act == FFVC_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}
__device__ __forceinline__ float apply_act_grad(float x, int act) {
  if (act == FFVC_ACT_GELU) return gelu_grad_f(x);
  if (act == FFVC_ACT_QUICKGELU) return quick_gelu_grad_f(x);
  if (act == FFVC_ACT_SWISH) return swish_grad_f(x);
  if (act == FFVC_ACT_RELU) return x > 0.f ? 1.0f : 0.f;
  return 1.0f;
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* v) {
  uint4 pk;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 h1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]);
  __nv_bfloat162 h3 = __floats2bfloat162_rn(v[6], v[7]);
  pk.x = *reinterpret_cast<uint32_t*>(&h0);
  pk.y = *reinterpret_cast<uint32_t*>(&h1);
  pk.z = *reinterpret_cast<uint32_t*>(&h2);
  pk.w = *reinterpret_cast<uint32_t*>(&h3);
  return pk;
}

// activation / activation-gradient on a CW-wide register chunk held as CW/2 float2 pairs; the switch is hoisted out of
// the element loop so each case is straight-line code (a per-element runtime branch costs more than the math).  GELU and
// its gradient run on the packed two-element forms (FFMA2 polynomial, no MUFU in the forward): these epilogues are
// issue-slot bound for the K = 256 token-mixing GEMMs.
template <int CW>
__device__ __forceinline__ void act_chunk(float2 (&v)[CW / 2], int act) {
  if (act == FFVC_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = gelu2(v[i]);
  } else if (act == FFVC_ACT_QUICKGELU) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = quick_gelu2(v[i]);
  } else if (act == FFVC_ACT_SWISH) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = make_float2(swish_f(v[i].x), swish_f(v[i].y));
  } else if (act == FFVC_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = make_float2(fmaxf(v[i].x, 0.f), fmaxf(v[i].y, 0.f));
  }
}
template <int CW>
__device__ __forceinline__ void mulgrad_chunk(float2 (&v)[CW / 2], const float2 (&x)[CW / 2], int act) {
  if (act == FFVC_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fmul2_rn(v[i], gelu_grad2(x[i]));
  } else if (act == FFVC_ACT_QUICKGELU) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fmul2_rn(v[i], quick_gelu_grad2(x[i]));
  } else if (act == FFVC_ACT_SWISH) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fmul2_rn(v[i], make_float2(swish_grad_f(x[i].x), swish_grad_f(x[i].y)));
  } else if (act == FFVC_ACT_RELU) {   // aux may be the pre- or the post-activation tensor: relu(x) > 0 <=> x > 0
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = make_float2(x[i].x > 0.f ? v[i].x : 0.f, x[i].y > 0.f ? v[i].y : 0.f);
  }
}
template <int CW>
__device__ __forceinline__ void unpack_bf16x2N(const uint4 (&pk)[CW / 8], float2 (&f)[CW / 2]) {
#pragma unroll
  for (int q = 0; q < CW / 8; ++q) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk[q]);
#pragma unroll
    for (int j = 0; j < 4; ++j) f[q * 4 + j] = __bfloat1622float2(h[j]);
  }
}
__device__ __forceinline__ uint4 pack_bf16x8_2(const float2* v) {
  uint4 pk;
  __nv_bfloat162 h0 = __float22bfloat162_rn(v[0]);
  __nv_bfloat162 h1 = __float22bfloat162_rn(v[1]);
  __nv_bfloat162 h2 = __float22bfloat162_rn(v[2]);
  __nv_bfloat162 h3 = __float22bfloat162_rn(v[3]);
  pk.x = *reinterpret_cast<uint32_t*>(&h0);
  pk.y = *reinterpret_cast<uint32_t*>(&h1);
  pk.z = *reinterpret_cast<uint32_t*>(&h2);
  pk.w = *reinterpret_cast<uint32_t*>(&h3);
  return pk;
}
template <int CW>
__device__ __forceinline__ void unpack_bf16xN(const uint4 (&pk)[CW / 8], float (&f)[CW]) {
#pragma unroll
  for (int q = 0; q < CW / 8; ++q) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk[q]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __bfloat1622float2(h[j]);
      f[q * 8 + 2 * j] = t.x;
      f[q * 8 + 2 * j + 1] = t.y;
    }
  }
}

// fused epilogue on 32 consecutive columns of one output row.
//   sbias : column bias of this tile staged in shared memory (already offset to this chunk), or nullptr
//   pf_aux / pf_res : aux / residual of this chunk prefetched into registers (valid when `vec` is true)
template <int CW>
__device__ __forceinline__ void epilogue_chunk(const GemmDev& p, const uint32_t (&r)[CW], int gn0, long long off, float rbias,
                                               bool vec, const float* sbias, const uint4 (&pf_aux)[CW / 8],
                                               const uint4 (&pf_res)[CW / 8]) {
  const int ncols = min(CW, p.N - gn0);
  float2 v[CW / 2];
  if (p.alpha != 1.0f) {
    const float2 al = make_float2(p.alpha, p.alpha);
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fmul2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), al);
  } else {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
  }
  if (p.bias_mode == 1) {
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(sbias + 4 * i);   // smem broadcast; zero beyond N
      v[2 * i] = __fadd2_rn(v[2 * i], make_float2(b4.x, b4.y));
      v[2 * i + 1] = __fadd2_rn(v[2 * i + 1], make_float2(b4.z, b4.w));
    }
  } else if (p.bias_mode == 2) {
    const float2 rb = make_float2(rbias, rbias);
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fadd2_rn(v[i], rb);
  }
  if (p.pre_out != nullptr) {
    __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(p.pre_out) + off;
    if (vec) {
#pragma unroll
      for (int i = 0; i < CW / 8; ++i) *reinterpret_cast<uint4*>(po + 8 * i) = pack_bf16x8_2(v + 4 * i);
    } else {
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) {
        if (2 * i < ncols) po[2 * i] = __float2bfloat16(v[i].x);
        if (2 * i + 1 < ncols) po[2 * i + 1] = __float2bfloat16(v[i].y);
      }
    }
  }
  act_chunk<CW>(v, p.act);
  if (p.mul_mode != FFVC_ACT_NONE) {
    float2 x[CW / 2];
    if (vec) {
      unpack_bf16x2N<CW>(pf_aux, x);
    } else {
      const __nv_bfloat16* ax = p.aux + off;
#pragma unroll
      for (int i = 0; i < CW / 2; ++i)
        x[i] = make_float2((2 * i < ncols) ? __bfloat162float(ax[2 * i]) : 0.f, (2 * i + 1 < ncols) ? __bfloat162float(ax[2 * i + 1]) : 0.f);
    }
    mulgrad_chunk<CW>(v, x, p.mul_mode);
  }
  if (p.res != nullptr) {
    float2 x[CW / 2];
    if (vec) {
      unpack_bf16x2N<CW>(pf_res, x);
    } else {
      const __nv_bfloat16* rs = p.res + off;
#pragma unroll
      for (int i = 0; i < CW / 2; ++i)
        x[i] = make_float2((2 * i < ncols) ? __bfloat162float(rs[2 * i]) : 0.f, (2 * i + 1 < ncols) ? __bfloat162float(rs[2 * i + 1]) : 0.f);
    }
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fadd2_rn(v[i], x[i]);
  }
  if (p.out_fp32) {
    float* o = reinterpret_cast<float*>(p.out) + off;
    if (p.atomic) {
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) {
        if (2 * i < ncols) atomicAdd(o + 2 * i, v[i].x);
        if (2 * i + 1 < ncols) atomicAdd(o + 2 * i + 1, v[i].y);
      }
    } else if (ncols == CW && (p.ldc % 4 == 0) && (p.out_bs % 4 == 0) && (p.out_bs_inner % 4 == 0)) {
#pragma unroll
      for (int i = 0; i < CW / 4; ++i)
        *reinterpret_cast<float4*>(o + 4 * i) = make_float4(v[2 * i].x, v[2 * i].y, v[2 * i + 1].x, v[2 * i + 1].y);
    } else {
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) {
        if (2 * i < ncols) o[2 * i] = v[i].x;
        if (2 * i + 1 < ncols) o[2 * i + 1] = v[i].y;
      }
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
    if (vec) {
#pragma unroll
      for (int i = 0; i < CW / 8; ++i) *reinterpret_cast<uint4*>(o + 8 * i) = pack_bf16x8_2(v + 4 * i);
    } else {
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) {
        if (2 * i < ncols) o[2 * i] = __float2bfloat16(v[i].x);
        if (2 * i + 1 < ncols) o[2 * i + 1] = __float2bfloat16(v[i].y);
      }
    }
  }
}


}  // namespace ffvc
