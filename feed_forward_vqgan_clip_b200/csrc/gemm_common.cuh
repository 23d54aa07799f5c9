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
  unsigned long long* argmin;   // optional: per-row arg-min epilogue (see ffvc_gemm_params.argmin_out)
  int quad;                     // CTA-pair kernel in clusters of 4: two M-adjacent pair tiles share the B tile through TMA multicast
  int stream_k;                 // 1: contiguous (tile, k-block) ranges per worker instead of whole tiles (atomic fp32 output)
  int tma_store;                // 1: the epilogue warps stage their results in shared memory and TMA-store them
  double* gn_ws;                // conv3x3_halo only: per-(image, group) sum / sum of squares of the stored output (GroupNorm(32) statistics)
  // conv3x3_halo only, gn_bwd = 1: the stored output is dy of a Normalize + swish whose INPUT is `aux`; the epilogue accumulates the
  // backward statistics sum g, sum g * xhat (g = dy * swish'(u) * gamma, u = gamma * xhat + beta) into gn_ws instead
  int gn_bwd;
  const float* gn_mean;         // [images][32]
  const float* gn_rstd;
  const float* gn_gamma;        // [128]
  const float* gn_beta;
  // conv3x3_halo only, xf_mean != nullptr: the INPUT tensor is the raw input x of a Normalize (GroupNorm(xf_groups) + swish); four
  // transform warps apply it to every halo tile in shared memory between the TMA load and the MMAs (the normalised tensor never
  // exists in HBM).  xf_mean / xf_rstd: [images][xf_groups]; xf_gamma / xf_beta: [Cin].
  const float* xf_mean;
  const float* xf_rstd;
  const float* xf_gamma;
  const float* xf_beta;
  int xf_groups, xf_cpg;
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

// Compile-time epilogue configuration.  The fused epilogue is driven by runtime flags (activation, activation-gradient
// multiplier, bias mode, second output, residual, fp32 / atomic output, alpha); evaluated per 16-column chunk they cost more
// instructions than the math itself (ncu source page of the K = 256 token-mixing GEMM: 375 warp instructions per chunk,
// 112 of them arithmetic — PRMT copies of unused prefetch registers, predicated-off loads, an indirect branch on the
// activation code, spill reloads).  kEpi >= 0 bakes the hot combinations in (bf16 output, alpha = 1, no atomics):
//   bits [0:3) activation, [3:6) activation-gradient multiplier, [6:8) bias mode, 8 second (pre-activation) output, 9 residual
// kEpi = -1 keeps every flag at run time (all other call sites, incl. the halo-reuse conv kernel).
constexpr int epi_code(int act, int mul, int bias_mode, int pre, int res) {
  return act | (mul << 3) | (bias_mode << 6) | (pre << 8) | (res << 9);
}
template <int kEpi>
struct Epi {
  static constexpr bool kGen = kEpi < 0;
  __device__ __forceinline__ static int act(const GemmDev& p) { if constexpr (kGen) return p.act; else return kEpi & 7; }
  __device__ __forceinline__ static int mul(const GemmDev& p) { if constexpr (kGen) return p.mul_mode; else return (kEpi >> 3) & 7; }
  __device__ __forceinline__ static int bias(const GemmDev& p) { if constexpr (kGen) return p.bias_mode; else return (kEpi >> 6) & 3; }
  __device__ __forceinline__ static bool pre(const GemmDev& p) { if constexpr (kGen) return p.pre_out != nullptr; else return ((kEpi >> 8) & 1) != 0; }
  __device__ __forceinline__ static bool res(const GemmDev& p) { if constexpr (kGen) return p.res != nullptr; else return ((kEpi >> 9) & 1) != 0; }
  __device__ __forceinline__ static bool f32(const GemmDev& p) { if constexpr (kGen) return p.out_fp32 != 0; else return false; }
  __device__ __forceinline__ static bool atomic(const GemmDev& p) { if constexpr (kGen) return p.atomic != 0; else return false; }
  __device__ __forceinline__ static bool scaled(const GemmDev& p) { if constexpr (kGen) return p.alpha != 1.0f; else return false; }
  __device__ __forceinline__ static bool argmin(const GemmDev& p) { if constexpr (kGen) return p.argmin != nullptr; else return false; }
};

// order-preserving map float -> uint32 (smaller float <-> smaller unsigned)
__device__ __forceinline__ uint32_t float_order_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// arg-min epilogue on one CW-wide chunk of one output row: v = alpha * acc + bias[n]; running (best, n) in registers
template <int CW>
__device__ __forceinline__ void argmin_chunk(const GemmDev& p, const uint32_t (&r)[CW], int gn0, const float* sbias, float& best,
                                             int& besti) {
#pragma unroll
  for (int i = 0; i < CW; ++i) {
    const float v = fmaf(__uint_as_float(r[i]), p.alpha, sbias ? sbias[i] : 0.f);
    if (gn0 + i < p.N && v < best) {    // strict <: ties keep the lowest column (torch.argmin)
      best = v;
      besti = gn0 + i;
    }
  }
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
// 256-bit global accesses (sm_100: STG.E.ENL2.256 / LDG.E.ENL2.256): one full 32-byte sector per lane.  In the TMEM register
// layout a lane owns one output ROW, so a warp store touches 32 different lines whatever the width; 32 B per lane halves the
// number of L1 store requests (and partial-sector writes) of the epilogue relative to 16-byte stores.
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void ld_global_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
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
template <int CW, int kEpi = -1>
__device__ __forceinline__ void epilogue_chunk(const GemmDev& p, const uint32_t (&r)[CW], int gn0, long long off, float rbias,
                                               bool vec, const float* sbias, const uint4 (&pf_aux)[CW / 8],
                                               const uint4 (&pf_res)[CW / 8], bool vec32 = false, float2* vfinal = nullptr) {
  using E = Epi<kEpi>;
  const int ncols = min(CW, p.N - gn0);
  float2 v[CW / 2];
  if (E::scaled(p)) {
    const float2 al = make_float2(p.alpha, p.alpha);
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fmul2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), al);
  } else {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
  }
  if (E::bias(p) == 1) {
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(sbias + 4 * i);   // smem broadcast; zero beyond N
      v[2 * i] = __fadd2_rn(v[2 * i], make_float2(b4.x, b4.y));
      v[2 * i + 1] = __fadd2_rn(v[2 * i + 1], make_float2(b4.z, b4.w));
    }
  } else if (E::bias(p) == 2) {
    const float2 rb = make_float2(rbias, rbias);
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fadd2_rn(v[i], rb);
  }
  if (E::pre(p)) {
    __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(p.pre_out) + off;
    if (vec32) {
#pragma unroll
      for (int i = 0; i < CW / 16; ++i) st_global_256(po + 16 * i, pack_bf16x8_2(v + 8 * i), pack_bf16x8_2(v + 8 * i + 4));
    } else if (vec) {
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
  act_chunk<CW>(v, E::act(p));
  if (E::mul(p) != FFVC_ACT_NONE) {
    float2 x[CW / 2];
    if (vec) {
      unpack_bf16x2N<CW>(pf_aux, x);
    } else {
      const __nv_bfloat16* ax = p.aux + off;
#pragma unroll
      for (int i = 0; i < CW / 2; ++i)
        x[i] = make_float2((2 * i < ncols) ? __bfloat162float(ax[2 * i]) : 0.f, (2 * i + 1 < ncols) ? __bfloat162float(ax[2 * i + 1]) : 0.f);
    }
    mulgrad_chunk<CW>(v, x, E::mul(p));
  }
  if (E::res(p)) {
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
  if (vfinal) {                 // hand the final values to the caller (GroupNorm statistics of the output)
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) vfinal[i] = v[i];
  }
  if (E::f32(p)) {
    float* o = reinterpret_cast<float*>(p.out) + off;
    if (E::atomic(p)) {
      if (ncols == CW && (p.ldc % 4 == 0) && (p.out_bs % 4 == 0) && (p.out_bs_inner % 4 == 0) &&
          (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
        // one 16-byte vector reduction per 4 columns: a quarter of the L2 atomic requests of the scalar form
#pragma unroll
        for (int i = 0; i < CW / 4; ++i)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * i), "f"(v[2 * i].x), "f"(v[2 * i].y),
                       "f"(v[2 * i + 1].x), "f"(v[2 * i + 1].y)
                       : "memory");
      } else {
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) {
          if (2 * i < ncols) atomicAdd(o + 2 * i, v[i].x);
          if (2 * i + 1 < ncols) atomicAdd(o + 2 * i + 1, v[i].y);
        }
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
    if (vec32) {
#pragma unroll
      for (int i = 0; i < CW / 16; ++i) st_global_256(o + 16 * i, pack_bf16x8_2(v + 8 * i), pack_bf16x8_2(v + 8 * i + 4));
    } else if (vec) {
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


// Compute-only form of the fused epilogue for the compile-time configurations (bf16 output, alpha = 1, full vector chunk):
// returns the packed bf16 results (and the packed pre-activation when the configuration has one) instead of storing them —
// the TMA-store epilogue stages them in shared memory.
template <int CW, int kEpi>
__device__ __forceinline__ void epilogue_chunk_pack(const GemmDev& p, const uint32_t (&r)[CW], float rbias, const float* sbias,
                                                    const uint4 (&pf_aux)[CW / 8], const uint4 (&pf_res)[CW / 8],
                                                    uint4 (&out_pk)[CW / 8], uint4 (&pre_pk)[CW / 8]) {
  using E = Epi<kEpi>;
  static_assert(kEpi >= 0, "compile-time epilogue configurations only");
  float2 v[CW / 2];
#pragma unroll
  for (int i = 0; i < CW / 2; ++i) v[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
  if (E::bias(p) == 1) {
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(sbias + 4 * i);
      v[2 * i] = __fadd2_rn(v[2 * i], make_float2(b4.x, b4.y));
      v[2 * i + 1] = __fadd2_rn(v[2 * i + 1], make_float2(b4.z, b4.w));
    }
  } else if (E::bias(p) == 2) {
    const float2 rb = make_float2(rbias, rbias);
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fadd2_rn(v[i], rb);
  }
  if (E::pre(p)) {
#pragma unroll
    for (int i = 0; i < CW / 8; ++i) pre_pk[i] = pack_bf16x8_2(v + 4 * i);
  }
  act_chunk<CW>(v, E::act(p));
  if (E::mul(p) != FFVC_ACT_NONE) {
    float2 x[CW / 2];
    unpack_bf16x2N<CW>(pf_aux, x);
    mulgrad_chunk<CW>(v, x, E::mul(p));
  }
  if (E::res(p)) {
    float2 x[CW / 2];
    unpack_bf16x2N<CW>(pf_res, x);
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) v[i] = __fadd2_rn(v[i], x[i]);
  }
#pragma unroll
  for (int i = 0; i < CW / 8; ++i) out_pk[i] = pack_bf16x8_2(v + 4 * i);
}

}  // namespace ffvc
