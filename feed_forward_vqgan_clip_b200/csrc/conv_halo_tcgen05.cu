// 3x3 convolution (pad 1, stride 1) on NHWC bf16 as an implicit GEMM with HALO REUSE in shared memory, for wide images
// (W a multiple of 128) and Cout <= 128 — the 128-channel layers of the VQGAN decoder at 128x128 and 256x256, which hold
// ~60 % of the decoder's FLOPs (taming Decoder, call site main.py:142).
//
// The tap-by-tap form (ffvc_gemm, CONV3X3 mode) loads one shifted 4-D TMA box per filter tap: every input pixel crosses
// L2 -> SM nine times and those layers are L2-bandwidth bound (profiles/r01_ncu_full_gemm_summary.md).  Here a CTA owns
// 2 output rows x 128 pixels; per 64-channel chunk it loads the (2+2) x (128+2) pixel HALO once (one TMA box, 128B
// swizzle, zero fill = padding) and the nine taps are nine UMMA descriptors into that same buffer: output row `sub`, tap
// (r, s) starts at halo row (sub + r), pixel s — 128 consecutive 128-byte rows.  Those start addresses are 128 B- but not
// 1024 B-aligned.  Measured on B200: the tensor core applies the 128B-swizzle XOR to the ABSOLUTE shared-memory address
// (bits [4:6] ^= bits [7:9]), exactly like TMA does when it writes the box, so the descriptor needs NO base-offset: with
// the field left 0 the results are exact, with ((addr >> 7) & 7) in it they are garbage (both variants were run).
// A-operand traffic drops 9x (66.5 KB per 64-channel chunk instead of 9 x 32 KB); weights stream through a 4-stage ring.
//
// Same warp roles, TMEM double buffering and fused epilogue as gemm_tcgen05_kernel (gemm_common.cuh).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "ffvc_internal.h"
#include "ptx.cuh"
#include "gemm_common.cuh"

namespace ffvc {

static constexpr int kHaloRows = 4;                                   // 2 output rows + 1 above + 1 below
static constexpr int kHaloPix = 130;                                  // 128 output pixels + 1 left + 1 right
static constexpr int kHaloBytes = kHaloRows * kHaloPix * 128;         // 66,560 B per 64-channel chunk
static constexpr int kHaloStride = (kHaloBytes + 1023) / 1024 * 1024; // 67,584 B (keeps every buffer 1024 B-aligned)
static constexpr int kBStages = 4;
static constexpr int kBStageBytes = 128 * 64 * 2;                     // up to 128 output channels x 64 input channels
static constexpr int kHaloSmem = 2 * kHaloStride + kBStages * kBStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*bias*/ +
                                 4096 /*GroupNorm partial sums: [2 tiles][8 warps][4 chunks][16]*/;
// epilogue warps: 8 (one per 32-pixel row group, 32-column chunks) or 16 (two per row group, each 64 of the 128 columns in
// 16-column chunks): the GroupNorm-statistics epilogues do ~20 instructions per element and, on 8 warps (2 per scheduler), pace
// the tile; 16 warps halve the work per warp and double the latency hiding

// descriptor for a K-major, 128B-swizzled operand whose start is 128 B- (not 1024 B-) aligned: base-offset field stays 0
__device__ __forceinline__ uint64_t umma_smem_desc_sw128_off(uint32_t saddr) {
  return umma_smem_desc_sw128(saddr, 16u, 1024u);
}

// kXf: kXfWarps TRANSFORM warps (the last threads of the CTA) sit between the TMA producer and the MMA issuer: when a halo tile has landed
// they rewrite it in place, y = swish(gamma * (x - mean) * rstd + beta) per (image, channel) — taming's Normalize + nonlinearity
// in front of this conv — skip the zero-filled padding pixels (the conv pads the ACTIVATED tensor), fence the
// generic-proxy writes for the tensor core and release the tile to the MMA warp (hready barrier).  A thread owns physical 16-byte
// chunk (t & 7) of lines (t >> 3) + 4 kXfWarps k: because the line step is a multiple of 8, the 128B-swizzle phase (line & 7) — hence the
// LOGICAL chunk, hence the 8 channels and their scale / shift — is fixed per thread.  swish(t) = h + h tanh(h), h = t / 2: two FFMA
// and one MUFU.TANH per element; 33 K elements per tile chunk = 2080 clk of MUFU against ~4700 clk of MMAs on that chunk.
static constexpr int kXfWarps = 4;   // transform warps.  8 (two per scheduler) measured the same: the transform is not issue-bound, it
                                     // competes with the tensor core for the shared-memory port (profiles/r02_gn_apply_fusion.md)

template <int kEW, bool kXf>
__global__ void __launch_bounds__(64 + 32 * kEW + (kXf ? 32 * kXfWarps : 0), 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t halo_base = smem_base;                               // [2][kHaloStride]
  const uint32_t b_base = smem_base + 2 * kHaloStride;                // [kBStages][kBStageBytes]
  const uint32_t bar_base = b_base + kBStages * kBStageBytes;
  auto hfull_bar = [&](int i) { return bar_base + 8u * i; };          // 2
  auto hempty_bar = [&](int i) { return bar_base + 8u * (2 + i); };   // 2
  auto bfull_bar = [&](int s) { return bar_base + 8u * (4 + s); };    // 4
  auto bempty_bar = [&](int s) { return bar_base + 8u * (8 + s); };   // 4
  auto tfull_bar = [&](int a) { return bar_base + 8u * (12 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (14 + a); };
  auto hready_bar = [&](int i) { return bar_base + 8u * (18 + i); };  // 2: halo tile transformed (kXf)
  const uint32_t tmem_slot = bar_base + 8u * 16;
  const uint32_t bias_smem = bar_base + 256u;
  const uint32_t gn_smem = bias_smem + 2048u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(hfull_bar(i), 1);
      mbar_init(hempty_bar(i), 1);
      mbar_init(hready_bar(i), kXfWarps);       // one arrival per transform warp
    }
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(bfull_bar(s), 1);
      mbar_init(bempty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEW);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // tile = image n, output rows (2*ty, 2*ty+1), pixels [128*tx, 128*tx + 128)
  const int H = p.conv_h, W = p.conv_w;
  const int tiles_x = W / 128, tiles_y = H / 2;
  const long long total_tiles = (long long)p.batch * tiles_y * tiles_x;   // p.batch = number of images
  const int cblocks = p.conv_cblocks;                                     // Cin / 64
  const uint32_t b_bytes = (uint32_t)p.block_n * 128u;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int hb = 0, bs = 0;
      uint32_t hphase = 0, bphase = 0;
      for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int tx = (int)(t % tiles_x);
        const int ty = (int)((t / tiles_x) % tiles_y);
        const int img = (int)(t / ((long long)tiles_x * tiles_y));
        for (int c = 0; c < cblocks; ++c) {
          mbar_wait(hempty_bar(hb), hphase ^ 1u);
          mbar_expect_tx(hfull_bar(hb), (uint32_t)kHaloBytes);
          tma_load_4d(halo_base + hb * kHaloStride, &tmap_x, hfull_bar(hb), c * 64, tx * 128 - 1, ty * 2 - 1, img);
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(bempty_bar(bs), bphase ^ 1u);
            mbar_expect_tx(bfull_bar(bs), b_bytes);
            tma_load_4d(b_base + bs * kBStageBytes, &tmap_w, bfull_bar(bs), (tap * cblocks + c) * 64, 0, 0, 0);
            if (++bs == kBStages) {
              bs = 0;
              bphase ^= 1u;
            }
          }
          if (++hb == 2) {
            hb = 0;
            hphase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m(128, p.block_n, 0, 0);
      int hb = 0, bs = 0, acc = 0;
      uint32_t hphase = 0, bphase = 0, acc_phase = 0;
      for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
        for (int c = 0; c < cblocks; ++c) {
          mbar_wait(kXf ? hready_bar(hb) : hfull_bar(hb), hphase);
          tc_fence_after();
          const uint32_t halo = halo_base + hb * kHaloStride;
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(bfull_bar(bs), bphase);
            tc_fence_after();
            const uint32_t sb = b_base + bs * kBStageBytes;
            const int r = tap / 3, s = tap % 3;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t db = umma_smem_desc_sw128(sb + k * 32u, 16u, 1024u);
              const uint32_t accum = (c > 0 || tap > 0 || k > 0) ? 1u : 0u;
#pragma unroll
              for (int sub = 0; sub < 2; ++sub) {
                const uint32_t arow = (uint32_t)((sub + r) * kHaloPix + s);       // first halo row of this A tile
                const uint64_t da = umma_smem_desc_sw128_off(halo + arow * 128u + k * 32u);
                umma_bf16(tmem_d + sub * 128, da, db, idesc, accum);
              }
            }
            umma_commit(bempty_bar(bs));
            if (++bs == kBStages) {
              bs = 0;
              bphase ^= 1u;
            }
          }
          umma_commit(hempty_bar(hb));   // halo buffer free once its 9 x 8 MMAs retired
          if (++hb == 2) {
            hb = 0;
            hphase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (kXf && warp >= 2 + kEW) {
    // ------------------------------------------------------------------ transform warps: GroupNorm + swish on the halo tile, in place
    const int xt = threadIdx.x - (64 + 32 * kEW);        // 0 .. 32 * kXfWarps - 1
    const int pch = xt & 7, lrow = xt >> 3;              // physical 16-byte chunk, first line
    const int j = pch ^ (lrow & 7);                      // logical chunk: channels 8 j .. 8 j + 7 of the 64-channel block
    int hb = 0;
    uint32_t hphase = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int tx = (int)(t % tiles_x);
      const int ty = (int)((t / tiles_x) % tiles_y);
      const int img = (int)(t / ((long long)tiles_x * tiles_y));
      const int y0 = ty * 2 - 1, x0 = tx * 128 - 1;
      for (int c = 0; c < cblocks; ++c) {
        // half-scale / half-shift of this thread's 8 channels (fetched while the tile is still in flight)
        float sc[8], sh[8];
        const int ch0 = c * 64 + j * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ch = ch0 + k;
          const int g = img * p.xf_groups + ch / p.xf_cpg;
          const float rs = p.xf_rstd[g], ga = p.xf_gamma[ch];
          sc[k] = 0.5f * rs * ga;
          sh[k] = 0.5f * (p.xf_beta[ch] - p.xf_mean[g] * rs * ga);
        }
        mbar_wait(hfull_bar(hb), hphase);
        const uint32_t tile = halo_base + hb * kHaloStride + (uint32_t)pch * 16u;
        // four lines per trip (loads first, then the arithmetic, then the stores): a thread's 33 lines are a dependent
        // LDS -> FFMA -> MUFU -> FFMA -> STS chain each, and with one transform warp per scheduler nothing else hides its latency
        constexpr int kLineStep = 4 * kXfWarps;            // lines between a thread's consecutive lines (a multiple of 8)
        for (int line0 = lrow; line0 < kHaloRows * kHaloPix; line0 += 4 * kLineStep) {
          uint4 v[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int line = line0 + kLineStep * u;
            const int r = line / kHaloPix, px = line - r * kHaloPix;
            const int yy = y0 + r, xx = x0 + px;
            ok[u] = line < kHaloRows * kHaloPix && yy >= 0 && yy < H && xx >= 0 && xx < W;      // padding pixels stay zero
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            if (ok[u]) {
              const uint32_t a = tile + (uint32_t)line * 128u;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "r"(a));
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[k]));
              const float h0 = fmaf(f.x, sc[2 * k], sh[2 * k]), h1 = fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]);
              const __nv_bfloat162 o = __floats2bfloat162_rn(fmaf(h0, tanh_approx(h0), h0), fmaf(h1, tanh_approx(h1), h1));
              w4[k] = *reinterpret_cast<const uint32_t*>(&o);
            }
            v[u] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (ok[u]) st_shared_v4(tile + (uint32_t)(line0 + kLineStep * u) * 128u, v[u]);
        }
        fence_proxy_async_smem();                          // generic-proxy writes -> visible to tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(hready_bar(hb));
        if (++hb == 2) {
          hb = 0;
          hphase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9): sub-tile = output row
    const int q = warp & 3;
    const int ew = warp - 2;                  // 0 .. kEW-1
    const int sub = (ew >> 2) & 1;            // output row of the tile
    const int rowwarp = ew & 7;               // 32-pixel row group (sub, q)
    const int c_begin = (ew >> 3) * (p.block_n / (kEW / 8));   // kEW = 16: the upper 8 warps take the upper half of the columns
    const int c_end = c_begin + p.block_n / (kEW / 8);
    const int etid = threadIdx.x - 64;
    float* sbias_all = reinterpret_cast<float*>(smem_raw + (bias_smem - smem_u32(smem_raw)));
    float* gn_part = reinterpret_cast<float*>(smem_raw + (gn_smem - smem_u32(smem_raw)));   // [2 (acc)][8 row groups][2 kinds][32 groups]
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (p.ldc % 8 == 0);
    // 32-byte (256-bit) row accesses when every row starts on a 32-byte boundary (bf16 output): half the L1 requests
    const bool vec32_ok = vec_ok && !p.out_fp32 && (p.ldc % 16 == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.aux) | reinterpret_cast<uintptr_t>(p.res)) & 31) == 0;
    const bool gnb = !kXf && p.gn_ws && p.gn_bwd;          // (the forward-only transform variant never takes backward statistics)
    const bool want_aux = p.mul_mode != FFVC_ACT_NONE || gnb, want_res = p.res != nullptr;
    if (gnb && etid < 128) {                  // gamma / beta of the Normalize, once per CTA (the bias staging area is free: dgrad)
      sbias_all[etid] = p.gn_gamma[etid];
      sbias_all[128 + etid] = p.gn_beta[etid];
    }
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int tx = (int)(t % tiles_x);
      const int ty = (int)((t / tiles_x) % tiles_y);
      const int img = (int)(t / ((long long)tiles_x * tiles_y));
      const long long gm = ((long long)img * H + (ty * 2 + sub)) * W + tx * 128 + q * 32 + lane;   // flattened pixel index
      const long long row_off = gm * p.ldc;
      float* sbias = sbias_all + acc * 256;
      if (p.bias_mode == 1) {
        if (etid < p.block_n) sbias[etid] = (etid < p.N) ? p.bias[etid] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
      }
      float* gstat = sbias_all + 256 + acc * 64;       // this tile's image: [32] rstd, [32] -mean * rstd
      if (gnb) {
        if (etid < 32) {
          const float rs = p.gn_rstd[(long long)img * 32 + etid];
          gstat[etid] = rs;
          gstat[32 + etid] = -p.gn_mean[(long long)img * 32 + etid] * rs;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
      }
      constexpr int CW = kEW == 16 ? 16 : 32;
      constexpr int GPC = CW / 4;              // GroupNorm groups (4 channels) per chunk
      uint4 pf_aux[CW / 8], pf_res[CW / 8];
      auto prefetch = [&](int c) {
        if (vec_ok && c + CW <= p.N) {
          if (want_aux) {
            if (vec32_ok) {
#pragma unroll
              for (int j = 0; j < CW / 16; ++j) ld_global_256(p.aux + row_off + c + 16 * j, pf_aux[2 * j], pf_aux[2 * j + 1]);
            } else {
              const uint4* ax = reinterpret_cast<const uint4*>(p.aux + row_off + c);
#pragma unroll
              for (int j = 0; j < CW / 8; ++j) pf_aux[j] = ax[j];
            }
          }
          if (want_res) {
            if (vec32_ok) {
#pragma unroll
              for (int j = 0; j < CW / 16; ++j) ld_global_256(p.res + row_off + c + 16 * j, pf_res[2 * j], pf_res[2 * j + 1]);
            } else {
              const uint4* rs = reinterpret_cast<const uint4*>(p.res + row_off + c);
#pragma unroll
              for (int j = 0; j < CW / 8; ++j) pf_res[j] = rs[j];
            }
          }
        }
      };
      if (want_aux || want_res) prefetch(c_begin);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      for (int c = c_begin; c < c_end; c += CW) {
        uint32_t r[CW];
        const uint32_t taddr = tmem_base + (uint32_t)(acc * 256 + sub * 128 + c) + ((uint32_t)(q * 32) << 16);
        if constexpr (CW == 32) tmem_ld_32x32(taddr, r);
        else tmem_ld_32x16(taddr, r);
        tmem_ld_wait();
        uint4 cur_aux[CW / 8], cur_res[CW / 8];
        if (want_aux) {
#pragma unroll
          for (int j = 0; j < CW / 8; ++j) cur_aux[j] = pf_aux[j];
        }
        if (want_res) {
#pragma unroll
          for (int j = 0; j < CW / 8; ++j) cur_res[j] = pf_res[j];
        }
        if ((want_aux || want_res) && c + CW < c_end) prefetch(c + CW);
        if (!p.gn_ws) {
          if (c < p.N)
            epilogue_chunk<CW>(p, r, c, row_off + c, 0.f, vec_ok && (c + CW <= p.N), sbias + c, cur_aux, cur_res, vec32_ok && (c + CW <= p.N));
        } else {
          // GroupNorm(32) statistics of the tensor being written (Cout = 128: 4 channels per group, 8 groups per 32-column
          // chunk), taken from the bf16-rounded values a later pass over the tensor would read.  Per lane 8 sums + 8 sums of
          // squares; a recursive-halving butterfly (8+4+2+1+1 = 16 shuffles) leaves lane L with kind (L>>4) of group (L>>1)&7
          // summed over the warp's 32 pixels; one slot per warp and chunk in shared memory, folded per tile below.
          float2 vf[CW / 2];
          epilogue_chunk<CW>(p, r, c, row_off + c, 0.f, vec_ok, sbias + c, cur_aux, cur_res, vec32_ok, vf);
          float vals[2 * GPC];
          if (!gnb) {
#pragma unroll
            for (int g = 0; g < GPC; ++g) {
              const float2 a = __bfloat1622float2(__floats2bfloat162_rn(vf[2 * g].x, vf[2 * g].y));
              const float2 b = __bfloat1622float2(__floats2bfloat162_rn(vf[2 * g + 1].x, vf[2 * g + 1].y));
              vals[g] = (a.x + a.y) + (b.x + b.y);
              vals[GPC + g] = fmaf(a.x, a.x, a.y * a.y) + fmaf(b.x, b.x, b.y * b.y);
            }
          } else {
            // backward statistics of the Normalize + swish whose output gradient this conv just produced (vf = dy, aux = its
            // input x): g = dy * swish'(gamma * xhat + beta) * gamma;  vals = sum g | sum g * xhat per group (groupnorm_bwd_stats_kernel)
            float2 xx[CW / 2];
            unpack_bf16x2N<CW>(cur_aux, xx);
            // packed (FFMA2) arithmetic, gamma / beta of a group as one 16-byte shared-memory broadcast each: the epilogue warps
            // pace the tile in this form, so instructions per element are what counts
            const float4* sg4 = reinterpret_cast<const float4*>(sbias_all + c);
            const float4* sb4 = reinterpret_cast<const float4*>(sbias_all + 128 + c);
            const float2 half2 = make_float2(0.5f, 0.5f);
#pragma unroll
            for (int g = 0; g < GPC; ++g) {
              const float rs = gstat[(c >> 2) + g], mrs = gstat[32 + (c >> 2) + g];
              const float2 rs2 = make_float2(rs, rs), mrs2 = make_float2(mrs, mrs);
              const float4 G4 = sg4[g], B4 = sb4[g];
              float2 as = make_float2(0.f, 0.f), aq = make_float2(0.f, 0.f);
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const float2 gam = h2 ? make_float2(G4.z, G4.w) : make_float2(G4.x, G4.y);
                const float2 bet = h2 ? make_float2(B4.z, B4.w) : make_float2(B4.x, B4.y);
                const float2 d = __bfloat1622float2(__floats2bfloat162_rn(vf[2 * g + h2].x, vf[2 * g + h2].y));
                const float2 xh = __ffma2_rn(xx[2 * g + h2], rs2, mrs2);
                const float2 u = __ffma2_rn(xh, gam, bet);
                const float2 hu = __fmul2_rn(u, half2);
                const float2 sg = __ffma2_rn(make_float2(tanh_approx(hu.x), tanh_approx(hu.y)), half2, half2);   // sigmoid(u)
                const float2 w2 = __ffma2_rn(make_float2(-sg.x, -sg.y), u, u);                                  // u (1 - s)
                const float2 sw = __ffma2_rn(sg, w2, sg);                                                       // swish'(u)
                const float2 gg = __fmul2_rn(__fmul2_rn(d, sw), gam);
                as = __fadd2_rn(as, gg);
                aq = __ffma2_rn(gg, xh, aq);
              }
              vals[g] = as.x + as.y;
              vals[GPC + g] = aq.x + aq.y;
            }
          }
          // recursive halving over the warp's 32 pixels: after the exchange with lane ^ 16 a lane keeps the sums (bit 4 = 0) or the
          // second kind (bit 4 = 1); every further step halves the groups it keeps; the last steps are plain butterfly sums
          const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
          float h1[GPC];
#pragma unroll
          for (int i = 0; i < GPC; ++i) {
            const float send = b4 ? vals[i] : vals[i + GPC], keep = b4 ? vals[i + GPC] : vals[i];
            h1[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
          float h2[GPC / 2];
#pragma unroll
          for (int i = 0; i < GPC / 2; ++i) {
            const float send = b3 ? h1[i] : h1[i + GPC / 2], keep = b3 ? h1[i + GPC / 2] : h1[i];
            h2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          float h3[GPC / 4];
#pragma unroll
          for (int i = 0; i < GPC / 4; ++i) {
            const float send = b2 ? h2[i] : h2[i + GPC / 4], keep = b2 ? h2[i + GPC / 4] : h2[i];
            h3[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
          float fin;
          int gl;                                     // group of the chunk this lane ends up with
          if constexpr (GPC == 8) {
            fin = (b1 ? h3[1] : h3[0]) + __shfl_xor_sync(0xffffffffu, b1 ? h3[0] : h3[1], 2);
            gl = (lane >> 1) & 7;
          } else {
            fin = h3[0] + __shfl_xor_sync(0xffffffffu, h3[0], 2);
            gl = (lane >> 2) & 3;
          }
          fin += __shfl_xor_sync(0xffffffffu, fin, 1);
          // slot [tile parity][row group][kind][group 0..31]: written exactly once per tile, by the warp that owns these columns
          if ((lane & (GPC == 8 ? 1 : 3)) == 0) gn_part[(((acc * 8 + rowwarp) * 2 + (lane >> 4)) << 5) + (c >> 2) + gl] = fin;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));           // the accumulator is free for the MMA warp before the fold below
      if (p.gn_ws) {
        // fold the 8 warps' slots of this tile (fixed order): thread (kind, group) -> one partial per (image, tile, kind, group)
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kEW) : "memory");
        if (etid < 64) {
          const int kind = etid >> 5, g = etid & 31;
          float tot = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) tot += gn_part[(((acc * 8 + w8) * 2 + kind) << 5) + g];
          // this tile's slot of the partials [n][tile][kind][group] behind the [n][group][2] block: exactly one writer, no atomics —
          // the fold over the tiles runs later in a fixed order (groupnorm_finalize / groupnorm_bwd_apply), so the statistics
          // are reproducible bit for bit and independent of the batch the image is part of
          const long long tile_in_img = (long long)ty * tiles_x + tx;
          p.gn_ws[(long long)p.batch * 64 + (((long long)img * tiles_y * tiles_x + tile_in_img) * 2 + kind) * 32 + g] = (double)tot;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*PFN_encodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode4(CUtensorMap* m, const void* base, const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* box) {
  static PFN_encodeTiled2 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr)
      return set_error(FFVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    fn = reinterpret_cast<PFN_encodeTiled2>(ptr);
  }
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bdim[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
  }
  for (int i = 1; i < 4; ++i) gstr[i - 1] = strides_elems[i] * 2ull;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(FFVC_ERR_CUDA, "conv_halo: cuTensorMapEncodeTiled failed");
  return FFVC_OK;
}

}  // namespace ffvc

using namespace ffvc;

// x: [n][h][w][cin] bf16 NHWC; w: [cout][9][cin] bf16 (tap-major); out: [n*h*w][ldc] bf16 (fp32 when out_fp32).  Epilogue fields as ffvc_gemm.
static int conv3x3_halo_launch(const void* x, const void* w, void* out, int n, int h, int wd, int cin, int cout, long long ldc,
                               const float* bias, const void* res, const void* aux, int mul_mode, int act, int out_fp32,
                               double* gn_ws, void* stream, const float* gnb_mean = nullptr, const float* gnb_rstd = nullptr,
                               const float* gnb_gamma = nullptr, const float* gnb_beta = nullptr, const float* xf_mean = nullptr,
                               const float* xf_rstd = nullptr, const float* xf_gamma = nullptr, const float* xf_beta = nullptr,
                               int xf_groups = 0) {
  if (!x || !w || !out) return set_error(FFVC_ERR_ARG, "conv_halo: null pointer");
  if (gn_ws && (cout != 128 || out_fp32 || ldc % 16 != 0))
    return set_error(FFVC_ERR_UNSUPPORTED, "conv_halo: epilogue GroupNorm statistics need Cout = 128 (32 groups of 4), bf16 output");
  if (wd % 128 != 0 || h % 2 != 0) return set_error(FFVC_ERR_UNSUPPORTED, "conv_halo: needs W % 128 == 0 and even H");
  if (cin % 64 != 0 || cout < 1 || cout > 128) return set_error(FFVC_ERR_UNSUPPORTED, "conv_halo: needs Cin % 64 == 0, Cout <= 128");
  static bool attr_done = false;
  static int num_sms = 0;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem);
    if (e != cudaSuccess) return set_error(FFVC_ERR_CUDA, cudaGetErrorString(e));
    e = cudaFuncSetAttribute(conv3x3_halo_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem);
    if (e != cudaSuccess) return set_error(FFVC_ERR_CUDA, cudaGetErrorString(e));
    e = cudaFuncSetAttribute(conv3x3_halo_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem);
    if (e != cudaSuccess) return set_error(FFVC_ERR_CUDA, cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_done = true;
  }
  const int block_n = cout > 64 ? 128 : (cout > 32 ? 64 : 32);
  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {(uint64_t)cin, (uint64_t)wd, (uint64_t)h, (uint64_t)n};
    uint64_t str[4] = {1, (uint64_t)cin, (uint64_t)wd * cin, (uint64_t)h * wd * cin};
    uint32_t box[4] = {64, (uint32_t)kHaloPix, (uint32_t)kHaloRows, 1};
    int rc = encode4(&tx, x, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)9 * cin, (uint64_t)cout, 1, 1};
    uint64_t str[4] = {1, (uint64_t)9 * cin, (uint64_t)9 * cin * 8, (uint64_t)9 * cin * 8};
    uint32_t box[4] = {64, (uint32_t)block_n, 1, 1};
    int rc = encode4(&tw, w, dims, str, box);
    if (rc) return rc;
  }
  GemmDev p;
  memset(&p, 0, sizeof(p));
  p.M = n * h * wd;
  p.N = cout;
  p.batch = n;
  p.batch_inner = 1;
  p.tile_m = 256;
  p.block_n = block_n;
  p.conv_h = h;
  p.conv_w = wd;
  p.conv_cblocks = cin / 64;
  p.out = out;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(aux);
  p.res = reinterpret_cast<const __nv_bfloat16*>(res);
  p.bias = bias;
  p.ldc = ldc;
  p.out_fp32 = out_fp32 ? 1 : 0;
  p.bias_mode = bias ? 1 : 0;
  p.act = act;
  p.mul_mode = aux ? mul_mode : 0;
  p.alpha = 1.0f;
  p.gn_ws = gn_ws;
  if (gnb_mean) {
    p.gn_bwd = 1;
    p.gn_mean = gnb_mean;
    p.gn_rstd = gnb_rstd;
    p.gn_gamma = gnb_gamma;
    p.gn_beta = gnb_beta;
  }
  if (xf_mean) {
    if (!xf_rstd || !xf_gamma || !xf_beta || xf_groups <= 0 || cin % xf_groups != 0)
      return set_error(FFVC_ERR_ARG, "conv_halo_xf: statistics / affine pointers and a group count dividing Cin are required");
    p.xf_mean = xf_mean;
    p.xf_rstd = xf_rstd;
    p.xf_gamma = xf_gamma;
    p.xf_beta = xf_beta;
    p.xf_groups = xf_groups;
    p.xf_cpg = cin / xf_groups;
  }
  const long long tiles = (long long)n * (h / 2) * (wd / 128);
  const int grid = (int)(tiles < sm_budget(num_sms) ? tiles : sm_budget(num_sms));
  // 16 epilogue warps for the GroupNorm-statistics epilogues (option "halo_epi16": 1 = backward statistics, 2 = forward too)
  const int epi16 = gn_ws ? option(OPT_HALO_EPI16) : 0;
  if (xf_mean)
    conv3x3_halo_kernel<8, true><<<grid, 64 + 32 * 8 + 32 * kXfWarps, kHaloSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tx, tw, p);
  else if (block_n == 128 && ((epi16 >= 1 && p.gn_bwd) || epi16 >= 2))
    conv3x3_halo_kernel<16, false><<<grid, 64 + 32 * 16, kHaloSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tx, tw, p);
  else
    conv3x3_halo_kernel<8, false><<<grid, 64 + 32 * 8, kHaloSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tx, tw, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFVC_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return FFVC_OK;
}

extern "C" int ffvc_conv3x3_halo(const void* x, const void* w, void* out, int n, int h, int wd, int cin, int cout, long long ldc,
                                 const float* bias, const void* res, const void* aux, int mul_mode, int act, int out_fp32, void* stream) {
  return conv3x3_halo_launch(x, w, out, n, h, wd, cin, cout, ldc, bias, res, aux, mul_mode, act, out_fp32, nullptr, stream);
}

// Same convolution; the epilogue also accumulates, per (image, group), the sum and the sum of squares of the tensor it stores
// (GroupNorm(32, 128) statistics: taming `Normalize` of the NEXT layer) into gn_ws[n][32][2] doubles
// (zeroed here) — ffvc_groupnorm_finalize turns them into mean / rstd, and the separate statistics pass over the tensor
// (one full read) is not needed.
extern "C" int ffvc_conv3x3_halo_gn(const void* x, const void* w, void* out, int n, int h, int wd, int cin, int cout, long long ldc,
                                    const float* bias, const void* res, double* gn_ws, void* stream) {
  if (!gn_ws) return set_error(FFVC_ERR_ARG, "conv_halo_gn: null statistics workspace");
  return conv3x3_halo_launch(x, w, out, n, h, wd, cin, cout, ldc, bias, res, nullptr, 0, 0, 0, gn_ws, stream);
}

// dgrad form with the backward statistics of the Normalize + swish in FRONT of the forward conv: `out` (= dy of that layer,
// Cout = 128 channels) is stored as usual; the epilogue also reads the layer's input gn_x at the same positions and
// accumulates sum g and sum g * xhat per (image, group), g = dy * swish'(gamma * xhat + beta) * gamma, into gn_ws[n][32][2]
// doubles (zeroed here) — exactly what ffvc_groupnorm_bwd's first pass computes from two more reads of dy and x;
// ffvc_groupnorm_bwd_apply finishes the job.
extern "C" int ffvc_conv3x3_halo_gnbwd(const void* x, const void* w, void* out, int n, int h, int wd, int cin, int cout, long long ldc,
                                       const void* res, const void* gn_x, const float* gn_mean, const float* gn_rstd,
                                       const float* gn_gamma, const float* gn_beta, double* gn_ws, void* stream) {
  if (!gn_ws || !gn_x || !gn_mean || !gn_rstd || !gn_gamma || !gn_beta) return set_error(FFVC_ERR_ARG, "conv_halo_gnbwd: null pointer");
  return conv3x3_halo_launch(x, w, out, n, h, wd, cin, cout, ldc, nullptr, res, gn_x, 0, 0, 0, gn_ws, stream, gn_mean, gn_rstd,
                             gn_gamma, gn_beta);
}

// The conv applied to swish(GroupNorm(x)) without that tensor ever existing: x is the RAW input of taming's Normalize + nonlinearity
// in front of this conv; four transform warps normalise every halo tile in shared memory between the TMA load and the MMAs (see
// conv3x3_halo_kernel).  xf_mean / xf_rstd: [n][xf_groups] statistics of x (ffvc_groupnorm_stats / _finalize), xf_gamma / xf_beta:
// [cin].  gn_ws (optional, Cout = 128): also take the GroupNorm(32) statistics of the OUTPUT in the epilogue, as ffvc_conv3x3_halo_gn.
extern "C" int ffvc_conv3x3_halo_xf(const void* x, const void* w, void* out, int n, int h, int wd, int cin, int cout, long long ldc,
                                    const float* bias, const void* res, const float* xf_mean, const float* xf_rstd,
                                    const float* xf_gamma, const float* xf_beta, int xf_groups, double* gn_ws, void* stream) {
  if (!xf_mean) return set_error(FFVC_ERR_ARG, "conv_halo_xf: null statistics");
  return conv3x3_halo_launch(x, w, out, n, h, wd, cin, cout, ldc, bias, res, nullptr, 0, 0, 0, gn_ws, stream, nullptr, nullptr, nullptr,
                             nullptr, xf_mean, xf_rstd, xf_gamma, xf_beta, xf_groups);
}
