// tcgen05 / TMEM / TMA GEMM for sm_100a — the tensor-core workhorse of the train step.
//
//   D[b][m][n] (+)= alpha * sum_k A[m][k] * B[n][k]      (bf16 operands, fp32 accumulate in TMEM)
//
// One persistent CTA per SM, 320 threads, warp-specialised:
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 4-stage smem ring, 128B swizzle)
//   warp 1      : MMA issuer    (one thread issues tcgen05.mma M=128, N=BLOCK_N, K=16; 4 (x2) per stage)
//   warps 2..9  : epilogue      (tcgen05.ld 32x32b -> registers -> bias/activation/residual -> global); two warps per
//                 TMEM lane quarter so the fused epilogue math never paces the tensor pipe
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the main loop of
// tile i+1.  Two tile shapes: 128 x {32..256} (one MMA per k-step) and 256 x {32..128} (two MMAs sharing the B
// tile: for N <= 128 the 128-row tile is shared-memory-bandwidth bound, the 256-row tile halves B traffic per FLOP).
//
// Operand modes (runtime, resolved by the producer / descriptor builder only):
//   K-major   : operand stored [rows][K] with K contiguous        (fwd Linear:  X[M,K], W[N,K])
//   MN-major  : operand stored [K][rows] with rows contiguous     (dgrad: W as B; wgrad: dY, X)
//   CONV3X3   : A is an NHWC image; the 9 taps x Cin/64 channel chunks form the K loop, each
//               A tile is one 4-D TMA box (zero fill outside the image = padding 1)
// Batching: output batch index b = (b_outer, b_inner); operands are 4-D tensor maps (inner, rows, b_inner, b_outer)
// so attention heads ([N][T][heads*64] slices) and per-sample token-mixing are plain strides.
// Replaces, on the reference's path, every nn.Linear / Conv1d(k=1) of the mappers
// (mlp_mixer_pytorch.py:16-23,32; vitgan.py:31-33,64-67), taming's Conv2d 3x3 / 1x1 in the VQGAN decoder
// (call site main.py:142), and the CLIP ViT linears (cloob.py:188-196,224,249).
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

static constexpr int kBlockK = 64;
static constexpr int kStages = 4;
static constexpr int kStageABytes = 256 * kBlockK * 2;      // up to 256 rows of A  (32 KB)
static constexpr int kStageBytes = 48 * 1024;               // A (16|32 KB) + B (32|16 KB)
static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*bias*/;
static constexpr int kNumThreads = 320;
static constexpr int kNumEpiWarps = 8;
static constexpr int kTmemCols = 512;

// Work decomposition shared by the three warp roles (they must walk the same sequence).
//   regular : output tiles (x split-K slices) are dealt round-robin to the persistent CTAs / CTA pairs;
//   stream-K: the linearised (tile, k-block) space is cut into one CONTIGUOUS range per worker, so every worker gets the
//             same number of k-blocks whatever the tile count (64 pair tiles on 74 CTA pairs ran at 86 %); a worker touches
//             at most two partial tiles and accumulates them with the fp32 atomic epilogue (wgrad GEMMs only).
struct WorkIter {
  long long t, t_step, t_end, tiles_all;
  long long u, u_end;
  int total_kb, splits, stream_k;
  __device__ __forceinline__ WorkIter(const GemmDev& p, long long worker, long long workers, long long tiles_all_, int total_kb_)
      : t(worker), t_step(workers), t_end(tiles_all_ * p.splits), tiles_all(tiles_all_), total_kb(total_kb_), splits(p.splits),
        stream_k(p.stream_k) {
    const long long units = tiles_all_ * total_kb_;
    u = units * worker / workers;
    u_end = units * (worker + 1) / workers;
  }
  // next work item: tile index `rem` (over tiles x batches) and its k-block range; false when this worker is done
  __device__ __forceinline__ bool next(int& rem, int& kb_begin, int& kb_end) {
    if (stream_k) {
      if (u >= u_end) return false;
      rem = (int)(u / total_kb);
      kb_begin = (int)(u % total_kb);
      const long long left = u_end - u;
      kb_end = (left < (long long)(total_kb - kb_begin)) ? kb_begin + (int)left : total_kb;
      u += kb_end - kb_begin;
      return true;
    }
    if (t >= t_end) return false;
    const int split = (int)(t / tiles_all);
    rem = (int)(t % tiles_all);
    kb_begin = (int)((long long)total_kb * split / splits);
    kb_end = (int)((long long)total_kb * (split + 1) / splits);
    t += t_step;
    return true;
  }
};

// kTwoCta = false: one CTA per tile (128 or 256 rows).
// kTwoCta = true : a CTA PAIR (cluster of 2, cta_group::2) per 256-row tile — each CTA stages its own 128 rows of A and
//   HALF of the B tile; the leader CTA issues tcgen05.mma M=256 that reads both CTAs' shared memory, so every SM
//   reads / is written only (A + B/2) per k-step instead of (A + B): the 1-CTA kernel is shared-memory-bandwidth
//   bound (128 B/clk/SM) at ~55-65 % of the tensor pipe, the pair removes a third of that traffic.  6-stage ring.
template <bool kTwoCta, int kEpiWarps, int kEpi, bool kTma, bool kQuad = false>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_pre, const GemmDev p) {
  // TMA-store epilogue (kTma: compile-time epilogues on the CTA-pair kernel, short K): the ring shrinks to 4 stages and the
  // freed 64 KB hold one [pre-activation | output] staging buffer (2 x 2 KB) per epilogue warp.  Measured: it pays for
  // K <= 512 (the epilogue dominates and 4 stages cover the whole K loop); for K = 1024 the shallower ring costs more than the
  // store offload returns, so the host only selects it for short K.
  constexpr bool kCanTmaStore = kTma && kTwoCta && kEpiWarps == 16 && kEpi >= 0;
  constexpr int kStagesT = kTwoCta ? (kCanTmaStore ? 4 : 6) : kStages;
  constexpr int kStageBytesT = kTwoCta ? 32 * 1024 : kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;   // same offset for both variants (6*32 KB == 4*48 KB)
  // barrier layout (8 bytes each): full[6], empty[6], tmem_full[2], tmem_empty[2], then tmem ptr slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (6 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (12 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (14 + a); };
  const uint32_t tmem_slot = bar_base + 8u * 16;
  const uint32_t bias_smem = bar_base + 256u;   // [2][256] floats, 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair: rank = 0 (leader, issues the MMAs) / 1.  Quad (kQuad, launched in clusters of 4 when p.quad): pair_id = 0 / 1 are two pair tiles adjacent in M
  // that share one B tile: every CTA fetches only HALF of its B rows and TMA-multicasts them to its counterpart in the other pair
  // (24 KB instead of 32 KB of L2 -> SM traffic per CTA and k-block; the large GEMMs run at the chip's L2 -> SM limit, ~6300 B/clk:
  // tensor pipe 65 % active with every other unit below 40 %, profiles/r02_ncu_gemm_family.md)
  const uint32_t rank4 = kTwoCta ? cluster_ctarank() : 0u;
  const uint32_t rank = kQuad ? (rank4 & 1u) : rank4;
  // kQuad is a compile-time variant (generic epilogue only): as a run-time flag it cost the default pair kernels 1 % of the step
  constexpr bool quad = kTwoCta && kQuad;
  const uint32_t pair_id = quad ? (rank4 >> 1) : 0u;
  const uint32_t leader_cta = quad ? (rank4 & ~1u) : 0u;
  const uint16_t pair_mask = quad ? (uint16_t)(3u << (2u * pair_id)) : (uint16_t)3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStagesT; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), quad ? 2 : 1);     // quad: a stage is free once BOTH pairs' MMAs have retired (multicast destinations)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kTwoCta ? 2 * kEpiWarps : kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kTwoCta) {
      tmem_alloc_2sm(tmem_slot, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (kTwoCta) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tiles_m = (p.M + p.tile_m - 1) / p.tile_m;
  const int tiles_n = (p.N + p.block_n - 1) / p.block_n;
  const int tiles_m_it = quad ? (tiles_m >> 1) : tiles_m;     // quad: one work item = two M-adjacent pair tiles
  const int tiles_per_batch = tiles_m_it * tiles_n;
  const int total_kb = p.kb_per_seg * p.k_segs;
  const long long tiles_all_batches = (long long)tiles_per_batch * p.batch;
  const int wshift = kTwoCta ? (quad ? 2 : 1) : 0;
  const long long tile_first = blockIdx.x >> wshift;
  const long long tile_step = gridDim.x >> wshift;
  const int rows_cta = kTwoCta ? 128 : p.tile_m;             // rows of A this CTA stages
  const int bn_cta = kTwoCta ? (p.block_n >> 1) : p.block_n;  // rows of B this CTA stages
  const uint32_t a_bytes = (uint32_t)rows_cta * kBlockK * 2;
  const uint32_t b_off = a_bytes;  // B follows A inside a stage

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t b_bytes = (uint32_t)bn_cta * kBlockK * 2;
      auto tma4 = [&](uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2, int c3) {
        if (kTwoCta) tma_load_4d_2sm(dst, desc, bar, c0, c1, c2, c3);
        else tma_load_4d(dst, desc, bar, c0, c1, c2, c3);
      };
      WorkIter it(p, tile_first, tile_step, tiles_all_batches, total_kb);
      int rem, kb_begin, kb_end;
      while (it.next(rem, kb_begin, kb_end)) {
        const int bi = rem / tiles_per_batch;
        int tm = (rem % tiles_per_batch) / tiles_n;
        if (quad) tm = 2 * tm + (int)pair_id;
        const int tn = rem % tiles_n;
        const int m0 = tm * p.tile_m + (kTwoCta ? (int)rank * 128 : 0);
        const int n0 = tn * p.block_n + (kTwoCta ? (int)rank * bn_cta : 0);
        const int bi_in = bi % p.batch_inner, bi_out = bi / p.batch_inner;
        // conv: decompose the flattened pixel index of the tile origin
        int cimg = 0, cy0 = 0, cx0 = 0;
        if (p.a_mode == FFVC_OP_CONV3X3) {
          const int hw = p.conv_h * p.conv_w;
          cimg = m0 / hw;
          const int r = m0 % hw;
          cy0 = r / p.conv_w;
          cx0 = r % p.conv_w;
        }
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * kStageBytesT;
          const uint32_t sb = sa + b_off;
          const uint32_t fb = full_bar(stage);
          if (!kTwoCta) mbar_expect_tx(fb, a_bytes + b_bytes);
          else if (rank == 0) mbar_expect_tx(fb, 2u * (a_bytes + b_bytes));   // both CTAs' loads land on the leader's barrier
          const int seg = kb / p.kb_per_seg;
          const int kk = kb % p.kb_per_seg;
          const int k0 = kk * kBlockK;
          // ---- A
          if (p.a_mode == FFVC_OP_CONV3X3) {
            const int tap = kk / p.conv_cblocks;
            const int c0 = (kk % p.conv_cblocks) * kBlockK;
            tma4(sa, &tmap_a, fb, c0, cx0 + (tap % 3 - 1), cy0 + (tap / 3 - 1), cimg);
          } else {
            const int c2 = (p.a_role == FFVC_ROLE_OUT_BATCH) ? bi_in : 0;
            const int c3 = (p.a_role == FFVC_ROLE_OUT_BATCH) ? bi_out : (p.a_role == FFVC_ROLE_K_SEGMENT ? seg : 0);
            if (p.a_mode == FFVC_OP_KMAJOR) {
              tma4(sa, &tmap_a, fb, k0, m0, c2, c3);
            } else {
              for (int j = 0; j < rows_cta / 64; ++j) tma4(sa + j * 8192, &tmap_a, fb, m0 + 64 * j, k0, c2, c3);
            }
          }
          // ---- B
          {
            const int c2 = (p.b_role == FFVC_ROLE_OUT_BATCH) ? bi_in : 0;
            const int c3 = (p.b_role == FFVC_ROLE_OUT_BATCH) ? bi_out : (p.b_role == FFVC_ROLE_K_SEGMENT ? seg : 0);
            if (quad) {
              // this CTA's half of the B rows it shares with CTA (rank4 ^ 2): rows [n0 + pair_id * bn_cta / 2, + bn_cta / 2)
              const uint16_t mc = (uint16_t)((1u << rank) | (1u << (2u + rank)));
              const int hrows = bn_cta >> 1;
              if (p.b_mode == FFVC_OP_KMAJOR) {
                tma_load_4d_2sm_mc(sb + pair_id * (uint32_t)hrows * 128u, &tmap_b, fb, k0, n0 + (int)pair_id * hrows, c2, c3, mc);
              } else {
                const int nch = hrows / 64;                       // 64-row chunks in this CTA's share
                for (int j = 0; j < nch; ++j) {
                  const int jj = (int)pair_id * nch + j;
                  tma_load_4d_2sm_mc(sb + jj * 8192, &tmap_b, fb, n0 + 64 * jj, k0, c2, c3, mc);
                }
              }
            } else if (p.b_mode == FFVC_OP_KMAJOR) {
              tma4(sb, &tmap_b, fb, k0, n0, c2, c3);
            } else {
              for (int j = 0; j < bn_cta / 64; ++j) tma4(sb + j * 8192, &tmap_b, fb, n0 + 64 * j, k0, c2, c3);
            }
          }
          if (++stage == kStagesT) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only in 2-CTA mode)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16_m(kTwoCta ? 256 : 128, p.block_n, p.a_mode == FFVC_OP_MNMAJOR,
                                               p.b_mode == FFVC_OP_MNMAJOR);
      // K-major : 8-row groups 1024 B apart (SBO), one swizzle atom along K (LBO unused), K step = 32 B
      // MN-major: 64-wide MN chunks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO), K step = 2048 B
      const uint32_t a_lbo = (p.a_mode == FFVC_OP_MNMAJOR) ? 8192u : 16u;
      const uint32_t b_lbo = (p.b_mode == FFVC_OP_MNMAJOR) ? 8192u : 16u;
      const uint32_t a_kstep = (p.a_mode == FFVC_OP_MNMAJOR) ? 2048u : 32u;
      const uint32_t b_kstep = (p.b_mode == FFVC_OP_MNMAJOR) ? 2048u : 32u;
      const int subs = kTwoCta ? 1 : p.tile_m / 128;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      WorkIter it(p, tile_first, tile_step, tiles_all_batches, total_kb);
      int rem, kb_begin, kb_end;
      while (it.next(rem, kb_begin, kb_end)) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytesT;
          const uint32_t sb = sa + b_off;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t db = umma_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024u);
            const uint32_t accum = (kb > kb_begin || k > 0) ? 1u : 0u;
            if (kTwoCta) {
              umma_bf16_2sm(tmem_d, umma_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024u), db, idesc, accum);
            } else {
              for (int sub = 0; sub < subs; ++sub) {
                const uint64_t da = umma_smem_desc_sw128(sa + sub * 16384 + k * a_kstep, a_lbo, 1024u);
                umma_bf16(tmem_d + sub * 128, da, db, idesc, accum);
              }
            }
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if (kTwoCta) umma_commit_2sm(empty_bar(stage), quad ? (uint16_t)0xF : pair_mask); else umma_commit(empty_bar(stage));
          if (++stage == kStagesT) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // accumulator ready for the epilogue warps (of both CTAs)
        if (kTwoCta) umma_commit_2sm(tfull_bar(acc), pair_mask); else umma_commit(tfull_bar(acc));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int slice = (warp - 2) >> 2;      // 0 .. kEpiWarps/4-1: which row sub-tile / column part this warp covers
    constexpr int kSlices = kEpiWarps / 4;
    const int etid = threadIdx.x - 64;      // 0 .. 32*kEpiWarps-1 among the epilogue threads
    const bool two_sub = !kTwoCta && p.tile_m == 256;
    const int sub = two_sub ? (slice & 1) : 0;                       // row sub-tile (1-CTA 256-row tiles)
    const int cparts = two_sub ? kSlices / 2 : kSlices;              // column parts the tile is split into
    const int cpart = two_sub ? (slice >> 1) : slice;
    const int row_in_tile = (kTwoCta ? (int)rank * 128 : 0) + sub * 128 + q * 32 + lane;
    // columns this warp covers inside the tile: block_n / cparts when that is a multiple of 32, else the first
    // warps take 32-column chunks and the rest idle
    int c_begin, c_end;
    {
      int per = p.block_n / cparts;
      if (per < 32) per = 32;
      c_begin = cpart * per;
      c_end = min(p.block_n, c_begin + per);
      if (c_begin >= p.block_n) c_begin = c_end = 0;
    }
    const int tmem_col0 = sub * 128;
    float* sbias_all = reinterpret_cast<float*>(smem_raw + (bias_smem - smem_u32(smem_raw)));  // [2][256] floats
    int acc = 0;
    uint32_t acc_phase = 0;
    using E = Epi<kEpi>;
    const bool vec_ok = (p.ldc % 8 == 0) && (p.out_bs % 8 == 0) && (p.out_bs_inner % 8 == 0);
    // 32-byte accesses: every row / batch stride a multiple of 16 elements and 32-byte aligned base pointers (bf16 outputs)
    const bool vec32_ok = vec_ok && !E::f32(p) && (p.ldc % 16 == 0) && (p.out_bs % 16 == 0) && (p.out_bs_inner % 16 == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.pre_out) |
                            reinterpret_cast<uintptr_t>(p.aux) | reinterpret_cast<uintptr_t>(p.res)) & 31) == 0;
    const bool want_aux = E::mul(p) != FFVC_ACT_NONE, want_res = E::res(p);
    // TMA-store path: this warp's staging buffers ([32 rows][32 columns] bf16 = 2 KB each, 64B-swizzled: 16-byte chunk j of
    // row r lives at chunk j ^ ((r >> 1) & 3) — what the tensor map expects and conflict-free for one-row-per-lane writes)
    const bool tma_store = kCanTmaStore && p.tma_store;
    const uint32_t stg_pre = smem_base + 4u * 32u * 1024u + (uint32_t)(warp - 2) * 4096u;
    const uint32_t stg_out = stg_pre + 2048u;
    const uint32_t stg_row = (uint32_t)lane * 64u, stg_sw = (uint32_t)((lane >> 1) & 3);
    bool stg_busy = false;     // a bulk store of this warp may still be reading the staging buffers
    WorkIter it(p, tile_first, tile_step, tiles_all_batches, total_kb);
    int rem, kb_begin_unused, kb_end_unused;
    while (it.next(rem, kb_begin_unused, kb_end_unused)) {
      const int bi = rem / tiles_per_batch;
      int tm = (rem % tiles_per_batch) / tiles_n;
      if (quad) tm = 2 * tm + (int)pair_id;
      const int tn = rem % tiles_n;
      const int gm = tm * p.tile_m + row_in_tile;
      const int n0 = tn * p.block_n;
      const int bi_in = bi % p.batch_inner, bi_out = bi / p.batch_inner;
      const long long row_off = (long long)bi_out * p.out_bs + (long long)bi_in * p.out_bs_inner + (long long)gm * p.ldc;
      const bool row_ok = gm < p.M;
      float* sbias = sbias_all + acc * 256;
      // stage this tile's column bias in shared memory (double-buffered with the accumulator stage) and prefetch the first
      // chunk's aux / residual: all of it overlaps the wait for the MMA warp
      if (E::bias(p) == 1) {
        if (etid < p.block_n) sbias[etid] = (n0 + etid < p.N) ? p.bias[n0 + etid] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      }
      const float rbias = (E::bias(p) == 2 && row_ok) ? p.bias[gm] : 0.0f;
      constexpr int CW = (kEpiWarps == 16) ? 16 : 32;   // columns per register chunk
      uint4 pf_aux[CW / 8], pf_res[CW / 8];
      auto prefetch = [&](int c) {
        const int gn0 = n0 + c;
        if (row_ok && vec_ok && gn0 + CW <= p.N) {
          if (want_aux) {
            if (vec32_ok) {
#pragma unroll
              for (int j = 0; j < CW / 16; ++j) ld_global_256(p.aux + row_off + gn0 + 16 * j, pf_aux[2 * j], pf_aux[2 * j + 1]);
            } else {
              const uint4* ax = reinterpret_cast<const uint4*>(p.aux + row_off + gn0);
#pragma unroll
              for (int j = 0; j < CW / 8; ++j) pf_aux[j] = ax[j];
            }
          }
          if (want_res) {
            if (vec32_ok) {
#pragma unroll
              for (int j = 0; j < CW / 16; ++j) ld_global_256(p.res + row_off + gn0 + 16 * j, pf_res[2 * j], pf_res[2 * j + 1]);
            } else {
              const uint4* rs = reinterpret_cast<const uint4*>(p.res + row_off + gn0);
#pragma unroll
              for (int j = 0; j < CW / 8; ++j) pf_res[j] = rs[j];
            }
          }
        }
      };
      if ((want_aux || want_res) && c_begin < c_end) prefetch(c_begin);
      float am_best = 3.402823466e+38f;
      int am_idx = -1;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      for (int c = c_begin; c < c_end; c += CW) {
        uint32_t r[CW];
        const uint32_t taddr = tmem_base + (uint32_t)(acc * 256 + tmem_col0 + c) + ((uint32_t)(q * 32) << 16);
        if constexpr (CW == 16) tmem_ld_32x16(taddr, r); else tmem_ld_32x32(taddr, r);
        tmem_ld_wait();
        const int gn0 = n0 + c;
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
        if constexpr (kCanTmaStore) {
          // (a 32-column group that sticks out of N takes the direct-store path below: its partial chunk needs element guards)
          if (tma_store && n0 + c_begin + ((c - c_begin) & ~31) + 32 <= p.N) {
            // registers -> swizzled shared memory -> one TMA store per 32-column group (rows / columns beyond M / N are
            // clipped by the tensor map); no global store instruction is issued by the epilogue warps
            uint4 o_pk[CW / 8], p_pk[CW / 8];
            epilogue_chunk_pack<CW, kEpi>(p, r, rbias, sbias + c, cur_aux, cur_res, o_pk, p_pk);
            const int half = ((c - c_begin) >> 4) & 1;                 // which half of the 32-column group
            if (half == 0 && stg_busy) {
              if (lane == 0) bulk_wait_read0();
              __syncwarp();
              stg_busy = false;
            }
            const uint32_t ch0 = ((uint32_t)(2 * half) ^ stg_sw) * 16u, ch1 = ((uint32_t)(2 * half + 1) ^ stg_sw) * 16u;
            st_shared_v4(stg_out + stg_row + ch0, o_pk[0]);
            st_shared_v4(stg_out + stg_row + ch1, o_pk[1]);
            if (E::pre(p)) {
              st_shared_v4(stg_pre + stg_row + ch0, p_pk[0]);
              st_shared_v4(stg_pre + stg_row + ch1, p_pk[1]);
            }
            if (half == 1) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                const int col0 = n0 + c - 16, row0 = tm * p.tile_m + (int)rank * 128 + q * 32;
                tma_store_4d(&tmap_out, stg_out, col0, row0, bi_in, bi_out);
                if (E::pre(p)) tma_store_4d(&tmap_pre, stg_pre, col0, row0, bi_in, bi_out);
                bulk_commit_group();
              }
              stg_busy = true;
            }
            __syncwarp();
            continue;
          }
        }
        if (E::argmin(p)) {
          if (row_ok && gn0 < p.N) argmin_chunk<CW>(p, r, gn0, E::bias(p) == 1 ? sbias + c : nullptr, am_best, am_idx);
        } else if (row_ok && gn0 < p.N) {
          epilogue_chunk<CW, kEpi>(p, r, gn0, row_off + gn0, rbias, vec_ok && (gn0 + CW <= p.N), sbias + c, cur_aux, cur_res,
                                   vec32_ok && (gn0 + CW <= p.N));
        }
        __syncwarp();
      }
      if (E::argmin(p) && row_ok && am_idx >= 0) {
        const unsigned long long key = ((unsigned long long)float_order_bits(am_best) << 32) | (unsigned)am_idx;
        atomicMin(p.argmin + (long long)bi * p.M + gm, key);
      }
      // all TMEM reads of this accumulator by this warp are done -> hand it back to the MMA warp (of the leader CTA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kTwoCta) mbar_arrive_cluster(tempty_bar(acc), leader_cta); else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (kCanTmaStore && stg_busy && lane == 0) bulk_wait_all();   // the staging buffers must outlive the stores reading them
  }

  tc_fence_before();
  if (kTwoCta) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kTwoCta) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// dims/strides in elements (bf16); rank 4; dim0 contiguous.
static int make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                     const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return set_error(FFVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 1; i < rank; ++i) {
    gstr[i - 1] = strides_elems[i] * 2ull;
    if (gstr[i - 1] % 16 != 0) return set_error(FFVC_ERR_ARG, "gemm: operand stride not a multiple of 16 bytes");
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return set_error(FFVC_ERR_ARG, "gemm: operand not 16B aligned");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
             (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
             rank > 3 ? box[3] : 0);
    return set_error(FFVC_ERR_CUDA, buf);
  }
  return FFVC_OK;
}

static int g_num_sms = 0;
static int g_max_quads = -1;          // co-resident clusters of 4 CTAs (GPC sizes are not all multiples of 4): queried once
static bool g_attr_set = false;
static int g_stream_k_enabled = 0;    // ffvc_gemm_set_stream_k(1): measured no better than split-K on this workload (the wgrad
                                      // GEMMs are L2-operand-bandwidth bound, not tail bound), so it is opt-in
static int g_tma_store_enabled = 1;   // ffvc_gemm_set_tma_store(0) keeps the direct-store epilogue (A/B measurements, tests)


// ---- kernel instantiations: generic epilogue x {1-CTA, CTA pair} x {8, 16 epilogue warps}, plus the compile-time epilogues
// of the mapper / CLIP activation GEMMs on the 16-warp form
static constexpr int kNumEpiCodes = 5;
static constexpr int kEpiCodes[kNumEpiCodes] = {
    epi_code(FFVC_ACT_GELU, FFVC_ACT_NONE, 1, 1, 0),        // channel-mix Linear 1: + bias, GELU, pre-activation saved
    epi_code(FFVC_ACT_GELU, FFVC_ACT_NONE, 2, 1, 0),        // token-mix Conv1d 1: + row bias, GELU, pre-activation saved
    epi_code(FFVC_ACT_NONE, FFVC_ACT_GELU, 0, 0, 0),        // dgrad through GELU: x gelu'(pre-activation)
    epi_code(FFVC_ACT_QUICKGELU, FFVC_ACT_NONE, 1, 1, 0),   // CLIP c_fc: + bias, QuickGELU, pre-activation saved
    epi_code(FFVC_ACT_NONE, FFVC_ACT_QUICKGELU, 0, 0, 0),   // CLIP dgrad through QuickGELU
};

template <bool kTwoCta, int kEpiWarps, int kEpi>
static cudaError_t launch_one(const cudaLaunchConfig_t& cfg, const CUtensorMap* tm, const GemmDev& p) {
  if constexpr (kTwoCta && kEpiWarps == 16 && kEpi >= 0) {
    if (p.tma_store) return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<kTwoCta, kEpiWarps, kEpi, true>, tm[0], tm[1], tm[2], tm[3], p);
  }
  return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<kTwoCta, kEpiWarps, kEpi, false>, tm[0], tm[1], tm[2], tm[3], p);
}
template <bool kTwoCta>
static cudaError_t launch_wide(const cudaLaunchConfig_t& cfg, int epi, const CUtensorMap* tm, const GemmDev& p) {
  if (epi == kEpiCodes[0]) return launch_one<kTwoCta, 16, kEpiCodes[0]>(cfg, tm, p);
  if (epi == kEpiCodes[1]) return launch_one<kTwoCta, 16, kEpiCodes[1]>(cfg, tm, p);
  if (epi == kEpiCodes[2]) return launch_one<kTwoCta, 16, kEpiCodes[2]>(cfg, tm, p);
  if (epi == kEpiCodes[3]) return launch_one<kTwoCta, 16, kEpiCodes[3]>(cfg, tm, p);
  if (epi == kEpiCodes[4]) return launch_one<kTwoCta, 16, kEpiCodes[4]>(cfg, tm, p);
  return launch_one<kTwoCta, 16, -1>(cfg, tm, p);
}
static cudaError_t launch_gemm(const cudaLaunchConfig_t& cfg, bool two_cta, bool wide_epi, int epi, const CUtensorMap* tm,
                               const GemmDev& p) {
  if (two_cta && p.quad) {
    if (wide_epi) return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<true, 16, -1, false, true>, tm[0], tm[1], tm[2], tm[3], p);
    return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<true, 8, -1, false, true>, tm[0], tm[1], tm[2], tm[3], p);
  }
  if (wide_epi) return two_cta ? launch_wide<true>(cfg, epi, tm, p) : launch_wide<false>(cfg, epi, tm, p);
  return two_cta ? launch_one<true, 8, -1>(cfg, tm, p) : launch_one<false, 8, -1>(cfg, tm, p);
}
template <bool kTwoCta, int kEpiWarps, int kEpi>
static cudaError_t set_attr_one() {
  cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<kTwoCta, kEpiWarps, kEpi, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if constexpr (kTwoCta && kEpiWarps == 16 && kEpi >= 0) {
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tcgen05_kernel<kTwoCta, kEpiWarps, kEpi, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  }
  return e;
}
template <bool kTwoCta>
static cudaError_t set_attr_all() {
  cudaError_t e = set_attr_one<kTwoCta, 8, -1>();
  if (e == cudaSuccess) e = set_attr_one<kTwoCta, 16, -1>();
  if (e == cudaSuccess) e = set_attr_one<kTwoCta, 16, kEpiCodes[0]>();
  if (e == cudaSuccess) e = set_attr_one<kTwoCta, 16, kEpiCodes[1]>();
  if (e == cudaSuccess) e = set_attr_one<kTwoCta, 16, kEpiCodes[2]>();
  if (e == cudaSuccess) e = set_attr_one<kTwoCta, 16, kEpiCodes[3]>();
  if (e == cudaSuccess) e = set_attr_one<kTwoCta, 16, kEpiCodes[4]>();
  return e;
}
static int set_gemm_attrs() {
  cudaError_t e = set_attr_all<false>();
  if (e == cudaSuccess) e = set_attr_all<true>();
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gemm_tcgen05_kernel<true, 8, -1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gemm_tcgen05_kernel<true, 16, -1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return set_error(FFVC_ERR_CUDA, cudaGetErrorString(e));
  return FFVC_OK;
}

}  // namespace ffvc

using namespace ffvc;

/* co-resident clusters of 4 CTAs of the CTA-pair GEMM kernel (-1 until the first quad launch queried it) */
extern "C" int ffvc_gemm_max_quads(void) { return g_max_quads; }

extern "C" int ffvc_gemm_set_stream_k(int on) {
  g_stream_k_enabled = on ? 1 : 0;
  return FFVC_OK;
}

extern "C" int ffvc_gemm_set_tma_store(int on) {
  g_tma_store_enabled = on < 0 ? 0 : (on > 2 ? 2 : on);
  return FFVC_OK;
}

extern "C" int ffvc_gemm(const ffvc_gemm_params* g, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (!g || !g->a || !g->b || (!g->out && !g->argmin_out)) return set_error(FFVC_ERR_ARG, "gemm: null pointer");
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return set_error(FFVC_ERR_ARG, "gemm: non-positive dimension");
  const int batch = g->batch > 0 ? g->batch : 1;
  const int batch_inner = g->batch_inner > 0 ? g->batch_inner : 1;
  if (batch % batch_inner) return set_error(FFVC_ERR_ARG, "gemm: batch must be a multiple of batch_inner");
  const int batch_outer = batch / batch_inner;
  const int k_segs = g->k_segs > 0 ? g->k_segs : 1;
  int splits = g->splits > 0 ? g->splits : 1;
  if (splits > 1 && !(g->out_fp32 && g->atomic)) return set_error(FFVC_ERR_ARG, "gemm: split-K needs fp32 atomic output");
  if (g->b_mode == FFVC_OP_CONV3X3) return set_error(FFVC_ERR_ARG, "gemm: conv mode is for operand A only");

  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) return set_error(FFVC_ERR_CUDA, "gemm: no CUDA device");
  }
  if (!g_attr_set) {
    int rc_attr = set_gemm_attrs();
    if (rc_attr) return rc_attr;
    g_attr_set = true;
  }

  // ---- tile shape
  int block_n = g->block_n;
  int tile_m = g->tile_m;
  if (block_n <= 0) {
    if (g->N > 128) block_n = 256;
    else if (g->N > 64) block_n = 128;
    else if (g->N > 32) block_n = 64;
    else block_n = 32;
    if (tile_m <= 0) {
      // prefer more tiles when the problem is small: fill the SMs
      const long long tm = (g->M + 127) / 128;
      while (block_n > 64 && tm * ((g->N + block_n - 1) / block_n) * batch * splits < g_num_sms) block_n >>= 1;
    }
  }
  if (block_n != 32 && block_n != 64 && block_n != 128 && block_n != 256)
    return set_error(FFVC_ERR_ARG, "gemm: block_n must be 32/64/128/256");
  if (g->b_mode == FFVC_OP_MNMAJOR && block_n < 64) block_n = 64;
  if (tile_m <= 0) {
    // 256-row tiles share one B tile between two MMAs: worthwhile whenever N <= 128 and there is enough M to go round
    const long long t256 = (long long)((g->M + 255) / 256) * ((g->N + block_n - 1) / block_n) * batch * splits;
    tile_m = (block_n <= 128 && g->M >= 256 && t256 >= g_num_sms) ? 256 : 128;
  }
  // ---- CTA-pair (cta_group::2) variant: 256-row pair tiles, block_n 128 / 256; auto-selected for large problems
  int two_cta = g->two_cta;   // 0 auto, 1 force, -1 never
  if (two_cta == 0) {
    // pair tiles are 256 x 256; with N <= 128 the pair's MMAs are too short (64 cycles) for the single issuing thread to
    // hide its per-stage barrier round trip, so those shapes stay on the 1-CTA 256-row tile
    const long long pair_tiles = (long long)((g->M + 255) / 256) * ((g->N + 255) / 256) * batch * splits;
    const bool conv_ok = g->a_mode != FFVC_OP_CONV3X3 || (((long long)g->conv_h * g->conv_w) % 256 == 0);
    two_cta = (g->block_n <= 0 && g->tile_m <= 0 && g->M >= 256 && g->N > 128 && conv_ok && pair_tiles >= g_num_sms / 4) ? 1 : -1;
  }
  if (two_cta == 1) {
    if (g->block_n <= 0) block_n = g->N > 128 ? 256 : 128;
    if (block_n != 128 && block_n != 256) return set_error(FFVC_ERR_ARG, "gemm: the CTA-pair kernel needs block_n 128 or 256");
    tile_m = 256;
  }
  if (tile_m != 128 && tile_m != 256) return set_error(FFVC_ERR_ARG, "gemm: tile_m must be 128 or 256");
  if (two_cta != 1 && tile_m == 256 && block_n > 128) return set_error(FFVC_ERR_ARG, "gemm: tile_m 256 needs block_n <= 128");
  const int rows_cta = (two_cta == 1) ? 128 : tile_m;          // A rows staged per CTA
  const int bn_cta = (two_cta == 1) ? block_n / 2 : block_n;   // B rows staged per CTA
  // quad: clusters of 4 = two M-adjacent pair tiles sharing the B tile by TMA multicast (option "gemm_quad")
  const int tiles_m_host = (g->M + tile_m - 1) / tile_m;
  bool quad = two_cta == 1 && option(OPT_GEMM_QUAD) != 0 && tiles_m_host % 2 == 0 && bn_cta >= 128 && !g->argmin_out;
  if (quad && g_max_quads < 0) {
    // how many clusters of 4 fit the GPU at once (a persistent grid must not exceed it: the excess clusters would run as a second wave)
    cudaLaunchConfig_t qc;
    memset(&qc, 0, sizeof(qc));
    cudaLaunchAttribute qa[1];
    qc.gridDim = dim3((unsigned)(g_num_sms / 4 * 4));
    qc.blockDim = dim3(64 + 32 * 16);
    qc.dynamicSmemBytes = kSmemBytes;
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 4;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    qc.attrs = qa;
    qc.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_kernel<true, 16, -1, false, true>, &qc) != cudaSuccess) n = 0;
    cudaGetLastError();
    g_max_quads = n;
  }
  if (quad && g_max_quads < 8) quad = false;       // no room for clusters of 4: stay on pairs

  GemmDev p;
  memset(&p, 0, sizeof(p));
  p.M = g->M;
  p.N = g->N;
  p.batch = batch;
  p.batch_inner = batch_inner;
  p.tile_m = tile_m;
  p.block_n = block_n;
  p.a_mode = g->a_mode;
  p.b_mode = g->b_mode;
  p.k_segs = k_segs;
  p.splits = splits;
  p.a_role = g->a_batch_role;
  p.b_role = g->b_batch_role;

  CUtensorMap ta, tb;
  int rc;
  // ---- operand A
  if (g->a_mode == FFVC_OP_CONV3X3) {
    const int H = g->conv_h, W = g->conv_w, C = g->conv_c;
    if (H <= 0 || W <= 0 || C <= 0 || C % 64 != 0) return set_error(FFVC_ERR_ARG, "conv: Cin must be a multiple of 64");
    if (tile_m == 256 && ((long long)H * W) % 256 != 0) {
      if (two_cta == 1) return set_error(FFVC_ERR_ARG, "conv: the CTA-pair kernel needs H*W % 256 == 0");
      tile_m = p.tile_m = 128;
    }
    const int ptile = (two_cta == 1) ? 128 : tile_m;           // pixels staged per CTA
    const int tile_w = W < ptile ? W : ptile;
    if (ptile % tile_w != 0 || W % tile_w != 0) return set_error(FFVC_ERR_ARG, "conv: W must divide or be a multiple of the pixel tile");
    const int tile_h = ptile / tile_w;
    if (H % tile_h != 0) return set_error(FFVC_ERR_ARG, "conv: H*W must tile by the pixel tile");
    const long long npix = (long long)g->conv_n * H * W;
    if (npix != g->M) return set_error(FFVC_ERR_ARG, "conv: M must equal N_img*H*W");
    if (g->K != 9 * C) return set_error(FFVC_ERR_ARG, "conv: K must equal 9*Cin");
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)g->conv_n};
    uint64_t str[4] = {1, (uint64_t)C, (uint64_t)W * C, (uint64_t)H * W * C};
    uint32_t box[4] = {64, (uint32_t)tile_w, (uint32_t)tile_h, 1};
    if ((rc = make_tmap(&ta, g->a, 4, dims, str, box)) != FFVC_OK) return rc;
    p.conv_h = H;
    p.conv_w = W;
    p.conv_cblocks = C / 64;
    p.kb_per_seg = 9 * (C / 64);
  } else {
    p.kb_per_seg = (g->K + kBlockK - 1) / kBlockK;
    uint64_t n_in = 1, n_out = 1, s_in = (uint64_t)g->a_ld * 8, s_out = (uint64_t)g->a_ld * 8;  // any valid stride for size-1 dims
    if (g->a_batch_role == FFVC_ROLE_OUT_BATCH) {
      n_in = batch_inner;
      n_out = batch_outer;
      if (n_in > 1) s_in = (uint64_t)g->a_batch_stride_inner;
      if (n_out > 1) s_out = (uint64_t)g->a_batch_stride;
    } else if (g->a_batch_role == FFVC_ROLE_K_SEGMENT) {
      n_out = k_segs;
      if (n_out > 1) s_out = (uint64_t)g->a_batch_stride;
    }
    if (g->a_mode == FFVC_OP_KMAJOR) {
      uint64_t dims[4] = {(uint64_t)g->K, (uint64_t)g->M, n_in, n_out};
      uint64_t str[4] = {1, (uint64_t)g->a_ld, s_in, s_out};
      uint32_t box[4] = {64, (uint32_t)rows_cta, 1, 1};
      if ((rc = make_tmap(&ta, g->a, 4, dims, str, box)) != FFVC_OK) return rc;
    } else {
      uint64_t dims[4] = {(uint64_t)g->M, (uint64_t)g->K, n_in, n_out};
      uint64_t str[4] = {1, (uint64_t)g->a_ld, s_in, s_out};
      uint32_t box[4] = {64, 64, 1, 1};
      if ((rc = make_tmap(&ta, g->a, 4, dims, str, box)) != FFVC_OK) return rc;
    }
  }
  // ---- operand B
  {
    uint64_t n_in = 1, n_out = 1, s_in = (uint64_t)g->b_ld * 8, s_out = (uint64_t)g->b_ld * 8;
    if (g->b_batch_role == FFVC_ROLE_OUT_BATCH) {
      n_in = batch_inner;
      n_out = batch_outer;
      if (n_in > 1) s_in = (uint64_t)g->b_batch_stride_inner;
      if (n_out > 1) s_out = (uint64_t)g->b_batch_stride;
    } else if (g->b_batch_role == FFVC_ROLE_K_SEGMENT) {
      n_out = k_segs;
      if (n_out > 1) s_out = (uint64_t)g->b_batch_stride;
    }
    if (g->b_mode == FFVC_OP_KMAJOR) {
      uint64_t dims[4] = {(uint64_t)g->K, (uint64_t)g->N, n_in, n_out};
      uint64_t str[4] = {1, (uint64_t)g->b_ld, s_in, s_out};
      uint32_t box[4] = {64, (uint32_t)(quad ? bn_cta / 2 : bn_cta), 1, 1};      // quad: every CTA fetches half of its rows
      if ((rc = make_tmap(&tb, g->b, 4, dims, str, box)) != FFVC_OK) return rc;
    } else {
      uint64_t dims[4] = {(uint64_t)g->N, (uint64_t)g->K, n_in, n_out};
      uint64_t str[4] = {1, (uint64_t)g->b_ld, s_in, s_out};
      uint32_t box[4] = {64, 64, 1, 1};
      if ((rc = make_tmap(&tb, g->b, 4, dims, str, box)) != FFVC_OK) return rc;
    }
  }
  if (splits > p.kb_per_seg * k_segs) splits = p.splits = p.kb_per_seg * k_segs;
  // stream-K for the fp32-atomic (wgrad) GEMMs whose tile count does not fill the machine evenly: replaces split-K
  {
    const long long tiles0 = (long long)((g->M + tile_m - 1) / tile_m) * ((g->N + block_n - 1) / block_n) * batch;
    const long long workers = (two_cta == 1) ? sm_budget(g_num_sms) / 2 : sm_budget(g_num_sms);
    const long long total_kb_h = (long long)p.kb_per_seg * k_segs;
    const long long waves_x100 = tiles0 * splits * 100 / workers;            // work items per worker, in percent
    const bool uneven = (waves_x100 % 100) != 0 && waves_x100 < 800;         // a fractional last wave that matters
    p.quad = quad ? 1 : 0;
    p.stream_k = (!quad && g_stream_k_enabled && g->out_fp32 && g->atomic && !g->argmin_out && !g->bias && !g->res && !g->aux && !g->pre_out &&
                  g->act == FFVC_ACT_NONE && uneven && tiles0 * total_kb_h >= 4 * workers) ? 1 : 0;
    if (p.stream_k) splits = p.splits = 1;
  }

  p.out = g->out;
  p.pre_out = g->pre_out;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(g->aux);
  p.res = reinterpret_cast<const __nv_bfloat16*>(g->res);
  p.bias = g->bias;
  p.ldc = g->ldc;
  p.out_bs = g->out_batch_stride;
  p.out_bs_inner = g->out_batch_stride_inner;
  p.out_fp32 = g->out_fp32;
  p.atomic = g->atomic;
  p.bias_mode = g->bias ? g->bias_mode : 0;
  p.act = g->act;
  p.mul_mode = g->aux ? g->mul_mode : 0;
  p.alpha = g->alpha == 0.0f ? 1.0f : g->alpha;
  p.argmin = reinterpret_cast<unsigned long long*>(g->argmin_out);
  if (p.argmin && (p.bias_mode == 2 || splits > 1)) return set_error(FFVC_ERR_ARG, "gemm: the arg-min epilogue takes a per-column bias only and no split-K");

  const long long tiles = (long long)((g->M + tile_m - 1) / tile_m) * ((g->N + block_n - 1) / block_n) * batch * splits;
  cudaError_t e;
  // 16 epilogue warps when the epilogue does real math (activation / activation gradient): with 8 warps (2 per scheduler)
  // those epilogues run at IPC ~0.4 and pace the whole kernel for short-K GEMMs
  const bool wide_epi = (p.act != FFVC_ACT_NONE || p.mul_mode != FFVC_ACT_NONE) && block_n >= 128 && g->epi_warps != 8;
  const int nthreads = wide_epi ? 64 + 32 * 16 : kNumThreads;
  // compile-time epilogue for the hot activation epilogues (see Epi<> in gemm_common.cuh); everything else stays generic
  int epi = -1;
  if (wide_epi && !p.out_fp32 && !p.atomic && p.alpha == 1.0f && !p.argmin && g->epi_warps != 16) {
    const int code = epi_code(p.act, p.mul_mode, p.bias_mode, p.pre_out != nullptr, p.res != nullptr);
    for (int i = 0; i < kNumEpiCodes; ++i)
      if (kEpiCodes[i] == code) epi = code;
  }
  if (quad) epi = -1;        // the cluster-of-4 variant is instantiated with the run-time epilogue only
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(nthreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  if (two_cta == 1 && quad) {
    // one cluster of 4 (two pairs) per two M-adjacent tiles, persistent over min(items, SMs/4) clusters
    const int sms = sm_budget(g_num_sms);
    const long long items = tiles / 2;
    int cap = sms / 4 < g_max_quads ? sms / 4 : g_max_quads;
    if (cap < 1) cap = 1;
    const long long quads = items >= cap ? cap : items;
    cfg.gridDim = dim3((unsigned)(4 * quads));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else if (two_cta == 1) {
    // one CTA pair (cluster of 2) per tile, persistent over min(tiles, SMs/2) pairs
    const int sms = sm_budget(g_num_sms);     // tile shapes above are chosen for the whole GPU (same arithmetic); only the grid shrinks
    const long long pairs = (p.stream_k || tiles >= sms / 2) ? sms / 2 : tiles;
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    const int sms = sm_budget(g_num_sms);
    cfg.gridDim = dim3((unsigned)((p.stream_k || tiles >= sms) ? sms : tiles));
  }
  // TMA-store epilogue: compile-time epilogue on the pair kernel, bf16 output(s) whose rows / batches start on 16-byte boundaries
  CUtensorMap tms[4];
  tms[0] = ta;
  tms[1] = tb;
  tms[2] = ta;   // placeholders when the direct-store epilogue runs
  tms[3] = ta;
  p.tma_store = 0;
  const int total_kb_host = p.kb_per_seg * k_segs;
  if (epi >= 0 && two_cta == 1 && g_tma_store_enabled && (total_kb_host <= 8 || g_tma_store_enabled == 2) && p.ldc % 8 == 0 && p.out_bs % 8 == 0 && p.out_bs_inner % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.pre_out) & 15) == 0) {
    uint64_t dims[4] = {(uint64_t)g->N, (uint64_t)g->M, (uint64_t)batch_inner, (uint64_t)batch_outer};
    uint64_t str[4] = {1, (uint64_t)p.ldc, batch_inner > 1 ? (uint64_t)p.out_bs_inner : (uint64_t)p.ldc * 8,
                       batch_outer > 1 ? (uint64_t)p.out_bs : (uint64_t)p.ldc * 8};
    uint32_t box[4] = {32, 32, 1, 1};
    if ((rc = make_tmap(&tms[2], p.out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) != FFVC_OK) return rc;
    if (p.pre_out && (rc = make_tmap(&tms[3], p.pre_out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) != FFVC_OK) return rc;
    p.tma_store = 1;
  }
  e = launch_gemm(cfg, two_cta == 1, wide_epi, epi, tms, p);
  if (e != cudaSuccess) return set_error(FFVC_ERR_CUDA, cudaGetErrorString(e));
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFVC_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return FFVC_OK;
}
