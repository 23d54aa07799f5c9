// LayerNorm (mapper + CLIP ViT) and GroupNorm(32)+swish (VQGAN decoder) — HBM-bound, vectorised,
// warp-shuffle reductions.  bf16 activations, fp32 statistics.
//   LayerNorm : mlp_mixer_pytorch.py:11,37 (nn.LayerNorm), cloob.py:170-176,244,250 (fp32 LayerNorm)
//   GroupNorm : taming Normalize = GroupNorm(32, C, eps=1e-6, affine) followed by x*sigmoid(x)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "ffvc_internal.h"
#include "ptx.cuh"

namespace ffvc {

__device__ __forceinline__ void unpack8(const uint4& pk, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 pk;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 h1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]);
  __nv_bfloat162 h3 = __floats2bfloat162_rn(f[6], f[7]);
  pk.x = *reinterpret_cast<uint32_t*>(&h0);
  pk.y = *reinterpret_cast<uint32_t*>(&h1);
  pk.z = *reinterpret_cast<uint32_t*>(&h2);
  pk.w = *reinterpret_cast<uint32_t*>(&h3);
  return pk;
}

// ------------------------------------------------------------------------------------ LayerNorm forward
// One warp per row, rows dealt round-robin to the resident warps; D % 8 == 0, D <= 32*8*kMaxV.  gamma / beta are staged in
// shared memory once per CTA: re-reading them from global for every row made the kernel L1-bound (ncu: 24 sectors per
// load request, 8 KB of parameter traffic per 2 KB row, 66 % L1 throughput at 1.5 TB/s of useful traffic).
template <int kMaxV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
                                                            float* __restrict__ rstd, long long rows, int D, float eps) {
  extern __shared__ float sgb[];   // [gamma D][beta D]
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    sgb[i] = gamma[i];
    sgb[D + i] = beta[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nvec = D >> 3;
  const float4* sg4 = reinterpret_cast<const float4*>(sgb);
  const float4* sb4 = reinterpret_cast<const float4*>(sgb + D);
  for (long long row = (long long)blockIdx.x * wpb + warp; row < rows; row += (long long)gridDim.x * wpb) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
    float v[kMaxV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        unpack8(xr[c], v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
    s = warp_sum(s);
    const float mu = s / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mu;
          q += d * d;
        }
      }
    }
    q = warp_sum(q);
    const float rs = rsqrtf(q / D + eps);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
    uint4* yr = reinterpret_cast<uint4*>(y + row * D);
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float o[8];
        const float4 g0 = sg4[2 * c], g1 = sg4[2 * c + 1];
        const float4 b0 = sb4[2 * c], b1 = sb4[2 * c + 1];
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mu) * rs * gg[j] + bb[j];
        yr[c] = pack8(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ LayerNorm backward
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  optional dx += add (residual path);
// optional dgamma += sum_rows dy * xhat, dbeta += sum_rows dy (fp32 atomics, one per column per CTA).
template <int kMaxV, bool kWgrad>
__global__ void __launch_bounds__(256, kWgrad ? 2 : 3) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                               const __nv_bfloat16* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               const __nv_bfloat16* __restrict__ add,
                                                               __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, long long rows, int D) {
  extern __shared__ float sm_all[];  // [gamma D] then, when dgamma != nullptr, [2][D] accumulators
  float* sgam = sm_all;
  float* sm = sm_all + D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) sgam[i] = gamma[i];   // re-reading gamma from global per row was L1-bound
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int nvec = D >> 3;
  constexpr bool wgrad = kWgrad;
  float accg[wgrad ? kMaxV : 1][8], accb[wgrad ? kMaxV : 1][8];
  if (wgrad) {
#pragma unroll
    for (int i = 0; i < kMaxV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) accg[i][j] = accb[i][j] = 0.f;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
  }
  for (long long row = (long long)blockIdx.x * nwarps + warp; row < rows; row += (long long)gridDim.x * nwarps) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * D);
    const uint4* dr = reinterpret_cast<const uint4*>(dy + row * D);
    const float mu = mean[row], rs = rstd[row];
    // the row stays packed (bf16) in registers between the two passes: 8 registers per 16 values instead of 16;
    // the residual-path gradient is fetched with the same batch of loads (it is only needed after the reductions)
    // (only in the no-wgrad form: the wgrad accumulators leave no registers for it)
    uint4 xp[kMaxV], dp[kMaxV], ap[kWgrad ? 1 : kMaxV];
    const uint4* ar = add ? reinterpret_cast<const uint4*>(add + row * D) : nullptr;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        xp[i] = xr[c];
        dp[i] = dr[c];
        if (!kWgrad && ar) ap[kWgrad ? 0 : i] = ar[c];
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float xv[8], dv[8];
        unpack8(xp[i], xv);
        unpack8(dp[i], dv);
        const float4 g0 = reinterpret_cast<const float4*>(sgam)[2 * c], g1 = reinterpret_cast<const float4*>(sgam)[2 * c + 1];
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mu) * rs;
          const float g = dv[j] * gg[j];
          s1 += g;
          s2 += g * xh;
          if (wgrad) {
            accg[wgrad ? i : 0][j] += dv[j] * xh;
            accb[wgrad ? i : 0][j] += dv[j];
          }
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    uint4* ox = reinterpret_cast<uint4*>(dx + row * D);
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float xv[8], dv[8], o[8];
        unpack8(xp[i], xv);
        unpack8(dp[i], dv);
        const float4 g0 = reinterpret_cast<const float4*>(sgam)[2 * c], g1 = reinterpret_cast<const float4*>(sgam)[2 * c + 1];
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mu) * rs;
          o[j] = rs * (dv[j] * gg[j] - s1 - xh * s2);
        }
        if (ar) {
          float av[8];
          unpack8(kWgrad ? ar[c] : ap[kWgrad ? 0 : i], av);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += av[j];
        }
        ox[c] = pack8(o);
      }
    }
  }
  if (wgrad) {
#pragma unroll
    for (int i = 0; i < (wgrad ? kMaxV : 1); ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          atomicAdd(&sm[c * 8 + j], accg[i][j]);
          atomicAdd(&sm[D + c * 8 + j], accb[i][j]);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      atomicAdd(&dgamma[i], sm[i]);
      atomicAdd(&dbeta[i], sm[D + i]);
    }
  }
}

// ------------------------------------------------------------------------------------ LayerNorm, column-owning pipelined form
// A CTA has exactly D/8 threads; thread c owns columns [8c, 8c+8) of every row the CTA visits.  gamma / beta and the
// per-column accumulators (dgamma, dbeta, column sums of dx) are then 8 registers each whatever D is (the warp-per-row
// form above needs D/32 of them per lane: 64 accumulator registers at D = 1024, two CTAs per SM).  Rows travel through a
// cp.async ring in shared memory: each thread copies ITS OWN 16-byte column slice of kRows rows per stage and is the only
// reader of it, so a stage needs cp.async.wait_group but no barrier, the loads of the next kStages-1 row groups are in
// flight while the current one is reduced, and nothing in flight occupies registers.  (A first version kept the rows in
// registers: under the 128-register cap ptxas serialised the loads between the shuffles — 79 us against 60 us for the
// warp-per-row kernel + separate sum kernels, profiles/r01_ab_kernels.md.)  Row statistics: warp shuffles, then one
// shared-memory exchange between the CTA's warps per row group (double-buffered: ONE __syncthreads per group).
// The backward kernel can also emit, from the values it stores, the two bias gradients the mixer takes from dx
// (mlp_mixer_pytorch.py:16-23): column sums (Linear bias) and per-token row sums (Conv1d bias) — otherwise two more passes.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? 16 : 0;                  // src-size 0: nothing is read, the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int kThreads, int kRows, int kStages>
__global__ void __launch_bounds__(kThreads) layernorm_fwd_pipe_kernel(const __nv_bfloat16* __restrict__ x,
                                                                      const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta,
                                                                      __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
                                                                      float* __restrict__ rstd, long long rows, float eps) {
  constexpr int D = kThreads * 8;
  constexpr int NW = kThreads / 32;
  extern __shared__ __align__(16) unsigned char ln_smem[];
  uint4* stage = reinterpret_cast<uint4*>(ln_smem);          // [kStages][kRows][kThreads]
  __shared__ float red[2][2][NW][kRows];                      // [group parity][sum | centred sumsq][warp][row]
  const int c = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long stride = (long long)gridDim.x * kRows;
  long long rnext = (long long)blockIdx.x * kRows;
  auto issue = [&](int s) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const long long row = rnext + r;
      const bool ok = row < rows;
      cp_async16(smem_u32(&stage[(s * kRows + r) * kThreads + c]), ok ? (const void*)(x + row * D + 8 * c) : (const void*)x, ok);
    }
    cp_async_commit();
    rnext += stride;
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) issue(s);
  const float4 g0 = reinterpret_cast<const float4*>(gamma)[2 * c], g1 = reinterpret_cast<const float4*>(gamma)[2 * c + 1];
  const float4 b0 = reinterpret_cast<const float4*>(beta)[2 * c], b1 = reinterpret_cast<const float4*>(beta)[2 * c + 1];
  const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bet[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  int it = 0, cur = 0;
  for (long long r0 = (long long)blockIdx.x * kRows; r0 < rows; r0 += stride, it ^= 1) {
    issue(cur == 0 ? kStages - 1 : cur - 1);                  // refill the stage consumed by the previous group
    cp_async_wait<kStages - 1>();                             // this thread's slice of the current stage has landed
    float v[kRows][8], mu[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      unpack8(stage[(cur * kRows + r) * kThreads + c], v[r]);
      float sacc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sacc += v[r][j];
      sacc = warp_sum(sacc);
      if (lane == 0) red[it][0][warp][r] = sacc;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += red[it][0][w][r];
      mu[r] = t * (1.0f / D);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[r][j] - mu[r];
        q += d * d;
      }
      q = warp_sum(q);
      if (lane == 0) red[it][1][warp][r] = q;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += red[it][1][w][r];
      const float rs = rsqrtf(t * (1.0f / D) + eps);
      if (r0 + r < rows) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[r][j] - mu[r]) * rs * gam[j] + bet[j];
        reinterpret_cast<uint4*>(y + (r0 + r) * D)[c] = pack8(o);
        if (c == 0) {
          if (mean) mean[r0 + r] = mu[r];
          if (rstd) rstd[r0 + r] = rs;
        }
      }
    }
    cur = (cur + 1 == kStages) ? 0 : cur + 1;
  }
  cp_async_wait<0>();
}

template <int kThreads, int kRows, int kStages, bool kWgrad>
__global__ void __launch_bounds__(kThreads) layernorm_bwd_pipe_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ add,
    __nv_bfloat16* __restrict__ dx, float* __restrict__ ws, bool want_colsum, bool want_rowsum, int rowsum_T, long long rows) {
  constexpr int D = kThreads * 8;
  constexpr int NW = kThreads / 32;
  extern __shared__ __align__(16) unsigned char ln_smem[];
  uint4* stage = reinterpret_cast<uint4*>(ln_smem);                                   // [kStages][kRows][3: x, dy, add][kThreads]
  float* srow = reinterpret_cast<float*>(ln_smem + (size_t)kStages * kRows * 3 * kThreads * 16);   // [rowsum_T]
  __shared__ float red[2][NW][2 * kRows];
  const int c = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long stride = (long long)gridDim.x * kRows;
  long long rnext = (long long)blockIdx.x * kRows;
  auto issue = [&](int s) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const long long row = rnext + r;
      const bool ok = row < rows;
      uint4* slot = &stage[((s * kRows + r) * 3) * kThreads + c];
      cp_async16(smem_u32(slot), ok ? (const void*)(x + row * D + 8 * c) : (const void*)x, ok);
      cp_async16(smem_u32(slot + kThreads), ok ? (const void*)(dy + row * D + 8 * c) : (const void*)dy, ok);
      if (add) cp_async16(smem_u32(slot + 2 * kThreads), ok ? (const void*)(add + row * D + 8 * c) : (const void*)add, ok);
    }
    cp_async_commit();
    rnext += stride;
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) issue(s);
  if (want_rowsum) {
    for (int i = threadIdx.x; i < rowsum_T; i += kThreads) srow[i] = 0.f;
    __syncthreads();
  }
  const float4 g0 = reinterpret_cast<const float4*>(gamma)[2 * c], g1 = reinterpret_cast<const float4*>(gamma)[2 * c + 1];
  const float gam[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  float ag[kWgrad ? 8 : 1], ab[kWgrad ? 8 : 1], ac[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ac[j] = 0.f;
    if (kWgrad) ag[kWgrad ? j : 0] = ab[kWgrad ? j : 0] = 0.f;
  }
  // per-row statistics of the NEXT row group are fetched one group ahead (registers), like the ring hides the row loads
  float mu_n[kRows], rs_n[kRows];
  auto fetch_stats = [&](long long r0) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const bool ok = r0 + r < rows;
      mu_n[r] = ok ? mean[r0 + r] : 0.f;
      rs_n[r] = ok ? rstd[r0 + r] : 0.f;      // rstd = 0 (and zero-filled data) makes a row past the end contribute nothing
    }
  };
  fetch_stats((long long)blockIdx.x * kRows);
  int it = 0, cur = 0;
  for (long long r0 = (long long)blockIdx.x * kRows; r0 < rows; r0 += stride, it ^= 1) {
    issue(cur == 0 ? kStages - 1 : cur - 1);                  // refill the stage consumed by the previous group
    float mu[kRows], rs[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      mu[r] = mu_n[r];
      rs[r] = rs_n[r];
    }
    fetch_stats(r0 + stride);
    cp_async_wait<kStages - 1>();                             // this thread's slices of the current stage have landed
    const uint4* st = &stage[(cur * kRows * 3) * kThreads + c];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float xv[8], dv[8];
      unpack8(st[(r * 3) * kThreads], xv);
      unpack8(st[(r * 3 + 1) * kThreads], dv);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mu[r]) * rs[r];
        const float g = dv[j] * gam[j];
        s1 += g;
        s2 += g * xh;
        if (kWgrad) {
          ag[kWgrad ? j : 0] += dv[j] * xh;
          ab[kWgrad ? j : 0] += dv[j];
        }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        red[it][warp][2 * r] = s1;
        red[it][warp][2 * r + 1] = s2;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        s1 += red[it][w][2 * r];
        s2 += red[it][w][2 * r + 1];
      }
      s1 *= (1.0f / D);
      s2 *= (1.0f / D);
      float xv[8], dv[8], o[8];
      unpack8(st[(r * 3) * kThreads], xv);
      unpack8(st[(r * 3 + 1) * kThreads], dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mu[r]) * rs[r];
        o[j] = rs[r] * (dv[j] * gam[j] - s1 - xh * s2);
      }
      if (add) {
        float av[8];
        unpack8(st[(r * 3 + 2) * kThreads], av);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += av[j];
      }
      const uint4 pk = pack8(o);
      const bool ok = r0 + r < rows;
      if (ok) reinterpret_cast<uint4*>(dx + (r0 + r) * D)[c] = pk;
      if (want_colsum || want_rowsum) {         // sums of the values as stored (bf16-rounded), like a pass over dx would see
        float q[8];
        unpack8(pk, q);
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ac[j] += q[j];
          t += q[j];
        }
        if (want_rowsum) {
          t = warp_sum(t);
          if (lane == 0 && ok) atomicAdd(&srow[(int)((r0 + r) % rowsum_T)], t);
        }
      }
    }
    cur = (cur + 1 == kStages) ? 0 : cur + 1;
  }
  cp_async_wait<0>();
  // per-CTA partial sums go to this CTA's row of the workspace [gridDim.x][3*D + rowsum_T]; ln_bwd_finalize_kernel folds the
  // rows.  (Direct atomicAdd from every CTA serialises gridDim.x deep on each of the 3*D addresses when the CTAs finish
  // together: measured ~45 us of a 79 us kernel at 592 CTAs, profiles/r01_ab_kernels.md.)
  float* my = ws + (size_t)blockIdx.x * (3 * D + rowsum_T);
  if (kWgrad) {
    reinterpret_cast<float4*>(my)[2 * c] = make_float4(ag[0], ag[kWgrad ? 1 : 0], ag[kWgrad ? 2 : 0], ag[kWgrad ? 3 : 0]);
    reinterpret_cast<float4*>(my)[2 * c + 1] = make_float4(ag[kWgrad ? 4 : 0], ag[kWgrad ? 5 : 0], ag[kWgrad ? 6 : 0], ag[kWgrad ? 7 : 0]);
    reinterpret_cast<float4*>(my + D)[2 * c] = make_float4(ab[0], ab[kWgrad ? 1 : 0], ab[kWgrad ? 2 : 0], ab[kWgrad ? 3 : 0]);
    reinterpret_cast<float4*>(my + D)[2 * c + 1] = make_float4(ab[kWgrad ? 4 : 0], ab[kWgrad ? 5 : 0], ab[kWgrad ? 6 : 0], ab[kWgrad ? 7 : 0]);
  }
  if (want_colsum) {
    reinterpret_cast<float4*>(my + 2 * D)[2 * c] = make_float4(ac[0], ac[1], ac[2], ac[3]);
    reinterpret_cast<float4*>(my + 2 * D)[2 * c + 1] = make_float4(ac[4], ac[5], ac[6], ac[7]);
  }
  if (want_rowsum) {
    __syncthreads();
    for (int i = threadIdx.x; i < rowsum_T; i += kThreads) my[3 * D + i] = srow[i];
  }
}

// out[i] += sum_p ws[p][i] for the sections that were requested: [0,D) dgamma, [D,2D) dbeta, [2D,3D) colsum, [3D,3D+T) rowsum.
// gridDim.y chunks of parts per column block: the atomic depth per address is gridDim.y (32), not the number of CTAs.
__global__ void __launch_bounds__(256) ln_bwd_finalize_kernel(const float* __restrict__ ws, int nparts, int D, int T,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              float* __restrict__ colsum_out, float* __restrict__ rowsum_out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int width = 3 * D + T;
  if (i >= width) return;
  float* out;
  if (i < D) out = dgamma ? dgamma + i : nullptr;
  else if (i < 2 * D) out = dbeta ? dbeta + (i - D) : nullptr;
  else if (i < 3 * D) out = colsum_out ? colsum_out + (i - 2 * D) : nullptr;
  else out = rowsum_out ? rowsum_out + (i - 3 * D) : nullptr;
  if (!out) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int p = blockIdx.y;
  const int step = gridDim.y;
  for (; p + 3 * step < nparts; p += 4 * step) {
    a0 += ws[(size_t)p * width + i];
    a1 += ws[(size_t)(p + step) * width + i];
    a2 += ws[(size_t)(p + 2 * step) * width + i];
    a3 += ws[(size_t)(p + 3 * step) * width + i];
  }
  for (; p < nparts; p += step) a0 += ws[(size_t)p * width + i];
  atomicAdd(out, (a0 + a1) + (a2 + a3));
}

// ------------------------------------------------------------------------------------ GroupNorm
// REPRODUCIBLE statistics (round 2): every reduction below runs in a FIXED order that depends on the sample's shape only —
// warp shuffles, one shared-memory slot per warp, a fixed-order fold over the warps, one workspace slot per CTA
// (ws partials [n][part][2][G]) and a fixed-order fold over the parts (gn_fold_parts) — so (mean, rstd) and the backward
// sums are bit-identical from run to run and for any batch size or sharding of the batch.  Round 1 used float shared-memory
// atomics inside the CTA and double atomics across CTAs: the last bit of mean / rstd flickered between identical runs, a
// bf16 rounding flipped here and there, and 60 layers later the image differed by 0.85 % — which the reference's
// discontinuous gradient (max-pool arg-max, HSV sectors) turned into a 16 % gradient difference (profiles/r02_parity_fullsize.md).
//
// Per-thread 8-channel partial sums (s, q) -> per-group sums of this WARP in sm[warp][0..G) / sm[warp][G..2G) (plain stores,
// every (warp, group) slot has exactly one writer; slots of groups the warp does not cover must be zero beforehand).
__device__ __forceinline__ void gn_fold_to_slots(float (&s)[8], float (&q)[8], float* sm, int G, int cpg, int vc,
                                                 int vec_per_pix) {
  const int lane = threadIdx.x & 31;
  float* slot = sm + (threadIdx.x >> 5) * 2 * G;
  if (vec_per_pix < 32) {      // lanes of the warp that own the same channel vector: butterfly (same tree every time)
    for (int o = vec_per_pix; o < 32; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
      }
    }
  }
  if (cpg >= 8) {
    float ss = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    float qq = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
    const int vpg = cpg >> 3;                     // channel vectors per group (power of two; neighbouring lanes)
    for (int o = 1; o < vpg; o <<= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      qq += __shfl_xor_sync(0xffffffffu, qq, o);
    }
    if ((vec_per_pix >= 32 || lane < vec_per_pix) && (vc & (vpg - 1)) == 0) {
      const int g = (vc * 8) / cpg;
      slot[g] = ss;
      slot[G + g] = qq;
    }
  } else {
    if (vec_per_pix < 32 && lane >= vec_per_pix) return;
    // cpg in {1, 2, 4}: fold neighbours with static register indices (no dynamic indexing -> no local memory)
    if (cpg >= 2) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        s[j] += s[j + 1];
        q[j] += q[j + 1];
      }
    }
    if (cpg >= 4) {
      s[0] += s[2];
      q[0] += q[2];
      s[4] += s[6];
      q[4] += q[6];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j % cpg == 0) {
        const int g = (vc * 8 + j) / cpg;
        slot[g] = s[j];
        slot[G + g] = q[j];
      }
    }
  }
}
// after __syncthreads(): thread i < 2G folds the warps' slots in order and stores this CTA's partial (no atomics)
__device__ __forceinline__ void gn_store_part(const float* sm, double* part, int G, int nwarps) {
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += sm[w * 2 * G + i];
    part[i] = (double)t;
  }
}
// 8 lanes per output fold `parts` partials (stride 2G doubles) in a fixed order: lane l takes parts l, l+8, ..., then a butterfly
__device__ __forceinline__ double gn_fold_parts(const double* p, int parts, int stride, int l) {
  double a = 0.0;
  for (int k = l; k < parts; k += 8) a += p[(long long)k * stride];
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  a += __shfl_xor_sync(0xffffffffu, a, 4);
  return a;
}
// ws = [N][G][2] folded sums (written here) followed by the partials [N][parts][2][G]
__global__ void __launch_bounds__(256) gn_fold_kernel(double* __restrict__ ws, int N, int G, int parts) {
  const int o = (blockIdx.x * 256 + threadIdx.x) >> 3, l = threadIdx.x & 7;
  const int total = N * 2 * G;
  const int oc = o < total ? o : total - 1;                // keep every lane in the shuffles
  const int n = oc / (2 * G), kind = (oc / G) & 1, g = oc % G;
  const double* part = ws + (long long)N * 2 * G + ((long long)n * parts * 2 + kind) * G + g;
  const double a = gn_fold_parts(part, parts, 2 * G, l);
  if (l == 0 && o < total) ws[((long long)n * G + g) * 2 + kind] = a;
}

// The single-kernel forms (off by default) keep round 1's atomic fold: their results are NOT reproducible bit for bit.
// Per-thread 8-channel partial sums (s, q) -> per-group sums in shared memory sm[0..G) / sm[G..2G).
// Channels are first combined per group in registers, then lanes of the warp that own the same channel vector
// (vec_per_pix < 32) are folded with shuffles, so each warp issues one shared atomic per (vector, group).
__device__ __forceinline__ void gn_fold_to_smem(float (&s)[8], float (&q)[8], float* sm, int G, int cpg, int vc,
                                                int vec_per_pix) {
  const int lane = threadIdx.x & 31;
  if (vec_per_pix < 32) {
    for (int o = vec_per_pix; o < 32; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
      }
    }
    if (lane >= vec_per_pix) return;
  }
  if (cpg >= 8) {
    float ss = 0.f, qq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ss += s[j];
      qq += q[j];
    }
    const int g = (vc * 8) / cpg;
    atomicAdd(&sm[g], ss);
    atomicAdd(&sm[G + g], qq);
  } else {
    // cpg in {1, 2, 4}: fold neighbours with static register indices (no dynamic indexing -> no local memory)
    if (cpg >= 2) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        s[j] += s[j + 1];
        q[j] += q[j + 1];
      }
    }
    if (cpg >= 4) {
      s[0] += s[2];
      q[0] += q[2];
      s[4] += s[6];
      q[4] += q[6];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j % cpg == 0) {
        const int g = (vc * 8 + j) / cpg;
        atomicAdd(&sm[g], s[j]);
        atomicAdd(&sm[G + g], q[j]);
      }
    }
  }
}

// x: NHWC bf16, G groups of cpg = C/G consecutive channels.  Statistics per (n, g) over HW*cpg elements.
// Pass 1: partial (sum, sumsq) per CTA -> its own slot of the partials in ws (see above).  C % 8 == 0.
// Each thread owns one 8-channel vector column (fixed group set) and strides over pixels.
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ x,
                                                              double* __restrict__ ws_parts, int HW, int C, int G,
                                                              int pix_per_cta) {
  extern __shared__ float sm[];  // [8 warps][2][G]
  const int n = blockIdx.y;
  const int cpg = C / G;
  const int vec_per_pix = C >> 3;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  for (int i = threadIdx.x; i < 16 * G; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  // thread -> (vector column vc, pixel lane pl); blockDim.x is a multiple of vec_per_pix or vice versa handled by stride
  const int vc = threadIdx.x % vec_per_pix;
  const int pl = threadIdx.x / vec_per_pix;
  const int pstride = blockDim.x / vec_per_pix;
  if (pstride > 0) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const __nv_bfloat16* base = x + (long long)n * HW * C;
    for (int p = p0 + pl; p < p1; p += pstride) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(base + (long long)p * C + vc * 8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += v[j];
        q[j] += v[j] * v[j];
      }
    }
    gn_fold_to_slots(s, q, sm, G, cpg, vc, vec_per_pix);
  }
  __syncthreads();
  gn_store_part(sm, ws_parts + ((long long)n * gridDim.x + blockIdx.x) * 2 * G, G, 8);
}

// folds the partials (fixed order), leaves the sums in ws[n][g][2] and turns them into mean / rstd; 8 lanes per (n, g)
__global__ void __launch_bounds__(256) groupnorm_finalize_kernel(double* __restrict__ ws, float* __restrict__ mean,
                                                                 float* __restrict__ rstd, int N, int G, int parts, double count,
                                                                 float eps) {
  const int o = (blockIdx.x * 256 + threadIdx.x) >> 3, l = threadIdx.x & 7;
  const int total = N * G;
  const int oc = o < total ? o : total - 1;
  const int n = oc / G, g = oc % G;
  const double* part = ws + (long long)N * 2 * G + (long long)n * parts * 2 * G + g;
  const double s1 = gn_fold_parts(part, parts, 2 * G, l);
  const double s2 = gn_fold_parts(part + G, parts, 2 * G, l);
  if (l != 0 || o >= total) return;
  ws[2 * (long long)o] = s1;
  ws[2 * (long long)o + 1] = s2;
  const double m = s1 / count;
  double var = s2 / count - m * m;
  if (var < 0) var = 0;
  mean[o] = (float)m;
  rstd[o] = (float)(1.0 / sqrt(var + (double)eps));
}

// y = act((x - mean) * rstd * gamma + beta), act = swish or identity.
// grid (pixel chunks, N); each thread owns one 8-channel vector column (its affine + statistics live in registers)
// and strides over the pixels of the chunk: no per-element index arithmetic, 16-byte coalesced loads/stores.
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              __nv_bfloat16* __restrict__ y, int HW, int C, int G,
                                                              int pix_per_cta, int swish) {
  const int n = blockIdx.y;
  const int cpg = C / G;
  const int vec_per_pix = C >> 3;
  const int vc = threadIdx.x % vec_per_pix;
  const int pl = threadIdx.x / vec_per_pix;
  const int pstride = blockDim.x / vec_per_pix;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vc * 8 + j;
    const int g = n * G + c / cpg;
    sc[j] = rstd[g] * gamma[c];
    sh[j] = beta[c] - mean[g] * sc[j];
  }
  const long long base = (long long)n * HW * C + vc * 8;
  for (int p = p0 + pl; p < p1; p += pstride) {
    float v[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(x + base + (long long)p * C), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = fmaf(v[j], sc[j], sh[j]);
      o[j] = swish ? swish_fast_f(u) : u;
    }
    *reinterpret_cast<uint4*>(y + base + (long long)p * C) = pack8(o);
  }
}

// backward pass 1: per (n,g) sums of g = dy*act'(u)*gamma and g*xhat  (double atomics into ws).
// Register diet -> occupancy: the statistics of a thread's 8 channels belong to only GPV = max(1, 8 / cpg) distinct groups,
// so the per-group coefficients (xhat = v*R + M) are held per group, not per channel (template on GPV); with two pixels per
// thread in flight the kernels fit 64 registers -> 4 CTAs (32 warps) per SM, which is what keeps HBM busy here: measured,
// many light warps beat few warps with deep unrolling on this part.
template <int GPV>
__global__ void __launch_bounds__(256, 4) groupnorm_bwd_stats_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                     const __nv_bfloat16* __restrict__ x,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta,
                                                                     double* __restrict__ ws_parts, int HW, int C, int G,
                                                                     int pix_per_cta, int swish) {
  extern __shared__ float sm[];  // [8 warps][2][G]
  constexpr int CPV = 8 / GPV;   // channels of the vector that share a group
  const int n = blockIdx.y;
  const int cpg = C / G;
  const int vec_per_pix = C >> 3;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  for (int i = threadIdx.x; i < 16 * G; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int vc = threadIdx.x % vec_per_pix;
  const int pl = threadIdx.x / vec_per_pix;
  const int pstride = blockDim.x / vec_per_pix;
  if (pstride > 0 && pl < pstride) {
    float s[8], q[8], gm[8], bt[8], R[GPV], M[GPV];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] = q[j] = 0.f;
      gm[j] = gamma[vc * 8 + j];
      bt[j] = beta[vc * 8 + j];
    }
#pragma unroll
    for (int k = 0; k < GPV; ++k) {
      const int g = n * G + (vc * 8 + k * CPV) / cpg;
      R[k] = rstd[g];
      M[k] = -mean[g] * R[k];
    }
    const long long base = (long long)n * HW * C + vc * 8;
    for (int p = p0 + pl; p < p1; p += pstride * 2) {
      uint4 xk[2], dk[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pp = p + u * pstride;
        if (pp < p1) {
          xk[u] = *reinterpret_cast<const uint4*>(x + base + (long long)pp * C);
          dk[u] = *reinterpret_cast<const uint4*>(dy + base + (long long)pp * C);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (p + u * pstride < p1) {
          float v[8], d[8];
          unpack8(xk[u], v);
          unpack8(dk[u], d);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = fmaf(v[j], R[j / CPV], M[j / CPV]);
            const float uu = fmaf(xh, gm[j], bt[j]);
            const float g = d[j] * (swish ? swish_grad_fast_f(uu) : 1.0f) * gm[j];
            s[j] += g;
            q[j] = fmaf(g, xh, q[j]);
          }
        }
      }
    }
    gn_fold_to_slots(s, q, sm, G, cpg, vc, vec_per_pix);
  }
  __syncthreads();
  gn_store_part(sm, ws_parts + ((long long)n * gridDim.x + blockIdx.x) * 2 * G, G, 8);
}

// backward pass 2: dx = rstd * (g - S1/cnt - xhat * S2/cnt) (+ add); same thread mapping as the apply kernel
template <int GPV>
__global__ void __launch_bounds__(256, 4) groupnorm_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                     const __nv_bfloat16* __restrict__ x,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta,
                                                                     const double* __restrict__ ws,
                                                                     const __nv_bfloat16* __restrict__ add,
                                                                     __nv_bfloat16* __restrict__ dx, int HW, int C, int G,
                                                                     int pix_per_cta, float inv_count, int swish) {
  constexpr int CPV = 8 / GPV;
  const int n = blockIdx.y;
  const int cpg = C / G;
  const int vec_per_pix = C >> 3;
  const int vc = threadIdx.x % vec_per_pix;
  const int pl = threadIdx.x / vec_per_pix;
  const int pstride = blockDim.x / vec_per_pix;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  float R[GPV], M[GPV], c1[GPV], c2[GPV], gm[8], bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gm[j] = gamma[vc * 8 + j];
    bt[j] = beta[vc * 8 + j];
  }
#pragma unroll
  for (int k = 0; k < GPV; ++k) {
    const int g = n * G + (vc * 8 + k * CPV) / cpg;
    R[k] = rstd[g];
    M[k] = -mean[g] * R[k];
    c1[k] = (float)ws[2 * g] * inv_count * R[k];        // rstd * S1 / cnt
    c2[k] = (float)ws[2 * g + 1] * inv_count * R[k];    // rstd * S2 / cnt
  }
  const long long base = (long long)n * HW * C + vc * 8;
  for (int p = p0 + pl; p < p1; p += pstride) {
    const long long off = base + (long long)p * C;
    const uint4 xk = __ldcs(reinterpret_cast<const uint4*>(x + off));
    const uint4 dk = __ldcs(reinterpret_cast<const uint4*>(dy + off));
    uint4 ak = make_uint4(0u, 0u, 0u, 0u);
    if (add) ak = __ldcs(reinterpret_cast<const uint4*>(add + off));
    float v[8], d[8], o[8], a[8];
    unpack8(xk, v);
    unpack8(dk, d);
    unpack8(ak, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = fmaf(v[j], R[j / CPV], M[j / CPV]);
      const float uu = fmaf(xh, gm[j], bt[j]);
      const float g = d[j] * (swish ? swish_grad_fast_f(uu) : 1.0f) * gm[j];
      o[j] = fmaf(g, R[j / CPV], a[j] - fmaf(xh, c2[j / CPV], c1[j / CPV]));
    }
    __stcs(reinterpret_cast<uint4*>(dx + off), pack8(o));
  }
}

// ------------------------------------------------------------------------------------ GroupNorm, single-kernel forms
// The two-kernel forms above read every tensor twice from HBM (statistics pass + apply pass).  The fused forms keep the
// second read in the 126 MB L2: ONE persistent grid (one 512-thread CTA per SM) walks the batch sample by sample; for
// sample n every CTA accumulates the statistics of its pixel slice (phase A), publishes them with double atomics and
// bumps a per-sample arrival counter; once the counter shows the whole grid, it normalises the SAME slice (phase B) —
// which it, and only it, streamed a moment ago, so the re-read hits L2 (one sample of the 256x256x128 tensor is 16.8 MB).
// Phase A of sample n+1 is issued BEFORE the wait for sample n (software pipelining), so the grid-wide arrival is
// practically always complete when a CTA looks at it and nobody idles at the barrier.
// HBM traffic: forward 4 B/element (was 6), backward 6 B/element (was 10).
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_stream(const __nv_bfloat16* p) {   // last use of the line: evict-first
  return __ldcs(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void st_stream(__nv_bfloat16* p, const uint4& v) { __stcs(reinterpret_cast<uint4*>(p), v); }

// Per-thread cp.async ring over this thread's strided sequence of 16-byte vectors (up to NT tensors per item).  Slots are
// private to the thread ([slot][tensor][thread] in shared memory), so consuming one needs cp.async.wait_group only — no
// barrier — and kDepth-1 items (x NT x 16 B x 512 threads) are in flight per CTA without occupying registers.  This is what
// the single-kernel GroupNorm forms lacked with register-staged loads (32 KB in flight per SM).
template <int kDepth, int NT>
struct VecRing {
  uint4* smem;                       // [kDepth][NT][blockDim.x]
  const __nv_bfloat16* ptr[NT];      // this thread's first item of every tensor (nullptr: tensor absent)
  long long stride;                  // elements between consecutive items
  int issued, total;
  __device__ __forceinline__ void issue() {
    if (issued < total) {
      const int slot = issued & (kDepth - 1);
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (ptr[t]) cp_async16(smem_u32(smem + (slot * NT + t) * blockDim.x + threadIdx.x), ptr[t] + (long long)issued * stride, true);
    }
    cp_async_commit();
    ++issued;
  }
  __device__ __forceinline__ void start(int n_items) {
    total = n_items;
    issued = 0;
#pragma unroll
    for (int i = 0; i < kDepth - 1; ++i) issue();
  }
  // item k (consumed in order): top up the ring, wait for the oldest group, hand out the slot
  __device__ __forceinline__ const uint4* next(int k) {
    issue();
    cp_async_wait<kDepth - 1>();
    return smem + ((k & (kDepth - 1)) * NT) * blockDim.x + threadIdx.x;
  }
  __device__ __forceinline__ void drain() { cp_async_wait<0>(); }
};
static constexpr int kGnRingDepthFwd = 16, kGnRingDepthBwd = 8;

static constexpr int kGnThreads = 512;
static constexpr int kGnUnroll = 4;
static constexpr int kGnUnrollBwd = 2;   // backward holds 6 per-channel coefficient vectors: 2 x (x, dy, add) loads in flight fit 128 registers

// publish this CTA's per-group partial sums (shared memory, [2][G] floats) for sample n and signal arrival
__device__ __forceinline__ void gn_publish(const float* part, double* ws, int* cnt, int n, int G) {
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    atomicAdd(&ws[((long long)n * G + i) * 2 + 0], (double)part[i]);
    atomicAdd(&ws[((long long)n * G + i) * 2 + 1], (double)part[G + i]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&cnt[n], 1);
}
__device__ __forceinline__ void gn_wait(const int* cnt, int n, int expected) {
  if (threadIdx.x == 0) {
    while (ld_acquire_gpu(&cnt[n]) < expected) __nanosleep(40);
  }
  __syncthreads();
}

template <bool kRing>
__global__ void __launch_bounds__(kGnThreads, 1)
groupnorm_fused_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                           __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                           double* __restrict__ ws, int* __restrict__ cnt, int N, int HW, int C, int G, int ppc, int swish,
                           float eps, int pipeline) {
  extern __shared__ float sm[];          // part[2][2G] (by sample parity) + stat[2G]
  float* stat = sm + 4 * G;
  const int cpg = C / G, vpp = C >> 3;
  const int vc = threadIdx.x % vpp, pl = threadIdx.x / vpp, pstride = kGnThreads / vpp;
  const int p0 = blockIdx.x * ppc, p1 = min(HW, p0 + ppc);
  const int my_items = (p0 + pl < p1) ? (p1 - p0 - pl + pstride - 1) / pstride : 0;      // pixel vectors this thread visits per sample
  uint4* ring_smem = reinterpret_cast<uint4*>(sm + ((6 * G + 3) & ~3));                   // 16-byte aligned, after part[] / stat[]
  const double inv_count = 1.0 / ((double)HW * cpg);
  float gm[8], bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gm[j] = gamma[vc * 8 + j];
    bt[j] = beta[vc * 8 + j];
  }
  auto phase_a = [&](int n) {
    float* part = sm + (n & 1) * 2 * G;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) part[i] = 0.f;
    __syncthreads();
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const __nv_bfloat16* base = x + (long long)n * HW * C + vc * 8;
    if (kRing) {
      VecRing<kGnRingDepthFwd, 1> ring;
      ring.smem = ring_smem;
      ring.ptr[0] = base + (long long)(p0 + pl) * C;
      ring.stride = (long long)pstride * C;
      ring.start(my_items);
      for (int k = 0; k < my_items; ++k) {
        float v[8];
        unpack8(*ring.next(k), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += v[j];
          q[j] = fmaf(v[j], v[j], q[j]);
        }
      }
      ring.drain();
    } else
    for (int p = p0 + pl; p < p1; p += pstride * kGnUnroll) {
      uint4 pk[kGnUnroll];
#pragma unroll
      for (int u = 0; u < kGnUnroll; ++u) {
        const int pp = p + u * pstride;
        pk[u] = pp < p1 ? *reinterpret_cast<const uint4*>(base + (long long)pp * C) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < kGnUnroll; ++u) {
        float v[8];
        unpack8(pk[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += v[j];
          q[j] = fmaf(v[j], v[j], q[j]);
        }
      }
    }
    gn_fold_to_smem(s, q, part, G, cpg, vc, vpp);
    gn_publish(part, ws, cnt, n, G);
  };
  auto phase_b = [&](int n) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (vc * 8 + j) / cpg;
      sc[j] = stat[G + g] * gm[j];
      sh[j] = bt[j] - stat[g] * sc[j];
    }
    const long long base = (long long)n * HW * C + vc * 8;
    if (kRing) {
      VecRing<kGnRingDepthFwd, 1> ring;
      ring.smem = ring_smem;
      ring.ptr[0] = x + base + (long long)(p0 + pl) * C;
      ring.stride = (long long)pstride * C;
      ring.start(my_items);
      __nv_bfloat16* yo = y + base + (long long)(p0 + pl) * C;
      for (int k = 0; k < my_items; ++k) {
        float v[8], o[8];
        unpack8(*ring.next(k), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(v[j], sc[j], sh[j]);
          o[j] = swish ? swish_fast_f(t) : t;
        }
        st_stream(yo + (long long)k * pstride * C, pack8(o));
      }
      ring.drain();
    } else
    for (int p = p0 + pl; p < p1; p += pstride * kGnUnroll) {
      uint4 pk[kGnUnroll];
#pragma unroll
      for (int u = 0; u < kGnUnroll; ++u) {
        const int pp = p + u * pstride;
        if (pp < p1) pk[u] = ld_stream(x + base + (long long)pp * C);
      }
#pragma unroll
      for (int u = 0; u < kGnUnroll; ++u) {
        const int pp = p + u * pstride;
        if (pp < p1) {
          float v[8], o[8];
          unpack8(pk[u], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = fmaf(v[j], sc[j], sh[j]);
            o[j] = swish ? swish_fast_f(t) : t;
          }
          st_stream(y + base + (long long)pp * C, pack8(o));
        }
      }
    }
  };
  if (pipeline) phase_a(0);
  for (int n = 0; n < N; ++n) {
    if (pipeline) {
      if (n + 1 < N) phase_a(n + 1);
    } else {
      phase_a(n);
    }
    gn_wait(cnt, n, (int)gridDim.x);
    if (threadIdx.x < G) {
      const long long gi = (long long)n * G + threadIdx.x;
      const double m = __ldcg(&ws[2 * gi]) * inv_count;
      double var = __ldcg(&ws[2 * gi + 1]) * inv_count - m * m;
      if (var < 0) var = 0;
      const float mu = (float)m, rs = (float)(1.0 / sqrt(var + (double)eps));
      stat[threadIdx.x] = mu;
      stat[G + threadIdx.x] = rs;
      if (blockIdx.x == 0) {
        mean_out[gi] = mu;
        rstd_out[gi] = rs;
      }
    }
    __syncthreads();
    phase_b(n);
    __syncthreads();   // stat is rewritten in the next iteration
  }
}

template <bool kRing>
__global__ void __launch_bounds__(kGnThreads, 1)
groupnorm_fused_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                           const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                           const float* __restrict__ beta, const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ dx,
                           double* __restrict__ ws, int* __restrict__ cnt, int N, int HW, int C, int G, int ppc, int swish,
                           int pipeline) {
  extern __shared__ float sm[];          // part[2][2G] + stat[2G]
  float* stat = sm + 4 * G;
  const int cpg = C / G, vpp = C >> 3;
  const int vc = threadIdx.x % vpp, pl = threadIdx.x / vpp, pstride = kGnThreads / vpp;
  const int p0 = blockIdx.x * ppc, p1 = min(HW, p0 + ppc);
  const int my_items = (p0 + pl < p1) ? (p1 - p0 - pl + pstride - 1) / pstride : 0;
  uint4* ring_smem = reinterpret_cast<uint4*>(sm + ((6 * G + 3) & ~3));
  const float inv_count = 1.0f / ((float)HW * cpg);
  float gm[8], bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gm[j] = gamma[vc * 8 + j];
    bt[j] = beta[vc * 8 + j];
  }
  // g = dy * act'(u) * gamma and xhat for one 8-channel vector
  auto grad8 = [&](const uint4& xk, const uint4& dk, const float (&mu)[8], const float (&rs)[8], float (&g)[8], float (&xh)[8]) {
    float v[8], d[8];
    unpack8(xk, v);
    unpack8(dk, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[j] = (v[j] - mu[j]) * rs[j];
      const float u = fmaf(xh[j], gm[j], bt[j]);
      g[j] = d[j] * (swish ? swish_grad_fast_f(u) : 1.0f) * gm[j];
    }
  };
  auto load_stats = [&](int n, float (&mu)[8], float (&rs)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = n * G + (vc * 8 + j) / cpg;
      mu[j] = mean[g];
      rs[j] = rstd[g];
    }
  };
  auto phase_a = [&](int n) {
    float* part = sm + (n & 1) * 2 * G;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) part[i] = 0.f;
    __syncthreads();
    float s[8], q[8], mu[8], rs[8];
    load_stats(n, mu, rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const long long base = (long long)n * HW * C + vc * 8;
    if (kRing) {
      VecRing<kGnRingDepthBwd, 3> ring;                     // same slot geometry as phase B (x, dy, add); add unused here
      ring.smem = ring_smem;
      ring.ptr[0] = x + base + (long long)(p0 + pl) * C;
      ring.ptr[1] = dy + base + (long long)(p0 + pl) * C;
      ring.ptr[2] = nullptr;
      ring.stride = (long long)pstride * C;
      ring.start(my_items);
      for (int k = 0; k < my_items; ++k) {
        const uint4* sl = ring.next(k);
        float g[8], xh[8];
        grad8(sl[0], sl[blockDim.x], mu, rs, g, xh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += g[j];
          q[j] = fmaf(g[j], xh[j], q[j]);
        }
      }
      ring.drain();
    } else
    for (int p = p0 + pl; p < p1; p += pstride * kGnUnrollBwd) {
      uint4 xk[kGnUnrollBwd], dk[kGnUnrollBwd];
#pragma unroll
      for (int u = 0; u < kGnUnrollBwd; ++u) {
        const int pp = p + u * pstride;
        if (pp < p1) {
          xk[u] = *reinterpret_cast<const uint4*>(x + base + (long long)pp * C);
          dk[u] = *reinterpret_cast<const uint4*>(dy + base + (long long)pp * C);
        }
      }
#pragma unroll
      for (int u = 0; u < kGnUnrollBwd; ++u) {
        const int pp = p + u * pstride;
        if (pp < p1) {
          float g[8], xh[8];
          grad8(xk[u], dk[u], mu, rs, g, xh);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s[j] += g[j];
            q[j] = fmaf(g[j], xh[j], q[j]);
          }
        }
      }
    }
    gn_fold_to_smem(s, q, part, G, cpg, vc, vpp);
    gn_publish(part, ws, cnt, n, G);
  };
  auto phase_b = [&](int n) {
    float mu[8], rs[8], s1[8], s2[8];
    load_stats(n, mu, rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (vc * 8 + j) / cpg;
      s1[j] = stat[g];
      s2[j] = stat[G + g];
    }
    const long long base = (long long)n * HW * C + vc * 8;
    if (kRing) {
      VecRing<kGnRingDepthBwd, 3> ring;
      ring.smem = ring_smem;
      ring.ptr[0] = x + base + (long long)(p0 + pl) * C;
      ring.ptr[1] = dy + base + (long long)(p0 + pl) * C;
      ring.ptr[2] = add ? add + base + (long long)(p0 + pl) * C : nullptr;
      ring.stride = (long long)pstride * C;
      ring.start(my_items);
      __nv_bfloat16* dxo = dx + base + (long long)(p0 + pl) * C;
      for (int k = 0; k < my_items; ++k) {
        const uint4* sl = ring.next(k);
        float g[8], xh[8], o[8];
        grad8(sl[0], sl[blockDim.x], mu, rs, g, xh);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rs[j] * (g[j] - s1[j] - xh[j] * s2[j]);
        if (add) {
          float a[8];
          unpack8(sl[2 * blockDim.x], a);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += a[j];
        }
        st_stream(dxo + (long long)k * pstride * C, pack8(o));
      }
      ring.drain();
    } else
    for (int p = p0 + pl; p < p1; p += pstride * kGnUnrollBwd) {
      uint4 xk[kGnUnrollBwd], dk[kGnUnrollBwd], ak[kGnUnrollBwd];
#pragma unroll
      for (int u = 0; u < kGnUnrollBwd; ++u) {
        const int pp = p + u * pstride;
        if (pp < p1) {
          xk[u] = ld_stream(x + base + (long long)pp * C);
          dk[u] = ld_stream(dy + base + (long long)pp * C);
          if (add) ak[u] = ld_stream(add + base + (long long)pp * C);
        }
      }
#pragma unroll
      for (int u = 0; u < kGnUnrollBwd; ++u) {
        const int pp = p + u * pstride;
        if (pp < p1) {
          float g[8], xh[8], o[8];
          grad8(xk[u], dk[u], mu, rs, g, xh);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rs[j] * (g[j] - s1[j] - xh[j] * s2[j]);
          if (add) {
            float a[8];
            unpack8(ak[u], a);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += a[j];
          }
          st_stream(dx + base + (long long)pp * C, pack8(o));
        }
      }
    }
  };
  if (pipeline) phase_a(0);
  for (int n = 0; n < N; ++n) {
    if (pipeline) {
      if (n + 1 < N) phase_a(n + 1);
    } else {
      phase_a(n);
    }
    gn_wait(cnt, n, (int)gridDim.x);
    if (threadIdx.x < G) {
      const long long gi = (long long)n * G + threadIdx.x;
      stat[threadIdx.x] = (float)__ldcg(&ws[2 * gi]) * inv_count;
      stat[G + threadIdx.x] = (float)__ldcg(&ws[2 * gi + 1]) * inv_count;
    }
    __syncthreads();
    phase_b(n);
    __syncthreads();
  }
}

}  // namespace ffvc

using namespace ffvc;

template <int kThreads, int kRows, int kStages>
static void ln_fwd_pipe_launch_r(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y, float* mean,
                                 float* rstd, long long rows, float eps, cudaStream_t st) {
  constexpr int kSmem = kStages * kRows * kThreads * 16;
  static int per_sm = 0;                       // resident CTAs per SM of this instantiation: the grid is one full wave
  if (per_sm == 0) {
    cudaFuncSetAttribute(layernorm_fwd_pipe_kernel<kThreads, kRows, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, layernorm_fwd_pipe_kernel<kThreads, kRows, kStages>, kThreads, kSmem);
    per_sm = n > 0 ? n : 1;
  }
  const long long want = (rows + kRows - 1) / kRows;
  const unsigned grid = (unsigned)(want < 148LL * per_sm ? want : 148LL * per_sm);
  layernorm_fwd_pipe_kernel<kThreads, kRows, kStages><<<grid, kThreads, kSmem, st>>>(x, gamma, beta, y, mean, rstd, rows, eps);
}
// option value 1: 8 rows per group, 2 stages;  2: 4 rows per group, 4 stages (same shared memory, finer-grained ring)
template <int kThreads>
static void ln_fwd_cols_launch(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y, float* mean,
                               float* rstd, long long rows, float eps, cudaStream_t st) {
  if (option(OPT_LN_FWD_V2) == 2) ln_fwd_pipe_launch_r<kThreads, 4, 4>(x, gamma, beta, y, mean, rstd, rows, eps, st);
  else ln_fwd_pipe_launch_r<kThreads, 8, 2>(x, gamma, beta, y, mean, rstd, rows, eps, st);
}
static bool ln_cols_ok(int D, const void* a, const void* b, const void* c, const void* d) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                       reinterpret_cast<uintptr_t>(d);
  return (D == 256 || D == 512 || D == 768 || D == 1024) && (al & 15) == 0;
}

extern "C" int ffvc_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                  float* rstd, long long rows, int D, float eps, void* stream) {
  if (D % 8 != 0 || D > 2048) return set_error(FFVC_ERR_ARG, "layernorm: D must be a multiple of 8 and <= 2048");
  if (rows <= 0) return FFVC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto yb = reinterpret_cast<__nv_bfloat16*>(y);
  if (option(OPT_LN_FWD_V2) && ln_cols_ok(D, x, y, gamma, beta)) {
    switch (D) {
      case 256: ln_fwd_cols_launch<32>(xb, gamma, beta, yb, mean, rstd, rows, eps, st); break;
      case 512: ln_fwd_cols_launch<64>(xb, gamma, beta, yb, mean, rstd, rows, eps, st); break;
      case 768: ln_fwd_cols_launch<96>(xb, gamma, beta, yb, mean, rstd, rows, eps, st); break;
      default: ln_fwd_cols_launch<128>(xb, gamma, beta, yb, mean, rstd, rows, eps, st); break;
    }
    FFVC_CHECK_LAUNCH();
    return FFVC_OK;
  }
  const int wpb = 8;
  long long want = (rows + wpb - 1) / wpb;
  const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);      // resident warps loop over the rows
  const size_t smem = 2 * (size_t)D * sizeof(float);
  if (D <= 1024)
    layernorm_fwd_kernel<4><<<grid, wpb * 32, smem, st>>>(xb, gamma, beta, yb, mean, rstd, rows, D, eps);
  else
    layernorm_fwd_kernel<8><<<grid, wpb * 32, smem, st>>>(xb, gamma, beta, yb, mean, rstd, rows, D, eps);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                  const float* rstd, const void* add, void* dx, float* dgamma, float* dbeta,
                                  long long rows, int D, void* stream) {
  if (D % 8 != 0 || D > 2048) return set_error(FFVC_ERR_ARG, "layernorm: D must be a multiple of 8 and <= 2048");
  if (rows <= 0) return FFVC_OK;
  if ((dgamma == nullptr) != (dbeta == nullptr)) return set_error(FFVC_ERR_ARG, "layernorm_bwd: dgamma/dbeta both or none");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int wpb = 8;
  long long want = (rows + wpb - 1) / wpb;
  const int per_sm = dgamma ? 2 : 3;
  const unsigned grid = (unsigned)(want < 148 * per_sm ? want : 148 * per_sm);
  const size_t smem = (dgamma ? 3 : 1) * (size_t)D * sizeof(float);
  auto dyb = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto ab = reinterpret_cast<const __nv_bfloat16*>(add);
  auto dxb = reinterpret_cast<__nv_bfloat16*>(dx);
  if (D <= 1024) {
    if (dgamma) layernorm_bwd_kernel<4, true><<<grid, wpb * 32, smem, st>>>(dyb, xb, gamma, mean, rstd, ab, dxb, dgamma, dbeta, rows, D);
    else layernorm_bwd_kernel<4, false><<<grid, wpb * 32, smem, st>>>(dyb, xb, gamma, mean, rstd, ab, dxb, dgamma, dbeta, rows, D);
  } else {
    if (dgamma) layernorm_bwd_kernel<8, true><<<grid, wpb * 32, smem, st>>>(dyb, xb, gamma, mean, rstd, ab, dxb, dgamma, dbeta, rows, D);
    else layernorm_bwd_kernel<8, false><<<grid, wpb * 32, smem, st>>>(dyb, xb, gamma, mean, rstd, ab, dxb, dgamma, dbeta, rows, D);
  }
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

static const int kLnBwdMaxCtasPerSm = 8;
template <int kThreads, int kRows, int kStages>
static void ln_bwd_pipe_launch_r(const __nv_bfloat16* dy, const __nv_bfloat16* x, const float* gamma, const float* mean,
                                 const float* rstd, const __nv_bfloat16* add, __nv_bfloat16* dx, float* dgamma, float* dbeta,
                                 float* colsum_out, float* rowsum_out, int rowsum_T, long long rows, float* ws, cudaStream_t st) {
  constexpr int D = kThreads * 8;
  constexpr int kRing = kStages * kRows * 3 * kThreads * 16;
  constexpr int kMaxSmem = kRing + 2048 * 4;                 // ring + per-token sums (rowsum_T <= 2048)
  static int per_sm_w = 0, per_sm_n = 0;       // resident CTAs per SM (wgrad / no-wgrad form): the grid is one full wave
  if (per_sm_w == 0) {
    cudaFuncSetAttribute(layernorm_bwd_pipe_kernel<kThreads, kRows, kStages, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaFuncSetAttribute(layernorm_bwd_pipe_kernel<kThreads, kRows, kStages, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, layernorm_bwd_pipe_kernel<kThreads, kRows, kStages, true>, kThreads, kRing + 1024);
    per_sm_w = n > 0 ? (n < kLnBwdMaxCtasPerSm ? n : kLnBwdMaxCtasPerSm) : 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, layernorm_bwd_pipe_kernel<kThreads, kRows, kStages, false>, kThreads, kRing + 1024);
    per_sm_n = n > 0 ? (n < kLnBwdMaxCtasPerSm ? n : kLnBwdMaxCtasPerSm) : 1;
  }
  const bool want_row = rowsum_out != nullptr, want_col = colsum_out != nullptr;
  const int T = want_row ? rowsum_T : 0;
  const size_t smem = (size_t)kRing + (size_t)T * sizeof(float);
  const int per_sm = dgamma ? per_sm_w : per_sm_n;
  const long long want = (rows + kRows - 1) / kRows;
  const unsigned grid = (unsigned)(want < 148LL * per_sm ? want : 148LL * per_sm);
  if (dgamma)
    layernorm_bwd_pipe_kernel<kThreads, kRows, kStages, true><<<grid, kThreads, smem, st>>>(
        dy, x, gamma, mean, rstd, add, dx, ws, want_col, want_row, T, rows);
  else
    layernorm_bwd_pipe_kernel<kThreads, kRows, kStages, false><<<grid, kThreads, smem, st>>>(
        dy, x, gamma, mean, rstd, add, dx, ws, want_col, want_row, T, rows);
  if (dgamma || want_col || want_row) {
    count_launch();
    const int width = 3 * D + T;
    // 32 chunks of partial rows per column block: ~18 sequential loads per thread instead of ~74 (the kernel is pure latency),
    // atomic depth 32 per address
    ln_bwd_finalize_kernel<<<dim3((unsigned)((width + 255) / 256), 32), 256, 0, st>>>(ws, (int)grid, D, T, dgamma, dbeta, colsum_out,
                                                                                       rowsum_out);
  }
}
// option value 1: 4 rows per group, 2 stages;  2: 2 rows per group, 4 stages (same shared memory, finer-grained ring)
template <int kThreads>
static void ln_bwd_cols_launch(const __nv_bfloat16* dy, const __nv_bfloat16* x, const float* gamma, const float* mean,
                               const float* rstd, const __nv_bfloat16* add, __nv_bfloat16* dx, float* dgamma, float* dbeta,
                               float* colsum_out, float* rowsum_out, int rowsum_T, long long rows, float* ws, cudaStream_t st) {
  if (option(OPT_LN_BWD_V2) == 2)
    ln_bwd_pipe_launch_r<kThreads, 2, 4>(dy, x, gamma, mean, rstd, add, dx, dgamma, dbeta, colsum_out, rowsum_out, rowsum_T, rows, ws, st);
  else
    ln_bwd_pipe_launch_r<kThreads, 4, 2>(dy, x, gamma, mean, rstd, add, dx, dgamma, dbeta, colsum_out, rowsum_out, rowsum_T, rows, ws, st);
}

// LayerNorm backward that also accumulates bias gradients taken from its own output:
//   colsum_out[d] += sum_rows dx[row][d]           (optional)
//   rowsum_out[t] += sum_{row % T == t, d} dx[row][d]   (optional; rows = B*T token-major)
// One fused kernel when the column-owning form applies (D in {256,512,768,1024}, option "ln_bwd_v2"), otherwise the
// warp-per-row kernel followed by ffvc_colsum / ffvc_rowsum over dx.
extern "C" int ffvc_layernorm_bwd_sums(const void* dy, const void* x, const float* gamma, const float* mean,
                                       const float* rstd, const void* add, void* dx, float* dgamma, float* dbeta,
                                       float* colsum_out, float* rowsum_out, int rowsum_T, float* ws, long long rows, int D,
                                       void* stream) {
  if (D % 8 != 0 || D > 2048) return set_error(FFVC_ERR_ARG, "layernorm: D must be a multiple of 8 and <= 2048");
  if ((dgamma == nullptr) != (dbeta == nullptr)) return set_error(FFVC_ERR_ARG, "layernorm_bwd: dgamma/dbeta both or none");
  if (rowsum_out && (rowsum_T <= 0 || rows % rowsum_T != 0))
    return set_error(FFVC_ERR_ARG, "layernorm_bwd_sums: rows must be a multiple of rowsum_T");
  if (rows <= 0) return FFVC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (option(OPT_LN_BWD_V2) && ws && ln_cols_ok(D, dy, x, dx, add) && ln_cols_ok(D, gamma, ws, nullptr, nullptr) &&
      (!rowsum_out || (rowsum_T <= 2048 && rowsum_T % 4 == 0))) {
    auto dyb = reinterpret_cast<const __nv_bfloat16*>(dy);
    auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
    auto ab = reinterpret_cast<const __nv_bfloat16*>(add);
    auto dxb = reinterpret_cast<__nv_bfloat16*>(dx);
    switch (D) {
      case 256: ln_bwd_cols_launch<32>(dyb, xb, gamma, mean, rstd, ab, dxb, dgamma, dbeta, colsum_out, rowsum_out, rowsum_T, rows, ws, st); break;
      case 512: ln_bwd_cols_launch<64>(dyb, xb, gamma, mean, rstd, ab, dxb, dgamma, dbeta, colsum_out, rowsum_out, rowsum_T, rows, ws, st); break;
      case 768: ln_bwd_cols_launch<96>(dyb, xb, gamma, mean, rstd, ab, dxb, dgamma, dbeta, colsum_out, rowsum_out, rowsum_T, rows, ws, st); break;
      default: ln_bwd_cols_launch<128>(dyb, xb, gamma, mean, rstd, ab, dxb, dgamma, dbeta, colsum_out, rowsum_out, rowsum_T, rows, ws, st); break;
    }
    FFVC_CHECK_LAUNCH();
    return FFVC_OK;
  }
  int rc = ffvc_layernorm_bwd(dy, x, gamma, mean, rstd, add, dx, dgamma, dbeta, rows, D, stream);
  if (rc) return rc;
  if (colsum_out && (rc = ffvc_colsum(dx, colsum_out, rows, D, stream))) return rc;
  if (rowsum_out && (rc = ffvc_rowsum(dx, rowsum_out, (int)(rows / rowsum_T), rowsum_T, D, stream))) return rc;
  return FFVC_OK;
}

// bytes of scratch ffvc_layernorm_bwd_sums needs for the fused form (one row of partial sums per resident CTA)
extern "C" long long ffvc_layernorm_bwd_ws_bytes(int D, int rowsum_T) {
  return (long long)sizeof(float) * 148 * kLnBwdMaxCtasPerSm * (3LL * D + (rowsum_T > 0 ? rowsum_T : 0));
}

// pixels per CTA: aim for >= 8 CTAs per SM worth of work, at least 64 pixels each
static int gn_pix_per_cta(int N, int HW) {
  int ppc = 1024;
  while (ppc > 64 && (long long)N * ((HW + ppc - 1) / ppc) < 148 * 8) ppc >>= 1;
  return ppc;
}
static int gn_check(int C, int G) {
  if (C % 8 != 0 || G <= 0 || C % G != 0 || C / 8 > 256 || 256 % (C / 8) != 0)
    return set_error(FFVC_ERR_ARG, "groupnorm: C must be a multiple of 8, divisible by G, C/8 must divide 256");
  return FFVC_OK;
}

// Statistics passes: pixels per CTA depend on HW ONLY (never on N), so the partial sums of a sample — and with them mean / rstd —
// are the same numbers whatever batch the sample is part of (4 .. 128 parts per sample).
static int gn_stats_ppc(int HW) {
  int ppc = HW / 128;
  if (ppc < 64) ppc = 64;
  if (ppc > 512) ppc = 512;
  return ppc;
}
static int gn_stats_parts(int HW) { return (HW + gn_stats_ppc(HW) - 1) / gn_stats_ppc(HW); }
// parts per sample the conv epilogues produce (one per 2-row x 128-pixel tile), 0 when the halo conv does not apply
static int gn_epi_parts(int HW) { return HW % 256 == 0 ? HW / 256 : 0; }

// doubles of statistics workspace for N samples of HW pixels: [N][G][2] folded sums + [N][parts][2][G] per-CTA partials
extern "C" long long ffvc_groupnorm_ws_doubles(int N, int HW, int G) {
  const int parts = gn_stats_parts(HW) > gn_epi_parts(HW) ? gn_stats_parts(HW) : gn_epi_parts(HW);
  return (long long)N * 2 * G * (1 + parts);
}

// ws: ffvc_groupnorm_ws_doubles(N, HW, G) doubles.  Produces mean/rstd [N*G]; leaves (sum, sum of squares) in ws[n][g][2].
extern "C" int ffvc_groupnorm_stats(const void* x, double* ws, float* mean, float* rstd, int N, int HW, int C, int G,
                                    float eps, void* stream) {
  int rc = gn_check(C, G);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int pix_per_cta = gn_stats_ppc(HW), parts = gn_stats_parts(HW);
  dim3 grid(parts, N);
  groupnorm_stats_kernel<<<grid, 256, 16 * G * sizeof(float), st>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                   ws + (long long)N * 2 * G, HW, C, G, pix_per_cta);
  FFVC_CHECK_LAUNCH();
  groupnorm_finalize_kernel<<<(N * G * 8 + 255) / 256, 256, 0, st>>>(ws, mean, rstd, N, G, parts, (double)HW * (C / G), eps);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

// statistics produced by the conv epilogue (ffvc_conv3x3_halo_gn: one partial per 256-pixel tile behind the [N][G][2] block of
// ws) -> mean / rstd; the folded sums are left in ws[n][g][2]
extern "C" int ffvc_groupnorm_finalize(double* ws, float* mean, float* rstd, int N, int HW, int C, int G, float eps,
                                       void* stream) {
  if (G <= 0 || C % G != 0) return set_error(FFVC_ERR_ARG, "groupnorm_finalize: C must be divisible by G");
  if (gn_epi_parts(HW) == 0) return set_error(FFVC_ERR_ARG, "groupnorm_finalize: HW must be a multiple of 256 (conv epilogue tiles)");
  groupnorm_finalize_kernel<<<(N * G * 8 + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ws, mean, rstd, N, G, gn_epi_parts(HW), (double)HW * (C / G), eps);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_groupnorm_apply(const void* x, const float* mean, const float* rstd, const float* gamma,
                                    const float* beta, void* y, int N, int HW, int C, int G, int swish, void* stream) {
  int rc = gn_check(C, G);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int pix_per_cta = gn_pix_per_cta(N, HW);
  dim3 grid((HW + pix_per_cta - 1) / pix_per_cta, N);
  groupnorm_apply_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), mean, rstd, gamma, beta,
                                               reinterpret_cast<__nv_bfloat16*>(y), HW, C, G, pix_per_cta, swish);
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

// dx = d/dx [ act(GN(x)) ] . dy  (+ add).  ws: statistics workspace (ffvc_groupnorm_ws_doubles).
template <int GPV>
static void gn_bwd_launch(dim3 grid, cudaStream_t st, const __nv_bfloat16* dyb, const __nv_bfloat16* xb, const float* mean,
                          const float* rstd, const float* gamma, const float* beta, double* ws, const __nv_bfloat16* add,
                          __nv_bfloat16* dx, int N, int HW, int C, int G, int ppc, int swish) {
  // pass 1 on the batch-invariant grid: per-CTA partials; fold (fixed order) into ws[n][g][2]; pass 2
  const int ppc_s = gn_stats_ppc(HW), parts = gn_stats_parts(HW);
  groupnorm_bwd_stats_kernel<GPV><<<dim3(parts, N), 256, 16 * G * sizeof(float), st>>>(dyb, xb, mean, rstd, gamma, beta,
                                                                                         ws + (long long)N * 2 * G, HW, C, G, ppc_s, swish);
  gn_fold_kernel<<<(N * 2 * G * 8 + 255) / 256, 256, 0, st>>>(ws, N, G, parts);
  groupnorm_bwd_apply_kernel<GPV><<<grid, 256, 0, st>>>(dyb, xb, mean, rstd, gamma, beta, ws, add, dx, HW, C, G, ppc,
                                                         1.0f / ((float)HW * (C / G)), swish);
}

template <int GPV>
static void gn_bwd_apply_launch(dim3 grid, cudaStream_t st, const __nv_bfloat16* dyb, const __nv_bfloat16* xb, const float* mean,
                                const float* rstd, const float* gamma, const float* beta, const double* ws, const __nv_bfloat16* add,
                                __nv_bfloat16* dx, int HW, int C, int G, int ppc, int swish) {
  groupnorm_bwd_apply_kernel<GPV><<<grid, 256, 0, st>>>(dyb, xb, mean, rstd, gamma, beta, ws, add, dx, HW, C, G, ppc,
                                                         1.0f / ((float)HW * (C / G)), swish);
}

// second pass of ffvc_groupnorm_bwd alone: sums[N*G][2] = (sum g, sum g * xhat) were produced elsewhere (the dgrad conv's
// epilogue, ffvc_conv3x3_halo_gnbwd).
extern "C" int ffvc_groupnorm_bwd_apply(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                                        const float* beta, double* sums, const void* add, void* dx, int N, int HW, int C,
                                        int G, int swish, void* stream) {
  int rc = gn_check(C, G);
  if (rc) return rc;
  if (gn_epi_parts(HW) == 0) return set_error(FFVC_ERR_ARG, "groupnorm_bwd_apply: HW must be a multiple of 256 (conv epilogue tiles)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  gn_fold_kernel<<<(N * 2 * G * 8 + 255) / 256, 256, 0, st>>>(sums, N, G, gn_epi_parts(HW));   // per-tile partials -> sums[n][g][2]
  int ppc = 1024;
  while (ppc > 64 && (long long)N * ((HW + ppc - 1) / ppc) < 148 * 16) ppc >>= 1;
  dim3 grid((HW + ppc - 1) / ppc, N);
  auto dyb = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto ab = reinterpret_cast<const __nv_bfloat16*>(add);
  auto dxb = reinterpret_cast<__nv_bfloat16*>(dx);
  const int cpg = C / G;
  if (cpg >= 8 && cpg % 8 == 0) gn_bwd_apply_launch<1>(grid, st, dyb, xb, mean, rstd, gamma, beta, sums, ab, dxb, HW, C, G, ppc, swish);
  else if (cpg == 4) gn_bwd_apply_launch<2>(grid, st, dyb, xb, mean, rstd, gamma, beta, sums, ab, dxb, HW, C, G, ppc, swish);
  else if (cpg == 2) gn_bwd_apply_launch<4>(grid, st, dyb, xb, mean, rstd, gamma, beta, sums, ab, dxb, HW, C, G, ppc, swish);
  else if (cpg == 1) gn_bwd_apply_launch<8>(grid, st, dyb, xb, mean, rstd, gamma, beta, sums, ab, dxb, HW, C, G, ppc, swish);
  else return set_error(FFVC_ERR_UNSUPPORTED, "groupnorm_bwd: channels per group must be 1, 2, 4 or a multiple of 8");
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

// dx = d/dx [ act(GN(x)) ] . dy  (+ add).  ws: statistics workspace (ffvc_groupnorm_ws_doubles); ends with (sum g, sum g xhat) in ws[n][g][2].
extern "C" int ffvc_groupnorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                                  const float* gamma, const float* beta, double* ws, const void* add, void* dx, int N,
                                  int HW, int C, int G, int swish, void* stream) {
  int rc = gn_check(C, G);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int ppc = 1024;   // pixels per CTA of the apply pass: aim for >= 16 CTAs per SM worth of work (4 resident), at least 64 pixels each
  while (ppc > 64 && (long long)N * ((HW + ppc - 1) / ppc) < 148 * 16) ppc >>= 1;
  dim3 grid((HW + ppc - 1) / ppc, N);
  auto dyb = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto xb = reinterpret_cast<const __nv_bfloat16*>(x);
  auto ab = reinterpret_cast<const __nv_bfloat16*>(add);
  auto dxb = reinterpret_cast<__nv_bfloat16*>(dx);
  const int cpg = C / G;
  // groups covered by one 8-channel vector: 8 / cpg when cpg divides 8, else (cpg a multiple of 8) exactly one
  if (cpg >= 8 && cpg % 8 == 0) gn_bwd_launch<1>(grid, st, dyb, xb, mean, rstd, gamma, beta, ws, ab, dxb, N, HW, C, G, ppc, swish);
  else if (cpg == 4) gn_bwd_launch<2>(grid, st, dyb, xb, mean, rstd, gamma, beta, ws, ab, dxb, N, HW, C, G, ppc, swish);
  else if (cpg == 2) gn_bwd_launch<4>(grid, st, dyb, xb, mean, rstd, gamma, beta, ws, ab, dxb, N, HW, C, G, ppc, swish);
  else if (cpg == 1) gn_bwd_launch<8>(grid, st, dyb, xb, mean, rstd, gamma, beta, ws, ab, dxb, N, HW, C, G, ppc, swish);
  else return set_error(FFVC_ERR_UNSUPPORTED, "groupnorm_bwd: channels per group must be 1, 2, 4 or a multiple of 8");
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

// ---- single-kernel forms (see groupnorm_fused_*_kernel).  ws: 2*N*G doubles followed by N ints (ffvc_groupnorm_ws_bytes).
static int g_gn_sms = 0;
static int gn_fused_grid(int HW, int C, int* ppc_out) {
  if (g_gn_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_gn_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_gn_sms <= 0) g_gn_sms = 148;
  }
  const int pstride = kGnThreads / (C >> 3);
  int ppc = (HW + g_gn_sms - 1) / g_gn_sms;
  ppc = (ppc + pstride - 1) / pstride * pstride;          // whole thread-rows of pixels per CTA
  *ppc_out = ppc;
  return (HW + ppc - 1) / ppc;                            // <= number of SMs: every CTA is resident (the kernel spins on arrivals)
}
static int gn_fused_check(int C, int G) {
  if (C % 8 != 0 || G <= 0 || G > 128 || C % G != 0 || kGnThreads % (C / 8) != 0)
    return set_error(FFVC_ERR_ARG, "groupnorm_fused: C must be a multiple of 8, divisible by G, C/8 must divide 512");
  return FFVC_OK;
}
static int g_gn_pipeline = 1;
static bool gn_pipeline_default() { return g_gn_pipeline != 0; }
// schedule of the single-kernel forms: 1 (default) = statistics of sample n+1 are issued before the wait for sample n
extern "C" int ffvc_groupnorm_set_pipeline(int on) {
  g_gn_pipeline = on ? 1 : 0;
  return FFVC_OK;
}

extern "C" long long ffvc_groupnorm_ws_bytes(int N, int G) { return (long long)sizeof(double) * 2 * N * G + (long long)sizeof(int) * N; }

extern "C" int ffvc_groupnorm_fused_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                        double* ws, int N, int HW, int C, int G, int swish, float eps, void* stream) {
  int rc = gn_fused_check(C, G);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(ws, 0, (size_t)ffvc_groupnorm_ws_bytes(N, G), st);
  int ppc;
  const int grid = gn_fused_grid(HW, C, &ppc);
  int* cnt = reinterpret_cast<int*>(ws + 2 * (long long)N * G);
  if (option(OPT_GN_RING)) {
    const size_t smem = (size_t)((6 * G + 3) & ~3) * sizeof(float) + (size_t)kGnRingDepthFwd * kGnThreads * 16;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(groupnorm_fused_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set = true;
    }
    groupnorm_fused_fwd_kernel<true><<<grid, kGnThreads, smem, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), mean, rstd, ws, cnt, N, HW, C, G,
        ppc, swish, eps, gn_pipeline_default() ? 1 : 0);
  } else {
    groupnorm_fused_fwd_kernel<false><<<grid, kGnThreads, 6 * G * sizeof(float), st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), mean, rstd, ws, cnt, N, HW, C, G,
        ppc, swish, eps, gn_pipeline_default() ? 1 : 0);
  }
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}

extern "C" int ffvc_groupnorm_fused_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                                        const float* beta, double* ws, const void* add, void* dx, int N, int HW, int C, int G,
                                        int swish, void* stream) {
  int rc = gn_fused_check(C, G);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(ws, 0, (size_t)ffvc_groupnorm_ws_bytes(N, G), st);
  int ppc;
  const int grid = gn_fused_grid(HW, C, &ppc);
  int* cnt = reinterpret_cast<int*>(ws + 2 * (long long)N * G);
  if (option(OPT_GN_RING)) {
    const size_t smem = (size_t)((6 * G + 3) & ~3) * sizeof(float) + (size_t)kGnRingDepthBwd * 3 * kGnThreads * 16;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(groupnorm_fused_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set = true;
    }
    groupnorm_fused_bwd_kernel<true><<<grid, kGnThreads, smem, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), mean, rstd, gamma, beta,
        reinterpret_cast<const __nv_bfloat16*>(add), reinterpret_cast<__nv_bfloat16*>(dx), ws, cnt, N, HW, C, G, ppc, swish,
        gn_pipeline_default() ? 1 : 0);
  } else {
    groupnorm_fused_bwd_kernel<false><<<grid, kGnThreads, 6 * G * sizeof(float), st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), mean, rstd, gamma, beta,
        reinterpret_cast<const __nv_bfloat16*>(add), reinterpret_cast<__nv_bfloat16*>(dx), ws, cnt, N, HW, C, G, ppc, swish,
        gn_pipeline_default() ? 1 : 0);
  }
  FFVC_CHECK_LAUNCH();
  return FFVC_OK;
}
