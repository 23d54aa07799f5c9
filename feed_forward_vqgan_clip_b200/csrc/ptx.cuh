// Inline-PTX wrappers for the sm_100a features used by the feed-forward VQGAN-CLIP
// train step: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld).
// Hand-written; no CUTLASS/CuTe dependency.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

namespace ffvc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- TMA loads
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- TMA stores (shared -> global, bulk async-group)
__device__ __forceinline__ void tma_store_4d(const void* desc, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (TMA) of this CTA
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane+i), columns c..c+31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16-column variant (lower register footprint for the 16-warp epilogue)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- descriptors
// UMMA shared-memory matrix descriptor, 128-byte swizzle, sm_100 (version field = 1).
// start address / LBO / SBO are byte quantities (multiples of 16).
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// UMMA instruction descriptor: kind::f16, A/B = bf16, D = fp32, M = 128.
__host__ __device__ inline uint32_t umma_idesc_bf16(int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                                  // D format: F32
  d |= 1u << 7;                                  // A format: BF16
  d |= 1u << 10;                                 // B format: BF16
  d |= static_cast<uint32_t>(a_mn_major) << 15;  // A major-ness (0 = K-major)
  d |= static_cast<uint32_t>(b_mn_major) << 16;  // B major-ness
  d |= static_cast<uint32_t>(n >> 3) << 17;      // N / 8
  d |= static_cast<uint32_t>(128 >> 4) << 24;    // M / 16
  return d;
}


// ---------------------------------------------------------------- 2-CTA (cta_group::2) forms
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load executed by either CTA of a pair; the transaction bytes are credited to the LEADER CTA's mbarrier
// (peer bit 24 of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Same, multicast: the box lands at the same shared-memory offset in every CTA of `mask` (cluster ranks), and the transaction
// bytes are credited, per destination CTA, to the mbarrier at `bar`'s offset in that CTA's PAIR LEADER (peer bit cleared).
__device__ __forceinline__ void tma_load_4d_2sm_mc(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2, int c3,
                                                   uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 across the CTA pair (128 rows per CTA), B split in halves of N/2 per CTA.  Issued by the leader CTA only.
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit -> arrive on the mbarrier at the same shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
// arrive on the mbarrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
__host__ __device__ inline uint32_t umma_idesc_bf16_m(int m, int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= static_cast<uint32_t>(a_mn_major) << 15;
  d |= static_cast<uint32_t>(b_mn_major) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------- small math helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact-erf GELU (nn.GELU() default).  Scalar form (SIMT kernels): Abramowitz-Stegun 7.1.26 rational erf
// (|abs err| <= 1.5e-7): one MUFU.RCP, one MUFU.EX2 and ~10 FMAs.  cdf = Phi(x), pdf = phi(x) share the exponential.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));  // MUFU.RCP (the IEEE __frcp_rn is a branchy software sequence)
  const float e = __expf(-z * z);
  const float poly =
      t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  const float tail = 0.5f * poly * e;  // = 0.5 * erfc(|x| / sqrt(2))
  cdf = (x >= 0.f) ? 1.0f - tail : tail;
  pdf = 0.39894228040143268f * e;
}
__device__ __forceinline__ float gelu_f(float x) {
  float c, d;
  gelu_parts(x, c, d);
  return x * c;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float c, d;
  gelu_parts(x, c, d);
  return fmaf(x, d, c);
}

// Packed form for the GEMM epilogues (two elements per instruction: Blackwell FFMA2 / FMUL2 / FADD2).  The epilogues of the
// short-K GEMMs are issue-slot bound, so the normal CDF is evaluated WITHOUT any MUFU op and without range clamps:
//   Phi(x) = sat(0.5 + x * P(x^2)),  P = degree-8 minimax polynomial fitted on |x| <= 4.2 (9 coefficients)
// The leading coefficient is positive, so outside the fitted range x*P(x^2) runs monotonically to +-inf and the free
// saturation of FFMA.SAT pins Phi to exactly 0 / 1.  max |Phi error| = 1.3e-5, max |gelu error| = 5.6e-5 over all x
// (fp32 Horner, checked on a 600k-point grid on |x| <= 12) — below the bf16 rounding of the stored activation.
__device__ __forceinline__ float2 gelu_phi2(const float2 x, float2& t) {
  t = __fmul2_rn(x, x);
  float2 p = make_float2(5.997774211e-11f, 5.997774211e-11f);
  p = __ffma2_rn(p, t, make_float2(-5.633105040e-09f, -5.633105040e-09f));
  p = __ffma2_rn(p, t, make_float2(2.343611383e-07f, 2.343611383e-07f));
  p = __ffma2_rn(p, t, make_float2(-5.760671296e-06f, -5.760671296e-06f));
  p = __ffma2_rn(p, t, make_float2(9.457352734e-05f, 9.457352734e-05f));
  p = __ffma2_rn(p, t, make_float2(-1.114143632e-03f, -1.114143632e-03f));
  p = __ffma2_rn(p, t, make_float2(9.830119561e-03f, 9.830119561e-03f));
  p = __ffma2_rn(p, t, make_float2(-6.636036497e-02f, -6.636036497e-02f));
  p = __ffma2_rn(p, t, make_float2(3.989074288e-01f, 3.989074288e-01f));
  float2 c;
  asm("fma.rn.ftz.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(c.x) : "f"(x.x), "f"(p.x));   // FFMA.SAT: clamp to [0, 1] for free
  asm("fma.rn.ftz.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(c.y) : "f"(x.y), "f"(p.y));
  return c;
}
__device__ __forceinline__ float2 gelu2(const float2 x) {
  float2 t;
  const float2 c = gelu_phi2(x, t);
  return __fmul2_rn(x, c);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// gelu'(x) = Phi(x) + x * phi(x),  phi(x) = exp(-x^2/2) / sqrt(2 pi): the polynomial CDF plus one MUFU.EX2 per element
__device__ __forceinline__ float2 gelu_grad2(const float2 x) {
  float2 t;
  const float2 c = gelu_phi2(x, t);
  const float2 a = __fmul2_rn(t, make_float2(-0.72134752044448170f, -0.72134752044448170f));   // -x^2/2 * log2(e)
  float2 e;
  e.x = ex2_approx(a.x);
  e.y = ex2_approx(a.y);
  const float2 xe = __fmul2_rn(x, e);
  return __ffma2_rn(xe, make_float2(0.39894228040143268f, 0.39894228040143268f), c);
}
// packed sigmoid(k * x): FMUL2, 2 x MUFU.EX2, FADD2, 2 x MUFU.RCP (same approx units as the scalar sigmoid_f)
__device__ __forceinline__ float2 sigmoid2_scaled(const float2 x, const float k) {
  const float kk = -k * 1.4426950408889634f;
  const float2 a = __fmul2_rn(x, make_float2(kk, kk));
  float2 e;
  e.x = ex2_approx(a.x);
  e.y = ex2_approx(a.y);
  const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
  return make_float2(rcp_approx(d.x), rcp_approx(d.y));
}
// QuickGELU x * sigmoid(1.702 x) (CLIP, cloob.py:179-181) and its derivative s * (1 + 1.702 x (1 - s)), two elements at a time
__device__ __forceinline__ float2 quick_gelu2(const float2 x) { return __fmul2_rn(x, sigmoid2_scaled(x, 1.702f)); }
__device__ __forceinline__ float2 quick_gelu_grad2(const float2 x) {
  const float2 s = sigmoid2_scaled(x, 1.702f);
  const float2 u = __fmul2_rn(x, make_float2(1.702f, 1.702f));
  const float2 w = __ffma2_rn(make_float2(-u.x, -u.y), s, u);      // u * (1 - s)
  return __ffma2_rn(s, w, s);
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_approx(1.0f + __expf(-x)); }
__device__ __forceinline__ float quick_gelu_f(float x) { return x * sigmoid_f(1.702f * x); }
__device__ __forceinline__ float quick_gelu_grad_f(float x) {
  const float s = sigmoid_f(1.702f * x);
  return s + 1.702f * x * s * (1.0f - s);
}
// sigmoid through ONE MUFU op: sigma(x) = 0.5 + 0.5 tanh(x / 2) (tanh.approx.f32, relative error 2^-11 -> |sigma error| <= 2.5e-4).
// Used by the GroupNorm + swish kernels, which are MUFU / issue bound with the two-op (EX2 + RCP) form; the error is 16x below
// the bf16 rounding of the value that is stored.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast_f(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float swish_fast_f(float x) { return x * sigmoid_fast_f(x); }
__device__ __forceinline__ float swish_grad_fast_f(float x) {
  const float s = sigmoid_fast_f(x);
  return s * fmaf(x, 1.0f - s, 1.0f);
}
__device__ __forceinline__ float swish_f(float x) { return x * sigmoid_f(x); }
__device__ __forceinline__ float swish_grad_f(float x) {
  const float s = sigmoid_f(x);
  return s * (1.0f + x * (1.0f - s));
}

}  // namespace ffvc
