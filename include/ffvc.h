/* libffvc_sm100.so — C ABI of the Blackwell-native feed-forward VQGAN-CLIP train step.
 *
 * The reference (mehdidc/feed_forward_vqgan_clip) has no FFI of its own: its hot path is the Python
 * call surface train() touches (main.py:715-837).  Every entry point below replaces the device work
 * of one piece of that surface; the reference site is cited next to each declaration.  The Python host
 * side (feed_forward_vqgan_clip_b200/*.py) mirrors the reference's module interface and calls only these
 * functions through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise
 *   - caller owns every buffer (outputs, saved-for-backward, workspaces); the library never allocates
 *     tensors the caller sees
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no hidden syncs,
 *     CUDA-graph capturable
 *   - return 0 (FFVC_OK) on success, negative ffvc_status otherwise; ffvc_last_error() returns a
 *     thread-local message.  No exceptions cross the ABI.
 *   - activations are bf16 (NHWC / token-major), accumulation and normalisation statistics fp32,
 *     trainable parameters / gradients / Adam state fp32.
 *
 * Shape contracts (a call outside them returns FFVC_ERR_ARG / FFVC_ERR_UNSUPPORTED and launches nothing; tests/abi_model.py asserts
 * the same conditions so that host code is checked against them without a GPU):
 *   ffvc_gemm            operands travel by TMA: 16-byte aligned bases, every row pitch / offset / batch stride a multiple of 8
 *                        elements; split-K needs the fp32 atomic output; block_n in {32, 64, 128, 256}, tile_m in {128, 256};
 *                        CONV3X3: see ffvc_gemm_params
 *   ffvc_conv3x3_halo*   w % 128 == 0, even h, cin % 64 == 0, cout <= 128 (the _gn / _gnbwd forms: cout == 128, bf16 out)
 *   ffvc_layernorm_*     D % 8 == 0, D <= 2048; dgamma and dbeta both or neither; _bwd_sums: rows % rowsum_T == 0
 *   ffvc_groupnorm_*     C % 8 == 0, C % G == 0, (C / 8) divides 256 (512 for the fused forms); backward: C / G in {1, 2, 4} or a
 *                        multiple of 8
 *   ffvc_vq_nearest      ncodes % 4 == 0, C in {64, 256};   ffvc_vq_nearest_tc: (3 * C) % 8 == 0, P < 2^31
 *   ffvc_mha_small_*     head_dim == 64, 1 <= T <= 64;   ffvc_vitgan_attn_*: 1 <= T <= 32, dh <= 256
 *   ffvc_upsample2x_*, ffvc_maxpool2x2_* (also even H, W), ffvc_conv3x3_cin3 (COUT <= 512): channels % 8 == 0
 *   ffvc_softmax_*       ld >= n;   ffvc_cutout_final_*: P % patch == 0;   ffvc_diversity_tap: C <= 512;   ffvc_tv_loss: H, W >= 2
 */
#ifndef FFVC_H_
#define FFVC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FFVC_OK = 0,
  FFVC_ERR_ARG = -1,
  FFVC_ERR_CUDA = -2,
  FFVC_ERR_UNSUPPORTED = -3
} ffvc_status;

/* operand storage modes for ffvc_gemm */
enum { FFVC_OP_KMAJOR = 0, FFVC_OP_MNMAJOR = 1, FFVC_OP_CONV3X3 = 2 };
/* role of an operand's 3rd (batch) dimension */
enum { FFVC_ROLE_BROADCAST = 0, FFVC_ROLE_OUT_BATCH = 1, FFVC_ROLE_K_SEGMENT = 2 };
/* activations */
enum { FFVC_ACT_NONE = 0, FFVC_ACT_GELU = 1, FFVC_ACT_QUICKGELU = 2, FFVC_ACT_SWISH = 3, FFVC_ACT_RELU = 4 };

const char* ffvc_last_error(void);
/* library / build information: returns the sm arch the kernels were compiled for (100). */
int ffvc_arch(void);
/* number of kernels this library has launched since the last reset (bench.py "gpu_launches"). */
long long ffvc_launch_count(void);
void ffvc_reset_launch_count(void);
/* kernel-selection switches (A/B measurement of alternative kernels for the same op; results are identical up to
 * summation order).  Names: "ln_fwd_v2" / "ln_bwd_v2" (column-owning LayerNorm kernels; 1 = 4 rows in flight per CTA, 2 = 8 rows fwd / 2 rows bwd),
 * "pool_v2" (row-mapped cutout-pool backward), "gn_ring" (cp.async rings in the single-kernel GroupNorm forms),
 * "halo_epi16" (16 epilogue warps in the halo conv's GroupNorm-statistics forms: 1 = backward statistics, 2 = forward too).
 * Initial values come from the environment variable FFVC_OPTS="name=0|1,...".  Returns the previous value, -1 if the
 * name is unknown. */
int ffvc_set_option(const char* name, int value);
int ffvc_get_option(const char* name);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05.mma, TMEM accumulators, TMA operand staging).
 *   out[b][m][n] (+)= epilogue( alpha * sum_{seg,k} A[m][k] * B[n][k] )
 * Replaces: nn.Linear / Conv1d(k=1) of the mappers (mlp_mixer_pytorch.py:16-23,32,76-78),
 *           Conv2d 3x3 / 1x1 of taming's VQGAN Decoder (call site main.py:142),
 *           the Linear layers and 32x32/32 patch-embed conv of the CLIP ViT (cloob.py:188-196,224,249),
 *           and their autograd backward (dgrad / wgrad) (main.py:832).
 * epilogue: v = alpha*acc; v += bias (per column / per row); pre_out = v; v = act(v);
 *           v *= act'(aux) (mul_mode); v += res; out = v  (bf16 store, fp32 store or fp32 atomic add)
 */
typedef struct {
  const void* a;           /* bf16 */
  const void* b;           /* bf16 */
  int a_mode, b_mode;      /* FFVC_OP_* (CONV3X3 only for A) */
  int64_t a_ld, b_ld;      /* elements between consecutive rows (K-major: rows of M/N; MN-major: rows of K) */
  int a_batch_role, b_batch_role; /* FFVC_ROLE_* */
  int64_t a_batch_stride, b_batch_stride; /* elements (outer batch / k-segment stride) */
  int64_t a_batch_stride_inner, b_batch_stride_inner; /* elements (inner batch stride; used when batch_inner > 1) */
  int M, N, K;             /* K = contraction length per segment */
  int batch;               /* output batches (>=1) = outer * inner */
  int batch_inner;         /* inner batch count (0/1 = none): b = b_outer * batch_inner + b_inner */
  int k_segs;              /* extra contraction over operand dim 2 (>=1) */
  int splits;              /* split-K (requires out_fp32 && atomic) */
  int block_n;             /* 0 = auto, else 32/64/128/256 */
  int tile_m;              /* 0 = auto, else 128/256 (256 needs block_n <= 128) */
  int two_cta;             /* 0 = auto, 1 = force the CTA-pair (cta_group::2) kernel, -1 = never */
  int epi_warps;           /* 0 = auto (16 for activation epilogues, compile-time epilogue where one exists),
                              8 = force the 8-warp epilogue, 16 = 16-warp epilogue with run-time flags only */
  /* CONV3X3 (A is NHWC [conv_n][conv_h][conv_w][conv_c], pad 1, stride 1; M = n*h*w; K = 9*c).  Shape contract: c % 64 == 0, and a
   * 128-pixel tile (one TMA box of whole rows) must tile one image: w % 128 == 0, or 128 % w == 0 and h % (128 / w) == 0 — i.e.
   * images of at least 128 pixels (16 x 8 upwards); smaller ones are rejected with FFVC_ERR_ARG. */
  int conv_n, conv_h, conv_w, conv_c;
  /* epilogue */
  void* out;               /* bf16 or fp32 */
  void* pre_out;           /* optional bf16: value before activation */
  const void* aux;         /* optional bf16, same layout as out: multiply by act'(aux) */
  const void* res;         /* optional bf16, same layout as out: residual add */
  const float* bias;       /* optional fp32 */
  int64_t ldc;
  int64_t out_batch_stride;
  int64_t out_batch_stride_inner;
  int out_fp32;
  int atomic;
  int bias_mode;           /* 1 = per column (n), 2 = per row (m) */
  int act;                 /* FFVC_ACT_* */
  int mul_mode;            /* FFVC_ACT_* whose derivative multiplies */
  float alpha;             /* 0 is treated as 1 */
  void* argmin_out;        /* optional uint64 [batch][M]: instead of storing, keep per row the minimum of v = alpha*acc + bias over n
                              as (order-preserving bits of v) << 32 | n  (atomicMin; caller presets 0xFF..); ties -> lowest n */
} ffvc_gemm_params;

int ffvc_gemm(const ffvc_gemm_params* p, void* stream);
/* 1 (default): compile-time-epilogue GEMMs on the CTA-pair kernel with K <= 512 write their bf16 outputs through shared memory
 * and TMA stores; 2: for every K; 0: every epilogue stores straight from registers (A/B measurements, tests). */
int ffvc_gemm_set_tma_store(int on);
/* 1: fp32-atomic GEMMs whose tiles do not fill the SMs evenly run stream-K (every CTA / CTA pair gets the same number of
 * k-blocks of the linearised (tile, k-block) space); 0 (default): split-K as requested by the caller.  Measured equal or
 * slightly slower than split-K on config #2 (those GEMMs are bound by L2 -> SM operand traffic), hence opt-in. */
int ffvc_gemm_set_stream_k(int on);
/* option "gemm_quad" (ffvc_set_option, default 0): the CTA-pair kernel in clusters of 4 — two M-adjacent pair tiles share their B
 * tile through TMA multicast (every CTA fetches half of its B rows and multicasts them to its counterpart in the other pair: 24 KB
 * instead of 32 KB of L2 -> SM traffic per CTA and k-block; a separate kernel instantiation with the run-time epilogue).  Same
 * results bit for bit.  ffvc_gemm_max_quads: how many such
 * clusters the GPU holds at once (-1 before the first quad launch). */
int ffvc_gemm_max_quads(void);

/* 3x3 conv (pad 1, stride 1) with shared-memory halo reuse: NHWC bf16 x [n][h][w][cin], packed weights [cout][9][cin]
 * (tap-major, as for FFVC_OP_CONV3X3), bf16 out [n*h*w][ldc].  Requires w % 128 == 0, even h, cin % 64 == 0, cout <= 128:
 * the wide 128-channel layers of the VQGAN decoder, where the tap-by-tap form is L2-bandwidth bound.
 * Epilogue: + bias[cout] (fp32, optional), act, * act'(aux) (mul_mode), + res (bf16, optional), like ffvc_gemm. */
int ffvc_conv3x3_halo(const void* x, const void* w, void* out, int n, int h, int wd, int cin, int cout, long long ldc,
                      const float* bias, const void* res, const void* aux, int mul_mode, int act, int out_fp32, void* stream);
/* GroupNorm statistics workspace (round 2: REPRODUCIBLE statistics).  Every statistics producer below — ffvc_groupnorm_stats,
 * the first pass of ffvc_groupnorm_bwd, the epilogues of ffvc_conv3x3_halo_gn / _gnbwd — writes one partial sum per CTA / conv
 * tile into its own slot ([N][parts][2][G] doubles behind the [N][G][2] block of folded sums; parts depends on HW only) and a
 * fixed-order fold finishes the job: no atomics, so mean / rstd and the backward sums are bit-identical from run to run and
 * for any batch size or sharding of the batch.  Returns the doubles one workspace must hold for N samples of HW pixels. */
long long ffvc_groupnorm_ws_doubles(int N, int HW, int G);
/* same conv (bias, optional residual, bf16 out, Cout = 128); its epilogue also takes the GroupNorm(32) statistics of the
 * tensor it writes — sum and sum of squares per (image, group) of the bf16-rounded output, one partial per 256-pixel tile in
 * gn_ws (layout above).  ffvc_groupnorm_finalize(gn_ws, ...) folds them (leaving the sums in gn_ws[n][32][2]) and gives the
 * mean / rstd taming's next `Normalize` needs, so ffvc_groupnorm_stats (one more read of the tensor) is skipped. */
int ffvc_conv3x3_halo_gn(const void* x, const void* w, void* out, int n, int h, int wd, int cin, int cout, long long ldc,
                         const float* bias, const void* res, double* gn_ws, void* stream);
int ffvc_groupnorm_finalize(double* ws, float* mean, float* rstd, int N, int HW, int C, int G, float eps, void* stream);
/* The same conv applied to swish(GroupNorm(x)) WITHOUT that tensor ever existing in HBM: x is the raw input of taming's
 * Normalize + nonlinearity in front of the conv (K6 of SURVEY 2.4).  Four transform warps normalise every halo
 * tile in shared memory between the TMA load and the MMAs — y = swish(gamma * (x - mean) * rstd + beta), padding pixels stay
 * zero — so the separate ffvc_groupnorm_apply pass (one read + one write of the tensor) and the re-read by the conv disappear.
 * xf_mean / xf_rstd: [n][xf_groups] statistics of x; xf_gamma / xf_beta: [cin].  gn_ws (optional, Cout = 128): the epilogue
 * also takes the GroupNorm(32) statistics of the output, as ffvc_conv3x3_halo_gn. */
int ffvc_conv3x3_halo_xf(const void* x, const void* w, void* out, int n, int h, int wd, int cin, int cout, long long ldc,
                         const float* bias, const void* res, const float* xf_mean, const float* xf_rstd, const float* xf_gamma,
                         const float* xf_beta, int xf_groups, double* gn_ws, void* stream);
/* dgrad form (no bias): `out` = dy of the Normalize + swish in front of the forward conv (Cout = 128 channels), stored as
 * usual; the epilogue also reads that layer's input gn_x at the same positions and takes the backward statistics
 * sum g, sum g * xhat per (image, group), g = dy * swish'(gamma * xhat + beta) * gamma (per-tile partials in gn_ws, layout
 * above) — the first pass of ffvc_groupnorm_bwd (two more reads of dy and x).  ffvc_groupnorm_bwd_apply folds them (leaving
 * the sums in sums[n][32][2]) and runs the second pass. */
int ffvc_conv3x3_halo_gnbwd(const void* x, const void* w, void* out, int n, int h, int wd, int cin, int cout, long long ldc,
                            const void* res, const void* gn_x, const float* gn_mean, const float* gn_rstd, const float* gn_gamma,
                            const float* gn_beta, double* gn_ws, void* stream);
int ffvc_groupnorm_bwd_apply(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                             const float* beta, double* sums, const void* add, void* dx, int N, int HW, int C, int G,
                             int swish, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (bf16 in/out, fp32 stats).  mlp_mixer_pytorch.py:11,37; cloob.py:170-176.
 * bwd: dx = LN'(dy) (+ add, the residual-path gradient); dgamma/dbeta (fp32, accumulated) optional.   */
int ffvc_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                       long long rows, int D, float eps, void* stream);
int ffvc_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       const void* add, void* dx, float* dgamma, float* dbeta, long long rows, int D, void* stream);
/* same, and the bias gradients the mixer takes from dx (mlp_mixer_pytorch.py:16-23,32-38; autograd, main.py:832) in the
 * same pass: colsum_out[d] += sum_rows dx[row][d] (Linear bias), rowsum_out[t] += sum_{row % rowsum_T == t, d} dx[row][d]
 * (token-mixing Conv1d bias; rows = B*rowsum_T token-major).  Either may be NULL.
 * ws: caller-owned scratch of ffvc_layernorm_bwd_ws_bytes(D, rowsum_T) bytes (per-CTA partial sums, folded by a second tiny
 * kernel instead of grid-deep atomics); NULL selects the unfused kernels. */
int ffvc_layernorm_bwd_sums(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                            const void* add, void* dx, float* dgamma, float* dbeta, float* colsum_out, float* rowsum_out,
                            int rowsum_T, float* ws, long long rows, int D, void* stream);
long long ffvc_layernorm_bwd_ws_bytes(int D, int rowsum_T);

/* GroupNorm(G groups, eps) [+ swish] on NHWC bf16 — taming Normalize + nonlinearity.
 * ws: ffvc_groupnorm_ws_doubles(N, HW, G) doubles of scratch; after the call ws[n][g][2] holds the folded sums
 * ((sum x, sum x^2) / (sum g, sum g xhat)).  bwd gives dx only (frozen affine), optionally + add.                 */
int ffvc_groupnorm_stats(const void* x, double* ws, float* mean, float* rstd, int N, int HW, int C, int G, float eps,
                         void* stream);
int ffvc_groupnorm_apply(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                         void* y, int N, int HW, int C, int G, int swish, void* stream);
int ffvc_groupnorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, double* ws, const void* add, void* dx, int N, int HW, int C, int G, int swish,
                       void* stream);

/* Single-kernel GroupNorm [+ swish]: one persistent grid walks the batch sample by sample, statistics pass and apply
 * pass of a sample back to back so the second read hits L2 (forward 4 B/element of HBM traffic instead of 6, backward 6
 * instead of 10).  Same arithmetic as the two-pass entry points; also writes mean / rstd [N*G] for the backward.
 * ws: ffvc_groupnorm_ws_bytes(N, G) bytes of scratch (2*N*G doubles + N arrival counters), zeroed by the call.       */
long long ffvc_groupnorm_ws_bytes(int N, int G);
int ffvc_groupnorm_set_pipeline(int on);   /* 1 (default): statistics of sample n+1 are issued before the wait for sample n */
int ffvc_groupnorm_fused_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                             double* ws, int N, int HW, int C, int G, int swish, float eps, void* stream);
int ffvc_groupnorm_fused_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                             const float* beta, double* ws, const void* add, void* dx, int N, int HW, int C, int G,
                             int swish, void* stream);

/* nearest-neighbour 2x upsample, NHWC bf16 (taming Upsample); bwd sums the 2x2 block. */
int ffvc_upsample2x_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream);
int ffvc_upsample2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, void* stream);

/* batched transpose in[b][R][Cc] -> out[b][Cc][R] (Rearrange 'b c h w -> b (h w) c', mlp_mixer_pytorch.py:31). */
int ffvc_transpose(const void* in, void* out, int B, int R, int Cc, int in_fp32, int out_fp32, void* stream);

/* row softmax (fp32 scores -> bf16 probabilities) and its backward (VQGAN AttnBlock, CLIP attention).
 * rows are ld elements apart, n valid columns; output padding [n, ld) is zero-filled. */
int ffvc_softmax_fwd(const float* s, void* p, long long rows, int n, int ld, void* stream);
int ffvc_softmax_bwd(const void* p, const float* dp, void* ds, long long rows, int n, int ld, float scale, void* stream);

/* CLIP text transformer pieces (encode_text, cloob.py:525-538): causal row softmax (row r attends to columns 0..r % T),
 * token-embedding gather + positional add (tok int64), and the gather of the EOT-token rows. */
int ffvc_softmax_causal_fwd(const float* s, void* p, long long rows, int T, int ld, void* stream);
int ffvc_embed_tokens(const long long* tok, const float* emb, const float* pos, void* x, long long rows, int T, int W, void* stream);
int ffvc_gather_rows(const void* src, const long long* idx, void* dst, int B, int T, int W, void* stream);

/* bias gradients: db[n] += sum_rows dy[row][n];  db[j] += sum_{b,d} dy[b][j][d]. */
int ffvc_colsum(const void* dy, float* db, long long rows, int n, void* stream);
int ffvc_rowsum(const void* dy, float* db, int B, int J, int D, void* stream);

int ffvc_cast_f32_bf16(const float* x, void* y, long long n, void* stream);
int ffvc_cast_bf16_f32(const void* x, float* y, long long n, void* stream);
int ffvc_add_bf16(const void* a, const void* b, void* y, long long n, void* stream);
int ffvc_rownorm2(const float* x, float* out, int rows, int C, void* stream);
int ffvc_sumsq(const float* x, float* out, long long n, void* stream);

/* clamp_with_grad(z, lo, hi) + vector_quantize (main.py:763,134-138), fp32: idx = argmin_c |z - c|^2, zq = codebook[idx].
 * codeT = codebook transposed [C][ncodes]; cnorm = |code|^2.  zc (optional) = clamped z. */
int ffvc_vq_nearest(const float* z, const float* codebook, const float* codeT, const float* cnorm, int* idx, void* zq_bf16,
                    float* zq_f32, float* zc, long long P, int C, int ncodes, float lo, float hi, void* stream);
/* Same result on the tensor cores: the distance search d(p, c) = |c|^2 - 2 z_p.c runs as ONE tcgen05 GEMM over a 3-way
 * bf16 split of both operands (z = hi + lo, c = hi + lo; K = 3*C: hi.hi + hi.lo + lo.hi, fp32 accumulation: dot-product
 * error ~1e-4, the same order as fp32 rounding of the reference's own |z|^2 + |c|^2 - 2 z.c) with a per-row arg-min epilogue.
 *   ffvc_vq_prepare_codebook: once per (frozen) codebook: csplit [ncodes][3*C] bf16 = [hi | lo | hi], cnorm [ncodes] fp32.
 *   ffvc_vq_nearest_tc: zsplit [P][3*C] bf16 and keys [P] uint64 are caller-provided scratch.                          */
int ffvc_vq_prepare_codebook(const float* codebook, void* csplit_bf16, float* cnorm, int ncodes, int C, void* stream);
int ffvc_vq_nearest_tc(const float* z, const float* codebook, const void* csplit_bf16, const float* cnorm, void* zsplit_bf16,
                       void* keys_u64, int* idx, void* zq_bf16, float* zq_f32, float* zc, long long P, int C, int ncodes,
                       float lo, float hi, void* stream);
/* ClampWithGrad.backward (main.py:126-129). */
int ffvc_clamp_bwd(const float* g, const float* x, float* gx, long long n, float lo, float hi, void* stream);
/* xr = clamp_with_grad((d + 1) / 2, 0, 1) and its backward (main.py:142). */
int ffvc_image_post_fwd(const float* d, float* xr, long long n, void* stream);
int ffvc_image_post_bwd(const float* g, const float* d, float* gd, long long n, void* stream);
/* 3x3 / pad-1 conv with COUT <= 3 output channels (the decoder's conv_out, taming Decoder; call site main.py:142) in two steps:
 * ffvc_gemm computes v[p][tap * COUT + co] = sum_c a[p][c] * w[co][tap][c] for every pixel ONCE (M = N*H*W, N = 32 columns of which
 * 9 * COUT are used, K = Cin, fp32 out, ldc = 32); this entry point adds the nine shifted taps, y[p][co] = bias[co] +
 * sum_tap v[p + off(tap)][tap * COUT + co] (fp32 [N*H*W][COUT]), and, when xr is given, also writes clamp((y + 1) / 2, 0, 1)
 * (ffvc_image_post_fwd).  The implicit-GEMM form read every pixel nine times through the tensor core for a 3-wide N. */
int ffvc_conv_taps_gather(const float* v, const float* bias, float* y, float* xr, int N, int H, int W, int COUT, void* stream);
/* 3x3 conv, Cin = 3 (fp32 NHWC in, [COUT][9][3] fp32 weights, bf16 NHWC out): dgrad of the decoder's conv_out. */
int ffvc_conv3x3_cin3(const float* x, const float* w, void* y, int N, int H, int W, int COUT, void* stream);

/* im2col of a 3-channel NHWC fp32 image for a 3x3 / pad-1 conv: col[p][tap*3+c], 32 bf16 per pixel (k >= 27 zero), so the
 * dgrad of conv_out runs as a tcgen05 GEMM with K = 32. */
int ffvc_im2col3x3_cin3(const float* x, void* col, int N, int H, int W, void* stream);

/* fused optimizer step over a flat fp32 arena: torch.optim.Adam (main.py:591,835) + optional clip_grad_norm_
 * (main.py:693,833-834) + optional CosineAnnealingLR (main.py:702-705,836-837) + optional torch_ema update
 * (main.py:524-525,843-844); refreshes the bf16 shadow the GEMMs read.
 * hyper_dev: DEVICE float[16] (device-resident so a captured CUDA graph sees each step's values):
 *   [0] lr of this step  [1] beta1  [2] beta2  [3] eps  [4] 1-beta1^t  [5] sqrt(1-beta2^t)  [6] grad_scale (1/world)
 *   [7] weight_decay  [8] t  [9] clip max_norm (0 = off)  [10] sum g^2 (caller: ffvc_sumsq(grad, hyper+10) before the tick)
 *   [11] clip coefficient (written by the tick)  [12] base lr  [13] cosine T_max (0 = constant lr)  [14] cosine eta_min
 *   [15] EMA decay (0 = off)
 * ffvc_adam_tick increments t and refreshes slots 4, 5, (0), (11) on the device; ffvc_adam_step[_ema] applies the update. */
int ffvc_adam_tick(float* hyper_dev, void* stream);
int ffvc_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, const float* hyper_dev,
                   void* stream);
int ffvc_adam_step_ema(float* p, const float* g, float* m, float* v, void* shadow_bf16, float* ema, long long n,
                       const float* hyper_dev, void* stream);

/* CLIP ViT multi-head attention for short sequences (T <= 64, head_dim 64); qkv [N][T][3W] bf16 (cloob.py:198-200). */
int ffvc_mha_small_fwd(const void* qkv, void* out, int N, int T, int heads, int head_dim, float scale, void* stream);
int ffvc_mha_small_bwd(const void* qkv, const void* dout, void* dqkv, int N, int T, int heads, int head_dim, float scale,
                       void* stream);
/* Tiled ("flash") attention for any sequence length, head_dim 64: the x-transformer mapper's causal attention at 1024 tokens
 * (transformer.py:11-20; x-transformers Attention) and ViTs beyond 64 tokens.  Same layouts as ffvc_mha_small_*; scores and
 * probabilities stay on the SM.  lse [N][heads][T] fp32: log2-domain log-sum-exp of the scaled (and masked) scores, written by
 * the forward, read by the backward (which recomputes the probabilities); delta_ws: [N][heads][T] fp32 scratch;
 * out (the forward's result) is read by the backward.  causal != 0: token i attends tokens j <= i. */
int ffvc_mha_flash_fwd(const void* qkv, void* out, float* lse, int N, int T, int heads, int head_dim, float scale, int causal,
                       void* stream);
int ffvc_mha_flash_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_ws, void* dqkv, int N,
                       int T, int heads, int head_dim, float scale, int causal, void* stream);

/* CLIP ViT token assembly: x[n][0] = cls + pos[0], x[n][1+p] = pe[n][p] + pos[1+p] (cloob.py:240-243); strided row copy. */
int ffvc_clip_assemble(const void* pe, const float* cls, const float* pos, void* x, int N, int T, int W, void* stream);
int ffvc_copy_rows(const void* src, void* dst, long long rows, int D, long long src_stride, long long dst_stride, void* stream);

/* MakeCutouts (main.py:212-229) on NHWC fp32 3-channel images, explicit augmentation parameters. */
int ffvc_cutout_pool_fwd(const float* x, float* y, int B, int H, int W, int P, void* stream);
/* Backward buffers between the three stages (dcut1 of ffvc_cutout_final_bwd, dout / din of ffvc_cutout_warp_bwd, dy of
 * ffvc_cutout_pool_bwd with dy_fixed != 0) are 64-bit FIXED-POINT accumulators, value x 2^40: the bilinear stages scatter, and
 * integer addition commutes, so d(image) is reproducible bit for bit (float atomics are not; resolution 9e-13, range +-8e6). */
int ffvc_cutout_pool_bwd(const float* x, const void* dy, float* dx, int B, int H, int W, int P, int dy_fixed, void* stream);
int ffvc_cutout_warp_fwd(const float* in, const float* hinv, float* out, int N, int n_src, int P, int border, void* stream);
int ffvc_cutout_warp_bwd(const long long* dout, const float* hinv, long long* din, int N, int n_src, int P, int border, void* stream);
/* erase: DEVICE int[4] (x0,y0,x1,y1); mean, std (3 floats each) are HOST pointers. patches: [N][(P/patch)^2][3*patch^2] bf16. */
int ffvc_cutout_final_fwd(const float* cut1, const float* hinv, const float* sat, const float* hue, const float* noise,
                          const float* facs, const int* erase, const float* mean, const float* std_, void* patches,
                          float* img_out, int N, int P, int patch, void* stream);
int ffvc_cutout_final_bwd(const float* cut1, const float* hinv, const float* sat, const float* hue, const int* erase,
                          const float* mean, const float* std_, const void* dpatches, long long* dcut1, int N, int P, int patch,
                          void* stream);

/* spherical distance loss fwd+bwd (main.py:801-811): loss_out (1 float), dembed fp32 and/or bf16 [N][D]. */
int ffvc_spherical_loss(const float* embed, const float* target, float* loss_out, float* dembed, void* dembed_bf16, int N,
                        int B, int D, float coef, void* stream);
/* same with the reference's `input_loss` term (main.py:812-824): + coef2 * the same distance to a second target [B][D] (the
 * source embeddings), one pass, one gradient */
int ffvc_spherical_loss2(const float* embed, const float* target, const float* target2, float* loss_out, float* dembed,
                         void* dembed_bf16, int N, int B, int D, float coef, float coef2, void* stream);
/* y[r] = x[r] / max(|x[r]|, 1e-12) — F.normalize(inp_feats, dim=1) of `normalize_input` (main.py:734-735); fp32 [rows][D] */
int ffvc_normalize_rows(const float* x, float* y, int rows, int D, void* stream);

/* total-variation loss (main.py:423-428) on NHWC fp32, fwd+bwd: loss_accum += coef*tv, dimg_accum += coef*d(tv)/d(img);
 * axpy for the z-L2 term's gradient (main.py:758-762). */
int ffvc_tv_loss(const float* img, float* loss_accum, float* dimg_accum, int B, int H, int W, int C, float coef, void* stream);
int ffvc_axpy_f32(const float* x, float* y, float a, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LPIPS-VGG16 diversity term (main.py:532-537,776-791): 2x2 max-pool on NHWC bf16, per-channel image normalisation on NHWC
 * fp32 x 3 (mean / std are HOST float[3]; the backward ACCUMULATES into dx), and the fused normalize_tensor + pairwise
 * squared difference of the `repeat` samples of each prompt for one VGG tap: feats [R*B][HW][C] bf16 (sample r*B + b),
 * loss_accum += scale * div, dfeat = scale * d(div)/d(feats). */
int ffvc_maxpool2x2_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream);
int ffvc_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, void* stream);
int ffvc_normalize3_fwd(const float* x, float* y, long long n, const float* mean, const float* std_, void* stream);
int ffvc_normalize3_bwd(const float* dy, float* dx_accum, long long n, const float* std_, void* stream);
int ffvc_relu_mask(const void* g, const void* post, void* out, long long n, void* stream);
int ffvc_diversity_tap(const void* feats, float* loss_accum, void* dfeat, int R, int B, int HW, int C, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * VitGAN mapper pieces (vitgan.py:8-21,44-97,254-260).
 * SLN modulation: s = w * (gamma * n + beta), n = LayerNorm(hl) from ffvc_layernorm_fwd; gamma / beta are DEVICE scalars.
 * bwd: dn (bf16), dw_acc += (fp32, shared by every SLN that consumes w), dgamma / dbeta += (fp32). */
int ffvc_sln_mod_fwd(const void* n, const void* w, const float* gamma, const float* beta, void* s, long long total, void* stream);
int ffvc_sln_mod_bwd(const void* ds, const void* n, const void* w, const float* gamma, const float* beta, void* dn, float* dw_acc,
                     float* dgamma, float* dbeta, long long total, void* stream);
/* attention with the reference's '(d k h)' interleaved projection layout: qkv [B][T][ld_qkv], column d*3H + k*H + h;
 * out [B][T][ld_out], column h*dh + d; probs [B*H][T][T] fp32 saved for the backward.  T <= 32, dh <= 256. */
int ffvc_vitgan_attn_fwd(const void* qkv, void* out, float* probs, int B, int T, int H, int dh, int ld_qkv, int ld_out,
                         float scale, void* stream);
int ffvc_vitgan_attn_bwd(const void* qkv, const float* probs, const void* dout, void* dqkv, int B, int T, int H, int dh, int ld_qkv,
                         int ld_out, float scale, void* stream);
/* fp32 [rows][cols] -> bf16 [rows][ld] (zero padded): gives 1020-wide operands a TMA-legal 16-byte row pitch. */
int ffvc_cast_f32_bf16_pitched(const float* src, void* dst, int rows, int cols, int ld, void* stream);
/* y[b][:] = bf16(x[:]) for b < B (pos_emb1D broadcast over the batch). */
int ffvc_broadcast_rows(const float* x, void* y, int B, long long n, void* stream);

/* SimpleGenerator (vitgan.py:262-305; build_model model_type "simple_vitgan", main.py:469-478): the same attention at
 * T = size*size tokens runs as batched ffvc_gemm launches, which need head-contiguous, 16-byte aligned q | k | v.  The projection
 * WEIGHTS are re-packed instead of the activations, the head dimension padded from dh to dhp (a multiple of 8):
 *   pack_qkv:   wp[(k*H + h)*dhp + d][c] = bf16(w[d*3H + k*H + h][c]) for d < dh, 0 in the padding   (to_qkv.weight [3*H*dh][D], '(d k h)' vitgan.py:82)
 *   unpack_qkv: dw[d*3H + k*H + h][c] += dwp[(k*H + h)*dhp + d][c]                                  (its fp32 gradient)
 *   pack_out:   wp[r][h*dhp + d] = bf16(w[r][h*dh + d]), 0 in the padding                           (w_out.weight [D][H*dh], vitgan.py:67,96-97)
 *   unpack_out: dw[r][h*dh + d] += dwp[r][h*dhp + d] */
int ffvc_vitgan_pack_qkv_weight(const float* w, void* wp, int H, int dh, int dhp, int D, void* stream);
int ffvc_vitgan_unpack_qkv_wgrad(const float* dwp, float* dw, int H, int dh, int dhp, int D, void* stream);
int ffvc_vitgan_pack_out_weight(const float* w, void* wp, int H, int dh, int dhp, int D, void* stream);
int ffvc_vitgan_unpack_out_wgrad(const float* dwp, float* dw, int H, int dh, int dhp, int D, void* stream);

/* sizeof() of the ABI structs, for binding self-checks. */
int ffvc_sizeof(const char* name);

#ifdef __cplusplus
}
#endif
#endif /* FFVC_H_ */
