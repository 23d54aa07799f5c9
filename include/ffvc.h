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
enum { FFVC_ACT_NONE = 0, FFVC_ACT_GELU = 1, FFVC_ACT_QUICKGELU = 2, FFVC_ACT_SWISH = 3 };

const char* ffvc_last_error(void);
/* library / build information: returns the sm arch the kernels were compiled for (100). */
int ffvc_arch(void);
/* number of kernels this library has launched since the last reset (bench.py "gpu_launches"). */
long long ffvc_launch_count(void);
void ffvc_reset_launch_count(void);

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
  int64_t a_batch_stride, b_batch_stride; /* elements */
  int M, N, K;             /* K = contraction length per segment */
  int batch;               /* output batches (>=1) */
  int k_segs;              /* extra contraction over operand dim 2 (>=1) */
  int splits;              /* split-K (requires out_fp32 && atomic) */
  int block_n;             /* 0 = auto, else 32/64/128/256 */
  /* CONV3X3 (A is NHWC [conv_n][conv_h][conv_w][conv_c], pad 1, stride 1; M = n*h*w; K = 9*c) */
  int conv_n, conv_h, conv_w, conv_c;
  /* epilogue */
  void* out;               /* bf16 or fp32 */
  void* pre_out;           /* optional bf16: value before activation */
  const void* aux;         /* optional bf16, same layout as out: multiply by act'(aux) */
  const void* res;         /* optional bf16, same layout as out: residual add */
  const float* bias;       /* optional fp32 */
  int64_t ldc;
  int64_t out_batch_stride;
  int out_fp32;
  int atomic;
  int bias_mode;           /* 1 = per column (n), 2 = per row (m) */
  int act;                 /* FFVC_ACT_* */
  int mul_mode;            /* FFVC_ACT_* whose derivative multiplies */
  float alpha;             /* 0 is treated as 1 */
} ffvc_gemm_params;

int ffvc_gemm(const ffvc_gemm_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FFVC_H_ */
