/* libln3b200 -- C ABI of the B200-native LN3Diff generation hot path.
 *
 * Drop-in boundary: every entry point takes a plain-C argument struct
 * of raw device pointers, explicit sizes/strides and enums, plus the CUDA stream as `void*`
 * (a cudaStream_t).  No torch types, no allocation, no retained pointers, no host
 * synchronisation: callers own every buffer (including workspaces).  All entry points return
 * LN3_OK (0) or a negative LN3_E* code; ln3_last_error() returns the thread-local message.
 * There is deliberately no CPU fallback: on a box without an sm_100 GPU every compute call
 * fails with LN3_ECUDA.
 *
 * Each entry point cites the reference code (NIRVANALAN/LN3Diff, paths relative to the
 * reference root) whose device work it replaces.  The reference has no FFI of its own -- it is
 * pure PyTorch -- so the "binding" is the ctypes stub in ln3diff_b200/_lib.py, mirrored for a
 * maintainer in INTEGRATION.md.
 */
#ifndef LN3B200_H_
#define LN3B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LN3_ABI_VERSION 1

#define LN3_OK 0
#define LN3_EINVAL (-1)       /* bad shape / alignment / enum */
#define LN3_ECUDA (-2)        /* CUDA runtime or driver error (includes: no GPU) */
#define LN3_EUNSUPPORTED (-3) /* configuration outside what the kernels implement */

int ln3_abi_version(void);
const char* ln3_last_error(void);
/* Number of kernels this library has launched in this process (for bench.py gpu_launches). */
unsigned long long ln3_launch_count(void);

/* ------------------------------------------------------------------ GEMM (tcgen05 + TMA)
 * out = epilogue(A[M,K] . W[N,K]^T): replaces every nn.Linear on the path
 *   dit/dit_models_xformers.py:231-323 (adaLN_modulation, FusedMLP), vit/vision_transformer.py:
 *   106-124 (qkv, proj), ldm/modules/attention.py:245-307 (to_q/k/v/out), dit/dit_decoder.py.
 * A, W bf16 row-major (K contiguous); fp32 accumulation in TMEM.
 * Epilogue: + bias[N] (fp32, optional) -> activation -> one of
 *   LN3_OUT_BF16       out bf16 [M, ldo]
 *   LN3_OUT_F32        out f32  [M, ldo]
 *   LN3_OUT_RESID_F32  out f32 residual stream updated in place:
 *                      out[m,n] += gate[(m / gate_rows) * gate_ld + n] * val   (gate NULL -> 1)
 *                      and, if out2 != NULL, out2 (bf16 [M, ldo2]) receives the updated row
 *                      (the un-normalised cross-attention query input of TextCondDiTBlock).
 * Constraints: K % 64 == 0, N % 128 == 0, 16-byte aligned pointers and leading dimensions.
 */
enum { LN3_ACT_NONE = 0, LN3_ACT_GELU_ERF = 1, LN3_ACT_GELU_TANH = 2, LN3_ACT_SILU = 3 };
enum { LN3_OUT_BF16 = 0, LN3_OUT_F32 = 1, LN3_OUT_RESID_F32 = 2 };

typedef struct ln3_gemm_args {
  const void* A;   /* bf16 [M, lda] */
  const void* W;   /* bf16 [N, ldw] */
  const float* bias;
  void* out;
  void* out2;
  const float* gate;
  int M, N, K;
  long long lda, ldw, ldo, ldo2, gate_ld;
  int gate_rows;
  int act;
  int out_kind;
} ln3_gemm_args;

int ln3_gemm_bf16(const ln3_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LN3B200_H_ */
