/*
 * mebt_b200 C ABI — hand-written sm_100a kernels for the MeBT latent-bottleneck transformer hot path.
 *
 * The reference (Ugness/MeBT) is pure PyTorch: it has no FFI layer, so each entry point below names the
 * reference call site (file:line under /root/reference) whose ATen/cuBLAS library calls it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, a non-zero MEBT_ERR_* code otherwise; mebt_last_error()
 *     returns a thread-local message.  No C++ exception crosses this boundary.
 *   - all pointers are DEVICE pointers into caller-owned (torch-owned) storage unless the name says
 *     `host`; inputs are const, outputs pre-allocated by the caller.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and never synchronise.
 *   - bf16 = __nv_bfloat16 bits, row-major; `ld*` are row strides in ELEMENTS.
 *   - there is no CPU fallback: without an sm_100 device every compute entry point fails.
 */
#ifndef MEBT_B200_H_
#define MEBT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  MEBT_OK = 0,
  MEBT_ERR_SHAPE = 1,
  MEBT_ERR_DTYPE = 2,
  MEBT_ERR_WORKSPACE = 3,
  MEBT_ERR_CUDA = 4,
  MEBT_ERR_UNSUPPORTED = 5,
  MEBT_ERR_DEVICE = 6
};

/* ---- runtime ------------------------------------------------------------------------------- */
const char* mebt_version(void);
const char* mebt_last_error(void);
/* 0 iff the current CUDA device is sm_100-class (B200). */
int mebt_device_check(void);

/* ---- K2 / K4 : dense contractions ---------------------------------------------------------- */
enum {
  MEBT_GEMM_GELU = 1,         /* exact erf GELU after bias (gpt.py:152 nn.GELU) */
  MEBT_GEMM_OUT_FP32 = 2,     /* C is float (logits, weight gradients); default bf16 */
  MEBT_GEMM_ACCUMULATE = 4,   /* C += result (fp32 C only; gradient accumulation) */
  MEBT_GEMM_FORCE_BN256 = 16, /* tile-width overrides, for tests and tuning */
  MEBT_GEMM_FORCE_BN128 = 32,
  MEBT_GEMM_FORCE_BN64 = 64
};
/*
 * C[M,N] = act( A * B^T + bias ) + residual, tcgen05/TMEM/TMA.
 *   a_mn_major = 0: A is [M,K] row-major (lda >= K).   1: A is stored [K,M] row-major (lda >= M).
 *   b_mn_major = 0: B is [N,K] row-major — a torch nn.Linear weight.   1: B is stored [K,N].
 * Replaces: nn.Linear forward at mebt/modules/gpt.py:126-128 (query/key/value), :140 (proj),
 * :150-155 (mlp fc1+GELU, fc2), :248 (head, no bias); with MN-major operands the dgrad
 * (dX = dY * W) and wgrad (dW = dY^T * X) GEMMs autograd runs for the same modules.
 * bias: fp32 [N] or NULL.  residual: bf16 [M, ldres] or NULL (gpt.py:184-185 residual adds).
 */
int mebt_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, void* C, int ldc,
                   int M, int N, int K, const float* bias, const void* residual, int ldres, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEBT_B200_H_ */
