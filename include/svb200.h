/* svb200.h -- C ABI of libsvb200.so, the B200 (sm_100a) GE2E speaker-verification hot path.
 *
 * Plain pointers, sizes and a cudaStream_t (as void*); no torch types.  All pointers are DEVICE pointers
 * unless a parameter is documented as host.  Every function returns SVB_OK (0) or a negative SVB_ERR_*;
 * svb_last_error() gives a message.  Nothing here synchronises the stream or touches the host copy of data.
 * Each entry point names the reference code it replaces (paths relative to
 * hwidong-na/PyTorch_Speaker_Verification).
 */
#ifndef SVB200_H
#define SVB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SVB_OK 0
#define SVB_ERR_ARG (-1)
#define SVB_ERR_CUDA (-2)
#define SVB_ERR_DRIVER (-3)
#define SVB_ERR_ALIGN (-4)
#define SVB_ERR_UNSUPPORTED (-5)

/* Version / build info: returns 100 for sm_100a builds. */
int svb_arch(void);
const char* svb_last_error(void);

/* Generic tensor-core GEMM used by the LSTM: C[M,N] = sum_t A_t[M,K] * B_t[N,K]^T (+ bias[n]).
 * bf16 operands, fp32 accumulate/out.  a_mn/b_mn = 1 when the operand is stored [K, rows] instead of [rows, K].
 * A and B are HOST arrays of nterms DEVICE pointers (nterms <= 3). Replaces the aten::addmm / cuDNN GEMMs that
 * nn.LSTM / nn.Linear issue for speech_embedder_net.py:28,31. */
int svb_gemm_bf16(const void* const* A, const void* const* B, int nterms, float* C, const float* bias, int M, int N,
                  int K, int64_t lda, int64_t ldb, int64_t ldc, int a_mn, int b_mn, void* stream);

#ifdef __cplusplus
}
#endif
#endif
