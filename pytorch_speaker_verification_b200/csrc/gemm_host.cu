// Host helpers for the tcgen05 GEMM core (tensor-map encoding) and the plain GEMM entry point.
#include "tc_gemm.cuh"
#include "epilogues.cuh"
#include "../../include/svb200.h"
#include <stdio.h>

namespace svb {

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
              uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1, int swizzle) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return SVB_ERR_DRIVER;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * elem_bytes, stride2_elems * elem_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15)) return SVB_ERR_ALIGN;
  const CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "svb: cuTensorMapEncodeTiled failed (%d) elem=%d dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u)\n",
            (int)r, elem_bytes, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
            (unsigned long long)strides[0], (unsigned long long)strides[1], box0, box1);
    return SVB_ERR_DRIVER;
  }
  return SVB_OK;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                   uint64_t stride2_elems, uint32_t box_rows) {
  return make_tmap(out, base, 2, d0, d1, d2, stride1_elems, stride2_elems, 64, box_rows, 3);
}

// Operand descriptor for one GEMM side.  K-major: rows x K, ld = row pitch.  MN-major: K x rows, ld = k-row pitch.
int make_operand_map(CUtensorMap* out, const void* p, int rows, int K, int64_t ld, int mn_major, int box_rows_kmajor) {
  if (mn_major) return make_tmap_bf16(out, p, (uint64_t)rows, (uint64_t)K, 1, (uint64_t)ld, (uint64_t)ld * K, kBK);
  return make_tmap_bf16(out, p, (uint64_t)K, (uint64_t)rows, 1, (uint64_t)ld, (uint64_t)ld * rows, box_rows_kmajor);
}

}  // namespace svb

using namespace svb;

extern "C" int svb_gemm_bf16(const void* const* A, const void* const* B, int nterms, float* C, const float* bias,
                             int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int a_mn, int b_mn,
                             void* stream) {
  if (nterms < 1 || nterms > kMaxTerms) return SVB_ERR_ARG;
  GemmOperands ops;
  memset(&ops, 0, sizeof(ops));
  ops.nterms = nterms; ops.M = M; ops.N = N; ops.K = K;
  for (int t = 0; t < nterms; ++t) {
    int e = make_operand_map(&ops.ta[t], A[t], M, K, lda, a_mn, kBM);
    if (e) return e;
    e = make_operand_map(&ops.tb[t], B[t], N, K, ldb, b_mn, 128);
    if (e) return e;
  }
  EpiStoreF32<128>::Params ep;
  {
    int e = make_store_params<128>(&ep, C, bias, M, N, ldc, 0);
    if (e) return e;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t err;
  if (!a_mn && !b_mn) err = launch_tc_gemm<128, 4, false, false, EpiStoreF32<128>>(ops, ep, s);
  else if (a_mn && b_mn) err = launch_tc_gemm<128, 4, true, true, EpiStoreF32<128>>(ops, ep, s);
  else if (!a_mn && b_mn) err = launch_tc_gemm<128, 4, false, true, EpiStoreF32<128>>(ops, ep, s);
  else err = launch_tc_gemm<128, 4, true, false, EpiStoreF32<128>>(ops, ep, s);
  return err == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}

// CTA-pair variant (tcgen05.mma.cta_group::2, 256 x 256 pair tiles): both operands K-major (mn = 0) or both
// MN-major (mn = 1).
extern "C" int svb_gemm_bf16_2cta(const void* const* A, const void* const* B, int nterms, float* C, const float* bias,
                                  int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int mn, void* stream) {
  if (nterms < 1 || nterms > kMaxTerms) return SVB_ERR_ARG;
  GemmOperands ops;
  memset(&ops, 0, sizeof(ops));
  ops.nterms = nterms; ops.M = M; ops.N = N; ops.K = K;
  for (int t = 0; t < nterms; ++t) {
    int e = make_operand_map(&ops.ta[t], A[t], M, K, lda, mn, kBM);
    if (e) return e;
    e = make_operand_map(&ops.tb[t], B[t], N, K, ldb, mn, 128);     // each CTA of the pair loads 128 of the 256 B rows
    if (e) return e;
  }
  if (mn) {
    EpiStoreF32<256>::Params ep;
    int e = make_store_params<256>(&ep, C, bias, M, N, ldc, 0);
    if (e) return e;
    cudaError_t err = launch_tc_gemm<256, 6, true, true, EpiStoreF32<256>, 8, 1, true>(
        ops, ep, reinterpret_cast<cudaStream_t>(stream));
    if (err != cudaSuccess) fprintf(stderr, "svb_gemm_bf16_2cta: %s\n", cudaGetErrorString(err));
    return err == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
  }
  EpiStoreF32<256>::Params ep;
  {
    int e = make_store_params<256>(&ep, C, bias, M, N, ldc, 0);
    if (e) return e;
  }
  cudaError_t err = launch_tc_gemm<256, 6, false, false, EpiStoreF32<256>, 8, 1, true>(
      ops, ep, reinterpret_cast<cudaStream_t>(stream));
  if (err != cudaSuccess) fprintf(stderr, "svb_gemm_bf16_2cta: %s\n", cudaGetErrorString(err));
  return err == cudaSuccess ? SVB_OK : SVB_ERR_CUDA;
}
