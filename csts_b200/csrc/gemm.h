// GEMM argument block shared by the tcgen05 kernel (gemm_tc.cu) and the generic strided/batched
// mma.sync kernel (gemm_mma.cu).  Mirrored field-for-field by include/csts_b200.h.
#pragma once
#include <stdint.h>

extern "C" {
typedef struct csts_gemm_args {
  const void* A;         // bf16
  const void* B;         // bf16
  void* C;               // f32 or bf16 (c_dtype)
  void* Z;               // bf16, same shape/ld as C (ldz): act==1 -> written with the pre-activation,
                         //                                  act==2 -> read (multiply by gelu'(Z))
  const float* bias;     // [N] or NULL
  const float* residual; // f32 [rows, N] (ldr) or NULL; row = m % res_mod when res_mod > 0
  const float* row_scale;// [ceil(M / rows_per_scale)] or NULL: result row m is multiplied by
                         // row_scale[m / rows_per_scale] before the residual is added (DropPath)
  int64_t lda, ldb, ldc, ldz, ldr;
  int64_t sA1, sA2, sB1, sB2, sC1, sC2;  // batch strides (elements): z -> (z / batch2, z % batch2)
  int32_t M, N, K;
  int32_t batch1, batch2;
  int32_t a_kmajor;      // 1: A[m*lda + k]   0: A[k*lda + m]
  int32_t b_kmajor;      // 1: B[n*ldb + k]   0: B[k*ldb + n]
  int32_t c_dtype;       // 0 f32, 1 bf16
  int32_t act;           // 0 none, 1 GELU(erf), 2 times GELU'(Z)
  int32_t accumulate;    // C += result
  int32_t res_mod;
  int32_t split_k;       // > 1: partial sums combined with f32 atomics (C must be f32)
  float alpha;
  int32_t backend;       // 0 auto, 1 mma.sync, 2 tcgen05
  int32_t rows_per_scale;
} csts_gemm_args;
}

int csts_gemm_mma_launch(const csts_gemm_args& a, cudaStream_t stream);
int csts_gemm_tc_launch(const csts_gemm_args& a, cudaStream_t stream);
bool csts_gemm_tc_supported(const csts_gemm_args& a);
