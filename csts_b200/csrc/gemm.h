// Internal launch entry points of the two GEMM kernels.  The argument block (csts_gemm_args) is part
// of the public C ABI: include/csts_b200.h.
#pragma once
#include <cuda_runtime.h>

#include "../../include/csts_b200.h"

int csts_gemm_mma_launch(const csts_gemm_args& a, cudaStream_t stream);
int csts_gemm_tc_launch(const csts_gemm_args& a, cudaStream_t stream);
bool csts_gemm_tc_supported(const csts_gemm_args& a);
int csts_gemm_tc_plan(const csts_gemm_args& a, int* bn, int* ctas, int* splits);
