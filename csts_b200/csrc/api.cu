// Library-level entry points: version, error reporting, device query, GEMM dispatch.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "gemm.h"

#include <atomic>

namespace {
thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};
}

void csts_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int csts_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    csts_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 4;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int csts_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

extern "C" {

int csts_version(void) { return 100; }

// number of kernels this library has launched since load (or since the last reset)
long long csts_launch_count(int reset) {
  long long n = g_launches.load(std::memory_order_relaxed);
  if (reset) g_launches.store(0, std::memory_order_relaxed);
  return n;
}

// copies the calling thread's last error message; returns its length
int csts_last_error(char* buf, int len) {
  if (!buf || len <= 0) return (int)strlen(g_err);
  strncpy(buf, g_err, (size_t)len - 1);
  buf[len - 1] = 0;
  return (int)strlen(buf);
}

// 0 when the current device is sm_100 (B200); non-zero with a message otherwise.  No other
// architecture is supported and there is no fallback.
int csts_check_device(void) {
  int dev = 0, major = 0, minor = 0;
  CSTS_CUDA(cudaGetDevice(&dev));
  CSTS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CSTS_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  CSTS_REQUIRE(major == 10, "libcsts_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
  return 0;
}

// which kernel csts_gemm would run for this problem: 2 = tcgen05, 1 = mma.sync
int csts_gemm_backend(const csts_gemm_args* a) {
  if (a->backend == 1) return 1;
  if (a->backend == 2) return 2;
  return csts_gemm_tc_supported(*a) ? 2 : 1;
}

int csts_gemm_plan(const csts_gemm_args* a, int* tile_n, int* ctas, int* splits) {
  CSTS_REQUIRE(a && tile_n && ctas && splits, "gemm_plan: null argument");
  return csts_gemm_tc_plan(*a, tile_n, ctas, splits);
}

int csts_gemm(const csts_gemm_args* a, void* stream) {
  CSTS_REQUIRE(a != nullptr, "gemm: null argument block");
  CSTS_REQUIRE((a->a_dtype == CSTS_BF16 || a->a_dtype == CSTS_F16) && (a->b_dtype == CSTS_BF16 || a->b_dtype == CSTS_F16),
               "gemm: a_dtype / b_dtype must be 1 (bf16) or 2 (f16), got %d / %d", a->a_dtype, a->b_dtype);
  CSTS_REQUIRE(a->c_dtype >= 0 && a->c_dtype <= 2, "gemm: c_dtype %d", a->c_dtype);
  if (a->Z) CSTS_REQUIRE(a->z_dtype == CSTS_BF16 || a->z_dtype == CSTS_F16, "gemm: z_dtype must be 1 (bf16) or 2 (f16)");
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = a->backend == 2 || (a->backend != 1 && csts_gemm_tc_supported(*a));
  if (a->rowsum) {
    CSTS_REQUIRE(!a->a_kmajor && a->batch1 * a->batch2 == 1, "gemm: rowsum needs an MN-major A operand and a single batch");
    const bool fused = tc && !a->b_kmajor && a->c_dtype == 0 && a->act == 0;   // the (MN, MN) f32 kernels carry the extra MMA
    if (!fused) {                                     // same result from a separate pass over A (K rows x M columns)
      csts_gemm_args b = *a;
      b.rowsum = nullptr;
      int rc = tc ? csts_gemm_tc_launch(b, st) : csts_gemm_mma_launch(b, st);
      if (rc) return rc;
      return csts_colsum(a->A, a->a_dtype, a->rowsum, a->K, a->M, a->lda, stream);
    }
  }
  return tc ? csts_gemm_tc_launch(*a, st) : csts_gemm_mma_launch(*a, st);
}

}  // extern "C"
