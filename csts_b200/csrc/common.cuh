// Shared device/host helpers for libcsts_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;
typedef __half f16;
typedef __half2 f162;

// Storage types (include/csts_b200.h): forward activations and operand copies of the weights are IEEE
// fp16 (10 mantissa bits — their rounding is what limits gradient parity), gradients are bf16 (fp32 range).
enum { CSTS_F32 = 0, CSTS_BF16 = 1, CSTS_F16 = 2 };

// ---- error plumbing (no exception crosses the C ABI) -----------------------------------------
void csts_set_error(const char* fmt, ...);
int csts_check_launch(const char* what);   // cudaGetLastError() -> 0 / error code with message

#define CSTS_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      csts_set_error(__VA_ARGS__);         \
      return 2;                            \
    }                                      \
  } while (0)

#define CSTS_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      csts_set_error("%s failed: %s", #call, cudaGetErrorString(e__));         \
      return 3;                                                                \
    }                                                                          \
  } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
int csts_num_sms();

// ---- programmatic dependent launch --------------------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute and
// calls pdl_wait() before its first global-memory access: the launch latency and CTA dispatch of
// kernel N+1 then overlap the tail of kernel N (the step is ~1100 short launches, so the gaps
// between kernels are a measurable share of it).  griddepcontrol.wait returns only once the
// preceding kernel has completed and flushed, so no data hazard is introduced; launch_dependents
// right after it lets the next kernel's CTAs become resident as soon as this one's are all running.
#ifndef CSTS_PDL_TRIGGER
#define CSTS_PDL_TRIGGER 1
#endif
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#if CSTS_PDL_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers ---------------------------------------------------------------------------
// index decode in 32-bit arithmetic (a 64-bit division costs ~100 instructions): q <- q / d, returns q % d.
// The host side guarantees that the flattened work-item count of such kernels stays below 2^32.
__device__ __forceinline__ uint32_t divmod(uint32_t& q, uint32_t d) {
  uint32_t n = q;
  q = n / d;
  return n - q * d;
}

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

// block-wide sum for blockDim.x <= 1024; `red` is a __shared__ float[33]
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// erf-based GELU (nn.GELU() default, ref slowfast/models/common.py:21) and its derivative.
// erf is evaluated with Abramowitz & Stegun 7.1.26 (|abs error| <= 1.5e-7 — below f32 round-off of the
// surrounding arithmetic and far below the bf16 storage of the result): one reciprocal, one ex2 and six
// FMAs instead of libm's branchy erff; exp(-x^2/2) is shared between erf(x/sqrt2) and the Gaussian pdf.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  // z = |x| sqrt(log2(e) / 2): exp(-x^2/2) = 2^(-z^2); the A&S argument y = |x| / sqrt2 = z / sqrt(log2 e).  The flush-to-
  // zero forms of ex2 / rcp are single MUFU instructions (the default ones wrap them in a denormal-scaling sequence).
  const float z = fabsf(x) * 0.84932180028801904f;
  const float e = ex2_ftz(-z * z);                                 // exp(-x^2/2)
  const float t = rcp_ftz(fmaf(0.3275911f * 0.83255461115769776f, z, 1.f));
  float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  poly = fmaf(t, poly, 0.5f * 1.421413741f);
  poly = fmaf(t, poly, 0.5f * -0.284496736f);
  poly = fmaf(t, poly, 0.5f * 0.254829592f);
  const float half_erf = fmaf(-poly * t, e, 0.5f);                 // erf(|x|/sqrt2) / 2
  cdf = 0.5f + copysignf(half_erf, x);
  pdf = 0.3989422804014327f * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, pdf;
  gelu_parts(x, cdf, pdf);
  return x * cdf;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float cdf, pdf;
  gelu_parts(x, cdf, pdf);
  return fmaf(x, pdf, cdf);
}

// typed element access: T in {float, bf16}
template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

template <> __device__ __forceinline__ float ld_f<f16>(const f16* p) { return __half2float(*p); }
template <> __device__ __forceinline__ void st_f<f16>(f16* p, float v) { *p = __float2half_rn(v); }

// 4-wide vector access (16 B for float, 8 B for bf16 / f16); pointer must be aligned accordingly
template <typename T> __device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<bf16>(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  bf162 a = *reinterpret_cast<bf162*>(&t.x), b = *reinterpret_cast<bf162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <typename T> __device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<bf16>(bf16* p, const float (&v)[4]) {
  bf162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

template <> __device__ __forceinline__ void ld4<f16>(const f16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  float2 a = __half22float2(*reinterpret_cast<f162*>(&t.x)), b = __half22float2(*reinterpret_cast<f162*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <> __device__ __forceinline__ void st4<f16>(f16* p, const float (&v)[4]) {
  f162 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  bf162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint32_t pack_f162(float a, float b) {
  f162 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// two 16-bit values in one register <-> two floats, for T in {bf16, f16}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<bf16>(float a, float b) { return pack_bf162(a, b); }
template <> __device__ __forceinline__ uint32_t pack2<f16>(float a, float b) { return pack_f162(a, b); }
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<bf16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack2<f16>(uint32_t w) { return __half22float2(*reinterpret_cast<f162*>(&w)); }

// Mixed-precision FMA of sm_100 (PTX fma.rn.f32.{f16,bf16} = SASS FHFMA): a0 += x.lo * w.lo, a1 += x.hi * w.hi with x, w packed
// 16-bit pairs and f32 accumulators.  The 16-bit operands are read straight from the halves of a 32-bit register (R.H0 / R.H1),
// so a kernel that multiplies two 16-bit tensors needs no unpack instructions at all; the product of two 11-bit (8-bit)
// significands is exact in f32, so the result is bit-identical to converting both operands first.
template <typename T> __device__ __forceinline__ void fhfma2(float& a0, float& a1, uint32_t x, uint32_t w);
template <> __device__ __forceinline__ void fhfma2<bf16>(float& a0, float& a1, uint32_t x, uint32_t w) {
  asm("{\n.reg .b16 x0, x1, w0, w1;\nmov.b32 {x0, x1}, %2;\nmov.b32 {w0, w1}, %3;\n"
      "fma.rn.f32.bf16 %0, x0, w0, %0;\nfma.rn.f32.bf16 %1, x1, w1, %1;\n}\n"
      : "+f"(a0), "+f"(a1)
      : "r"(x), "r"(w));
}
template <> __device__ __forceinline__ void fhfma2<f16>(float& a0, float& a1, uint32_t x, uint32_t w) {
  asm("{\n.reg .b16 x0, x1, w0, w1;\nmov.b32 {x0, x1}, %2;\nmov.b32 {w0, w1}, %3;\n"
      "fma.rn.f32.f16 %0, x0, w0, %0;\nfma.rn.f32.f16 %1, x1, w1, %1;\n}\n"
      : "+f"(a0), "+f"(a1)
      : "r"(x), "r"(w));
}
