// Optimizer step of the training loop as two multi-tensor launches (tools/train_avgaze_net.py:101-109,
// slowfast/models/optimizer.py:98-104): global gradient L2 norm, then GradScaler-unscale + clip_grad_norm_ +
// AdamW + refresh of the 16-bit operand copy of every weight, in ONE pass over the 188 M parameters
// (reads grad, param, exp_avg, exp_avg_sq once; writes param, exp_avg, exp_avg_sq and the 16-bit copy).
// The reference performs the same arithmetic as: unscale_ (1 pass), clip_grad_norm_ (2 passes), AdamW (1 pass),
// and the next forward's weight casts (1 pass).  HBM-bound: 4 B x (4 reads + 3 writes) + 2 B per parameter.
#include "common.cuh"
#include "optim.h"

namespace {

constexpr int CHUNK = 16384;          // elements per block-chunk (multiple of 4 * 256)

// sum of squares of every gradient -> *out_sq (double atomics: one per block)
__global__ void __launch_bounds__(256) mt_sqnorm_kernel(const csts_mt_tensor* __restrict__ tensors, const int2* __restrict__ chunks,
                                                        double* __restrict__ out_sq) {
  pdl_wait();
  __shared__ float red[33];
  const int2 ch = chunks[blockIdx.x];
  const csts_mt_tensor t = tensors[ch.x];
  const int64_t beg = (int64_t)ch.y * CHUNK;
  const int64_t end = beg + CHUNK < t.numel ? beg + CHUNK : t.numel;
  const float* g = reinterpret_cast<const float*>(t.grad);
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int64_t n4 = (end - beg) / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g + beg);
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 v = __ldg(g4 + i);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t i = beg + n4 * 4 + threadIdx.x; i < end; i += blockDim.x) s += g[i] * g[i];
  } else {
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) s += g[i] * g[i];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out_sq, (double)s);
}

struct AdamArgs {
  const double* total_sq;      // sum of squares of the (still scaled) gradients
  const float* grad_scale;     // GradScaler scale or NULL
  float* found_inf;            // receives 1.0 when the gradient norm is not finite (step skipped), else 0.0; may be NULL
  const float* step;           // number of steps taken so far (device scalar); this step is *step + 1
  const float* lr[2];          // learning rate of parameter group 0 / 1 (device scalars)
  double beta1, beta2;         // doubles: 1 - beta and the bias corrections are formed as torch forms them (in double)
  float eps, max_norm;
};

template <typename T16> __device__ __forceinline__ void store16x4(void* dst, int64_t i, const float (&p)[4]) {
  st4(reinterpret_cast<T16*>(dst) + i, p);
}

__global__ void __launch_bounds__(256) mt_adamw_kernel(const csts_mt_tensor* __restrict__ tensors, const int2* __restrict__ chunks, AdamArgs a) {
  pdl_wait();
  const float inv_scale = a.grad_scale ? 1.f / *a.grad_scale : 1.f;
  const float total_norm = (float)sqrt(*a.total_sq) * inv_scale;
  const bool bad = !isfinite(total_norm);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.found_inf) *a.found_inf = bad ? 1.f : 0.f;
  if (bad) return;                                   // GradScaler semantics: the whole step is skipped
  // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
  float gmul = inv_scale;
  if (a.max_norm > 0.f) gmul *= fminf(a.max_norm / (total_norm + 1e-6f), 1.f);
  const double stepd = (double)*a.step + 1.0;
  const float bc1 = (float)(1.0 - pow(a.beta1, stepd));
  const float bc2_sqrt = (float)sqrt(1.0 - pow(a.beta2, stepd));
  const float b2 = (float)a.beta2, omb1 = (float)(1.0 - a.beta1), omb2 = (float)(1.0 - a.beta2);
  const int2 ch = chunks[blockIdx.x];
  const csts_mt_tensor t = tensors[ch.x];
  const float lr = *a.lr[t.group];
  const float decay = 1.f - lr * t.weight_decay;
  const float step_size = lr / bc1;
  const int64_t beg = (int64_t)ch.y * CHUNK;
  const int64_t end = beg + CHUNK < t.numel ? beg + CHUNK : t.numel;
  float* p = reinterpret_cast<float*>(t.param);
  const float* g = reinterpret_cast<const float*>(t.grad);
  float* m = reinterpret_cast<float*>(t.exp_avg);
  float* v = reinterpret_cast<float*>(t.exp_avg_sq);
  auto update = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gmul;
    pp *= decay;                                     // param -= lr * wd * param
    mm += (gg - mm) * omb1;                          // lerp(exp_avg, grad, 1 - beta1)
    vv = b2 * vv + omb2 * gg * gg;
    const float denom = sqrtf(vv) / bc2_sqrt + a.eps;
    pp -= step_size * (mm / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (t.w16 == nullptr || (reinterpret_cast<uintptr_t>(t.w16) & 7) == 0);
  int64_t tail = beg;
  if (vec) {
    const int64_t n4 = (end - beg) / 4;
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      const int64_t e = beg + 4 * i;
      float pv[4], gv[4], mv[4], vv[4];
      ld4(p + e, pv); ld4(g + e, gv); ld4(m + e, mv); ld4(v + e, vv);
#pragma unroll
      for (int k = 0; k < 4; ++k) update(pv[k], gv[k], mv[k], vv[k]);
      st4(p + e, pv); st4(m + e, mv); st4(v + e, vv);
      if (t.w16) {
        if (t.w16_dtype == CSTS_F16) store16x4<f16>(t.w16, e, pv);
        else store16x4<bf16>(t.w16, e, pv);
      }
    }
    tail = beg + n4 * 4;
  }
  for (int64_t e = tail + threadIdx.x; e < end; e += blockDim.x) {
    float pv = p[e], mv = m[e], vv = v[e];
    update(pv, g[e], mv, vv);
    p[e] = pv; m[e] = mv; v[e] = vv;
    if (t.w16) {
      if (t.w16_dtype == CSTS_F16) reinterpret_cast<f16*>(t.w16)[e] = __float2half_rn(pv);
      else reinterpret_cast<bf16*>(t.w16)[e] = __float2bfloat16_rn(pv);
    }
  }
}

}  // namespace

extern "C" {

int csts_mt_chunk_elems(void) { return CHUNK; }

int csts_grad_sqnorm(const csts_mt_tensor* tensors, const int32_t* chunks, int n_chunks, double* out_sq, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CSTS_CUDA(cudaMemsetAsync(out_sq, 0, sizeof(double), st));
  if (n_chunks == 0) return 0;
  launch_pdl(mt_sqnorm_kernel, dim3(n_chunks), dim3(256), 0, st, tensors, reinterpret_cast<const int2*>(chunks), out_sq);
  return csts_check_launch("mt_sqnorm");
}

int csts_clip_adamw_step(const csts_mt_tensor* tensors, const int32_t* chunks, int n_chunks, const double* total_sq, const float* grad_scale,
                         float* found_inf, const float* step, const float* lr0, const float* lr1, double beta1, double beta2, float eps,
                         float max_norm, void* stream) {
  if (n_chunks == 0) return 0;
  CSTS_REQUIRE(total_sq && step && lr0 && lr1, "clip_adamw: total_sq / step / lr pointers are required");
  AdamArgs a;
  a.total_sq = total_sq; a.grad_scale = grad_scale; a.found_inf = found_inf; a.step = step;
  a.lr[0] = lr0; a.lr[1] = lr1;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm;
  launch_pdl(mt_adamw_kernel, dim3(n_chunks), dim3(256), 0, (cudaStream_t)stream, tensors, reinterpret_cast<const int2*>(chunks), a);
  return csts_check_launch("mt_adamw");
}

}  // extern "C"
