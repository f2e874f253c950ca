// Stem, fusion glue and head kernels of CSTS: patch-embed im2col, separable position embedding,
// fusion re-weighting, token mean, and the 1x1x1 classifier fused with the trilinear stem skip.
#include <algorithm>

#include "common.cuh"

namespace {

int grid_for(int64_t work_items, int per_block) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)csts_num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// ------------------------------------------------------------------------------------------------
// im2col for PatchEmbed: Conv3d k(3,7,7) s(2,4,4) p(1,3,3)      ref: stem_helper.py:27-38
// x f32 (B, Cin, T, H, W) -> patches (bf16 / f16) [B*To*Ho*Wo, Kp], column = ((c*3 + kt)*7 + kh)*7 + kw,
// columns >= Cin*147 are zero (Kp is the GEMM-friendly padded width).
// ------------------------------------------------------------------------------------------------
// One thread produces 8 consecutive columns of one patch row (a 16-byte store): the (c, kt, kh, kw) decode is
// done once per chunk and advanced incrementally, all index arithmetic is 32-bit.
template <typename TO>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, TO* __restrict__ out, int B, int Cin, int T, int H,
                                                     int W, int Kp) {
  pdl_wait();
  const int To = T / 2, Ho = H / 4, Wo = W / 4;
  const int K = Cin * 147;
  const uint32_t chunks = Kp / 8;
  const uint32_t total = (uint32_t)B * To * Ho * Wo * chunks;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t q = i;
    const int col0 = (int)divmod(q, chunks) * 8;
    const uint32_t tok = q;
    const int wo = (int)divmod(q, Wo), ho = (int)divmod(q, Ho), to = (int)divmod(q, To);
    const int b = (int)q;
    int kw = col0 % 7, kh = (col0 / 7) % 7, kt = (col0 / 49) % 3, c = col0 / 147;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = 0.f;
      if (col0 + e < K) {
        const int ti = to * 2 + kt - 1, hi = ho * 4 + kh - 3, wi = wo * 4 + kw - 3;
        if (ti >= 0 && ti < T && hi >= 0 && hi < H && wi >= 0 && wi < W)
          v[e] = __ldg(x + ((((int64_t)b * Cin + c) * T + ti) * H + hi) * W + wi);
      }
      if (++kw == 7) { kw = 0; if (++kh == 7) { kh = 0; if (++kt == 3) { kt = 0; ++c; } } }
    }
    uint4 pk = make_uint4(pack2<TO>(v[0], v[1]), pack2<TO>(v[2], v[3]), pack2<TO>(v[4], v[5]), pack2<TO>(v[6], v[7]));
    *reinterpret_cast<uint4*>(out + (int64_t)tok * Kp + col0) = pk;
  }
}

// pos[t*HW + s][c] = spatial[s][c] + temporal[t][c]       ref: custom_multimodal_builder.py:362-365
__global__ void pos_embed_kernel(const float* __restrict__ spatial, const float* __restrict__ temporal, float* __restrict__ pos, int T,
                                 int HW, int C) {
  pdl_wait();
  const int64_t total = (int64_t)T * HW * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t r = i / C;
    int s = (int)(r % HW), t = (int)(r / HW);
    pos[i] = spatial[(int64_t)s * C + c] + temporal[t * C + c];
  }
}
// dspatial[s][c] = sum_{b,t} dY[b][t][s][c];  dtemporal[t][c] += sum_{b,s} dY[b][t][s][c]
__global__ void __launch_bounds__(128) pos_embed_bwd_kernel(const float* __restrict__ dY, float* __restrict__ dspatial,
                                                            float* __restrict__ dtemporal, int B, int T, int HW, int C, int s_per_block) {
  pdl_wait();
  const int c = threadIdx.x;
  if (c >= C) return;
  int s0 = blockIdx.x * s_per_block, s1 = min(HW, s0 + s_per_block);
  float tacc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) tacc[t] = 0.f;
  for (int s = s0; s < s1; ++s) {
    float sacc = 0.f;
    for (int b = 0; b < B; ++b)
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (t < T) {
          float v = dY[(((int64_t)b * T + t) * HW + s) * C + c];
          sacc += v;
          tacc[t] += v;
        }
    dspatial[(int64_t)s * C + c] = sacc;
  }
#pragma unroll
  for (int t = 0; t < 8; ++t)
    if (t < T) atomicAdd(dtemporal + t * C + c, tacc[t]);
}

// ------------------------------------------------------------------------------------------------
// Fusion re-weighting: out[b][t][s][c] = x[b][t][s][c] * w[b][t][c]   ref: custom_multimodal_builder.py:454-461
// w rows are addressed with a batch stride so the (B, 8, C) temporal-fusion output is used in place.
// ------------------------------------------------------------------------------------------------
__global__ void reweight_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out, int B, int T, int S,
                                    int C, int64_t w_sB) {
  pdl_wait();
  const int C4 = C / 4;
  const int64_t total = (int64_t)B * T * S * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4;
    int64_t r = i / C4;
    int t = (int)((r / S) % T);
    int64_t b = r / ((int64_t)S * T);
    float xv[4], wv[4];
    ld4(x + r * C + c, xv);
    ld4(w + b * w_sB + (int64_t)t * C + c, wv);
#pragma unroll
    for (int k = 0; k < 4; ++k) xv[k] *= wv[k];
    st4(out + r * C + c, xv);
  }
}
// dx = dout * w ; dw[b][t][c] = sum_s dout * x        (one thread owns (b, t, 4 channels): no atomics)
__global__ void reweight_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ w,
                                    float* __restrict__ dx, float* __restrict__ dw, int B, int T, int S, int C, int64_t w_sB) {
  pdl_wait();
  const int C4 = C / 4;
  const int64_t total = (int64_t)B * T * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4;
    int64_t bt = i / C4;
    int t = (int)(bt % T);
    int64_t b = bt / T;
    float wv[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
    ld4(w + b * w_sB + (int64_t)t * C + c, wv);
    for (int s = 0; s < S; ++s) {
      int64_t off = (bt * S + s) * C + c;
      float dv[4], xv[4], o[4];
      ld4(dout + off, dv);
      ld4(x + off, xv);
#pragma unroll
      for (int k = 0; k < 4; ++k) { o[k] = dv[k] * wv[k]; acc[k] = fmaf(dv[k], xv[k], acc[k]); }
      if (dx) st4(dx + off, o);
    }
    st4(dw + b * w_sB + (int64_t)t * C + c, acc);
  }
}

// out[b][c] = mean_n x[b][n][c]  (16-bit out: it is the A operand of the NCE projection GEMM)
template <typename TO>
__global__ void token_mean_fwd_kernel(const float* __restrict__ x, TO* __restrict__ out, int B, int N, int C) {
  pdl_wait();
  const int64_t total = (int64_t)B * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t b = i / C;
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc += x[(b * N + n) * C + c];
    st_f(out + i, acc / (float)N);
  }
}
// dx[b][n][c] (+)= dout[b][c] / N
__global__ void token_mean_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int B, int N, int C, int accumulate) {
  pdl_wait();
  const int64_t total = (int64_t)B * N * C;
  const float inv = 1.f / (float)N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t b = i / ((int64_t)N * C);
    float v = dout[b * C + c] * inv;
    dx[i] = accumulate ? dx[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------------------
// Head: logits[b][to][s] = bias + sum_c w[c] * (feat[b][to][s][c] + lerp_t(stem)[b][to][s][c])
// ref: custom_multimodal_builder.py:476-481 (F.interpolate T -> 2T trilinear, align_corners=False,
// then Conv3d(96,1,1)).  One warp per output token.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void t_coef(int to, int Ti, int& i0, int& i1, float& lam) {
  float src = ((float)to + 0.5f) * 0.5f - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = i0 + 1 < Ti ? i0 + 1 : Ti - 1;
  lam = src - (float)i0;
}

__global__ void __launch_bounds__(256) classifier_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ stem,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ logits, int B, int Ti, int S, int C) {
  pdl_wait();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int To = 2 * Ti;
  const int64_t total = (int64_t)B * To * S;
  for (int64_t tok = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); tok < total; tok += (int64_t)gridDim.x * wpb) {
    int s = (int)(tok % S), to = (int)((tok / S) % To);
    int64_t b = tok / ((int64_t)S * To);
    int i0, i1; float lam;
    t_coef(to, Ti, i0, i1, lam);
    const float* f = feat + tok * C;
    const float* a0 = stem + ((b * Ti + i0) * S + s) * C;
    const float* a1 = stem + ((b * Ti + i1) * S + s) * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc += w[c] * (f[c] + (1.f - lam) * a0[c] + lam * a1[c]);
    acc = warp_sum(acc);
    if (lane == 0) logits[tok] = acc + bias[0];
  }
}
// dfeat = dlogit * w ; dw += sum dlogit * val ; dbias += sum dlogit
__global__ void __launch_bounds__(256) classifier_bwd_feat_kernel(const float* __restrict__ dlogits, const float* __restrict__ feat,
                                                                  const float* __restrict__ stem, const float* __restrict__ w,
                                                                  float* __restrict__ dfeat, float* __restrict__ dw, float* __restrict__ dbias,
                                                                  int B, int Ti, int S, int C) {
  pdl_wait();
  __shared__ float s_dw[256];
  __shared__ float s_db;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_dw[i] = 0.f;
  if (threadIdx.x == 0) s_db = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int To = 2 * Ti;
  const int64_t total = (int64_t)B * To * S;
  float adw[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) adw[k] = 0.f;
  float adb = 0.f;
  for (int64_t tok = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); tok < total; tok += (int64_t)gridDim.x * wpb) {
    int s = (int)(tok % S), to = (int)((tok / S) % To);
    int64_t b = tok / ((int64_t)S * To);
    int i0, i1; float lam;
    t_coef(to, Ti, i0, i1, lam);
    const float g = dlogits[tok];
    const float* f = feat + tok * C;
    const float* a0 = stem + ((b * Ti + i0) * S + s) * C;
    const float* a1 = stem + ((b * Ti + i1) * S + s) * C;
    float* df = dfeat + tok * C;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int c = lane + 32 * k;
      if (c < C) {
        df[c] = g * w[c];
        adw[k] += g * (f[c] + (1.f - lam) * a0[c] + lam * a1[c]);
      }
    }
    if (lane == 0) adb += g;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = lane + 32 * k;
    if (c < C) atomicAdd(&s_dw[c], adw[k]);
  }
  if (lane == 0) atomicAdd(&s_db, adb);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dw + i, s_dw[i]);
  if (threadIdx.x == 0) atomicAdd(dbias, s_db);
}
// dstem[b][ti][s][c] = w[c] * sum_{to} coef(to -> ti) * dlogit[b][to][s]
// One warp per (b, ti, s) token: the interpolation-adjoint sum is a per-token scalar (evaluated once, not once per
// channel), then the C channels are written as coalesced rows.
__global__ void __launch_bounds__(256) classifier_bwd_stem_kernel(const float* __restrict__ dlogits, const float* __restrict__ w,
                                                                  float* __restrict__ dstem, int B, int Ti, int S, int C) {
  pdl_wait();
  const int To = 2 * Ti;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const uint32_t total = (uint32_t)B * Ti * S;
  for (uint32_t tok = blockIdx.x * wpb + (threadIdx.x >> 5); tok < total; tok += gridDim.x * wpb) {
    uint32_t q = tok;
    const int s = (int)divmod(q, S), ti = (int)divmod(q, Ti);
    const int64_t b = q;
    float acc = 0.f;
    for (int to = max(2 * ti - 2, 0); to <= min(2 * ti + 3, To - 1); ++to) {
      int i0, i1; float lam;
      t_coef(to, Ti, i0, i1, lam);
      float cf = 0.f;
      if (i0 == ti) cf += 1.f - lam;
      if (i1 == ti) cf += lam;
      if (cf != 0.f) acc += cf * __ldg(dlogits + (b * To + to) * S + s);
    }
    float* d = dstem + (int64_t)tok * C;
    for (int c = lane; c < C; c += 32) d[c] = acc * w[c];
  }
}

// ---- C <= 96 (the model's 96): a token is owned by 8 lanes with 3 four-channel vectors each, a warp works on 4 tokens at once
// with 16-byte accesses (the kernels above walk one token per warp with scalar loads)
__global__ void __launch_bounds__(256) classifier96_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ stem,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               float* __restrict__ logits, int B, int Ti, int S, int C) {
  pdl_wait();
  const int lane = threadIdx.x & 31, li = lane & 7, grp = lane >> 3, wpb = blockDim.x >> 5;
  const int To = 2 * Ti;
  const uint32_t total = (uint32_t)B * To * S;
  float wv[3][4];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = (li + 8 * j) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) wv[j][k] = 0.f;
    if (c < C) ld4(w + c, wv[j]);
  }
  const float b0 = bias[0];
  for (uint32_t t0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * 4; t0 < total; t0 += gridDim.x * wpb * 4) {
    const uint32_t tok = t0 + grp;
    const bool ok = tok < total;
    uint32_t q = ok ? tok : 0;
    const int s = (int)divmod(q, S), to = (int)divmod(q, To);
    const int64_t b = q;
    int i0, i1; float lam;
    t_coef(to, Ti, i0, i1, lam);
    const float* f = feat + (int64_t)(ok ? tok : 0) * C;
    const float* a0 = stem + ((b * Ti + i0) * S + s) * C;
    const float* a1 = stem + ((b * Ti + i1) * S + s) * C;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = (li + 8 * j) * 4;
      if (c < C) {
        float fv[4], x0[4], x1[4];
        ld4(f + c, fv);
        ld4(a0 + c, x0);
        ld4(a1 + c, x1);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += wv[j][k] * (fv[k] + (1.f - lam) * x0[k] + lam * x1[k]);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (ok && li == 0) logits[tok] = acc + b0;
  }
}

__global__ void __launch_bounds__(256) classifier96_bwd_feat_kernel(const float* __restrict__ dlogits, const float* __restrict__ feat,
                                                                    const float* __restrict__ stem, const float* __restrict__ w,
                                                                    float* __restrict__ dfeat, float* __restrict__ dw, float* __restrict__ dbias,
                                                                    int B, int Ti, int S, int C) {
  pdl_wait();
  __shared__ float s_dw[96];
  __shared__ float s_db;
  for (int i = threadIdx.x; i < 96; i += blockDim.x) s_dw[i] = 0.f;
  if (threadIdx.x == 0) s_db = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, li = lane & 7, grp = lane >> 3, wpb = blockDim.x >> 5;
  const int To = 2 * Ti;
  const uint32_t total = (uint32_t)B * To * S;
  float wv[3][4], adw[3][4];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = (li + 8 * j) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) { wv[j][k] = 0.f; adw[j][k] = 0.f; }
    if (c < C) ld4(w + c, wv[j]);
  }
  float adb = 0.f;
  for (uint32_t t0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * 4; t0 < total; t0 += gridDim.x * wpb * 4) {
    const uint32_t tok = t0 + grp;
    if (tok >= total) continue;
    uint32_t q = tok;
    const int s = (int)divmod(q, S), to = (int)divmod(q, To);
    const int64_t b = q;
    int i0, i1; float lam;
    t_coef(to, Ti, i0, i1, lam);
    const float g = dlogits[tok];
    const float* f = feat + (int64_t)tok * C;
    const float* a0 = stem + ((b * Ti + i0) * S + s) * C;
    const float* a1 = stem + ((b * Ti + i1) * S + s) * C;
    float* df = dfeat + (int64_t)tok * C;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = (li + 8 * j) * 4;
      if (c < C) {
        float fv[4], x0[4], x1[4], o[4];
        ld4(f + c, fv);
        ld4(a0 + c, x0);
        ld4(a1 + c, x1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          o[k] = g * wv[j][k];
          adw[j][k] += g * (fv[k] + (1.f - lam) * x0[k] + lam * x1[k]);
        }
        st4(df + c, o);
      }
    }
    if (li == 0) adb += g;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      adw[j][k] += __shfl_xor_sync(0xffffffffu, adw[j][k], 8);
      adw[j][k] += __shfl_xor_sync(0xffffffffu, adw[j][k], 16);
    }
    const int c = (li + 8 * j) * 4;
    if (grp == 0 && c < C) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&s_dw[c + k], adw[j][k]);
    }
  }
  adb += __shfl_xor_sync(0xffffffffu, adb, 8);
  adb += __shfl_xor_sync(0xffffffffu, adb, 16);
  if (lane == 0) atomicAdd(&s_db, adb);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dw + i, s_dw[i]);
  if (threadIdx.x == 0) atomicAdd(dbias, s_db);
}

}  // namespace

extern "C" {

int csts_im2col_patch(const float* x, void* patches, int dtype, int B, int Cin, int T, int H, int W, int Kp, void* stream) {
  CSTS_REQUIRE(dtype == CSTS_BF16 || dtype == CSTS_F16, "im2col: patches must be bf16 or f16");
  CSTS_REQUIRE(T % 2 == 0 && H % 4 == 0 && W % 4 == 0 && Kp >= Cin * 147 && Kp % 8 == 0, "im2col: bad geometry");
  int64_t total = (int64_t)B * (T / 2) * (H / 4) * (W / 4) * (Kp / 8);          // 16-byte chunks
  if (total == 0) return 0;
  CSTS_REQUIRE(total < (1LL << 32) && ((uintptr_t)patches & 15) == 0, "im2col: too many patch chunks for 32-bit indexing / unaligned output");
  if (dtype == CSTS_F16) launch_pdl(im2col_kernel<f16>, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, (f16*)patches, B, Cin, T, H, W, Kp);
  else launch_pdl(im2col_kernel<bf16>, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, (bf16*)patches, B, Cin, T, H, W, Kp);
  return csts_check_launch("im2col_patch");
}
int csts_pos_embed(const float* spatial, const float* temporal, float* pos, int T, int HW, int C, void* stream) {
  launch_pdl(pos_embed_kernel, dim3(grid_for((int64_t)T * HW * C, 256)), dim3(256), 0, (cudaStream_t)stream, spatial, temporal, pos, T, HW, C);
  return csts_check_launch("pos_embed");
}
// dtemporal must be zeroed by the caller; dspatial is overwritten
int csts_pos_embed_bwd(const float* dY, float* dspatial, float* dtemporal, int B, int T, int HW, int C, void* stream) {
  CSTS_REQUIRE(C <= 128 && T <= 8, "pos_embed_bwd: C <= 128 and T <= 8 required");
  int s_per_block = 8;
  launch_pdl(pos_embed_bwd_kernel, dim3(ceil_div(HW, s_per_block)), dim3(128), 0, (cudaStream_t)stream, dY, dspatial, dtemporal, B, T, HW, C, s_per_block);
  return csts_check_launch("pos_embed_bwd");
}
int csts_reweight_fwd(const float* x, const float* w, float* out, int B, int T, int S, int C, int64_t w_sB, void* stream) {
  CSTS_REQUIRE(C % 4 == 0 && w_sB % 4 == 0, "reweight: C and w_sB must be multiples of 4");
  launch_pdl(reweight_fwd_kernel, dim3(grid_for((int64_t)B * T * S * C / 4, 256)), dim3(256), 0, (cudaStream_t)stream, x, w, out, B, T, S, C, w_sB);
  return csts_check_launch("reweight_fwd");
}
int csts_reweight_bwd(const float* dout, const float* x, const float* w, float* dx, float* dw, int B, int T, int S, int C, int64_t w_sB,
                      void* stream) {
  CSTS_REQUIRE(C % 4 == 0 && w_sB % 4 == 0, "reweight: C and w_sB must be multiples of 4");
  launch_pdl(reweight_bwd_kernel, dim3(grid_for((int64_t)B * T * C / 4, 64)), dim3(64), 0, (cudaStream_t)stream, dout, x, w, dx, dw, B, T, S, C, w_sB);
  return csts_check_launch("reweight_bwd");
}
int csts_token_mean_fwd(const float* x, void* out, int out_dtype, int B, int N, int C, void* stream) {
  CSTS_REQUIRE(out_dtype == CSTS_BF16 || out_dtype == CSTS_F16, "token_mean: out must be bf16 or f16");
  if (out_dtype == CSTS_F16) launch_pdl(token_mean_fwd_kernel<f16>, dim3(grid_for((int64_t)B * C, 64)), dim3(64), 0, (cudaStream_t)stream, x, (f16*)out, B, N, C);
  else launch_pdl(token_mean_fwd_kernel<bf16>, dim3(grid_for((int64_t)B * C, 64)), dim3(64), 0, (cudaStream_t)stream, x, (bf16*)out, B, N, C);
  return csts_check_launch("token_mean_fwd");
}
int csts_token_mean_bwd(const float* dout, float* dx, int B, int N, int C, int accumulate, void* stream) {
  launch_pdl(token_mean_bwd_kernel, dim3(grid_for((int64_t)B * N * C, 256)), dim3(256), 0, (cudaStream_t)stream, dout, dx, B, N, C, accumulate);
  return csts_check_launch("token_mean_bwd");
}
int csts_classifier_fwd(const float* feat, const float* stem, const float* w, const float* bias, float* logits, int B, int Ti, int S, int C,
                        void* stream) {
  if (C <= 96 && C % 4 == 0 && (int64_t)B * 2 * Ti * S < (1LL << 31)) {
    launch_pdl(classifier96_fwd_kernel, dim3(grid_for((int64_t)B * 2 * Ti * S, 32)), dim3(256), 0, (cudaStream_t)stream, feat, stem, w, bias, logits,
               B, Ti, S, C);
    return csts_check_launch("classifier_fwd");
  }
  launch_pdl(classifier_fwd_kernel, dim3(grid_for((int64_t)B * 2 * Ti * S, 8)), dim3(256), 0, (cudaStream_t)stream, feat, stem, w, bias, logits, B, Ti, S, C);
  return csts_check_launch("classifier_fwd");
}
// dw, dbias must be zeroed by the caller
int csts_classifier_bwd(const float* dlogits, const float* feat, const float* stem, const float* w, float* dfeat, float* dstem, float* dw,
                        float* dbias, int B, int Ti, int S, int C, void* stream) {
  CSTS_REQUIRE(C <= 256, "classifier_bwd: C <= 256 required");
  int64_t toks = (int64_t)B * 2 * Ti * S;
  int64_t blocks = (toks + 255) / 256;
  int grid = (int)(blocks < csts_num_sms() * 2 ? blocks : csts_num_sms() * 2);
  if (C <= 96 && C % 4 == 0 && toks < (1LL << 31)) {
    const int g96 = (int)std::min<int64_t>((toks + 127) / 128, (int64_t)csts_num_sms() * 4);       // >= 4 passes per warp
    launch_pdl(classifier96_bwd_feat_kernel, dim3(g96 > 0 ? g96 : 1), dim3(256), 0, (cudaStream_t)stream, dlogits, feat, stem, w, dfeat, dw, dbias,
               B, Ti, S, C);
  } else {
    launch_pdl(classifier_bwd_feat_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, dlogits, feat, stem, w, dfeat, dw, dbias, B, Ti, S, C);
  }
  int rc = csts_check_launch("classifier_bwd_feat");
  if (rc) return rc;
  CSTS_REQUIRE((int64_t)B * Ti * S * C < (1LL << 32), "classifier_bwd: tensor too large for 32-bit indexing");
  launch_pdl(classifier_bwd_stem_kernel, dim3(grid_for((int64_t)B * Ti * S, 8)), dim3(256), 0, (cudaStream_t)stream, dlogits, w, dstem, B, Ti, S, C);
  return csts_check_launch("classifier_bwd_stem");
}

}  // extern "C"
