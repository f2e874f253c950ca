// Argument blocks for the token-grid kernels (mirrored by include/csts_b200.h).
#pragma once
#include <stdint.h>

extern "C" {
// Depthwise 3x3x3 (transposed) conv over a token grid, optional fused LayerNorm(d).
// Element (b, head, pos, c) of `in` lives at in + b*in_sB + head*in_sH + pos*in_sP + c (bf16).
typedef struct csts_pool_args {
  const void* in;        // bf16
  void* out;             // bf16
  const float* w;        // (d,1,3,3,3) parameter, f32
  const float* gamma;    // LayerNorm(d) weight or NULL (no norm: `out` = raw conv)
  const float* beta;
  void* pre;             // bf16 dense (B, heads, Lo, d): raw conv output kept for backward (norm only)
  float* mean;           // [B*heads*Lo] (norm only)
  float* rstd;
  int64_t in_sB, in_sH, in_sP;
  int64_t out_sB, out_sH, out_sP;
  int32_t B, heads, d;
  int32_t Ti, Hi, Wi;    // input grid
  int32_t To, Ho, Wo;    // output grid
  int32_t st, sh, sw;    // stride of the (un-transposed) convolution
  int32_t transposed;    // 0: out[o] = sum_tap w[tap] in[o*s+tap-1]; 1: out[o] = sum_tap w[tap] in[(o+1-tap)/s]
  float eps;
} csts_pool_args;

// dw[c][tap] += sum small[o][c] * big[o*s + tap - 1][c]
typedef struct csts_wgrad_args {
  const void* small;     // bf16, grid (Ts,Hs,Ws)
  const void* big;       // bf16, grid (Tb,Hb,Wb)
  float* dw;             // (d,1,3,3,3) f32, accumulated into
  int64_t small_sB, small_sH, small_sP;
  int64_t big_sB, big_sH, big_sP;
  int32_t B, heads, d;
  int32_t Ts, Hs, Ws;
  int32_t Tb, Hb, Wb;
  int32_t st, sh, sw;
} csts_wgrad_args;
}
