// Memory-bound row-wise kernels: LayerNorm fwd/bwd, attention softmax fwd/bwd, casts, column sums,
// elementwise residual adds.  All HBM-bound: 128-bit (f32x4) / 64-bit (bf16x4) accesses, one warp per
// row with shuffle reductions, grids sized in multiples of the SM count.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int LN_MAX_WIDTH = 768;   // up to 6 column groups of 128 per warp

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * gamma + beta            ref: nn.LayerNorm (attention.py:239,243)
// ------------------------------------------------------------------------------------------------
template <typename TI, typename TO, int LN_MAXJ>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const TI* __restrict__ x, TO* __restrict__ y,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int rows, int width, float eps) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_w = 1.f / (float)width;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const TI* xr = x + (int64_t)row * width;
    float v[LN_MAXJ][4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) {
      int c = (lane + 32 * j) * 4;
      if (c < width) {
        ld4(xr + c, v[j]);
        s += v[j][0] + v[j][1] + v[j][2] + v[j][3];
      }
    }
    const float mean = warp_sum(s) * inv_w;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) {
      int c = (lane + 32 * j) * 4;
      if (c < width) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { float d = v[j][i] - mean; q += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_w + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    TO* yr = y + (int64_t)row * width;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) {
      int c = (lane + 32 * j) * 4;
      if (c < width) {
        float g[4], b[4], o[4];
        ld4(gamma + c, g);
        ld4(beta + c, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = (v[j][i] - mean) * rstd * g[i] + b[i];
        st4(yr + c, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  dx = [add +] rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));
// dgamma += sum_rows dy*xhat, dbeta += sum_rows dy.
//
// A row is owned by LPR lanes (8 / 16 / 32 for widths <= 96 / 192 / wider), each holding J vectors of 4 columns, so
// a warp works on 32/LPR rows at once with every lane busy (the 96-wide pool norms are 60 % of the launches), and U
// such passes are loaded before the first reduction: U * J * 3 independent 8/16-byte loads in flight per lane.  The
// grid is persistent (a few blocks per SM): dgamma / dbeta accumulate in registers over all rows of a lane, are folded
// across the warp's row groups by shuffles and across the block's warps in shared memory, and reach global memory as
// 2 * width atomics per BLOCK — a few hundred per address per launch instead of one per 32 rows.
// ------------------------------------------------------------------------------------------------
// second problem of a paired launch (blockIdx.y == 1): same geometry and types, other tensors — the norm_k / norm_v
// backward of a block run as one launch
struct LnSecond { const void* dy; const void* x; const float* mean; const float* rstd; const float* gamma; void* dx; float* dgamma; float* dbeta; };

template <typename TX, typename TDY, typename TDX, int LPR, int J, int U>
__global__ void __launch_bounds__(256, J <= 3 ? 2 : 1) layernorm_bwd_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ add,
                                                            TDX* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int rows, int width, void* __restrict__ dx16,
                                                            int dx16_half, const float* __restrict__ row_scale, int rows_per_scale,
                                                            LnSecond sec) {
  pdl_wait();
  if (blockIdx.y == 1) {
    dy = reinterpret_cast<const TDY*>(sec.dy); x = reinterpret_cast<const TX*>(sec.x); mean = sec.mean; rstd = sec.rstd; gamma = sec.gamma;
    dx = reinterpret_cast<TDX*>(sec.dx); dgamma = sec.dgamma; dbeta = sec.dbeta;
  }
  __shared__ float s_dg[768], s_db[768];
  for (int i = threadIdx.x; i < width; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  __syncthreads();
  constexpr int G = 32 / LPR;                       // rows per warp pass
  const int lane = threadIdx.x & 31;
  const int li = lane % LPR, grp = lane / LPR;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_w = 1.f / (float)width;
  float adg[J][4], adb[J][4], g[J][4];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c = (li + LPR * j) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) { adg[j][i] = 0.f; adb[j][i] = 0.f; g[j][i] = 0.f; }
    if (c < width) ld4(gamma + c, g[j]);
  }
  const int total_warps = gridDim.x * warps_per_block;
  const int passes = (rows + G - 1) / G;
  for (int p0 = blockIdx.x * warps_per_block + (threadIdx.x >> 5); p0 < passes; p0 += U * total_warps) {
    float xv[U][J][4], dv[U][J][4], av[U][J][4], mu[U], rs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = (p0 + u * total_warps) * G + grp;
      const bool rv = p0 + u * total_warps < passes && row < rows;
      mu[u] = rv ? mean[row] : 0.f;
      rs[u] = rv ? rstd[row] : 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c = (li + LPR * j) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) { xv[u][j][i] = 0.f; dv[u][j][i] = 0.f; av[u][j][i] = 0.f; }
        if (rv && c < width) {
          ld4(x + (int64_t)row * width + c, xv[u][j]);
          ld4(dy + (int64_t)row * width + c, dv[u][j]);
          if (add) ld4(add + (int64_t)row * width + c, av[u][j]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = (p0 + u * total_warps) * G + grp;
      const bool rv = p0 + u * total_warps < passes && row < rows;      // whole row groups take the shuffles together
      float xh[J][4], gd[J][4];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xh[j][i] = (xv[u][j][i] - mu[u]) * rs[u];       // lanes beyond the width / rows hold zeros: they add nothing
          adg[j][i] += dv[u][j][i] * xh[j][i];
          adb[j][i] += dv[u][j][i];
          gd[j][i] = dv[u][j][i] * g[j][i];
          s1 += gd[j][i];
          s2 += gd[j][i] * xh[j][i];
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      s1 *= inv_w;
      s2 *= inv_w;
      if (!rv) continue;
      TDX* dxr = dx + (int64_t)row * width;
      const float sc = (dx16 && row_scale) ? row_scale[row / rows_per_scale] : 1.f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c = (li + LPR * j) * 4;
        if (c < width) {
          float o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = rs[u] * (gd[j][i] - s1 - xh[j][i] * s2) + av[u][j][i];
          st4(dxr + c, o);
          if (dx16) {                                  // 16-bit (row-scaled) copy: the operand of the next backward GEMM
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] *= sc;
            if (dx16_half) st4(reinterpret_cast<f16*>(dx16) + (int64_t)row * width + c, o);
            else st4(reinterpret_cast<bf16*>(dx16) + (int64_t)row * width + c, o);
          }
        }
      }
    }
  }
  // fold the warp's row groups (they own the same columns), then the block's warps
#pragma unroll
  for (int j = 0; j < J; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int o = 16; o >= LPR; o >>= 1) {
        adg[j][i] += __shfl_xor_sync(0xffffffffu, adg[j][i], o);
        adb[j][i] += __shfl_xor_sync(0xffffffffu, adb[j][i], o);
      }
    }
    const int c = (li + LPR * j) * 4;
    if (grp == 0 && c < width) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { atomicAdd(&s_dg[c + i], adg[j][i]); atomicAdd(&s_db[c + i], adb[j][i]); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += blockDim.x) { atomicAdd(dgamma + i, s_dg[i]); atomicAdd(dbeta + i, s_db[i]); }
}

// ------------------------------------------------------------------------------------------------
// Attention softmax over rows of S (f32, already scaled).  P is bf16 with leading dim ldp; columns
// [n, ldp) are written as zero so that P can be a GEMM operand.  mask_hw > 0 selects the in-frame
// block mask of SpatialAttention (ref av_attention.py:336-347): token i belongs to frame
// i / mask_hw for i < T*mask_hw and to frame i - T*mask_hw otherwise; cross-frame logits get -1e8,
// i.e. probability exactly 0 in fp32.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int frame_of(int i, int thw, int hw) { return i < thw ? i / hw : i - thw; }

template <typename TP>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ S, TP* __restrict__ P, int64_t rows, int n,
                                                          int lds, int ldp, int nq, int mask_hw, int mask_t) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * wpb) {
    const float* s = S + row * lds;
    TP* p = P + row * ldp;
    const int thw = mask_t * mask_hw;
    const int fq = mask_hw > 0 ? frame_of((int)(row % nq), thw, mask_hw) : 0;
    float mx = -INFINITY;
    for (int c = lane; c < n; c += 32) {
      bool ok = mask_hw <= 0 || frame_of(c, thw, mask_hw) == fq;
      if (ok) mx = fmaxf(mx, s[c]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < n; c += 32) {
      bool ok = mask_hw <= 0 || frame_of(c, thw, mask_hw) == fq;
      if (ok) sum += __expf(s[c] - mx);
    }
    const float inv = 1.f / warp_sum(sum);
    for (int c = lane; c < ldp; c += 32) {
      float v = 0.f;
      if (c < n) {
        bool ok = mask_hw <= 0 || frame_of(c, thw, mask_hw) == fq;
        if (ok) v = __expf(s[c] - mx) * inv;
      }
      st_f(p + c, v);
    }
  }
}

// dS = scale * P o (dP - rowsum(dP o P)); dS 16-bit with zeroed pad columns
template <typename TP, typename TDS>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const TP* __restrict__ P, const float* __restrict__ dP, TDS* __restrict__ dS,
                                                          int64_t rows, int n, int ldp, int lddp, float scale) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * wpb) {
    const TP* p = P + row * ldp;
    const float* dp = dP + row * lddp;
    TDS* ds = dS + row * ldp;
    float dot = 0.f;
    for (int c = lane; c < n; c += 32) dot += ld_f(p + c) * dp[c];
    dot = warp_sum(dot);
    for (int c = lane; c < ldp; c += 32) {
      float v = c < n ? scale * ld_f(p + c) * (dp[c] - dot) : 0.f;
      st_f(ds + c, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// D[b][head][q] = sum_d dO[b][q][head][d] * O[b][q][head][d]: the row term of the softmax backward, rowsum(dP o P) = dO . O,
// so that dP itself never has to exist.  One warp per (b, q) token; 8 consecutive lanes own a 64-column stripe.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) rowdot_kernel(const T* __restrict__ dO, const T* __restrict__ O, float* __restrict__ D, int B, int Lq,
                                                     int heads, int d) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int Cn = heads * d;
  for (int64_t tok = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); tok < (int64_t)B * Lq; tok += (int64_t)gridDim.x * wpb) {
    const T* a = dO + tok * Cn;
    const T* b = O + tok * Cn;
    const int bi = (int)(tok / Lq), q = (int)(tok - (int64_t)bi * Lq);
    for (int h = 0; h < heads; ++h) {
      float acc = 0.f;
      for (int c = lane * 4; c < d; c += 128) {
        float x[4], y[4];
        ld4(a + h * d + c, x);
        ld4(b + h * d + c, y);
        acc += x[0] * y[0] + x[1] * y[1] + x[2] * y[2] + x[3] * y[3];
      }
      acc = warp_sum(acc);
      if (lane == 0) D[((int64_t)bi * heads + h) * Lq + q] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// casts / permutes / elementwise
// ------------------------------------------------------------------------------------------------
// f32 [rows, cols] -> 16-bit [rows, ld_out] (zero padded columns), T in {bf16, f16}
template <typename T>
__global__ void cast_pad_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t rows, int cols, int ld_out,
                                const float* __restrict__ row_scale, int rows_per_scale) {
  pdl_wait();
  int64_t total = rows * ld_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / ld_out;
    int c = (int)(i - r * ld_out);
    float v = c < cols ? src[r * cols + c] : 0.f;
    if (row_scale) v *= row_scale[r / rows_per_scale];
    st_f(dst + i, v);
  }
}
// cols must be a multiple of 4 (a 4-vector never straddles rows)
template <typename T>
__global__ void cast_vec_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n4, int cols4,
                                const float* __restrict__ row_scale, int rows_per_scale) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4];
    ld4(src + 4 * i, v);
    if (row_scale) {
      float rs = row_scale[(i / cols4) / rows_per_scale];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] *= rs;
    }
    st4(dst + 4 * i, v);
  }
}
// src [a][b][c] -> dst [a][c][b]   (tile transpose through shared memory), TO in {float, bf16}
template <typename TO>
__global__ void permute_021_kernel(const float* __restrict__ src, TO* __restrict__ dst, int a, int b, int c) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int64_t base = (int64_t)blockIdx.z * b * c;
  int b0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int bb = b0 + i, cc = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (bb < b && cc < c) ? src[base + (int64_t)bb * c + cc] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int cc = c0 + i, bb = b0 + threadIdx.x;
    if (cc < c && bb < b) st_f(dst + base + (int64_t)cc * b + bb, tile[threadIdx.x][i]);
  }
}

// out = a + b (f32), vectorised
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n4) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}
// out = a * (*scalar)   (device scalar; used to apply the upstream loss gradient)
__global__ void scale_kernel(const float* __restrict__ a, const float* __restrict__ scalar, float* __restrict__ out, int64_t n) {
  pdl_wait();
  const float s = *scalar;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = a[i] * s;
}

// column sums: out[n] += sum_m X[m][n]   (bias gradients).  Block = 32 column-quads x 8 row phases.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, float* __restrict__ out, int64_t M, int N, int64_t ld,
                                                     int rows_per_block) {
  pdl_wait();
  __shared__ float red[8][128];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 4;
  int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < N) {
#pragma unroll 4
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float v[4];
      ld4(X + r * ld + col, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) red[ty][tx * 4 + i] = acc[i];
  __syncthreads();
  if (ty == 0 && col < N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) t += red[y][tx * 4 + i];
      atomicAdd(out + col + i, t);
    }
  }
}

// 8 columns per thread (one 16-byte load for 16-bit inputs, two for f32).  A block covers GX column groups x
// RY = 256 / GX row phases, so narrow matrices (N = 96 ... 288 with M = 131072 rows) still keep every thread
// busy; four rows are in flight per thread.
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  float2 a = unpack2<T>(t.x), b = unpack2<T>(t.y), c = unpack2<T>(t.z), d = unpack2<T>(t.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T>
__global__ void __launch_bounds__(256) colsum8_kernel(const T* __restrict__ X, float* __restrict__ out, int64_t M, int N, int64_t ld,
                                                      int rows_per_block, int GX) {
  pdl_wait();
  __shared__ float red[256 * 8];
  const int tx = threadIdx.x % GX, ty = threadIdx.x / GX, RY = blockDim.x / GX;
  const int col = (blockIdx.x * GX + tx) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (ty < RY && col < N) {
    const T* base = X + col;
#pragma unroll 4
    for (int64_t r = r0 + ty; r < r1; r += RY) {
      float v[8];
      ld8(base + r * ld, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[i * 256 + threadIdx.x] = acc[i];     // [column-in-group][thread]: conflict-free both ways
  __syncthreads();
  // thread (i, tx) of the first 8 * GX threads sums column i of group tx over the RY row phases
  for (int e = threadIdx.x; e < 8 * GX; e += blockDim.x) {
    const int i = e / GX, gx = e - i * GX;
    const int c = (blockIdx.x * GX + gx) * 8 + i;
    if (c < N) {
      float t = 0.f;
      for (int y = 0; y < RY; ++y) t += red[i * 256 + y * GX + gx];
      atomicAdd(out + c, t);
    }
  }
}

int grid_for(int64_t work_items, int per_block) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)csts_num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

extern "C" {

// dtype codes: 0 = f32, 1 = bf16, 2 = f16
int csts_layernorm_fwd(const void* x, int x_dtype, void* y, int y_dtype, const float* gamma, const float* beta, float* mean,
                       float* rstd, int64_t rows, int width, float eps, void* stream) {
  CSTS_REQUIRE(width % 4 == 0 && width <= LN_MAX_WIDTH, "layernorm: width %d unsupported (multiple of 4, <= 768)", width);
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(rows, 8);
#define LN_FWD_J(TI, TO, J) launch_pdl(layernorm_fwd_kernel<TI, TO, J>, dim3(grid), dim3(256), 0, st, (const TI*)x, (TO*)y, gamma, beta, mean, rstd, (int)rows, width, eps)
#define LN_FWD(TI, TO)                                   \
  do {                                                   \
    if (width <= 128) LN_FWD_J(TI, TO, 1);               \
    else if (width <= 256) LN_FWD_J(TI, TO, 2);          \
    else if (width <= 384) LN_FWD_J(TI, TO, 3);          \
    else LN_FWD_J(TI, TO, 6);                            \
  } while (0)
  if (x_dtype == 0 && y_dtype == 2) LN_FWD(float, f16);
  else if (x_dtype == 0 && y_dtype == 1) LN_FWD(float, bf16);
  else if (x_dtype == 0 && y_dtype == 0) LN_FWD(float, float);
  else if (x_dtype == 2 && y_dtype == 2) LN_FWD(f16, f16);
  else if (x_dtype == 1 && y_dtype == 1) LN_FWD(bf16, bf16);
  else if (x_dtype == 1 && y_dtype == 0) LN_FWD(bf16, float);
  else CSTS_REQUIRE(false, "layernorm: bad dtype codes %d %d", x_dtype, y_dtype);
#undef LN_FWD
#undef LN_FWD_J
  return csts_check_launch("layernorm_fwd");
}

static int layernorm_bwd_launch(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                                const float* gamma, const float* add, void* dx, int dx_dtype, float* dgamma, float* dbeta, int64_t rows,
                                int width, void* dx16, int dx16_dtype, const float* row_scale, int rows_per_scale, const LnSecond* second,
                                void* stream) {
  CSTS_REQUIRE(width % 4 == 0 && width <= LN_MAX_WIDTH, "layernorm_bwd: width %d unsupported", width);
  if (dx16) CSTS_REQUIRE(dx16_dtype == CSTS_BF16 || dx16_dtype == CSTS_F16, "layernorm_bwd: dx16 must be bf16 or f16");
  if (dx16 && row_scale) CSTS_REQUIRE(rows_per_scale > 0, "layernorm_bwd: rows_per_scale must be positive");
  const int dx16_half = dx16_dtype == CSTS_F16;
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  LnSecond sec = {};
  if (second) sec = *second;
  const int ny = second ? 2 : 1;
  // persistent grid: enough warps to keep HBM busy, few enough blocks that the 2 * width closing atomics per block are noise
  const int blocks_per_sm = width <= 384 ? 2 : 1;          // what the register budget of the two variants allows
  const int lpr = width <= 96 ? 8 : (width <= 192 ? 16 : 32);
  const int64_t passes = (rows + 32 / lpr - 1) / (32 / lpr);
  const int64_t blocks = (passes + 8 * 4 - 1) / (8 * 4);                 // >= 4 passes per warp
  int grid = (int)(blocks < (int64_t)csts_num_sms() * blocks_per_sm ? (blocks > 0 ? blocks : 1) : (int64_t)csts_num_sms() * blocks_per_sm);
  if (ny == 2 && grid > csts_num_sms() * blocks_per_sm / 2) grid = csts_num_sms() * blocks_per_sm / 2;
#define LN_BWD_J(TX, TDY, TDX, LPR, J, U) \
  launch_pdl(layernorm_bwd_kernel<TX, TDY, TDX, LPR, J, U>, dim3(grid, ny), dim3(256), 0, st, (const TDY*)dy, (const TX*)x, mean, rstd, gamma, add, (TDX*)dx, dgamma, dbeta, (int)rows, width, dx16, dx16_half, row_scale, rows_per_scale, sec)
#define LN_BWD(TX, TDY, TDX)                                  \
  do {                                                        \
    if (width <= 96) LN_BWD_J(TX, TDY, TDX, 8, 3, 1);         \
    else if (width <= 192) LN_BWD_J(TX, TDY, TDX, 16, 3, 1);  \
    else if (width <= 384) LN_BWD_J(TX, TDY, TDX, 32, 3, 1);  \
    else LN_BWD_J(TX, TDY, TDX, 32, 6, 1);                    \
  } while (0)
  if (x_dtype == 0 && dy_dtype == 1 && dx_dtype == 0) LN_BWD(float, bf16, float);
  else if (x_dtype == 0 && dy_dtype == 2 && dx_dtype == 0) LN_BWD(float, f16, float);
  else if (x_dtype == 2 && dy_dtype == 2 && dx_dtype == 2) LN_BWD(f16, f16, f16);
  else if (x_dtype == 2 && dy_dtype == 1 && dx_dtype == 1) LN_BWD(f16, bf16, bf16);
  else if (x_dtype == 1 && dy_dtype == 1 && dx_dtype == 1) LN_BWD(bf16, bf16, bf16);
  else if (x_dtype == 1 && dy_dtype == 0 && dx_dtype == 1) LN_BWD(bf16, float, bf16);      // split-K (f32) dK / dV of the pooled keys
  else if (x_dtype == 2 && dy_dtype == 0 && dx_dtype == 2) LN_BWD(f16, float, f16);
  else if (x_dtype == 0 && dy_dtype == 0 && dx_dtype == 0) LN_BWD(float, float, float);
  else CSTS_REQUIRE(false, "layernorm_bwd: unsupported dtype combination x=%d dy=%d dx=%d", x_dtype, dy_dtype, dx_dtype);
#undef LN_BWD
#undef LN_BWD_J
  return csts_check_launch("layernorm_bwd");
}

int csts_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                       const float* gamma, const float* add, void* dx, int dx_dtype, float* dgamma, float* dbeta, int64_t rows,
                       int width, void* dx16, int dx16_dtype, const float* row_scale, int rows_per_scale, void* stream) {
  return layernorm_bwd_launch(dy, dy_dtype, x, x_dtype, mean, rstd, gamma, add, dx, dx_dtype, dgamma, dbeta, rows, width, dx16, dx16_dtype,
                              row_scale, rows_per_scale, nullptr, stream);
}

int csts_layernorm_bwd_pair(const void* dy0, const void* dy1, int dy_dtype, const void* x0, const void* x1, int x_dtype, const float* mean0,
                            const float* mean1, const float* rstd0, const float* rstd1, const float* gamma0, const float* gamma1, void* dx0,
                            void* dx1, int dx_dtype, float* dgamma0, float* dgamma1, float* dbeta0, float* dbeta1, int64_t rows, int width,
                            void* stream) {
  LnSecond sec = {dy1, x1, mean1, rstd1, gamma1, dx1, dgamma1, dbeta1};
  return layernorm_bwd_launch(dy0, dy_dtype, x0, x_dtype, mean0, rstd0, gamma0, nullptr, dx0, dx_dtype, dgamma0, dbeta0, rows, width, nullptr, 0,
                              nullptr, 0, &sec, stream);
}

int csts_rowdot(const void* dO, const void* O, int dtype, float* D, int B, int Lq, int heads, int d, void* stream) {
  CSTS_REQUIRE(dtype == CSTS_BF16 || dtype == CSTS_F16, "rowdot: 16-bit inputs only");
  CSTS_REQUIRE(d % 4 == 0, "rowdot: head dim %d must be a multiple of 4", d);
  if ((int64_t)B * Lq == 0) return 0;
  const int grid = grid_for((int64_t)B * Lq, 8);
  if (dtype == CSTS_F16)
    launch_pdl(rowdot_kernel<f16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const f16*)dO, (const f16*)O, D, B, Lq, heads, d);
  else
    launch_pdl(rowdot_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)dO, (const bf16*)O, D, B, Lq, heads, d);
  return csts_check_launch("rowdot");
}

int csts_softmax_fwd(const float* S, void* P, int p_dtype, int64_t rows, int n, int lds, int ldp, int nq, int mask_hw, int mask_t,
                     void* stream) {
  if (rows == 0) return 0;
  CSTS_REQUIRE(ldp >= n && lds >= n, "softmax: leading dims smaller than n");
  CSTS_REQUIRE(p_dtype == CSTS_BF16 || p_dtype == CSTS_F16, "softmax: P must be bf16 or f16");
  if (p_dtype == CSTS_F16)
    launch_pdl(softmax_fwd_kernel<f16>, dim3(grid_for(rows, 8)), dim3(256), 0, (cudaStream_t)stream, S, (f16*)P, rows, n, lds, ldp, nq, mask_hw, mask_t);
  else
    launch_pdl(softmax_fwd_kernel<bf16>, dim3(grid_for(rows, 8)), dim3(256), 0, (cudaStream_t)stream, S, (bf16*)P, rows, n, lds, ldp, nq, mask_hw, mask_t);
  return csts_check_launch("softmax_fwd");
}

int csts_softmax_bwd(const void* P, int p_dtype, const float* dP, void* dS, int ds_dtype, int64_t rows, int n, int ldp, int lddp,
                     float scale, void* stream) {
  if (rows == 0) return 0;
  CSTS_REQUIRE((p_dtype == CSTS_BF16 || p_dtype == CSTS_F16) && (ds_dtype == CSTS_BF16 || ds_dtype == CSTS_F16),
               "softmax_bwd: P and dS must be bf16 or f16");
  cudaStream_t st = (cudaStream_t)stream;
#define SM_BWD(TP, TDS) launch_pdl(softmax_bwd_kernel<TP, TDS>, dim3(grid_for(rows, 8)), dim3(256), 0, st, (const TP*)P, dP, (TDS*)dS, rows, n, ldp, lddp, scale)
  if (p_dtype == CSTS_F16 && ds_dtype == CSTS_F16) SM_BWD(f16, f16);
  else if (p_dtype == CSTS_F16) SM_BWD(f16, bf16);
  else if (ds_dtype == CSTS_F16) SM_BWD(bf16, f16);
  else SM_BWD(bf16, bf16);
#undef SM_BWD
  return csts_check_launch("softmax_bwd");
}

// row m of the result is multiplied by row_scale[m / rows_per_scale] when row_scale != NULL
int csts_cast16(const float* src, void* dst, int dst_dtype, int64_t rows, int cols, int ld_out, const float* row_scale,
                int rows_per_scale, void* stream) {
  if (rows * cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CSTS_REQUIRE(dst_dtype == CSTS_BF16 || dst_dtype == CSTS_F16, "cast: dst must be bf16 or f16");
  if (row_scale) CSTS_REQUIRE(rows_per_scale > 0, "cast: rows_per_scale must be positive");
  const bool half = dst_dtype == CSTS_F16;
  if (ld_out == cols && cols % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0) {
    int64_t n4 = rows * cols / 4;
    if (half) launch_pdl(cast_vec_kernel<f16>, dim3(grid_for(n4, 256)), dim3(256), 0, st, src, (f16*)dst, n4, cols / 4, row_scale, rows_per_scale);
    else launch_pdl(cast_vec_kernel<bf16>, dim3(grid_for(n4, 256)), dim3(256), 0, st, src, (bf16*)dst, n4, cols / 4, row_scale, rows_per_scale);
  } else {
    CSTS_REQUIRE(ld_out >= cols, "cast: ld_out < cols");
    if (half) launch_pdl(cast_pad_kernel<f16>, dim3(grid_for(rows * ld_out, 256)), dim3(256), 0, st, src, (f16*)dst, rows, cols, ld_out, row_scale, rows_per_scale);
    else launch_pdl(cast_pad_kernel<bf16>, dim3(grid_for(rows * ld_out, 256)), dim3(256), 0, st, src, (bf16*)dst, rows, cols, ld_out, row_scale, rows_per_scale);
  }
  return csts_check_launch("cast16");
}

// src f32 [a][b][c] -> dst [a][c][b], dst dtype 0 f32 / 1 bf16 / 2 f16
int csts_permute_021(const float* src, void* dst, int dst_dtype, int a, int b, int c, void* stream) {
  if ((int64_t)a * b * c == 0) return 0;
  dim3 grid(ceil_div(c, 32), ceil_div(b, 32), a), block(32, 8);
  CSTS_REQUIRE(a <= 65535 && grid.y <= 65535, "permute_021: dims too large");
  if (dst_dtype == 0) launch_pdl(permute_021_kernel<float>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, src, (float*)dst, a, b, c);
  else if (dst_dtype == CSTS_F16) launch_pdl(permute_021_kernel<f16>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, src, (f16*)dst, a, b, c);
  else launch_pdl(permute_021_kernel<bf16>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, src, (bf16*)dst, a, b, c);
  return csts_check_launch("permute_021");
}

int csts_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream) {
  if (n == 0) return 0;
  CSTS_REQUIRE(n % 4 == 0, "add: n must be a multiple of 4");
  launch_pdl(add_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, (cudaStream_t)stream, a, b, out, n / 4);
  return csts_check_launch("add_f32");
}

int csts_scale_f32(const float* a, const float* device_scalar, float* out, int64_t n, void* stream) {
  if (n == 0) return 0;
  launch_pdl(scale_kernel, dim3(grid_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, a, device_scalar, out, n);
  return csts_check_launch("scale_f32");
}

// out[N] (+)= column sums of X[M][N]; out must be zeroed by the caller unless accumulating
int csts_colsum(const void* X, int x_dtype, float* out, int64_t M, int N, int64_t ld, void* stream) {
  if (M == 0 || N == 0) return 0;
  CSTS_REQUIRE(N % 4 == 0 && ld % 4 == 0, "colsum: N and ld must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int esz = x_dtype == 0 ? 4 : 2;
  if (N % 8 == 0 && (ld * esz) % 16 == 0 && ((uintptr_t)X & 15) == 0) {
    const int groups = N / 8;
    const int bx = ceil_div(groups, 32), GX = ceil_div(groups, bx), RY = 256 / GX;
    int want_y = csts_num_sms() * 4 / bx;
    if (want_y < 1) want_y = 1;
    int rows_per_block = (int)((M + want_y - 1) / want_y);
    if (rows_per_block < 8 * RY) rows_per_block = 8 * RY;
    dim3 grid(bx, ceil_div(M, rows_per_block));
    if (x_dtype == 0) launch_pdl(colsum8_kernel<float>, dim3(grid), dim3(256), 0, st, (const float*)X, out, M, N, ld, rows_per_block, GX);
    else if (x_dtype == CSTS_F16) launch_pdl(colsum8_kernel<f16>, dim3(grid), dim3(256), 0, st, (const f16*)X, out, M, N, ld, rows_per_block, GX);
    else launch_pdl(colsum8_kernel<bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)X, out, M, N, ld, rows_per_block, GX);
    return csts_check_launch("colsum8");
  }
  int bx = ceil_div(N, 128);
  int want_y = csts_num_sms() * 4 / bx;
  if (want_y < 1) want_y = 1;
  int rows_per_block = (int)((M + want_y - 1) / want_y);
  if (rows_per_block < 64) rows_per_block = 64;
  dim3 grid(bx, ceil_div(M, rows_per_block));
  if (x_dtype == 0) launch_pdl(colsum_kernel<float>, dim3(grid), dim3(256), 0, st, (const float*)X, out, M, N, ld, rows_per_block);
  else if (x_dtype == CSTS_F16) launch_pdl(colsum_kernel<f16>, dim3(grid), dim3(256), 0, st, (const f16*)X, out, M, N, ld, rows_per_block);
  else launch_pdl(colsum_kernel<bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)X, out, M, N, ld, rows_per_block);
  return csts_check_launch("colsum");
}

}  // extern "C"
