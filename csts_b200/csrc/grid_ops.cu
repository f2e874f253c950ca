// Token-grid kernels: depthwise 3x3x3 conv / transposed conv over the (T,H,W) token grid with a fused
// LayerNorm(head_dim) epilogue, its weight gradient, the MaxPool3d and trilinear skip paths.
// They read and write the token-major layouts the GEMMs produce ((B, N, 3, heads, d) qkv slices,
// (B, heads, L, d) pooled tensors, (B, N, C) residual streams) directly through element strides, so
// the reference's permute+contiguous copies (attention.py:31,37) never happen.
#include <stdlib.h>

#include "common.cuh"
#include "grid_ops.h"

namespace {

// ------------------------------------------------------------------------------------------------
// Gather-form depthwise conv.
//   regular   : out[o] = sum_tap w[tap] * in[o*s + tap - 1]
//   transposed: out[o] = sum_tap w[tap] * in[(o + 1 - tap) / s]     (when divisible and in range)
// One warp per output position; lane owns channels [4*lane, 4*lane+4) and, for d = 192, also
// [128 + 4*lane, ...).  Weights live in shared memory as [tap][d].
// ------------------------------------------------------------------------------------------------
// Tap geometry for one output coordinate along one axis: which of the 3 taps are in range and the
// input coordinate each one reads.  Strides are powers of two (1,2,4,8,16): shifts, no division.
template <bool TRANSPOSED>
__device__ __forceinline__ void tap_coords(int o, int shift, int n_in, int (&idx)[3], bool (&ok)[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (TRANSPOSED) {
      int num = o + 1 - k;
      idx[k] = num >> shift;
      ok[k] = num >= 0 && (num & ((1 << shift) - 1)) == 0 && idx[k] < n_in;
    } else {
      idx[k] = (o << shift) + k - 1;
      ok[k] = idx[k] >= 0 && idx[k] < n_in;
    }
  }
}

template <int D, bool TRANSPOSED, bool NORM, typename T>
__global__ void __launch_bounds__(256) dwconv_kernel(csts_pool_args p, int lt, int lh, int lw) {
  pdl_wait();
  if (blockIdx.y == 1) {                           // second problem of the launch (same geometry, other tensors)
    p.in = p.in2; p.out = p.out2; p.w = p.w2; p.gamma = p.gamma2; p.beta = p.beta2; p.pre = p.pre2; p.mean = p.mean2; p.rstd = p.rstd2;
  }
  constexpr int NJ = (D + 127) / 128;
  __shared__ float s_w[27 * D];
  for (int i = threadIdx.x; i < 27 * D; i += blockDim.x) {
    int tap = i / D, c = i - tap * D;
    s_w[i] = p.w[c * 27 + tap];               // parameter layout (d,1,3,3,3) -> [tap][c]
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int Lo = p.To * p.Ho * p.Wo;
  const int64_t total = (int64_t)p.B * p.heads * Lo;
  const T* in = reinterpret_cast<const T*>(p.in);
  T* out = reinterpret_cast<T*>(p.out);
  T* pre = reinterpret_cast<T*>(p.pre);
  const int sW = (int)p.in_sP, sH = p.Wi * sW, sT = p.Hi * sH;      // input strides in elements (32-bit)

  // each warp walks a contiguous range of output positions: coordinates are decoded once and then
  // advanced incrementally (no per-position division)
  const int64_t nwarps = (int64_t)gridDim.x * wpb;
  const int64_t per_warp = (total + nwarps - 1) / nwarps;
  int64_t idx = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * per_warp;
  const int64_t idx_end = idx + per_warp < total ? idx + per_warp : total;
  int wo = 0, ho = 0, to = 0, hd = 0, b = 0;
  if (idx < idx_end) {
    int o = (int)(idx % Lo);
    int bh = (int)(idx / Lo);
    hd = bh % p.heads; b = bh / p.heads;
    wo = o % p.Wo; ho = (o / p.Wo) % p.Ho; to = o / (p.Wo * p.Ho);
  }
  for (; idx < idx_end; ++idx) {
    const int o = (to * p.Ho + ho) * p.Wo + wo;
    const T* in_bh = in + b * p.in_sB + hd * p.in_sH;
    float acc[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    int ti[3], hi[3], wi[3];
    bool vt[3], vh[3], vw[3];
    tap_coords<TRANSPOSED>(to, lt, p.Ti, ti, vt);
    tap_coords<TRANSPOSED>(ho, lh, p.Hi, hi, vh);
    tap_coords<TRANSPOSED>(wo, lw, p.Wi, wi, vw);
    const bool any_tap = (vt[0] | vt[1] | vt[2]) & (vh[0] | vh[1] | vh[2]) & (vw[0] | vw[1] | vw[2]);
    // 32-bit element offsets per axis (the whole (b, head) slab is < 2^31 elements, checked on the host):
    // one add per tap instead of 64-bit multiplies
    int ot[3], oh[3], ow[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ot[k] = ti[k] * sT; oh[k] = hi[k] * sH; ow[k] = wi[k] * sW; }
    const T* pl = in_bh + 4 * lane;
    // interior fast path (regular conv): all 9 (kh, kw) taps of a plane are in range -> no per-tap predicates,
    // the 9 loads of a plane are independent and issue back to back
    const bool hw_interior = !TRANSPOSED && vh[0] && vh[2] && vw[0] && vw[2];
    if (hw_interior) {
      const T* c0 = pl + (oh[0] + ow[0]);
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        if (!vt[kt]) continue;
        const T* pk = c0 + ot[kt];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (4 * lane + 128 * j < D) {
            float v[9][4];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) ld4(pk + (kh * sH + kw * sW) + 128 * j, v[kh * 3 + kw]);
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              float w4[4];
              ld4(s_w + (kt * 9 + t9) * D + 4 * lane + 128 * j, w4);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(v[t9][i], w4[i], acc[j][i]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      if (hw_interior || !any_tap || !vt[kt]) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        if (!vh[kh]) continue;
        const int orow = ot[kt] + oh[kh];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          if (!vw[kw]) continue;
          const T* src = pl + (orow + ow[kw]);
          const float* wt = s_w + ((kt * 3 + kh) * 3 + kw) * D + 4 * lane;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            if (4 * lane + 128 * j < D) {
              float v[4], w4[4];
              ld4(src + 128 * j, v);
              ld4(wt + 128 * j, w4);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(v[i], w4[i], acc[j][i]);
            }
          }
        }
      }
    }
    T* dst = out + b * p.out_sB + hd * p.out_sH + (int64_t)o * p.out_sP;
    if (!NORM) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int c = 4 * lane + 128 * j;
        if (c < D) st4(dst + c, acc[j]);
      }
    } else {
      // the raw conv result is rounded to its storage type first (it is what backward re-reads), and the
      // statistics are taken from the rounded values so forward and backward agree exactly
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int c = 4 * lane + 128 * j;
        if (c < D) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { T r16; st_f(&r16, acc[j][i]); acc[j][i] = ld_f(&r16); s += acc[j][i]; }
        }
      }
      const float mean = warp_sum(s) * (1.f / D);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int c = 4 * lane + 128 * j;
        if (c < D) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { float dlt = acc[j][i] - mean; q += dlt * dlt; }
        }
      }
      const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + p.eps);
      if (lane == 0) { p.mean[idx] = mean; p.rstd[idx] = rstd; }
      T* pdst = pre + idx * D;                 // pre is dense (B, heads, Lo, d)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        int c = 4 * lane + 128 * j;
        if (c < D) {
          float g[4], be[4], y[4];
          ld4(p.gamma + c, g);
          ld4(p.beta + c, be);
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = (acc[j][i] - mean) * rstd * g[i] + be[i];
          st4(pdst + c, acc[j]);
          st4(dst + c, y);
        }
      }
    }
    if (++wo == p.Wo) { wo = 0; if (++ho == p.Ho) { ho = 0; if (++to == p.To) { to = 0; if (++hd == p.heads) { hd = 0; ++b; } } } }
  }
}

// ------------------------------------------------------------------------------------------------
// T-column variant for compile-time temporal extents (TI input planes, TO output planes, temporal stride
// ST): every pooling conv of the model has T = 4 with temporal stride 1, and the decoder's last
// up-sampling conv goes 4 -> 8 planes with stride 2.  A warp owns one (b, head, ho, wo) column and produces
// its TO outputs together: each valid spatial tap loads and unpacks its TI input planes once and feeds up
// to three outputs per plane, the three kt weights of the tap are read from shared memory once per column,
// and the spatial tap geometry is decoded once per column instead of once per output.  The kernel is
// instruction-issue bound (16-bit -> f32 unpacking + FMAs on L2-resident data), so this is what counts.
//   regular   : ti = to * ST + kt - 1
//   transposed: ti = (to + 1 - kt) / ST  <=>  to = ti * ST + kt - 1
// ------------------------------------------------------------------------------------------------
template <int D, bool TRANSPOSED, bool NORM, typename T, int TI, int TO, int ST>
__global__ void __launch_bounds__(256) dwconv_tcol_kernel(csts_pool_args p, int lh, int lw) {
  pdl_wait();
  if (blockIdx.y == 1) {                           // second problem of the launch (same geometry, other tensors)
    p.in = p.in2; p.out = p.out2; p.w = p.w2; p.gamma = p.gamma2; p.beta = p.beta2; p.pre = p.pre2; p.mean = p.mean2; p.rstd = p.rstd2;
  }
  constexpr int NJ = (D + 127) / 128;
  // the kernel in the activations' 16-bit type (what conv3d computes with under the reference's autocast): products
  // of two 16-bit values then go through the mixed-precision FMA with no unpack instruction (common.cuh::fhfma2)
  __shared__ __align__(16) T s_w[27 * D];
  for (int i = threadIdx.x; i < 27 * D; i += blockDim.x) {
    int tap = i / D, c = i - tap * D;
    st_f(&s_w[i], p.w[c * 27 + tap]);         // parameter layout (d,1,3,3,3) -> [tap][c]
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int HWo = p.Ho * p.Wo;
  const int64_t total = (int64_t)p.B * p.heads * HWo;                // columns
  const T* in = reinterpret_cast<const T*>(p.in);
  T* out = reinterpret_cast<T*>(p.out);
  T* pre = reinterpret_cast<T*>(p.pre);
  const int sW = (int)p.in_sP, sH = p.Wi * sW, sT = p.Hi * sH;      // input strides in elements (32-bit)
  const int64_t oT = (int64_t)HWo * p.out_sP;                       // output stride between T planes

  const int64_t nwarps = (int64_t)gridDim.x * wpb;
  const int64_t per_warp = (total + nwarps - 1) / nwarps;
  int64_t idx = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * per_warp;
  const int64_t idx_end = idx + per_warp < total ? idx + per_warp : total;
  int wo = 0, ho = 0, hd = 0, b = 0;
  if (idx < idx_end) {
    int o = (int)(idx % HWo);
    int bh = (int)(idx / HWo);
    hd = bh % p.heads; b = bh / p.heads;
    wo = o % p.Wo; ho = o / p.Wo;
  }
  for (; idx < idx_end; ++idx) {
    float acc[TO][NJ][4];
#pragma unroll
    for (int t = 0; t < TO; ++t)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[t][j][i] = 0.f;
    int hi[3], wi[3];
    bool vh[3], vw[3];
    tap_coords<TRANSPOSED>(ho, lh, p.Hi, hi, vh);
    tap_coords<TRANSPOSED>(wo, lw, p.Wi, wi, vw);
    const T* pl = in + b * p.in_sB + hd * p.in_sH + 4 * lane;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      if (!vh[kh]) continue;                                       // warp-uniform
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        if (!vw[kw]) continue;
        const T* src = pl + (hi[kh] * sH + wi[kw] * sW);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (4 * lane + 128 * j < D) {
            uint2 v[TI], w3[3];
#pragma unroll
            for (int ti = 0; ti < TI; ++ti) v[ti] = *reinterpret_cast<const uint2*>(src + ti * sT + 128 * j);
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) w3[kt] = *reinterpret_cast<const uint2*>(s_w + ((kt * 3 + kh) * 3 + kw) * D + 4 * lane + 128 * j);
            // (plane, kt) -> output pairing, resolved at compile time
#pragma unroll
            for (int a = 0; a < (TRANSPOSED ? TI : TO); ++a)
#pragma unroll
              for (int kt = 0; kt < 3; ++kt) {
                const int other = a * ST + kt - 1;                 // transposed: a = ti, other = to;  regular: a = to, other = ti
                if (other < 0 || other >= (TRANSPOSED ? TO : TI)) continue;
                const int ti = TRANSPOSED ? a : other, to = TRANSPOSED ? other : a;
                fhfma2<T>(acc[to][j][0], acc[to][j][1], v[ti].x, w3[kt].x);
                fhfma2<T>(acc[to][j][2], acc[to][j][3], v[ti].y, w3[kt].y);
              }
          }
        }
      }
    }
    const int o_hw = ho * p.Wo + wo;
    T* dst0 = out + b * p.out_sB + hd * p.out_sH + (int64_t)o_hw * p.out_sP;
    const int64_t row0 = ((int64_t)(b * p.heads + hd) * TO) * HWo + o_hw;     // dense (B, heads, T*Ho*Wo) index of to = 0
#pragma unroll
    for (int to = 0; to < TO; ++to) {
      T* dst = dst0 + to * oT;
      if (!NORM) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          int c = 4 * lane + 128 * j;
          if (c < D) st4(dst + c, acc[to][j]);
        }
      } else {
        // statistics from the storage-rounded conv output (what backward re-reads), as in dwconv_kernel
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          int c = 4 * lane + 128 * j;
          if (c < D) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { T r16; st_f(&r16, acc[to][j][i]); acc[to][j][i] = ld_f(&r16); sum += acc[to][j][i]; }
          }
        }
        const float mean = warp_sum(sum) * (1.f / D);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          int c = 4 * lane + 128 * j;
          if (c < D) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { float dlt = acc[to][j][i] - mean; q += dlt * dlt; }
          }
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + p.eps);
        const int64_t row = row0 + (int64_t)to * HWo;
        if (lane == 0) { p.mean[row] = mean; p.rstd[row] = rstd; }
        T* pdst = pre + row * D;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          int c = 4 * lane + 128 * j;
          if (c < D) {
            float g[4], be[4], y[4];           // (re-loaded per output: keeping them in registers costs occupancy — measured slower)
            ld4(p.gamma + c, g);
            ld4(p.beta + c, be);
#pragma unroll
            for (int i = 0; i < 4; ++i) y[i] = (acc[to][j][i] - mean) * rstd * g[i] + be[i];
            st4(pdst + c, acc[to][j]);
            st4(dst + c, y);
          }
        }
      }
    }
    if (++wo == p.Wo) { wo = 0; if (++ho == p.Ho) { ho = 0; if (++hd == p.heads) { hd = 0; ++b; } } }
  }
}

// ------------------------------------------------------------------------------------------------
// Depthwise weight gradient:  dw[c][tap] += sum_{b,head,o} small[o][c] * big[o*s + tap - 1][c]
// (conv: small = d(conv out), big = conv in;  transposed conv: small = conv in, big = d(out)).
// ------------------------------------------------------------------------------------------------
// Block = 9 warps; warp w owns the tap row (kt, kh) = (w / 3, w % 3) with its 3 kw taps, so no
// cross-warp reduction is needed and each lane carries only 3 x 4 accumulators per channel group.
// A block walks a contiguous range of `small` ROWS (b, head, to, ho); inside a row the three kw taps
// slide over `big` through registers: stride 1 re-uses two of three loads, stride 2 one of three.
template <int D, int SW, typename TS, typename TB>
__global__ void __launch_bounds__(288) dwconv_wgrad_kernel(csts_wgrad_args p, int lt, int lh, int lw, int rows_per_block) {
  pdl_wait();
  if (blockIdx.y == 1) { p.small = p.small2; p.big = p.big2; p.dw = p.dw2; }   // second problem of the launch
  constexpr int NJ = (D + 127) / 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kt = warp / 3, kh = warp % 3;
  const int64_t rows_total = (int64_t)p.B * p.heads * p.Ts * p.Hs;
  const TS* small = reinterpret_cast<const TS*>(p.small);
  const TB* big = reinterpret_cast<const TB*>(p.big);
  float acc[NJ][3][4];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][t][i] = 0.f;
  const int64_t beg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t end = beg + rows_per_block < rows_total ? beg + rows_per_block : rows_total;
  int ho = 0, to = 0, hd = 0, b = 0;
  if (beg < end) {
    int64_t r = beg;
    ho = (int)(r % p.Hs); r /= p.Hs;
    to = (int)(r % p.Ts); r /= p.Ts;
    hd = (int)(r % p.heads); b = (int)(r / p.heads);
  }
  for (int64_t row = beg; row < end; ++row) {
    const int ti = (to << lt) + kt - 1, hi = (ho << lh) + kh - 1;
    if (ti >= 0 && ti < p.Tb && hi >= 0 && hi < p.Hb) {                  // warp-uniform
      const TS* srow = small + b * p.small_sB + hd * p.small_sH + (int64_t)((to * p.Hs + ho) * p.Ws) * p.small_sP;
      const TB* brow = big + b * p.big_sB + hd * p.big_sH + (int64_t)((ti * p.Hb + hi) * p.Wb) * p.big_sP;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = 4 * lane + 128 * j;
        if (c >= D) continue;
        const TB* bcol = brow + c;
        const TS* scol = srow + c;
        const int bsP = (int)p.big_sP, ssP = (int)p.small_sP;
        auto ld_big = [&](int wi, float (&v)[4]) {
          if (wi >= 0 && wi < p.Wb) ld4(bcol + wi * bsP, v);
          else { v[0] = v[1] = v[2] = v[3] = 0.f; }
        };
        float w0[4], w1[4], w2[4];           // big[wi0], big[wi0+1], big[wi0+2] with wi0 = wo*sw - 1
        if (SW == 1) { ld_big(-1, w0); ld_big(0, w1); }
        if (SW == 2) { ld_big(-1, w0); }
#pragma unroll 4
        for (int wo = 0; wo < p.Ws; ++wo) {
          float sv[4];
          ld4(scol + wo * ssP, sv);
          const int wi0 = (wo << lw) - 1;
          if (SW == 1) ld_big(wi0 + 2, w2);
          else if (SW == 2) { ld_big(wi0 + 1, w1); ld_big(wi0 + 2, w2); }
          else { ld_big(wi0, w0); ld_big(wi0 + 1, w1); ld_big(wi0 + 2, w2); }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[j][0][i] = fmaf(w0[i], sv[i], acc[j][0][i]);
            acc[j][1][i] = fmaf(w1[i], sv[i], acc[j][1][i]);
            acc[j][2][i] = fmaf(w2[i], sv[i], acc[j][2][i]);
          }
          if (SW == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { w0[i] = w1[i]; w1[i] = w2[i]; }
          } else if (SW == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) w0[i] = w2[i];
          }
        }
      }
    }
    if (++ho == p.Hs) { ho = 0; if (++to == p.Ts) { to = 0; if (++hd == p.heads) { hd = 0; ++b; } } }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    int c = 4 * lane + 128 * j;
    if (c < D) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(p.dw + (c + i) * 27 + (kt * 3 + kh) * 3 + kw, acc[j][kw][i]);
    }
  }
}

// T-column weight gradient (d = 96, compile-time temporal extents: TSM planes of `small`, TBG planes of `big`,
// temporal stride ST).  Three warps share one (b, head, hs, ws) column of `small`, one per kernel row kh: a warp
// issues its 4 + 3*TBG independent loads (the column of `small`, the kw = 0..2 columns of its `big` row) before
// touching any of them, then accumulates its 9 taps (kt x kw) x 4 channels in registers.  The kernel is
// latency-bound on L2-resident data, so what matters is loads in flight: ~16 per warp, 18 warps per SM.
// Warps of a block combine through shared memory, then one global atomic per weight per block.
template <typename TS, typename TB, int TSM, int TBG, int ST>
__global__ void __launch_bounds__(192) dwconv_wgrad_tcol_kernel(csts_wgrad_args p, int lh, int lw, int cols_per_slot) {
  pdl_wait();
  // blockIdx.y = (problem of the launch) * groups + (96-channel group of the head): depthwise channels are independent, so a
  // 192-channel head is two 96-channel problems whose operands start 96 elements further on
  const int groups = gridDim.y / (p.small2 ? 2 : 1);
  if ((int)blockIdx.y >= groups) { p.small = p.small2; p.big = p.big2; p.dw = p.dw2; }
  constexpr int D = 96;
  const int cg = blockIdx.y % groups;
  __shared__ float s_dw[27 * D];
  for (int i = threadIdx.x; i < 27 * D; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kh = warp % 3, slot = warp / 3, slots = (blockDim.x >> 5) / 3;
  const bool active = 4 * lane < D;
  const TS* small = reinterpret_cast<const TS*>(p.small) + cg * D;
  const TB* big = reinterpret_cast<const TB*>(p.big) + cg * D;
  float* dw = p.dw + cg * D * 27;
  const int HWs = p.Hs * p.Ws;
  const int64_t total = (int64_t)p.B * p.heads * HWs;
  const int ssP = (int)p.small_sP, bsP = (int)p.big_sP;
  const int ssT = HWs * ssP, bsH = p.Wb * bsP, bsT = p.Hb * bsH;
  float acc[3][3][4];                                                   // [kt][kw][channel]
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b2 = 0; b2 < 3; ++b2)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[a][b2][i] = 0.f;
  int64_t idx = ((int64_t)blockIdx.x * slots + slot) * cols_per_slot;
  const int64_t idx_end = idx + cols_per_slot < total ? idx + cols_per_slot : total;
  int ws = 0, hs = 0, hd = 0, b = 0;
  if (idx < idx_end) {
    int o = (int)(idx % HWs);
    int bh = (int)(idx / HWs);
    hd = bh % p.heads; b = bh / p.heads;
    ws = o % p.Ws; hs = o / p.Ws;
  }
  for (; idx < idx_end; ++idx) {
    const int hb = (hs << lh) + kh - 1;
    if (active && hb >= 0 && hb < p.Hb) {                               // warp-uniform
      const TS* sp = small + b * p.small_sB + hd * p.small_sH + (hs * p.Ws + ws) * ssP + 4 * lane;
      const TB* bp = big + b * p.big_sB + hd * p.big_sH + hb * bsH + 4 * lane;
      uint2 sraw[TSM], braw[3][TBG];
#pragma unroll
      for (int t = 0; t < TSM; ++t) sraw[t] = __ldg(reinterpret_cast<const uint2*>(sp + t * ssT));
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wb = (ws << lw) + kw - 1;
        const bool ok = wb >= 0 && wb < p.Wb;
#pragma unroll
        for (int t = 0; t < TBG; ++t) braw[kw][t] = ok ? __ldg(reinterpret_cast<const uint2*>(bp + wb * bsP + t * bsT)) : make_uint2(0u, 0u);
      }
      // both factors are 16-bit and of one type (the launcher guarantees it): mixed-precision FMAs on the packed pairs
      static_assert(sizeof(TS) == 2 && sizeof(TB) == 2, "16-bit operands");
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int kt = 0; kt < 3; ++kt)
#pragma unroll
          for (int ts = 0; ts < TSM; ++ts) {
            const int tb = ts * ST + kt - 1;
            if (tb < 0 || tb >= TBG) continue;                          // resolved at compile time
            fhfma2<TS>(acc[kt][kw][0], acc[kt][kw][1], sraw[ts].x, braw[kw][tb].x);
            fhfma2<TS>(acc[kt][kw][2], acc[kt][kw][3], sraw[ts].y, braw[kw][tb].y);
          }
    }
    if (++ws == p.Ws) { ws = 0; if (++hs == p.Hs) { hs = 0; if (++hd == p.heads) { hd = 0; ++b; } } }
  }
  if (active) {
#pragma unroll
    for (int kt = 0; kt < 3; ++kt)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(&s_dw[((kt * 3 + kh) * 3 + kw) * D + 4 * lane + i], acc[kt][kw][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * D; i += blockDim.x) {
    int tap = i / D, c = i - tap * D;
    atomicAdd(dw + c * 27 + tap, s_dw[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// MaxPool3d k(1,3,3) s(1,2,2) p(0,1,1) on a token-major f32 stream (B, T*H*W, C).
// ref: attention.py:225-236 (pool_skip).  `arg` records the winning tap (0..8) per output element.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ arg,
                                                          int B, int T, int H, int W, int C) {
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const uint32_t total = (uint32_t)B * T * Ho * Wo * C4;               // < 2^32, checked on the host
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t q = i;
    const int c = (int)divmod(q, C4) * 4;
    const int64_t pos = q;
    const int wo = (int)divmod(q, Wo), ho = (int)divmod(q, Ho);
    const int64_t bt = q;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {0, 0, 0, 0};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int hi = ho * 2 + kh - 1;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int wi = wo * 2 + kw - 1;
        if (wi < 0 || wi >= W) continue;
        float v[4];
        ld4(x + ((bt * H + hi) * W + wi) * C + c, v);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (v[k] > best[k]) { best[k] = v[k]; bi[k] = kh * 3 + kw; }
      }
    }
    st4(y + pos * C + c, best);
    if (arg) *reinterpret_cast<uchar4*>(arg + pos * C + c) = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
  }
}
// dx[i] = sum over the (<= 4) windows containing i of dy[o] * [arg[o] == tap(i, o)]
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ arg,
                                                          float* __restrict__ dx, int B, int T, int H, int W, int C) {
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const uint32_t total = (uint32_t)B * T * H * W * C4;                 // < 2^32, checked on the host
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t q = i;
    const int c = (int)divmod(q, C4) * 4;
    const int64_t pos = q;
    const int wi = (int)divmod(q, W), hi = (int)divmod(q, H);
    const int64_t bt = q;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int num = hi + 1 - kh;
      if (num < 0 || (num & 1)) continue;
      int ho = num >> 1;
      if (ho >= Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int numw = wi + 1 - kw;
        if (numw < 0 || (numw & 1)) continue;
        int wo = numw >> 1;
        if (wo >= Wo) continue;
        int64_t op = ((bt * Ho + ho) * Wo + wo) * C + c;
        uchar4 a = *reinterpret_cast<const uchar4*>(arg + op);
        float v[4];
        ld4(dy + op, v);
        int tap = kh * 3 + kw;
        if (a.x == tap) acc[0] += v[0];
        if (a.y == tap) acc[1] += v[1];
        if (a.z == tap) acc[2] += v[2];
        if (a.w == tap) acc[3] += v[3];
      }
    }
    st4(dx + pos * C + c, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// Trilinear up-sampling (align_corners=False) by integer factors (ft, fh, fw) on a token-major f32
// stream.  ref: attention.py:463-467 (nn.Upsample) and custom_multimodal_builder.py:479.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lin_coef(int o, int f, int n_in, int& i0, int& i1, float& lam) {
  if (f == 1) { i0 = i1 = o; lam = 0.f; return; }
  float src = ((float)o + 0.5f) / (float)f - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = i0 + 1 < n_in ? i0 + 1 : n_in - 1;
  lam = src - (float)i0;
}

__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int H, int W,
                                                           int C, int ft, int fh, int fw) {
  pdl_wait();
  const int To = T * ft, Ho = H * fh, Wo = W * fw, C4 = C / 4;
  const uint32_t total = (uint32_t)B * To * Ho * Wo * C4;              // < 2^32, checked on the host
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t q = i;
    const int c = (int)divmod(q, C4) * 4;
    const int64_t pos = q;
    const int wo = (int)divmod(q, Wo), ho = (int)divmod(q, Ho), to = (int)divmod(q, To);
    const int64_t b = q;
    int t0, t1, h0, h1, w0, w1;
    float lt, lh, lw;
    lin_coef(to, ft, T, t0, t1, lt);
    lin_coef(ho, fh, H, h0, h1, lh);
    lin_coef(wo, fw, W, w0, w1, lw);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      float ct = a ? lt : 1.f - lt;
      if (ct == 0.f) continue;
      int ti = a ? t1 : t0;
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        float ch = bb ? lh : 1.f - lh;
        if (ch == 0.f) continue;
        int hi = bb ? h1 : h0;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          float cw = cc ? lw : 1.f - lw;
          if (cw == 0.f) continue;
          int wi = cc ? w1 : w0;
          float v[4];
          ld4(x + (((b * T + ti) * H + hi) * W + wi) * C + c, v);
          float wgt = ct * ch * cw;
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[k] = fmaf(wgt, v[k], acc[k]);
        }
      }
    }
    st4(y + pos * C + c, acc);
  }
}
// gather form of the adjoint: each input position sums the outputs that interpolate from it
__device__ __forceinline__ float lin_adj(int i, int o, int f, int n_in) {
  int i0, i1; float lam;
  lin_coef(o, f, n_in, i0, i1, lam);
  float w = 0.f;
  if (i0 == i) w += 1.f - lam;
  if (i1 == i) w += lam;
  return w;
}
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int T, int H, int W,
                                                           int C, int ft, int fh, int fw, int accumulate) {
  pdl_wait();
  const int To = T * ft, Ho = H * fh, Wo = W * fw, C4 = C / 4;
  const uint32_t total = (uint32_t)B * T * H * W * C4;                 // < 2^32, checked on the host
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t q = i;
    const int c = (int)divmod(q, C4) * 4;
    const int64_t pos = q;
    const int wi = (int)divmod(q, W), hi = (int)divmod(q, H), ti = (int)divmod(q, T);
    const int64_t b = q;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int tlo = ft == 1 ? ti : max(ti * ft - ft, 0), thi = ft == 1 ? ti : min(ti * ft + 2 * ft - 1, To - 1);
    int hlo = fh == 1 ? hi : max(hi * fh - fh, 0), hhi = fh == 1 ? hi : min(hi * fh + 2 * fh - 1, Ho - 1);
    int wlo = fw == 1 ? wi : max(wi * fw - fw, 0), whi = fw == 1 ? wi : min(wi * fw + 2 * fw - 1, Wo - 1);
    // the adjoint weights depend on one axis each: evaluate them once per axis (<= 3 f candidates, f <= 2 on the path),
    // not once per (to, ho, wo) triple
    constexpr int MAXC = 6;
    if (hhi - hlo < MAXC && whi - wlo < MAXC) {
      float chv[MAXC], cwv[MAXC];
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        chv[k] = (hlo + k <= hhi) ? lin_adj(hi, hlo + k, fh, H) : 0.f;
        cwv[k] = (wlo + k <= whi) ? lin_adj(wi, wlo + k, fw, W) : 0.f;
      }
      for (int to = tlo; to <= thi; ++to) {
        const float ct = lin_adj(ti, to, ft, T);
        if (ct == 0.f) continue;
#pragma unroll
        for (int kh = 0; kh < MAXC; ++kh) {
          if (chv[kh] == 0.f) continue;
#pragma unroll
          for (int kw = 0; kw < MAXC; ++kw) {
            if (cwv[kw] == 0.f) continue;
            float v[4];
            ld4(dy + (((b * To + to) * Ho + (hlo + kh)) * Wo + (wlo + kw)) * C + c, v);
            const float wgt = ct * chv[kh] * cwv[kw];
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = fmaf(wgt, v[k], acc[k]);
          }
        }
      }
    } else {
      for (int to = tlo; to <= thi; ++to) {
        float ct = lin_adj(ti, to, ft, T);
        if (ct == 0.f) continue;
        for (int ho = hlo; ho <= hhi; ++ho) {
          float ch = lin_adj(hi, ho, fh, H);
          if (ch == 0.f) continue;
          for (int wo = wlo; wo <= whi; ++wo) {
            float cw = lin_adj(wi, wo, fw, W);
            if (cw == 0.f) continue;
            float v[4];
            ld4(dy + (((b * To + to) * Ho + ho) * Wo + wo) * C + c, v);
            float wgt = ct * ch * cw;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = fmaf(wgt, v[k], acc[k]);
          }
        }
      }
    }
    float* d = dx + pos * C + c;
    if (accumulate) {
      float o[4];
      ld4(d, o);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += o[k];
    }
    st4(d, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// The factors that occur on the path are (1,2,2) (decoder blocks 1-3) and (2,1,1) (decoder block 4).  For a factor of two
// the half-pixel (align_corners=False) weights are constants: along an axis of n inputs
//   y[2i]   = 0.25 x[i-1] + 0.75 x[i]   (y[0]    = x[0]),      y[2i+1] = 0.75 x[i] + 0.25 x[i+1]   (y[2n-1] = x[n-1]).
// Forward: a thread owns one INPUT cell (4 channels), loads its <= 3x3 (or 3) neighbourhood once and writes the 2x2 (or 2)
// outputs it centres.  Adjoint: a thread owns one input cell and gathers its 4 taps per doubled axis,
//   dx[i] = 0.25 dy[2i-1] + w0 dy[2i] + w1 dy[2i+1] + 0.25 dy[2i+2],  w0 = (i == 0 ? 1 : 0.75),  w1 = (i == n-1 ? 1 : 0.75).
// No divisions, no coefficient evaluation: the generic kernels above spend most of their time there.
// ------------------------------------------------------------------------------------------------
template <bool UP_T>       // false: (1,2,2)   true: (2,1,1)
__global__ void __launch_bounds__(256) upsample2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int H, int W,
                                                            int C) {
  pdl_wait();
  const int C4 = C / 4;
  const uint32_t total = (uint32_t)B * T * H * W * C4;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t q = i;
    const int c = (int)divmod(q, C4) * 4;
    const int wi = (int)divmod(q, W), hi = (int)divmod(q, H), ti = (int)divmod(q, T);
    const int64_t b = q;
    if (UP_T) {
      const int64_t plane = (int64_t)H * W * C;
      const float* px = x + ((b * T + ti) * H + hi) * (int64_t)W * C + (int64_t)wi * C + c;
      float m[4], lo[4], hi4[4];
      ld4(px, m);
      ld4(ti > 0 ? px - plane : px, lo);
      ld4(ti < T - 1 ? px + plane : px, hi4);
      float o0[4], o1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o0[k] = ti > 0 ? 0.25f * lo[k] + 0.75f * m[k] : m[k];
        o1[k] = ti < T - 1 ? 0.75f * m[k] + 0.25f * hi4[k] : m[k];
      }
      float* py = y + ((b * 2 * T + 2 * ti) * H + hi) * (int64_t)W * C + (int64_t)wi * C + c;
      st4(py, o0);
      st4(py + plane, o1);
    } else {
      const int64_t row = (int64_t)W * C;
      const float* pc = x + ((b * T + ti) * H + hi) * row + (int64_t)wi * C + c;
      const int dh0 = hi > 0 ? -1 : 0, dh1 = hi < H - 1 ? 1 : 0, dw0 = wi > 0 ? -1 : 0, dw1 = wi < W - 1 ? 1 : 0;
      float v[3][3][4];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) ld4(pc + (a == 0 ? dh0 : (a == 2 ? dh1 : 0)) * row + (bb == 0 ? dw0 : (bb == 2 ? dw1 : 0)) * C, v[a][bb]);
      // separable: rows first (two output rows), then columns (two output columns)
      const float ha0 = hi > 0 ? 0.25f : 0.f, ha1 = hi > 0 ? 0.75f : 1.f;          // output row 2hi   = ha0 * up + ha1 * mid
      const float hb1 = hi < H - 1 ? 0.75f : 1.f, hb2 = hi < H - 1 ? 0.25f : 0.f;  // output row 2hi+1 = hb1 * mid + hb2 * down
      const float wa0 = wi > 0 ? 0.25f : 0.f, wa1 = wi > 0 ? 0.75f : 1.f;
      const float wb1 = wi < W - 1 ? 0.75f : 1.f, wb2 = wi < W - 1 ? 0.25f : 0.f;
      float* py = y + ((b * T + ti) * 2 * H + 2 * hi) * (2 * row) + (int64_t)(2 * wi) * C + c;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float rowv[3][4];
#pragma unroll
        for (int bb = 0; bb < 3; ++bb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            rowv[bb][k] = r == 0 ? ha0 * v[0][bb][k] + ha1 * v[1][bb][k] : hb1 * v[1][bb][k] + hb2 * v[2][bb][k];
        float o0[4], o1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          o0[k] = wa0 * rowv[0][k] + wa1 * rowv[1][k];
          o1[k] = wb1 * rowv[1][k] + wb2 * rowv[2][k];
        }
        st4(py + r * (2 * row), o0);
        st4(py + r * (2 * row) + C, o1);
      }
    }
  }
}

template <bool UP_T>
__global__ void __launch_bounds__(256) upsample2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int T, int H, int W,
                                                            int C, int accumulate) {
  pdl_wait();
  const int C4 = C / 4;
  const uint32_t total = (uint32_t)B * T * H * W * C4;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t q = i;
    const int c = (int)divmod(q, C4) * 4;
    const int64_t pos = q;
    const int wi = (int)divmod(q, W), hi = (int)divmod(q, H), ti = (int)divmod(q, T);
    const int64_t b = q;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (UP_T) {
      const int64_t plane = (int64_t)H * W * C;
      const float* p0 = dy + ((b * 2 * T + 2 * ti) * H + hi) * (int64_t)W * C + (int64_t)wi * C + c;   // output plane 2ti
      const float w[4] = {ti > 0 ? 0.25f : 0.f, ti > 0 ? 0.75f : 1.f, ti < T - 1 ? 0.75f : 1.f, ti < T - 1 ? 0.25f : 0.f};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (w[a] == 0.f) continue;
        float v[4];
        ld4(p0 + (a - 1) * plane, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fmaf(w[a], v[k], acc[k]);
      }
    } else {
      const int64_t row = 2 * (int64_t)W * C;                       // one output row
      const float* p0 = dy + ((b * T + ti) * 2 * H + 2 * hi) * row + (int64_t)(2 * wi) * C + c;           // output (2hi, 2wi)
      const float wh[4] = {hi > 0 ? 0.25f : 0.f, hi > 0 ? 0.75f : 1.f, hi < H - 1 ? 0.75f : 1.f, hi < H - 1 ? 0.25f : 0.f};
      const float ww[4] = {wi > 0 ? 0.25f : 0.f, wi > 0 ? 0.75f : 1.f, wi < W - 1 ? 0.75f : 1.f, wi < W - 1 ? 0.25f : 0.f};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (wh[a] == 0.f) continue;
        float racc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          if (ww[bb] == 0.f) continue;
          float v[4];
          ld4(p0 + (a - 1) * row + (bb - 1) * C, v);
#pragma unroll
          for (int k = 0; k < 4; ++k) racc[k] = fmaf(ww[bb], v[k], racc[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fmaf(wh[a], racc[k], acc[k]);
      }
    }
    float* d = dx + pos * C + c;
    if (accumulate) {
      float o[4];
      ld4(d, o);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += o[k];
    }
    st4(d, acc);
  }
}

int grid_for(int64_t work_items, int per_block) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)csts_num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

int log2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

template <int D, typename T>
int launch_dwconv(const csts_pool_args& p, cudaStream_t st) {
  int64_t total = (int64_t)p.B * p.heads * p.To * p.Ho * p.Wo;
  int grid = grid_for(total, 8);
  bool norm = p.gamma != nullptr;
  const int ny = p.in2 ? 2 : 1;
  int lt = log2_exact(p.st), lh = log2_exact(p.sh), lw = log2_exact(p.sw);
  CSTS_REQUIRE(lt >= 0 && lh >= 0 && lw >= 0, "dwconv: strides must be powers of two (%d,%d,%d)", p.st, p.sh, p.sw);
  static const bool no_tcol = getenv("CSTS_NO_TCOL") != nullptr;       // A/B tuning runs only
  // T-column kernels: one warp pass per (h, w) column
  const int64_t cols = (int64_t)p.B * p.heads * p.Ho * p.Wo;
  const int cgrid = grid_for(cols, 8);
  if (p.st == 1 && p.Ti == 4 && p.To == 4 && !no_tcol) {
    if (p.transposed) {
      if (norm) launch_pdl(dwconv_tcol_kernel<D, true, true, T, 4, 4, 1>, dim3(cgrid, ny), dim3(256), 0, st, p, lh, lw);
      else launch_pdl(dwconv_tcol_kernel<D, true, false, T, 4, 4, 1>, dim3(cgrid, ny), dim3(256), 0, st, p, lh, lw);
    } else {
      if (norm) launch_pdl(dwconv_tcol_kernel<D, false, true, T, 4, 4, 1>, dim3(cgrid, ny), dim3(256), 0, st, p, lh, lw);
      else launch_pdl(dwconv_tcol_kernel<D, false, false, T, 4, 4, 1>, dim3(cgrid, ny), dim3(256), 0, st, p, lh, lw);
    }
    return csts_check_launch("dwconv_tcol");
  }
  if (D == 96 && p.st == 2 && !no_tcol) {                               // the decoder's temporal up-sampling conv and its adjoint
    if (p.transposed && norm && p.Ti == 4 && p.To == 8) {
      launch_pdl(dwconv_tcol_kernel<96, true, true, T, 4, 8, 2>, dim3(cgrid, ny), dim3(256), 0, st, p, lh, lw);
      return csts_check_launch("dwconv_tcol");
    }
    if (!p.transposed && !norm && p.Ti == 8 && p.To == 4) {
      launch_pdl(dwconv_tcol_kernel<96, false, false, T, 8, 4, 2>, dim3(cgrid, ny), dim3(256), 0, st, p, lh, lw);
      return csts_check_launch("dwconv_tcol");
    }
  }
  if (p.transposed) {
    if (norm) launch_pdl(dwconv_kernel<D, true, true, T>, dim3(grid, ny), dim3(256), 0, st, p, lt, lh, lw);
    else launch_pdl(dwconv_kernel<D, true, false, T>, dim3(grid, ny), dim3(256), 0, st, p, lt, lh, lw);
  } else {
    if (norm) launch_pdl(dwconv_kernel<D, false, true, T>, dim3(grid, ny), dim3(256), 0, st, p, lt, lh, lw);
    else launch_pdl(dwconv_kernel<D, false, false, T>, dim3(grid, ny), dim3(256), 0, st, p, lt, lh, lw);
  }
  return csts_check_launch("dwconv");
}

}  // namespace

extern "C" {

int csts_dwconv(const csts_pool_args* p, void* stream) {
  CSTS_REQUIRE(p->d == 96 || p->d == 192, "dwconv: head_dim %d unsupported (96 or 192)", p->d);
  CSTS_REQUIRE(p->in_sB % 4 == 0 && p->in_sH % 4 == 0 && p->in_sP % 4 == 0 && p->out_sB % 4 == 0 && p->out_sH % 4 == 0 &&
                   p->out_sP % 4 == 0, "dwconv: strides must be multiples of 4 elements");
  CSTS_REQUIRE(((uintptr_t)p->in & 7) == 0 && ((uintptr_t)p->out & 7) == 0, "dwconv: in/out must be 8-byte aligned");
  if (p->gamma) CSTS_REQUIRE(p->beta && p->pre && p->mean && p->rstd, "dwconv: norm epilogue needs beta/pre/mean/rstd");
  if (p->in2) {
    CSTS_REQUIRE(p->out2 && p->w2 && ((uintptr_t)p->in2 & 7) == 0 && ((uintptr_t)p->out2 & 7) == 0, "dwconv: second problem needs in2/out2/w2 (8-byte aligned)");
    if (p->gamma) CSTS_REQUIRE(p->gamma2 && p->beta2 && p->pre2 && p->mean2 && p->rstd2, "dwconv: second problem needs its own norm tensors");
  }
  if ((int64_t)p->B * p->heads * p->To * p->Ho * p->Wo == 0) return 0;
  CSTS_REQUIRE((int64_t)p->Ti * p->Hi * p->Wi * p->in_sP < (1LL << 31), "dwconv: one (batch, head) slab must stay below 2^31 elements");
  CSTS_REQUIRE(p->dtype == CSTS_BF16 || p->dtype == CSTS_F16, "dwconv: dtype must be 1 (bf16) or 2 (f16)");
  if (p->dtype == CSTS_F16)
    return p->d == 96 ? launch_dwconv<96, f16>(*p, (cudaStream_t)stream) : launch_dwconv<192, f16>(*p, (cudaStream_t)stream);
  return p->d == 96 ? launch_dwconv<96, bf16>(*p, (cudaStream_t)stream) : launch_dwconv<192, bf16>(*p, (cudaStream_t)stream);
}

int csts_dwconv_wgrad(const csts_wgrad_args* p, void* stream) {
  CSTS_REQUIRE(p->d == 96 || p->d == 192, "dwconv_wgrad: head_dim %d unsupported", p->d);
  int64_t total = (int64_t)p->B * p->heads * p->Ts * p->Hs * p->Ws;
  if (total == 0) return 0;
  if (p->small2) CSTS_REQUIRE(p->big2 && p->dw2, "dwconv_wgrad: second problem needs small2/big2/dw2");
  int lt = log2_exact(p->st), lh = log2_exact(p->sh), lw = log2_exact(p->sw);
  CSTS_REQUIRE(lt >= 0 && lh >= 0 && lw >= 0, "dwconv_wgrad: strides must be powers of two");
  CSTS_REQUIRE((p->small_dtype == CSTS_BF16 || p->small_dtype == CSTS_F16) && (p->big_dtype == CSTS_BF16 || p->big_dtype == CSTS_F16),
               "dwconv_wgrad: operand dtypes must be 1 (bf16) or 2 (f16)");
  cudaStream_t st = (cudaStream_t)stream;
  static const bool no_tcol = getenv("CSTS_NO_TCOL") != nullptr;       // A/B tuning runs only
  if ((p->d == 96 || p->d == 192) && p->small_dtype == p->big_dtype && !no_tcol &&
      ((p->st == 1 && p->Ts == 4 && p->Tb == 4) || (p->st == 2 && p->Ts == 4 && p->Tb == 8))) {
    // T-column kernel: 2 column slots x 3 kernel rows per block; every block ends with 27*d global atomics, so at
    // most 3 blocks per SM
    const int64_t cols = (int64_t)p->B * p->heads * p->Hs * p->Ws;
    static const int slot_mult = getenv("CSTS_WGRAD_SLOTS") ? atoi(getenv("CSTS_WGRAD_SLOTS")) : 2;      // tuning hook (A/B: 6 -> 29.04 ms, 2 -> 28.69 ms per step)
    int64_t nslots = cols < (int64_t)csts_num_sms() * slot_mult ? cols : (int64_t)csts_num_sms() * slot_mult;
    int cols_per_warp = (int)((cols + nslots - 1) / nslots);
    int cgrid = (int)((cols + (int64_t)cols_per_warp * 2 - 1) / ((int64_t)cols_per_warp * 2));
    const int ny = (p->small2 ? 2 : 1) * (p->d / 96);               // grid.y: [problem][96-channel group of the head]
#define WGRAD_TCOL(TS_, TB_)                                                                                          \
  do {                                                                                                                \
    if (p->st == 1) launch_pdl(dwconv_wgrad_tcol_kernel<TS_, TB_, 4, 4, 1>, dim3(cgrid, ny), dim3(192), 0, st, *p, lh, lw, cols_per_warp); \
    else launch_pdl(dwconv_wgrad_tcol_kernel<TS_, TB_, 4, 8, 2>, dim3(cgrid, ny), dim3(192), 0, st, *p, lh, lw, cols_per_warp);            \
  } while (0)
    if (p->small_dtype == CSTS_F16) WGRAD_TCOL(f16, f16);
    else WGRAD_TCOL(bf16, bf16);
#undef WGRAD_TCOL
    return csts_check_launch("dwconv_wgrad_tcol");
  }
  // every block ends with 27*d global atomics: cap the grid at 4 blocks per SM
  const int64_t rows_total = (int64_t)p->B * p->heads * p->Ts * p->Hs;
  int grid = (int)(rows_total < csts_num_sms() * 4 ? rows_total : csts_num_sms() * 4);
  int rows_per_block = (int)((rows_total + grid - 1) / grid);
  grid = (int)((rows_total + rows_per_block - 1) / rows_per_block);
#define WGRAD_T(D_, SW_, TS_, TB_) launch_pdl(dwconv_wgrad_kernel<D_, SW_, TS_, TB_>, dim3(grid, p->small2 ? 2 : 1), dim3(288), 0, st, *p, lt, lh, lw, rows_per_block)
#define WGRAD(D_, SW_)                                                                        \
  do {                                                                                        \
    if (p->small_dtype == CSTS_BF16 && p->big_dtype == CSTS_F16) WGRAD_T(D_, SW_, bf16, f16); \
    else if (p->small_dtype == CSTS_F16 && p->big_dtype == CSTS_BF16) WGRAD_T(D_, SW_, f16, bf16); \
    else if (p->small_dtype == CSTS_BF16) WGRAD_T(D_, SW_, bf16, bf16);                       \
    else WGRAD_T(D_, SW_, f16, f16);                                                          \
  } while (0)
  if (p->d == 96) { if (p->sw == 1) WGRAD(96, 1); else if (p->sw == 2) WGRAD(96, 2); else WGRAD(96, 0); }
  else { if (p->sw == 1) WGRAD(192, 1); else if (p->sw == 2) WGRAD(192, 2); else WGRAD(192, 0); }
#undef WGRAD
#undef WGRAD_T
  return csts_check_launch("dwconv_wgrad");
}

int csts_maxpool_fwd(const float* x, float* y, void* arg, int B, int T, int H, int W, int C, void* stream) {
  CSTS_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: C%%4, H%%2, W%%2 required");
  int64_t total = (int64_t)B * T * (H / 2) * (W / 2) * (C / 4);
  if (total == 0) return 0;
  CSTS_REQUIRE(total < (1LL << 32), "maxpool_fwd: tensor too large for 32-bit indexing");
  launch_pdl(maxpool_fwd_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, (uint8_t*)arg, B, T, H, W, C);
  return csts_check_launch("maxpool_fwd");
}
int csts_maxpool_bwd(const float* dy, const void* arg, float* dx, int B, int T, int H, int W, int C, void* stream) {
  int64_t total = (int64_t)B * T * H * W * (C / 4);
  if (total == 0) return 0;
  CSTS_REQUIRE(total < (1LL << 32), "maxpool_bwd: tensor too large for 32-bit indexing");
  launch_pdl(maxpool_bwd_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, dy, (const uint8_t*)arg, dx, B, T, H, W, C);
  return csts_check_launch("maxpool_bwd");
}
int csts_upsample_fwd(const float* x, float* y, int B, int T, int H, int W, int C, int ft, int fh, int fw, void* stream) {
  CSTS_REQUIRE(C % 4 == 0 && ft >= 1 && fh >= 1 && fw >= 1, "upsample: bad arguments");
  int64_t total = (int64_t)B * T * ft * H * fh * W * fw * (C / 4);
  if (total == 0) return 0;
  CSTS_REQUIRE(total < (1LL << 32), "upsample_fwd: tensor too large for 32-bit indexing");
  const int64_t cells = (int64_t)B * T * H * W * (C / 4);
  if (ft == 1 && fh == 2 && fw == 2 && H >= 2 && W >= 2) {
    launch_pdl(upsample2_fwd_kernel<false>, dim3(grid_for(cells, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, B, T, H, W, C);
    return csts_check_launch("upsample_fwd");
  }
  if (ft == 2 && fh == 1 && fw == 1 && T >= 2) {
    launch_pdl(upsample2_fwd_kernel<true>, dim3(grid_for(cells, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, B, T, H, W, C);
    return csts_check_launch("upsample_fwd");
  }
  launch_pdl(upsample_fwd_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, B, T, H, W, C, ft, fh, fw);
  return csts_check_launch("upsample_fwd");
}
int csts_upsample_bwd(const float* dy, float* dx, int B, int T, int H, int W, int C, int ft, int fh, int fw, int accumulate, void* stream) {
  CSTS_REQUIRE(C % 4 == 0 && ft >= 1 && fh >= 1 && fw >= 1, "upsample: bad arguments");
  int64_t total = (int64_t)B * T * H * W * (C / 4);
  if (total == 0) return 0;
  CSTS_REQUIRE(total < (1LL << 32), "upsample_bwd: tensor too large for 32-bit indexing");
  if (ft == 1 && fh == 2 && fw == 2 && H >= 2 && W >= 2) {
    launch_pdl(upsample2_bwd_kernel<false>, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, dy, dx, B, T, H, W, C, accumulate);
    return csts_check_launch("upsample_bwd");
  }
  if (ft == 2 && fh == 1 && fw == 1 && T >= 2) {
    launch_pdl(upsample2_bwd_kernel<true>, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, dy, dx, B, T, H, W, C, accumulate);
    return csts_check_launch("upsample_bwd");
  }
  launch_pdl(upsample_bwd_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, dy, dx, B, T, H, W, C, ft, fh, fw, accumulate);
  return csts_check_launch("upsample_bwd");
}

}  // extern "C"
