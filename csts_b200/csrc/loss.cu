// Loss kernels: frame-softmax + KL-divergence over heat-maps (fused forward + gradient), cosine
// similarity matrix, and the symmetric InfoNCE (EgoNCE) log-sum-exp loss.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// One block per (b, t) frame of HW logits.
//   p = softmax(logit / tau)                                   ref: slowfast/utils/utils.py:5-12
//   kl = sum p*log(p+1e-10) - sum p*log(q+1e-10)                ref: slowfast/models/losses.py:66-77
//   loss = sum_frames kl / (T*log(HW)) / B                      ref: losses.py:79-81
// Also emits p (the softmaxed prediction the caller/metrics use) and dloss/dlogit:
//   g_j = dkl/dp_j = log(p_j+eps) + p_j/(p_j+eps) - log(q_j+eps)
//   dlogit_j = norm/tau * p_j * (g_j - sum_k p_k g_k)
// ------------------------------------------------------------------------------------------------
template <int PER_THREAD>
__global__ void __launch_bounds__(256) kldiv_frame_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                          float* __restrict__ prob, float* __restrict__ frame_kl,
                                                          float* __restrict__ dlogits, int HW, float inv_tau, float norm) {
  pdl_wait();
  __shared__ float red[33];
  const int64_t base = (int64_t)blockIdx.x * HW;
  float l[PER_THREAD], q[PER_THREAD];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER_THREAD; ++i) {
    int c = threadIdx.x + i * 256;
    l[i] = c < HW ? logits[base + c] * inv_tau : -INFINITY;
    q[i] = (c < HW && target) ? target[base + c] : 0.f;
    mx = fmaxf(mx, l[i]);
  }
  mx = block_max(mx, red);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER_THREAD; ++i) {
    l[i] = (threadIdx.x + i * 256 < HW) ? expf(l[i] - mx) : 0.f;
    s += l[i];
  }
  const float inv = 1.f / block_sum(s, red);
  float kl = 0.f, pg = 0.f;
  float g[PER_THREAD];
#pragma unroll
  for (int i = 0; i < PER_THREAD; ++i) {
    int c = threadIdx.x + i * 256;
    float p = l[i] * inv;
    l[i] = p;
    float lp = logf(p + 1e-10f), lq = logf(q[i] + 1e-10f);
    g[i] = lp + p / (p + 1e-10f) - lq;
    if (c < HW) {
      kl += p * lp - p * lq;
      pg += p * g[i];
      if (prob) prob[base + c] = p;
    }
  }
  kl = block_sum(kl, red);
  pg = block_sum(pg, red);
  if (threadIdx.x == 0) frame_kl[blockIdx.x] = kl;
  if (dlogits) {
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      int c = threadIdx.x + i * 256;
      if (c < HW) dlogits[base + c] = norm * inv_tau * l[i] * (g[i] - pg);
    }
  }
}

// loss[0] = scale * sum_i v[i]     (single block, deterministic)
__global__ void reduce_scale_kernel(const float* __restrict__ v, int n, float scale, float* __restrict__ out) {
  pdl_wait();
  __shared__ float red[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// ------------------------------------------------------------------------------------------------
// sim[i][j] = <a_i/max(|a_i|,eps), b_j/max(|b_j|,eps)>            ref: slowfast/utils/utils.py:15-24
// Single CTA (Bg <= 1024 rows of D <= 1024): the problem is a few KB.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sim_matrix_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ sim,
                                                             float* __restrict__ na, float* __restrict__ nb, int n, int D, float eps) {
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < 2 * n; r += nw) {
    const float* v = r < n ? a + (int64_t)r * D : b + (int64_t)(r - n) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += v[c] * v[c];
    s = sqrtf(warp_sum(s));
    if (lane == 0) (r < n ? na[r] : nb[r - n]) = fmaxf(s, eps);
  }
  __syncthreads();
  for (int ij = warp; ij < n * n; ij += nw) {
    int i = ij / n, j = ij - i * n;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += a[(int64_t)i * D + c] * b[(int64_t)j * D + c];
    s = warp_sum(s);
    if (lane == 0) sim[ij] = s / (na[i] * nb[j]);
  }
}
// da_i = (sum_j dsim_ij bn_j - an_i * sum_j dsim_ij sim_ij) / na_i   (and symmetrically for b);
// exact for |a_i| >= eps (the clamp is never active for real embeddings).
__global__ void __launch_bounds__(256) sim_matrix_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             const float* __restrict__ sim, const float* __restrict__ dsim,
                                                             const float* __restrict__ na, const float* __restrict__ nb,
                                                             float* __restrict__ da, float* __restrict__ db, int n, int D) {
  pdl_wait();
  const int64_t total = (int64_t)2 * n * D;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % D);
    int r = (int)(idx / D);
    bool is_a = r < n;
    int i = is_a ? r : r - n;
    float acc = 0.f, dot = 0.f;
    for (int j = 0; j < n; ++j) {
      float ds = is_a ? dsim[i * n + j] : dsim[j * n + i];
      float sm = is_a ? sim[i * n + j] : sim[j * n + i];
      float other = is_a ? b[(int64_t)j * D + c] / nb[j] : a[(int64_t)j * D + c] / na[j];
      acc += ds * other;
      dot += ds * sm;
    }
    if (is_a) da[(int64_t)i * D + c] = (acc - a[(int64_t)i * D + c] / na[i] * dot) / na[i];
    else db[(int64_t)i * D + c] = (acc - b[(int64_t)i * D + c] / nb[i] * dot) / nb[i];
  }
}

// ------------------------------------------------------------------------------------------------
// EgoNCE: loss = -mean_i log softmax(x/T)_ii - mean_j log softmax(x^T/T)_jj   ref: losses.py:157-170
// dsim_ij = ((rowsoftmax_ij - d_ij) + (colsoftmax_ij - d_ij)) / (n*T)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) egonce_kernel(const float* __restrict__ sim, float* __restrict__ loss, float* __restrict__ dsim,
                                                     float* __restrict__ lse_row, float* __restrict__ lse_col, int n, float inv_temp) {
  pdl_wait();
  __shared__ float red[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < 2 * n; r += nw) {
    bool row = r < n;
    int i = row ? r : r - n;
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, (row ? sim[i * n + j] : sim[j * n + i]) * inv_temp);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += expf((row ? sim[i * n + j] : sim[j * n + i]) * inv_temp - mx);
    s = warp_sum(s);
    if (lane == 0) (row ? lse_row[i] : lse_col[i]) = mx + logf(s);
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float z = sim[i * n + i] * inv_temp;
    acc += (z - lse_row[i]) + (z - lse_col[i]);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = -acc / (float)n;
  if (dsim) {
    const float k = inv_temp / (float)n;
    for (int ij = threadIdx.x; ij < n * n; ij += blockDim.x) {
      int i = ij / n, j = ij - i * n;
      float z = sim[ij] * inv_temp;
      float d = (i == j) ? 2.f : 0.f;
      dsim[ij] = k * (expf(z - lse_row[i]) + expf(z - lse_col[j]) - d);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// adaptive_f1 (slowfast/utils/metrics.py:9-74) with the min-max rescale of the loops fused in
// (tools/train_avgaze_net.py:125-127, test_avgaze_net.py:66-68): the reference materialises two
// (thresholds, B, T, H, W) f32 tensors per call; here one block per frame counts, for every threshold at once,
//   tp = #(pred > th  and  label > 0.001), fg_pred = #(pred > th), fg_label = #(label > 0.001)
// and a single-block second kernel turns the counts of the tracked frames (labels[:, :, 2] == fixation_idx) into
// recall / precision / f1 per threshold and picks the best threshold.  Nothing is read back by the kernels.
// ------------------------------------------------------------------------------------------------
constexpr int F1_MAX_THR = 32;

__global__ void __launch_bounds__(256) f1_count_kernel(const float* __restrict__ preds, const float* __restrict__ labels_hm,
                                                       const float* __restrict__ thr, int n_thr, int HW, int rescale,
                                                       float* __restrict__ counts /* [frames][2*n_thr+1] */) {
  pdl_wait();
  __shared__ float red[33];
  __shared__ int s_cnt[2 * F1_MAX_THR + 1];
  for (int i = threadIdx.x; i < 2 * n_thr + 1; i += blockDim.x) s_cnt[i] = 0;
  const int64_t base = (int64_t)blockIdx.x * HW;
  float mn = INFINITY, mx = -INFINITY;
  if (rescale) {
    for (int c = threadIdx.x; c < HW; c += blockDim.x) {
      float v = preds[base + c];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    mn = -block_max(-mn, red);
  }
  __syncthreads();
  const float den = mx - mn + 1e-6f;
  int tp[F1_MAX_THR], fp[F1_MAX_THR], fl = 0;
#pragma unroll
  for (int i = 0; i < F1_MAX_THR; ++i) { tp[i] = 0; fp[i] = 0; }
  for (int c = threadIdx.x; c < HW; c += blockDim.x) {
    float v = preds[base + c];
    if (rescale) v = (v - mn) / den;
    const int lab = labels_hm[base + c] > 0.001f;
    fl += lab;
#pragma unroll
    for (int i = 0; i < F1_MAX_THR; ++i) {
      if (i < n_thr) {
        const int on = v > thr[i];
        fp[i] += on;
        tp[i] += on & lab;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < F1_MAX_THR; ++i) {
    if (i < n_thr) {
      int a = tp[i], b = fp[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt[2 * i], a); atomicAdd(&s_cnt[2 * i + 1], b); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fl += __shfl_xor_sync(0xffffffffu, fl, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt[2 * n_thr], fl);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * n_thr + 1; i += blockDim.x) counts[(int64_t)blockIdx.x * (2 * n_thr + 1) + i] = (float)s_cnt[i];
}

// out[0..4] = f1, recall, precision, threshold, index of the best threshold (first maximum, as torch.argmax)
__global__ void __launch_bounds__(256) f1_select_kernel(const float* __restrict__ counts, const float* __restrict__ labels,
                                                        const float* __restrict__ thr, int n_thr, int frames, int fixation_idx,
                                                        float* __restrict__ out) {
  pdl_wait();
  __shared__ float red[33];
  __shared__ float s_f1[F1_MAX_THR], s_rc[F1_MAX_THR], s_pr[F1_MAX_THR];
  const int stride = 2 * n_thr + 1;
  float ntr = 0.f;
  for (int f = threadIdx.x; f < frames; f += blockDim.x) ntr += (labels[3 * f + 2] == (float)fixation_idx) ? 1.f : 0.f;
  ntr = block_sum(ntr, red);
  for (int i = 0; i < n_thr; ++i) {
    float rc = 0.f, pr = 0.f;
    for (int f = threadIdx.x; f < frames; f += blockDim.x) {
      if (labels[3 * f + 2] == (float)fixation_idx) {
        const float tp = counts[(int64_t)f * stride + 2 * i], fgp = counts[(int64_t)f * stride + 2 * i + 1], fgl = counts[(int64_t)f * stride + 2 * n_thr];
        rc += tp / (fgl + 1e-6f);
        pr += tp / (fgp + 1e-6f);
      }
    }
    rc = block_sum(rc, red) / ntr;
    pr = block_sum(pr, red) / ntr;
    if (threadIdx.x == 0) { s_rc[i] = rc; s_pr[i] = pr; s_f1[i] = (2.f * rc * pr) / (rc + pr + 1e-6f); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = 0;
    for (int i = 1; i < n_thr; ++i)
      if (s_f1[i] > s_f1[best]) best = i;
    out[0] = s_f1[best]; out[1] = s_rc[best]; out[2] = s_pr[best]; out[3] = thr[best]; out[4] = (float)best;
  }
}

}  // namespace

extern "C" {

// counts: scratch f32 [frames * (2 * n_thr + 1)]; out: f32 [5] (f1, recall, precision, threshold, threshold index)
int csts_adaptive_f1(const float* preds, const float* labels_hm, const float* labels, const float* thresholds, int n_thr, int frames, int HW,
                     int fixation_idx, int rescale, float* counts, float* out, void* stream) {
  CSTS_REQUIRE(n_thr >= 1 && n_thr <= F1_MAX_THR, "adaptive_f1: 1..%d thresholds", F1_MAX_THR);
  CSTS_REQUIRE(frames > 0 && HW > 0, "adaptive_f1: empty input");
  launch_pdl(f1_count_kernel, dim3(frames), dim3(256), 0, (cudaStream_t)stream, preds, labels_hm, thresholds, n_thr, HW, rescale, counts);
  int rc = csts_check_launch("adaptive_f1 (count)");
  if (rc) return rc;
  launch_pdl(f1_select_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, (const float*)counts, labels, thresholds, n_thr, frames, fixation_idx, out);
  return csts_check_launch("adaptive_f1 (select)");
}


// logits/target/prob/dlogits: [frames, HW] f32; frame_kl: [frames] scratch; loss: [1]
// loss = (1 / (T * log(HW) * B)) * sum_frames kl ; frames = B*T.  dlogits already carries that factor.
int csts_kldiv_frame_softmax(const float* logits, const float* target, float* prob, float* frame_kl, float* loss, float* dlogits,
                             int frames, int HW, int T, float temperature, void* stream) {
  CSTS_REQUIRE(frames > 0 && HW > 0 && HW <= 16 * 256, "kldiv: HW %d unsupported (<= 4096)", HW);
  CSTS_REQUIRE(target != nullptr, "kldiv: target required (uniform-prior variant is not used by CSTS)");
  cudaStream_t st = (cudaStream_t)stream;
  float norm = 1.f / ((float)T * logf((float)HW) * ((float)frames / (float)T));
  if (HW <= 4 * 256) launch_pdl(kldiv_frame_kernel<4>, dim3(frames), dim3(256), 0, st, logits, target, prob, frame_kl, dlogits, HW, 1.f / temperature, norm);
  else launch_pdl(kldiv_frame_kernel<16>, dim3(frames), dim3(256), 0, st, logits, target, prob, frame_kl, dlogits, HW, 1.f / temperature, norm);
  int rc = csts_check_launch("kldiv_frame");
  if (rc) return rc;
  launch_pdl(reduce_scale_kernel, dim3(1), dim3(256), 0, st, frame_kl, frames, norm, loss);
  return csts_check_launch("kldiv_reduce");
}

int csts_sim_matrix_fwd(const float* a, const float* b, float* sim, float* na, float* nb, int n, int D, float eps, void* stream) {
  CSTS_REQUIRE(n > 0 && D > 0, "sim_matrix: empty input");
  launch_pdl(sim_matrix_fwd_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, a, b, sim, na, nb, n, D, eps);
  return csts_check_launch("sim_matrix_fwd");
}
int csts_sim_matrix_bwd(const float* a, const float* b, const float* sim, const float* dsim, const float* na, const float* nb, float* da,
                        float* db, int n, int D, void* stream) {
  int64_t total = (int64_t)2 * n * D;
  int grid = (int)((total + 255) / 256);
  launch_pdl(sim_matrix_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a, b, sim, dsim, na, nb, da, db, n, D);
  return csts_check_launch("sim_matrix_bwd");
}
// lse_scratch: [2*n] floats
int csts_egonce(const float* sim, float* loss, float* dsim, float* lse_scratch, int n, float temperature, void* stream) {
  CSTS_REQUIRE(n > 0, "egonce: empty similarity matrix");
  launch_pdl(egonce_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, sim, loss, dsim, lse_scratch, lse_scratch + n, n, 1.f / temperature);
  return csts_check_launch("egonce");
}

}  // extern "C"
