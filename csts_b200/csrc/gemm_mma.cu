// Generic strided / batched 16-bit GEMM on legacy tensor-core instructions (mma.sync m16n8k16).
//
// Role in the design: this is the *shape-agnostic* kernel — any operand
// majorness, two-level batch strides (batch, head), ragged M/N/K (Nk = 260, M = 8), split-K with
// f32 atomics.  The token-major Linear GEMMs that carry 87 % of the FLOPs run on the tcgen05
// kernel in gemm_tc.cu; this kernel serves the odd shapes and is the on-device cross-check for it.
//
// C[z][m][n] = epi( alpha * sum_k opA(A[z])[m][k] * opB(B[z])[k][n] )
#include "common.cuh"
#include "gemm.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 4, THREADS = 256;
constexpr int TILE_ELEMS = BM * BK;                  // 4096 bf16 = 8 KB per operand per stage
constexpr int SMEM_BYTES = STAGES * 2 * TILE_ELEMS * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// mma.sync needs one operand type: a mixed (bf16 gradient x f16 activation) product re-rounds the f16
// fragment to bf16 in registers (this kernel only carries the small odd-shaped GEMMs of the path)
__device__ __forceinline__ uint32_t f162_to_bf162(uint32_t h) {
  float2 f = unpack2<f16>(h);
  return pack_bf162(f.x, f.y);
}

// smem element offset of the 16-byte chunk holding (row, k) for a tile of extent 128 (mn) x 32 (k)
//   KMAJOR : stored [mn][k]   64 B rows, 4 chunks / row, chunk ^= (row >> 1) & 3
//   !KMAJOR: stored [k][mn]  256 B rows, 16 chunks / row, chunk ^= k & 7
template <bool KMAJOR> __device__ __forceinline__ int tile_off(int mn, int k) {
  if (KMAJOR) {
    int c = (k >> 3) ^ ((mn >> 1) & 3);
    return mn * 32 + c * 8 + (k & 7);
  } else {
    int c = (mn >> 3) ^ (k & 7);
    return k * 128 + c * 8 + (mn & 7);
  }
}

// Load one 128 x 32 operand tile (global -> smem) with zero fill outside [0,MN) x [0,K).
template <bool KMAJOR>
__device__ __forceinline__ void load_tile(bf16* s, const bf16* g, int64_t ld, int mn0, int k0, int MN, int K, int tid) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int q = tid + i * THREADS;           // 512 chunks
    int mn, k, valid;
    const bf16* src;
    if (KMAJOR) {
      mn = q >> 2; k = (q & 3) * 8;
      int gm = mn0 + mn, gk = k0 + k;
      valid = (gm < MN) ? min(max(K - gk, 0), 8) : 0;
      src = g + (int64_t)gm * ld + gk;
    } else {
      k = q >> 4; mn = (q & 15) * 8;
      int gm = mn0 + mn, gk = k0 + k;
      valid = (gk < K) ? min(max(MN - gm, 0), 8) : 0;
      src = g + (int64_t)gk * ld + gm;
    }
    if (valid == 0) src = g;             // keep the address legal; 0 bytes are read
    cp_async16(smem_u32(s + tile_off<KMAJOR>(mn, k)), src, valid * 2);
  }
}

struct Epi {
  void* C; void* Z; const float* bias; const float* residual; const float* row_scale; int rows_per_scale;
  int64_t ldc, ldz, ldr;
  int M, N, c_dtype, z_half, act, accumulate, res_mod, atomic, first_split;
  float alpha;
};

__device__ __forceinline__ void epi_store(const Epi& e, int m, int n, float v0, float v1) {
  if (m >= e.M || n >= e.N) return;
  bool two = (n + 1 < e.N);
  v0 *= e.alpha; v1 *= e.alpha;
  if (e.atomic) {
    float* c = reinterpret_cast<float*>(e.C) + (int64_t)m * e.ldc + n;
    if (e.first_split) {
      if (e.bias) { v0 += e.bias[n]; if (two) v1 += e.bias[n + 1]; }
      if (e.residual) {
        const float* r = e.residual + (int64_t)(e.res_mod > 0 ? m % e.res_mod : m) * e.ldr + n;
        v0 += r[0]; if (two) v1 += r[1];
      }
    }
    atomicAdd(c, v0);
    if (two) atomicAdd(c + 1, v1);
    return;
  }
  if (e.bias) { v0 += e.bias[n]; if (two) v1 += e.bias[n + 1]; }
  if (e.act == 1) {          // C = gelu(v), Z = gelu'(v)
    if (e.Z && e.z_half) {
      f16* z = reinterpret_cast<f16*>(e.Z) + (int64_t)m * e.ldz + n;
      z[0] = __float2half_rn(gelu_erf_grad(v0)); if (two) z[1] = __float2half_rn(gelu_erf_grad(v1));
    } else if (e.Z) {
      bf16* z = reinterpret_cast<bf16*>(e.Z) + (int64_t)m * e.ldz + n;
      z[0] = __float2bfloat16_rn(gelu_erf_grad(v0)); if (two) z[1] = __float2bfloat16_rn(gelu_erf_grad(v1));
    }
    v0 = gelu_erf(v0); v1 = gelu_erf(v1);
  } else if (e.act == 2 && e.z_half) {   // C = v * Z
    const f16* z = reinterpret_cast<const f16*>(e.Z) + (int64_t)m * e.ldz + n;
    v0 *= __half2float(z[0]);
    if (two) v1 *= __half2float(z[1]);
  } else if (e.act == 2) {
    const bf16* z = reinterpret_cast<const bf16*>(e.Z) + (int64_t)m * e.ldz + n;
    v0 *= __bfloat162float(z[0]);
    if (two) v1 *= __bfloat162float(z[1]);
  }
  if (e.row_scale) { float rs = e.row_scale[m / e.rows_per_scale]; v0 *= rs; v1 *= rs; }
  if (e.residual) {
    const float* r = e.residual + (int64_t)(e.res_mod > 0 ? m % e.res_mod : m) * e.ldr + n;
    v0 += r[0]; if (two) v1 += r[1];
  }
  if (e.c_dtype == 0) {
    float* c = reinterpret_cast<float*>(e.C) + (int64_t)m * e.ldc + n;
    if (e.accumulate) { v0 += c[0]; if (two) v1 += c[1]; }
    c[0] = v0; if (two) c[1] = v1;
  } else if (e.c_dtype == CSTS_F16) {
    f16* c = reinterpret_cast<f16*>(e.C) + (int64_t)m * e.ldc + n;
    if (e.accumulate) { v0 += __half2float(c[0]); if (two) v1 += __half2float(c[1]); }
    c[0] = __float2half_rn(v0); if (two) c[1] = __float2half_rn(v1);
  } else {
    bf16* c = reinterpret_cast<bf16*>(e.C) + (int64_t)m * e.ldc + n;
    if (e.accumulate) { v0 += __bfloat162float(c[0]); if (two) v1 += __bfloat162float(c[1]); }
    c[0] = __float2bfloat16_rn(v0); if (two) c[1] = __float2bfloat16_rn(v1);
  }
}

template <bool AK, bool BKM>
__global__ void __launch_bounds__(THREADS) gemm_mma_kernel(csts_gemm_args p, int k_per_split) {
  pdl_wait();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  bf16* sA = reinterpret_cast<bf16*>(smem_raw);
  bf16* sB = sA + STAGES * TILE_ELEMS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;            // 2 x 4 warps, warp tile 64 x 32
  const bool a_half = p.a_dtype == CSTS_F16, b_half = p.b_dtype == CSTS_F16;
  const bool both_half = a_half && b_half;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int z = blockIdx.z;
  const int split = z % p.split_k;
  z /= p.split_k;
  const int z1 = z / p.batch2, z2 = z % p.batch2;

  const bf16* A = reinterpret_cast<const bf16*>(p.A) + z1 * p.sA1 + z2 * p.sA2;
  const bf16* B = reinterpret_cast<const bf16*>(p.B) + z1 * p.sB1 + z2 * p.sB2;
  const int64_t coff = z1 * p.sC1 + z2 * p.sC2;

  const int kbeg = split * k_per_split;
  const int kend = min(p.K, kbeg + k_per_split);
  const int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      load_tile<AK>(sA + s * TILE_ELEMS, A, p.lda, m0, kbeg + s * BK, p.M, kend, tid);
      load_tile<BKM>(sB + s * TILE_ELEMS, B, p.ldb, n0, kbeg + s * BK, p.N, kend, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {  // prefetch tile kt + STAGES - 1 into the slot consumed at iteration kt - 1
      int nt = kt + STAGES - 1;
      if (nt < nk) {
        int s = nt % STAGES;
        load_tile<AK>(sA + s * TILE_ELEMS, A, p.lda, m0, kbeg + nt * BK, p.M, kend, tid);
        load_tile<BKM>(sB + s * TILE_ELEMS, B, p.ldb, n0, kbeg + nt * BK, p.N, kend, tid);
      }
      cp_async_commit();
    }
    const bf16* a_s = sA + (kt % STAGES) * TILE_ELEMS;
    const bf16* b_s = sB + (kt % STAGES) * TILE_ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 16) {
      uint32_t af[4][4], bfr[2][4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        int mb = wm * 64 + mi * 16;
        if (AK) {
          int row = mb + (lane & 7) + 8 * ((lane >> 3) & 1), k = kk + 8 * (lane >> 4);
          ldsm4(smem_u32(a_s + tile_off<true>(row, k)), af[mi]);
        } else {
          int k = kk + (lane & 7) + 8 * (lane >> 4), m = mb + 8 * ((lane >> 3) & 1);
          ldsm4t(smem_u32(a_s + tile_off<false>(m, k)), af[mi]);
        }
      }
#pragma unroll
      for (int nj = 0; nj < 2; ++nj) {    // each x4 covers two 8-wide n blocks
        int nb = wn * 32 + nj * 16;
        if (BKM) {
          int n = nb + (lane & 7) + 8 * (lane >> 4), k = kk + 8 * ((lane >> 3) & 1);
          ldsm4(smem_u32(b_s + tile_off<true>(n, k)), bfr[nj]);
        } else {
          int k = kk + (lane & 7) + 8 * ((lane >> 3) & 1), n = nb + 8 * (lane >> 4);
          ldsm4t(smem_u32(b_s + tile_off<false>(n, k)), bfr[nj]);
        }
      }
      if (both_half) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
            mma16816_f16(acc[mi][ni], af[mi], bfr[ni >> 1][(ni & 1) * 2], bfr[ni >> 1][(ni & 1) * 2 + 1]);
      } else {
        if (a_half) {
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int r = 0; r < 4; ++r) af[mi][r] = f162_to_bf162(af[mi][r]);
        }
        if (b_half) {
#pragma unroll
          for (int nj = 0; nj < 2; ++nj)
#pragma unroll
            for (int r = 0; r < 4; ++r) bfr[nj][r] = f162_to_bf162(bfr[nj][r]);
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
            mma16816(acc[mi][ni], af[mi], bfr[ni >> 1][(ni & 1) * 2], bfr[ni >> 1][(ni & 1) * 2 + 1]);
      }
    }
  }
  cp_async_wait<0>();

  Epi e;
  e.ldc = p.ldc; e.ldz = p.ldz; e.ldr = p.ldr;
  e.M = p.M; e.N = p.N; e.c_dtype = p.c_dtype; e.z_half = p.z_dtype == CSTS_F16; e.act = p.act; e.accumulate = p.accumulate;
  e.res_mod = p.res_mod; e.atomic = p.split_k > 1; e.first_split = (split == 0);
  e.alpha = p.alpha; e.bias = p.bias; e.residual = p.residual;
  e.row_scale = p.row_scale; e.rows_per_scale = p.rows_per_scale > 0 ? p.rows_per_scale : 1;
  e.C = p.c_dtype == 0 ? (void*)(reinterpret_cast<float*>(p.C) + coff) : (void*)(reinterpret_cast<bf16*>(p.C) + coff);
  e.Z = p.Z ? (void*)(reinterpret_cast<bf16*>(p.Z) + coff) : nullptr;
  if (nk == 0 && !(e.atomic && e.first_split)) return;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      int m = m0 + wm * 64 + mi * 16 + (lane >> 2);
      int n = n0 + wn * 32 + ni * 8 + 2 * (lane & 3);
      epi_store(e, m, n, acc[mi][ni][0], acc[mi][ni][1]);
      epi_store(e, m + 8, n, acc[mi][ni][2], acc[mi][ni][3]);
    }
}

template <bool AK, bool BKM>
int launch(const csts_gemm_args& a, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CSTS_CUDA(cudaFuncSetAttribute(gemm_mma_kernel<AK, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  int split = a.split_k > 1 ? a.split_k : 1;
  int kblocks = ceil_div(a.K, BK);
  int k_per_split = ceil_div(kblocks, split) * BK;
  csts_gemm_args p = a;
  p.split_k = split;
  dim3 grid(ceil_div(a.N, BN), ceil_div(a.M, BM), a.batch1 * a.batch2 * split);
  launch_pdl(gemm_mma_kernel<AK, BKM>, dim3(grid), dim3(THREADS), SMEM_BYTES, stream, p, k_per_split);
  return csts_check_launch("gemm_mma_kernel");
}

}  // namespace

int csts_gemm_mma_launch(const csts_gemm_args& a, cudaStream_t stream) {
  CSTS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && a.batch1 > 0 && a.batch2 > 0, "gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  CSTS_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements (lda=%ld ldb=%ld)", (long)a.lda, (long)a.ldb);
  CSTS_REQUIRE(((uintptr_t)a.A & 15) == 0 && ((uintptr_t)a.B & 15) == 0, "gemm: A/B must be 16-byte aligned");
  CSTS_REQUIRE((a.sA1 % 8 == 0) && (a.sA2 % 8 == 0) && (a.sB1 % 8 == 0) && (a.sB2 % 8 == 0), "gemm: batch strides must be multiples of 8");
  if (a.split_k > 1) {
    CSTS_REQUIRE(a.c_dtype == 0 && a.act == 0 && !a.row_scale, "gemm: split-K needs f32 output, no activation, no row scale");
    if (!a.accumulate) {
      CSTS_REQUIRE(a.batch1 * a.batch2 == 1, "gemm: split-K without accumulate supports a single batch");
      CSTS_CUDA(cudaMemset2DAsync(a.C, a.ldc * sizeof(float), 0, (size_t)a.N * sizeof(float), a.M, stream));
    }
  }
  if (a.act == 2) CSTS_REQUIRE(a.Z != nullptr, "gemm: act==2 needs Z");
  if (a.a_kmajor) return a.b_kmajor ? launch<true, true>(a, stream) : launch<true, false>(a, stream);
  return a.b_kmajor ? launch<false, true>(a, stream) : launch<false, false>(a, stream);
}
