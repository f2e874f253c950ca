// tcgen05 / TMEM / TMA GEMM for the token-major Linear layers (sm_100a).
//
//   C[M,N] = epi( A[M,K] . B[N,K]^T )        A, B bf16 K-major; f32 accumulate in TMEM
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, STAGES-deep mbarrier ring)
//   warp 1      MMA issuer     (one elected lane: tcgen05.mma cta_group::1 kind::f16, M=128, N=BN)
//               + TMEM allocator
//   warps 2..5  epilogue       (tcgen05.ld 32x32b -> bias / GELU / GELU' / residual -> global)
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the main loop of tile i+1.  K and M tails are handled by TMA out-of-bounds zero fill; N must be
// a multiple of BN (every Linear width in CSTS is a multiple of 96).
#include <cuda.h>

#include "common.cuh"
#include "gemm.h"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BUDGET = 200 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst_smem), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(COLS));
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type SWIZZLE_128B=2 [61,64))
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, both K-major,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct TcParams {
  void* C; void* Z; const float* bias; const float* residual; const float* row_scale; int rows_per_scale;
  int64_t ldc, ldz, ldr;
  int M, N, K;
  int c_dtype, act, accumulate, res_mod;
  float alpha;
};

template <int BN> struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;   // >= 2 accumulators, power of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, int m, int n, const uint32_t (&r)[16]) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
  if (p.bias) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(p.bias + n + i);
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (p.act == 1) {
    if (p.Z) {
      uint4* z = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.Z) + (int64_t)m * p.ldz + n);
      z[0] = make_uint4(pack_bf162(v[0], v[1]), pack_bf162(v[2], v[3]), pack_bf162(v[4], v[5]), pack_bf162(v[6], v[7]));
      z[1] = make_uint4(pack_bf162(v[8], v[9]), pack_bf162(v[10], v[11]), pack_bf162(v[12], v[13]), pack_bf162(v[14], v[15]));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
  } else if (p.act == 2) {
    const uint4* z = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.Z) + (int64_t)m * p.ldz + n);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 t = z[h];
      const bf162* zz = reinterpret_cast<const bf162*>(&t);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[h * 8 + 2 * i] *= gelu_erf_grad(__low2float(zz[i]));
        v[h * 8 + 2 * i + 1] *= gelu_erf_grad(__high2float(zz[i]));
      }
    }
  }
  if (p.row_scale) {
    const float rs = p.row_scale[m / p.rows_per_scale];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= rs;
  }
  if (p.residual) {
    const float4* r4 = reinterpret_cast<const float4*>(p.residual + (int64_t)(p.res_mod > 0 ? m % p.res_mod : m) * p.ldr + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = r4[i];
      v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
  }
  if (p.c_dtype == 0) {
    float4* c = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (int64_t)m * p.ldc + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      if (p.accumulate) { float4 o = c[i]; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
      c[i] = t;
    }
  } else {
    uint4* c = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.C) + (int64_t)m * p.ldc + n);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (p.accumulate) {
        uint4 o = c[h];
        const bf162* oo = reinterpret_cast<const bf162*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[h * 8 + 2 * i] += __low2float(oo[i]); v[h * 8 + 2 * i + 1] += __high2float(oo[i]); }
      }
      c[h] = make_uint4(pack_bf162(v[h * 8], v[h * 8 + 1]), pack_bf162(v[h * 8 + 2], v[h * 8 + 3]),
                        pack_bf162(v[h * 8 + 4], v[h * 8 + 5]), pack_bf162(v[h * 8 + 6], v[h * 8 + 7]));
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, TcParams p) {
  using C = Cfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + C::STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]
  uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * C::STAGES;
  uint32_t tfull0 = empty0 + 8 * C::STAGES, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = p.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int kblocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        // consecutive CTAs share the same B (weight) tile and walk M: weights stay L2-hot
        const int tm = t % tiles_m, tn = t / tiles_m;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          mbar_expect_tx(full0 + 8 * stage, C::STAGE_BYTES);
          tma_load_2d(smem_u32(sA + stage * C::A_BYTES), &tmap_a, full0 + 8 * stage, kb * BK, tm * BM);
          tma_load_2d(smem_u32(sB + stage * C::B_BYTES), &tmap_b, full0 + 8 * stage, kb * BK, tn * BN);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(tempty0 + 8 * as, aphase ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sA + stage * C::A_BYTES));
          const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sB + stage * C::B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the >>4 address field
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(empty0 + 8 * stage);           // smem slot free once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull0 + 8 * as);                // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    int as = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tm = t % tiles_m, tn = t / tiles_m;
      mbar_wait(tfull0 + 8 * as, aphase);
      tc_fence_after();
      const int m = tm * BM + quad * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        uint32_t r[16];
        tmem_ld16(trow + c, r);
        tmem_ld_wait();
        if (m < p.M) epilogue_chunk<BN>(p, m, tn * BN + c, r);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor with row pitch ld elements; box = [box_rows, 64 cols], 128B swizzle
int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  CSTS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old / no GPU)");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CSTS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%ld cols=%ld ld=%ld", (int)r, (long)rows, (long)cols, (long)ld);
  return 0;
}

template <int BN>
int launch(const csts_gemm_args& a, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CSTS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, a.A, a.M, a.K, a.lda, BM);
  if (rc) return rc;
  rc = make_tmap(&tb, a.B, a.N, a.K, a.ldb, BN);
  if (rc) return rc;
  TcParams p;
  p.C = a.C; p.Z = a.Z; p.bias = a.bias; p.residual = a.residual;
  p.row_scale = a.row_scale; p.rows_per_scale = a.rows_per_scale > 0 ? a.rows_per_scale : 1;
  p.ldc = a.ldc; p.ldz = a.ldz; p.ldr = a.ldr;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.c_dtype = a.c_dtype; p.act = a.act; p.accumulate = a.accumulate; p.res_mod = a.res_mod; p.alpha = a.alpha;
  int tiles = ceil_div(a.M, BM) * (a.N / BN);
  int grid = tiles < csts_num_sms() ? tiles : csts_num_sms();
  gemm_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  return csts_check_launch("gemm_tc_kernel");
}

int pick_bn(int M, int N) {
  // widest tile that divides N while still giving every SM work; prefer fewer, fatter tiles
  const int cands[4] = {256, 192, 128, 96};
  int sms = csts_num_sms();
  int best = 0;
  for (int i = 0; i < 4; ++i) {
    int bn = cands[i];
    if (N % bn) continue;
    if (!best) best = bn;                       // widest divisor as fallback
    if (ceil_div(M, BM) * (N / bn) >= sms) return bn;
  }
  // not enough tiles for a full wave with any width: take the narrowest divisor (most CTAs)
  for (int i = 3; i >= 0; --i)
    if (N % cands[i] == 0) return cands[i];
  return best;
}

}  // namespace

bool csts_gemm_tc_supported(const csts_gemm_args& a) {
  if (!a.a_kmajor || !a.b_kmajor) return false;
  if (a.batch1 * a.batch2 != 1 || a.split_k > 1) return false;
  if (a.N % 96 != 0 && a.N % 128 != 0) return false;
  if (a.K % 8 != 0 || a.lda % 8 != 0 || a.ldb % 8 != 0) return false;
  if (((uintptr_t)a.A & 15) || ((uintptr_t)a.B & 15)) return false;
  // vectorised epilogue: 16-column chunks, 16-byte aligned rows
  int cbytes = a.c_dtype == 0 ? 4 : 2;
  if (((uintptr_t)a.C & 15) || (a.ldc * cbytes) % 16) return false;
  if (a.Z && (((uintptr_t)a.Z & 15) || (a.ldz * 2) % 16)) return false;
  if (a.residual && (((uintptr_t)a.residual & 15) || (a.ldr * 4) % 16)) return false;
  if (a.bias && ((uintptr_t)a.bias & 15)) return false;
  if (a.M < 64) return false;                   // skinny problems: the generic kernel with split-K
  return true;
}

int csts_gemm_tc_launch(const csts_gemm_args& a, cudaStream_t stream) {
  CSTS_REQUIRE(csts_gemm_tc_supported(a), "gemm_tc: unsupported problem (M=%d N=%d K=%d)", a.M, a.N, a.K);
  if (a.act == 2) CSTS_REQUIRE(a.Z != nullptr, "gemm: act==2 needs Z");
  switch (pick_bn(a.M, a.N)) {
    case 256: return launch<256>(a, stream);
    case 192: return launch<192>(a, stream);
    case 128: return launch<128>(a, stream);
    case 96: return launch<96>(a, stream);
  }
  csts_set_error("gemm_tc: no tile width divides N=%d", a.N);
  return 2;
}
