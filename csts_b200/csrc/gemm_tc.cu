// tcgen05 / TMEM / TMA GEMM (sm_100a) for the Linear layers of CSTS: forward, data-gradient and
// weight-gradient products.
//
//   C[M,N] = epi( sum_k A[m,k] * B[n,k] )         bf16 / f16 operands (independently), f32 accumulate in TMEM
//
//   operand storage   K-major : X[mn][k]  (k contiguous)   forward  (x . W^T)  and dX (dY . W^T^T)
//                     MN-major: X[k][mn]  (mn contiguous)  weight gradients dW = dY^T . X  (both operands
//                               token-major, contraction over tokens) — no transposed copies are made
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, STAGES-deep mbarrier ring)
//   warp 1      MMA issuer     (one elected lane: tcgen05.mma cta_group::1 kind::f16, M=128, N=BN)
//               + TMEM allocator
//   warps 2..5  epilogue       tcgen05.ld 32x32b -> registers -> per-warp swizzled smem tile ->
//               fully coalesced 16-byte global stores (and coalesced residual / Z / C reads)
// The accumulator is double-buffered in TMEM (2 x BN columns): the epilogue of work item i overlaps
// the main loop of item i+1.  Split-K (weight gradients: few tiles, K = 10^5 tokens) distributes
// (tile, k-range) items over the persistent CTAs and combines with vectorised f32 atomics.
// K / M / N tails are handled by TMA out-of-bounds zero fill and row/column predicates.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm.h"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
// Two builds of the kernel (template parameter CTAS = CTAs that fit one SM):
//   CTAS 1: 12 epilogue warps (three per TMEM lane quadrant), ~215 KB of shared memory, all 512 TMEM columns — the
//           throughput variant for problems of several waves;
//   CTAS 2: 4 epilogue warps, <= 113 KB, 256 TMEM columns — two CTAs share an SM, so a problem of up to two "waves" is
//           resident at once, consecutive kernels overlap under programmatic dependent launch (the next kernel's CTAs
//           run their prologue next to the draining ones) and a weight-gradient GEMM on the second stream shares SMs
//           with the data-gradient chain.  Two thirds of the step's GEMMs are at most one wave: this is their variant.

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"      // %2: suspend-time hint (ns): wait in hardware, not in the issue slots
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}

// ---- TMA ----------------------------------------------------------------------------------------
// operands are described as rank-4 tensors (inner, rows, batch2, batch1); unbatched problems use 1 x 1
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Output tiles leave through TMA as well: registers -> swizzled staging tile -> one cp.async.bulk.tensor store (or
// f32 reduce-add for split-K partial sums) issued by one lane.  Rows / columns beyond M / N are clipped by the tensor
// map, so the store path carries no predicates and no per-row address arithmetic, and it is asynchronous: the warp
// goes on to the next accumulator slice while the copy engine drains the tile.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type SWIZZLE_128B=2 [61,64).
//   K-major  tile: rows of 64 bf16 (128 B); 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major tile: k-rows of 64 mn-elements (128 B); 8-k groups 1024 B apart (SBO); the next 64
//                  mn-elements start LBO bytes later (one TMA box = 64 k-rows x 128 B = 8 KB).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32 [4,6)=1, A format [7,10), B format [10,13) (0 = f16, 1 = bf16, chosen
// per operand), a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn, int a_bf16, int b_bf16) {
  return (1u << 4) | ((uint32_t)a_bf16 << 7) | ((uint32_t)b_bf16 << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct TcParams {
  void* C; void* Z; const float* bias; const float* residual; const float* row_scale;
  float* rowsum;                    // += sum_k A[k][m] (weight-gradient products: the bias gradient), or NULL
  const float* rowvec;              // per-row f32 input of the attention epilogues, [batch][M]
  int64_t ldc, ldz, ldr;
  int64_t sC1, sC2;                 // batch strides of C (elements)
  int M, N, K;
  int c_dtype, act, accumulate, res_mod, rows_per_scale;
  int a_bf16, b_bf16;               // operand formats of the MMA (0 = f16, 1 = bf16)
  int c_half, z_half;               // 16-bit C / Z stored as f16 (else bf16)
  int splits, kblocks_per_split;
  int batch, batch2;                // batch = batch1 * batch2; z -> (z / batch2, z % batch2)
  float alpha;
};

template <int BN, int CTAS> struct Cfg {
  static constexpr int NUM_EPI_WARPS = CTAS == 1 ? 12 : 4;
  static constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * 32 * 128;   // one 32-row x 128-byte tile per epilogue warp
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BOXES = (BN + 63) / 64;                 // MN-major B: 64-wide boxes
  static constexpr int B_BYTES = B_BOXES * 64 * BK * 2;          // >= BN * BK * 2, multiple of 8 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ONES_BYTES = 2048;                        // all-ones 16 x 64 operand tile (row-sum MMA)
  static constexpr int FIXED_BYTES = STAGING_BYTES + ONES_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  // CTAS 2: (228 KB - 1 KB reserved per CTA) / 2 = 113 KB each
  static constexpr int SMEM_BUDGET = (CTAS == 1 ? 168 * 1024 : 113 * 1024 - FIXED_BYTES);
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 512 / CTAS;
  static constexpr int NACC = (TMEM_COLS / BN) >= 4 ? 4 : ((TMEM_COLS / BN) >= 2 ? 2 : 1);   // accumulator stages in TMEM
  static constexpr int RS_COLS = 32;                             // TMEM columns per accumulator stage for the row sums
  static constexpr bool RS_FITS = NACC * (BN + RS_COLS) <= TMEM_COLS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
  static_assert(STAGES >= 2, "the TMA ring needs two stages");
  static_assert(CTAS == 1 || SMEM_BYTES <= 113 * 1024, "two CTAs must fit one SM");
};

// ---- epilogue helpers: a warp moves a 32-row x 128-byte tile between global memory (coalesced: 8 lanes
// cover one 128-byte row segment, 4 rows per instruction) and its swizzled staging tile, in which
// lane r then owns row r (16-byte chunk j of row r lives at chunk j ^ (r & 7): conflict-free both ways).
// Staging is addressed in the shared window explicitly (ld/st.shared), never through generic pointers.
__device__ __forceinline__ uint32_t stage_addr(uint32_t stage, int row, int chunk) {
  return stage + row * 128 + ((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// lane's 32 f32 values -> 16-bit row (64 B = chunks 0..3) of its staging tile
__device__ __forceinline__ void stage_row16(uint32_t stage, int lane, const float (&v)[32], bool half) {
  if (half) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(stage_addr(stage, lane, j), make_uint4(pack_f162(v[8 * j], v[8 * j + 1]), pack_f162(v[8 * j + 2], v[8 * j + 3]),
                                                    pack_f162(v[8 * j + 4], v[8 * j + 5]), pack_f162(v[8 * j + 6], v[8 * j + 7])));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(stage_addr(stage, lane, j), make_uint4(pack_bf162(v[8 * j], v[8 * j + 1]), pack_bf162(v[8 * j + 2], v[8 * j + 3]),
                                                    pack_bf162(v[8 * j + 4], v[8 * j + 5]), pack_bf162(v[8 * j + 6], v[8 * j + 7])));
  }
}
// Output tile of a 16-bit result: 32 rows x 64 bytes, dense, in the 64-byte TMA swizzle (16-byte chunk j of row r at
// chunk j ^ ((r >> 1) & 3)); `tile` must be 512-byte aligned.  A 16-byte store per lane is conflict-free in this layout.
__device__ __forceinline__ uint32_t out16_addr(uint32_t tile, int row, int chunk) {
  return tile + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ void stage_out16(uint32_t tile, int lane, const float (&v)[32], bool half) {
  if (half) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(out16_addr(tile, lane, j), make_uint4(pack_f162(v[8 * j], v[8 * j + 1]), pack_f162(v[8 * j + 2], v[8 * j + 3]),
                                                   pack_f162(v[8 * j + 4], v[8 * j + 5]), pack_f162(v[8 * j + 6], v[8 * j + 7])));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(out16_addr(tile, lane, j), make_uint4(pack_bf162(v[8 * j], v[8 * j + 1]), pack_bf162(v[8 * j + 2], v[8 * j + 3]),
                                                   pack_bf162(v[8 * j + 4], v[8 * j + 5]), pack_bf162(v[8 * j + 6], v[8 * j + 7])));
  }
}
// lane's 16-bit staging row -> 32 floats
__device__ __forceinline__ void unstage_row16(uint32_t stage, int lane, float (&o)[32], bool half) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 t = lds128(stage_addr(stage, lane, j));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = half ? unpack2<f16>(w[i]) : unpack2<bf16>(w[i]);
      o[8 * j + 2 * i] = f.x;
      o[8 * j + 2 * i + 1] = f.y;
    }
  }
}
// global -> registers -> staging.  `gbase` points at (row 0, first byte) of the 32 x 128 B window; rows beyond
// `rows_valid` and bytes beyond `bytes_valid` read as zero.  The fetch is separated from the put so that the
// next slice's tile can be in flight while the current slice is processed (software prefetch).
__device__ __forceinline__ void stage_fetch(uint4 (&v)[8], const unsigned char* gbase, int64_t pitch_bytes, int rows_valid,
                                            int bytes_valid, int lane) {
  const int chunk = lane & 7;
  const bool col_ok = chunk * 16 < bytes_valid;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int row = 4 * i + (lane >> 3);
    v[i] = make_uint4(0, 0, 0, 0);
    if (col_ok && row < rows_valid) v[i] = __ldg(reinterpret_cast<const uint4*>(gbase + row * pitch_bytes + chunk * 16));
  }
}
__device__ __forceinline__ void stage_put(uint32_t stage, const uint4 (&v)[8], int lane) {
  const int chunk = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) sts128(stage_addr(stage, 4 * i + (lane >> 3), chunk), v[i]);
}
__device__ __forceinline__ void stage_load(uint32_t stage, const unsigned char* gbase, int64_t pitch_bytes, int rows_valid,
                                           int bytes_valid, int lane) {
  uint4 v[8];
  stage_fetch(v, gbase, pitch_bytes, rows_valid, bytes_valid, lane);
  stage_put(stage, v, lane);
}
template <bool ATOMIC>
__device__ __forceinline__ void stage_store(uint32_t stage, unsigned char* gbase, int64_t pitch_bytes, int rows_valid, int bytes_valid,
                                            int lane) {
  const int chunk = lane & 7;
  const bool col_ok = chunk * 16 < bytes_valid;
  uint4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = lds128(stage_addr(stage, 4 * i + (lane >> 3), chunk));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int row = 4 * i + (lane >> 3);
    if (col_ok && row < rows_valid) {
      if (ATOMIC) {
        float4 f = make_float4(__uint_as_float(v[i].x), __uint_as_float(v[i].y), __uint_as_float(v[i].z), __uint_as_float(v[i].w));
        atomicAdd(reinterpret_cast<float4*>(gbase + row * pitch_bytes + chunk * 16), f);
      } else {
        *reinterpret_cast<uint4*>(gbase + row * pitch_bytes + chunk * 16) = v[i];
      }
    }
  }
}

enum { EPI_BF16 = 0, EPI_BF16_GELU = 1, EPI_BF16_DGELU = 2, EPI_F32 = 3, EPI_F32_ATOMIC = 4, EPI_SOFTMAX = 5, EPI_DSOFTMAX = 6,
       // attention with more keys than one tile holds (Lk > 256), in two passes over q.k^T so that neither the f32 scores
       // nor dP ever reach HBM (the contraction is only d = 96 deep, evaluating it twice is cheap):
       EPI_LSE = 7,        // pass 1: C[z][m] = logsumexp_n(alpha * acc)  — one CTA walks all n-tiles of its rows, online max / sum
       EPI_EXPSUB = 8,     // pass 2: C = P = exp(alpha * acc - rowvec[m])                (16-bit)
       EPI_DSOFTMAX_D = 9  // backward: C = dS = alpha * Z o (acc - rowvec[m]),  Z = P, rowvec = rowsum(dO o O)   (16-bit)
};

// One 32-column slice of the accumulator rows owned by this warp -> epilogue math -> global memory.
// m0 = first row of the warp's 32-row band, n0 = first column, `stage` = the warp's staging tile.
// Which tile (if any) a slice reads from global memory before its math: the saved pre-activation Z (GELU
// backward), the f32 residual, or the old C (accumulate).
template <int EPI>
__device__ __forceinline__ bool slice_side_input(const TcParams& p, bool first_split) {
  constexpr bool F32 = EPI == EPI_F32 || EPI == EPI_F32_ATOMIC;
  if (EPI == EPI_BF16_DGELU || EPI == EPI_DSOFTMAX_D) return true;
  if (F32 && p.residual != nullptr && first_split) return true;
  return (EPI == EPI_F32 || EPI == EPI_BF16) && p.accumulate;
}
template <int EPI>
__device__ __forceinline__ void slice_prefetch(const TcParams& p, int64_t coff, int m0, int n0, int cols_in_tile, int lane,
                                               bool first_split, uint4 (&pre)[8]) {
  constexpr bool F32 = EPI == EPI_F32 || EPI == EPI_F32_ATOMIC;
  constexpr int ESZ = F32 ? 4 : 2;
  const int rows_valid = min(32, p.M - m0);
  const int cols_valid = min(min(32, cols_in_tile), p.N - n0);
  if (EPI == EPI_BF16_DGELU) {
    stage_fetch(pre, reinterpret_cast<const unsigned char*>(p.Z) + ((int64_t)m0 * p.ldz + n0) * 2, p.ldz * 2, rows_valid, cols_valid * 2, lane);
  } else if (EPI == EPI_DSOFTMAX_D) {          // P has C's layout, batch strides included
    stage_fetch(pre, reinterpret_cast<const unsigned char*>(p.Z) + (coff + (int64_t)m0 * p.ldz + n0) * 2, p.ldz * 2, rows_valid, cols_valid * 2,
                lane);
  } else if (F32 && p.residual != nullptr && first_split) {
    int rrow = p.res_mod > 0 ? m0 % p.res_mod : m0;
    stage_fetch(pre, reinterpret_cast<const unsigned char*>(p.residual + (int64_t)rrow * p.ldr + n0), p.ldr * 4, rows_valid, cols_valid * 4, lane);
  } else {
    stage_fetch(pre, reinterpret_cast<const unsigned char*>(p.C) + (coff + (int64_t)m0 * p.ldc + n0) * ESZ, p.ldc * ESZ, rows_valid,
                cols_valid * ESZ, lane);
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_slice(const TcParams& p, int64_t coff, uint32_t stage, int m0, int n0, int cols_in_tile,
                                               int lane, uint32_t taddr, float rs, float rv, bool first_split, const uint4 (&pre)[8],
                                               const CUtensorMap* tmc, const CUtensorMap* tmz, int z2, int z1, int& parity) {
  constexpr bool F32 = EPI == EPI_F32 || EPI == EPI_F32_ATOMIC;
  constexpr int ESZ = F32 ? 4 : 2;
  const int rows_valid = min(32, p.M - m0);
  const int cols_valid = min(min(32, cols_in_tile), p.N - n0);
  unsigned char* cg = reinterpret_cast<unsigned char*>(p.C) + (coff + (int64_t)m0 * p.ldc + n0) * ESZ;
  // A 16-bit output tile is 2 KB, half of the warp's staging: plain 16-bit epilogues (no second output, no side tile)
  // alternate between the halves, so only the store issued TWO slices ago has to have read its tile; the others wait for
  // the previous slice's store before the staging is written again.
  constexpr bool ALT = EPI == EPI_BF16 || EPI == EPI_EXPSUB;
  const bool alt = ALT && !(EPI == EPI_BF16 && p.accumulate);
  if (lane == 0) {
    if (alt) bulk_wait_read1();
    else bulk_wait_read0();
  }
  __syncwarp();
  const uint32_t out_tile = stage + ((alt && (parity & 1)) ? 2048u : 0u);
  parity ^= 1;
  // global reads first: their latency overlaps the TMEM load
  float4 bv[8];
  const bool use_bias = p.bias != nullptr && first_split;
  if (use_bias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) bv[i] = (4 * i < cols_valid) ? __ldg(reinterpret_cast<const float4*>(p.bias + n0) + i) : make_float4(0, 0, 0, 0);
  }
  const bool has_res = F32 && p.residual != nullptr && first_split;
  const bool has_acc = (EPI == EPI_F32 || EPI == EPI_BF16) && p.accumulate;
  if (EPI == EPI_BF16_DGELU || EPI == EPI_DSOFTMAX_D || has_res || has_acc) stage_put(stage, pre, lane);      // tile prefetched by the caller
  uint32_t r[32];
  tmem_ld32(taddr, r);
  tmem_ld_wait();
  float v[32];
  if (EPI == EPI_EXPSUB) {                       // P = exp(alpha * s - lse): rv holds lse * log2(e)
    const float a2 = p.alpha * 1.4426950408889634f;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = ex2_ftz(fmaf(__uint_as_float(r[i]), a2, -rv));
  } else if (EPI == EPI_DSOFTMAX_D) {            // dS = alpha * P o (dP - D)
    __syncwarp();
    float pf[32];
    unstage_row16(stage, lane, pf, p.z_half);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = p.alpha * pf[i] * (__uint_as_float(r[i]) - rv);
    __syncwarp();
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
  }
  if (use_bias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[4 * i] += bv[i].x; v[4 * i + 1] += bv[i].y; v[4 * i + 2] += bv[i].z; v[4 * i + 3] += bv[i].w; }
  }
  if (EPI == EPI_BF16_GELU) {
    // one evaluation of (cdf, pdf) yields both gelu(v) and gelu'(v); the derivative is what backward needs,
    // so it is stored (bf16) in Z and the backward epilogue is a plain multiply
    float gp[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float cdf, pdf;
      gelu_parts(v[i], cdf, pdf);
      gp[i] = fmaf(v[i], pdf, cdf);
      v[i] = v[i] * cdf;
    }
    if (p.Z) stage_out16(stage + 2048, lane, gp, p.z_half);     // second 2 KB tile; stored together with C below
  }
  if (EPI == EPI_BF16_DGELU) {
    __syncwarp();
    float zf[32];
    unstage_row16(stage, lane, zf, p.z_half);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= zf[i];
    __syncwarp();
  }
  if (p.row_scale) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= rs;
  }
  if (has_res || (F32 && has_acc)) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 t = lds128(stage_addr(stage, lane, j));
      v[4 * j] += __uint_as_float(t.x); v[4 * j + 1] += __uint_as_float(t.y);
      v[4 * j + 2] += __uint_as_float(t.z); v[4 * j + 3] += __uint_as_float(t.w);
    }
    __syncwarp();
    if (has_res && has_acc) {                      // residual and accumulate together: second pass for the old C
      stage_load(stage, cg, p.ldc * 4, rows_valid, cols_valid * 4, lane);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 t = lds128(stage_addr(stage, lane, j));
        v[4 * j] += __uint_as_float(t.x); v[4 * j + 1] += __uint_as_float(t.y);
        v[4 * j + 2] += __uint_as_float(t.z); v[4 * j + 3] += __uint_as_float(t.w);
      }
      __syncwarp();
    }
  } else if (!F32 && has_acc) {
    __syncwarp();
    float of[32];
    unstage_row16(stage, lane, of, p.c_half);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += of[i];
    __syncwarp();
  }
  if (F32) {                                     // 32 rows x 128 bytes in the 128-byte TMA swizzle (= stage_addr)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts128(stage_addr(stage, lane, j), make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                    __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
  } else {
    stage_out16(out_tile, lane, v, p.c_half);
  }
  fence_async_smem();                            // generic-proxy writes -> visible to the copy engine
  __syncwarp();
  if (lane == 0) {
    if (EPI == EPI_BF16_GELU && p.Z) tma_store_4d(tmz, stage + 2048, n0, m0, 0, 0);
    if (EPI == EPI_F32_ATOMIC) tma_reduce_add_4d(tmc, stage, n0, m0, z2, z1);
    else tma_store_4d(tmc, F32 ? stage : out_tile, n0, m0, z2, z1);
    bulk_commit();
  }
}

// Fused attention-softmax epilogues for key counts that fit one tile (N = Lk <= BN <= 256): a warp owns
// 32 complete rows (lane = row), so the row reductions need no cross-lane traffic at all.
//   EPI_SOFTMAX : acc = q.k^T          -> C = P  = softmax(alpha * acc)                (16-bit, pad columns zero)
//   EPI_DSOFTMAX: acc = dO.v^T (= dP)  -> C = dS = alpha * P o (dP - rowsum(dP o P))   (P read from Z)
// replacing the f32 S / dP round trips through HBM and the separate softmax kernels (attention.py:154-155).
template <int EPI>
__device__ __forceinline__ void softmax_rows(const TcParams& p, int64_t coff, uint32_t stage, int m0, int lane, uint32_t trow) {
  const int rows_valid = min(32, p.M - m0);
  const int ncols = p.N, nstore = (int)p.ldc;
  unsigned char* cg = reinterpret_cast<unsigned char*>(p.C) + (coff + (int64_t)m0 * p.ldc) * 2;
  const unsigned char* pg = reinterpret_cast<const unsigned char*>(p.Z) + (coff + (int64_t)m0 * p.ldz) * 2;
  float m = -INFINITY, l = 0.f, dot = 0.f;
  // the softmax runs in the base-2 domain: t = acc * alpha * log2(e), p = 2^(t - max) — one FMUL/FFMA and one
  // single-instruction ex2.approx.ftz per element (expf's default form wraps the MUFU in a denormal-scaling sequence)
  const float a2 = p.alpha * 1.4426950408889634f;
#pragma unroll 1
  for (int c = 0; c < ncols; c += 32) {
    uint32_t r[32];
    if (EPI == EPI_DSOFTMAX) stage_load(stage, pg + c * 2, p.ldz * 2, rows_valid, min(32, nstore - c) * 2, lane);
    tmem_ld32(trow + c, r);
    tmem_ld_wait();
    if (EPI == EPI_SOFTMAX) {
      float mloc = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = (c + i < ncols) ? __uint_as_float(r[i]) * a2 : -INFINITY;
        r[i] = __float_as_uint(v);
        mloc = fmaxf(mloc, v);
      }
      const float mnew = fmaxf(m, mloc);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) sum += ex2_ftz(__uint_as_float(r[i]) - mnew);
      l = l * ex2_ftz(m - mnew) + sum;
      m = mnew;
    } else {
      __syncwarp();
      float pf[32];
      unstage_row16(stage, lane, pf, p.z_half);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c + i < ncols) dot = fmaf(pf[i], __uint_as_float(r[i]), dot);
      __syncwarp();
    }
  }
  const float inv = 1.f / l;
#pragma unroll 1
  for (int c = 0; c < nstore; c += 32) {
    uint32_t r[32];
    float v[32];
    if (EPI == EPI_DSOFTMAX) stage_load(stage, pg + c * 2, p.ldz * 2, rows_valid, min(32, nstore - c) * 2, lane);
    tmem_ld32(trow + c, r);
    tmem_ld_wait();
    if (EPI == EPI_SOFTMAX) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (c + i < ncols) ? ex2_ftz(fmaf(__uint_as_float(r[i]), a2, -m)) * inv : 0.f;
    } else {
      __syncwarp();
      float pf[32];
      unstage_row16(stage, lane, pf, p.z_half);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (c + i < ncols) ? p.alpha * pf[i] * (__uint_as_float(r[i]) - dot) : 0.f;
      __syncwarp();
    }
    stage_row16(stage, lane, v, p.c_half);
    __syncwarp();
    stage_store<false>(stage, cg + c * 2, p.ldc * 2, rows_valid, min(32, nstore - c) * 2, lane);
    __syncwarp();
  }
}

// One n-tile of the logsumexp pass: lane = row, running (max, sum) of t = alpha * log2(e) * acc in the base-2 domain.
__device__ __forceinline__ void lse_tile(const TcParams& p, uint32_t trow, int n_base, int bn, float& m, float& l) {
  const float a2 = p.alpha * 1.4426950408889634f;
#pragma unroll 1
  for (int c = 0; c < bn && n_base + c < p.N; c += 32) {
    uint32_t r[32];
    tmem_ld32(trow + c, r);
    tmem_ld_wait();
    float t[32], mloc = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      t[i] = (n_base + c + i < p.N) ? __uint_as_float(r[i]) * a2 : -INFINITY;
      mloc = fmaxf(mloc, t[i]);
    }
    const float mnew = fmaxf(m, mloc);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) sum += ex2_ftz(t[i] - mnew);
    l = l * ex2_ftz(m - mnew) + sum;
    m = mnew;
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CTAS>
__global__ void __launch_bounds__((Cfg<BN, CTAS>::NUM_THREADS), CTAS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_z, TcParams p) {
  using C = Cfg<BN, CTAS>;
  constexpr int NUM_EPI_WARPS = C::NUM_EPI_WARPS;
  constexpr int STAGING_BYTES = C::STAGING_BYTES;
  extern __shared__ unsigned char smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + C::STAGES * C::A_BYTES;
  unsigned char* sStage = smem + C::STAGES * C::STAGE_BYTES;
  unsigned char* sOnes = sStage + STAGING_BYTES;                 // 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + C::ONES_BYTES);
  // Row sums of the A operand (the bias gradient that belongs to a weight-gradient product dW = dY^T . X): one extra
  // N = 16 MMA per k-step against an all-ones tile, into 32 spare TMEM columns per accumulator stage, for the
  // tile_n == 0 tiles only.  Any layout of an all-ones tile is an all-ones tile, so one 2 KB region serves all k.
  constexpr bool RS_OK = A_MN && B_MN && (EPI == EPI_F32 || EPI == EPI_F32_ATOMIC) && C::RS_FITS;
  const bool rowsum_on = RS_OK && p.rowsum != nullptr;
  // bars: full[STAGES], empty[STAGES], tmem_full[NACC], tmem_empty[NACC]
  uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * C::STAGES;
  uint32_t tfull0 = empty0 + 8 * C::STAGES, tempty0 = tfull0 + 8 * C::NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 2 * C::NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  // EPI_LSE: a work item is a row band and the CTA walks its n-tiles itself (the row statistics live in registers)
  constexpr bool NINNER = EPI == EPI_LSE;
  const int n_rep = NINNER ? tiles_n : 1;
  const int num_items = tiles_m * (NINNER ? 1 : tiles_n) * p.splits * p.batch;
  const int kblocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    constexpr bool ROWS = EPI == EPI_SOFTMAX || EPI == EPI_DSOFTMAX || EPI == EPI_LSE;      // one warp per quadrant drains a tile
    for (int s = 0; s < C::NACC; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, ROWS ? 4 : NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
  if (rowsum_on) {
    const uint32_t one2 = p.a_bf16 ? 0x3F803F80u : 0x3C003C00u;   // 1.0 twice, in the operand format
    for (int i = threadIdx.x; i < C::ONES_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sOnes)[i] = one2;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch)
  // overlaps the tail of the previous kernel in the stream; global memory is first touched below.
  pdl_wait();

  // work item -> (tile_m, tile_n, split); consecutive items (= CTAs running at the same time) walk N first: the CTAs that
  // share an A tile read it within microseconds of each other, so it comes from HBM once and from L2 for the other
  // n-tiles.  B (a weight, or one head's keys) is small and stays in L2 throughout.  (Walking M first re-read a
  // 100 MB activation once per n-tile: measured 402 MB of DRAM reads for a 100 MB operand.)
  auto decode = [&](int item, int& tm, int& tn, int& z, int& kb0, int& kb1) {
    int split = item % p.splits;
    int t = item / p.splits;
    tn = 0;
    if (!NINNER) {
      tn = t % tiles_n;
      t /= tiles_n;
    }
    tm = t % tiles_m;
    z = t / tiles_m;
    kb0 = split * p.kblocks_per_split;
    kb1 = min(kblocks, kb0 + p.kblocks_per_split);
    return split;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int tm, tn, z, kb0, kb1;
        decode(item, tm, tn, z, kb0, kb1);
        const int z1 = z / p.batch2, z2 = z % p.batch2;
        for (int rep = 0; rep < n_rep; ++rep) {
        if (NINNER) tn = rep;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          mbar_expect_tx(full0 + 8 * stage, C::A_BYTES + (B_MN ? C::B_BOXES * 8192 : BN * BK * 2));
          const uint32_t a_dst = smem_u32(sA + stage * C::A_BYTES), b_dst = smem_u32(sB + stage * C::B_BYTES);
          const uint32_t bar = full0 + 8 * stage;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_4d(a_dst + j * 8192, &tmap_a, bar, tm * BM + j * 64, kb * BK, z2, z1);
          } else {
            tma_load_4d(a_dst, &tmap_a, bar, kb * BK, tm * BM, z2, z1);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < C::B_BOXES; ++j) tma_load_4d(b_dst + j * 8192, &tmap_b, bar, tn * BN + j * 64, kb * BK, z2, z1);
          } else {
            tma_load_4d(b_dst, &tmap_b, bar, kb * BK, tn * BN, z2, z1);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0, p.a_bf16, p.b_bf16);
      const uint32_t idesc_rs = make_idesc(BM, 16, A_MN ? 1 : 0, 0, p.a_bf16, p.a_bf16);
      const uint64_t odesc = make_sw128_desc(smem_u32(sOnes), 16, 1024);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int tm, tn, z, kb0, kb1;
        decode(item, tm, tn, z, kb0, kb1);
        for (int rep = 0; rep < n_rep; ++rep) {
        mbar_wait(tempty0 + 8 * as, aphase ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        const uint32_t tmem_rs = tmem_base + C::NACC * BN + as * C::RS_COLS;
        const bool rs_tile = rowsum_on && tn == 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * C::A_BYTES), b_addr = smem_u32(sB + stage * C::B_BYTES);
          // K-major : advance 16 elements (32 B) inside the swizzle atom per UMMA_K
          // MN-major: advance 16 k-rows (16 x 128 B = 2048 B) per UMMA_K
          const uint64_t adesc = A_MN ? make_sw128_desc(a_addr, 8192, 1024) : make_sw128_desc(a_addr, 16, 1024);
          const uint64_t bdesc = B_MN ? make_sw128_desc(b_addr, 8192, 1024) : make_sw128_desc(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_bf16(tmem_d, adesc + (uint64_t)((A_MN ? 2048 : 32) >> 4) * k, bdesc + (uint64_t)((B_MN ? 2048 : 32) >> 4) * k, idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
            if (rs_tile) umma_bf16(tmem_rs, adesc + (uint64_t)((A_MN ? 2048 : 32) >> 4) * k, odesc, idesc_rs, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * stage);           // smem slot free once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull0 + 8 * as);                // accumulator complete
        if (++as == C::NACC) { as = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..13): three warps per TMEM lane quadrant =====================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int sub = (warp - 2) >> 2;                 // which of the quadrant's three warps
    const uint32_t stage_buf = smem_u32(sStage) + (warp - 2) * (32 * 128);
    if (EPI == EPI_LSE) {
      // logsumexp pass: the first warp of every quadrant follows its 32 rows across the n-tiles of the band
      int as = 0; uint32_t aphase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int tm, tn, z, kb0, kb1;
        decode(item, tm, tn, z, kb0, kb1);
        const int m0 = tm * BM + quad * 32;
        float m = -INFINITY, l = 0.f;
        for (int rep = 0; rep < n_rep; ++rep) {
          if (sub == 0) {
            mbar_wait(tfull0 + 8 * as, aphase);
            tc_fence_after();
            if (m0 < p.M) lse_tile(p, tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN, rep * BN, BN, m, l);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * as);
          }
          if (++as == C::NACC) { as = 0; aphase ^= 1; }
        }
        if (sub == 0 && m0 + lane < p.M)
          reinterpret_cast<float*>(p.C)[(int64_t)z * p.M + m0 + lane] = (m + __log2f(l)) * 0.6931471805599453f;
      }
    } else if (EPI == EPI_SOFTMAX || EPI == EPI_DSOFTMAX) {
      // whole-row epilogues: the quadrant's warps take turns on successive tiles (tile seq -> warp seq % NSUB),
      // so up to NSUB tiles per quadrant are drained concurrently and no cross-warp reduction exists
      constexpr int NSUB = (C::NACC < NUM_EPI_WARPS / 4 ? C::NACC : NUM_EPI_WARPS / 4);
      int seq = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++seq) {
        if (sub >= NSUB || seq % NSUB != sub) continue;
        const int as = seq % C::NACC;
        const uint32_t aphase = (uint32_t)(seq / C::NACC) & 1u;
        int tm, tn, z, kb0, kb1;
        decode(item, tm, tn, z, kb0, kb1);
        const int64_t coff = (int64_t)(z / p.batch2) * p.sC1 + (int64_t)(z % p.batch2) * p.sC2;
        const int m0 = tm * BM + quad * 32;
        mbar_wait(tfull0 + 8 * as, aphase);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
        if (m0 < p.M) softmax_rows<EPI>(p, coff, stage_buf, m0, lane, trow);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      }
    } else {
    int as = 0; uint32_t aphase = 0;
    int out_parity = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int tm, tn, z, kb0, kb1;
      const int split = decode(item, tm, tn, z, kb0, kb1);
      const int64_t coff = (int64_t)(z / p.batch2) * p.sC1 + (int64_t)(z % p.batch2) * p.sC2;
      const int m0 = tm * BM + quad * 32;
      float rs = 1.f, rv = 0.f;
      if (p.row_scale) rs = p.row_scale[min(m0 + lane, p.M - 1) / p.rows_per_scale];
      if (EPI == EPI_EXPSUB || EPI == EPI_DSOFTMAX_D) {
        rv = p.rowvec[(int64_t)z * p.M + min(m0 + lane, p.M - 1)];
        if (EPI == EPI_EXPSUB) rv *= 1.4426950408889634f;
      }
      constexpr int EPI_S = (EPI == EPI_SOFTMAX || EPI == EPI_DSOFTMAX || EPI == EPI_LSE) ? EPI_BF16 : EPI;
      constexpr int CSTEP = (NUM_EPI_WARPS / 4) * 32;
      const bool live = m0 < p.M && kb1 > kb0;       // rows beyond M / an empty k-range contribute nothing
      const bool side = live && slice_side_input<EPI_S>(p, split == 0);
      // side input (Z / residual / old C) of the first slice is requested before the accumulator wait, the
      // next slice's while the current one is being processed
      uint4 pre[8];
      {
        const int c = sub * 32, n0 = tn * BN + c;
        if (side && c < BN && n0 < p.N) slice_prefetch<EPI_S>(p, coff, m0, n0, BN - c, lane, split == 0, pre);
      }
      mbar_wait(tfull0 + 8 * as, aphase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = sub * 32; c < BN; c += CSTEP) {
        const int n0 = tn * BN + c;
        if (!live || n0 >= p.N) continue;
        uint4 cur[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = pre[i];
        const int cn = c + CSTEP, n1 = tn * BN + cn;
        if (side && cn < BN && n1 < p.N) slice_prefetch<EPI_S>(p, coff, m0, n1, BN - cn, lane, split == 0, pre);
        epilogue_slice<EPI_S>(p, coff, stage_buf, m0, n0, BN - c, lane, trow + c, rs, rv, split == 0, cur, &tmap_c, &tmap_z, z % p.batch2,
                              z / p.batch2, out_parity);
      }
      if (rowsum_on && tn == 0 && sub == 0 && live) {      // lane = output row: its sum over this item's k-range
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + C::NACC * BN + as * C::RS_COLS, r);
        tmem_ld_wait();
        if (m0 + lane < p.M) atomicAdd(p.rowsum + m0 + lane, __uint_as_float(r[0]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      if (++as == C::NACC) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) bulk_wait0();                     // every output tile has reached global memory before the CTA retires
    __syncwarp();
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

// [batch1][batch2][rows][cols] tensor, row pitch ld, batch strides s1 / s2 (elements); out-of-bounds elements read as
// zero / are not written.  kind 0: 16-bit GEMM operand, box [box_rows x 64 cols], 128-byte swizzle;
// kind 1: 16-bit output, box [32 x 32] (64-byte rows, 64-byte swizzle); kind 2: f32 output, box [32 x 32] (128-byte rows,
// 128-byte swizzle).
int encode_tmap(CUtensorMap* map, int kind, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int batch1, int batch2,
                int64_t s1, int64_t s2) {
  EncodeTiledFn fn = get_encode_fn();
  CSTS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old / no GPU)");
  const cuuint64_t esz = kind == 2 ? 4 : 2;
  cuuint64_t gdim[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch2, (cuuint64_t)batch1};
  // a dimension of extent 1 still needs a legal (16-byte multiple) stride
  cuuint64_t gstride[3] = {(cuuint64_t)ld * esz, (cuuint64_t)(batch2 > 1 ? s2 * esz : ld * esz), (cuuint64_t)(batch1 > 1 ? s1 * esz : ld * esz)};
  cuuint32_t box[4] = {kind == 0 ? 64u : 32u, (cuuint32_t)box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = kind == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  const CUtensorMapSwizzle sw = kind == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(map, dt, 4, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CSTS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) kind=%d rows=%ld cols=%ld ld=%ld batch=%dx%d strides %ld %ld", (int)r,
               kind, (long)rows, (long)cols, (long)ld, batch1, batch2, (long)s1, (long)s2);
  return 0;
}

// A descriptor is a pure function of (base, extents, strides, box): the training step presents the same few hundred
// tensors every iteration (the caching allocator hands back the same blocks), so encoded maps are kept, keyed by those
// values.  The cache is bounded (cleared when full) and guarded by a mutex: autograd's backward thread and the forward
// thread both launch GEMMs.
struct TmapKey {
  const void* base; int64_t rows, cols, ld, s1, s2; int box_rows, batch1, batch2, kind;
  bool operator==(const TmapKey& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return (size_t)h;
  }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed word-wise");

int make_tmap(CUtensorMap* map, int kind, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int batch1, int batch2,
              int64_t s1, int64_t s2) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key;
  std::memset(&key, 0, sizeof(key));
  key.base = base; key.rows = rows; key.cols = cols; key.ld = ld; key.s1 = batch1 > 1 ? s1 : 0; key.s2 = batch2 > 1 ? s2 : 0;
  key.box_rows = box_rows; key.batch1 = batch1; key.batch2 = batch2; key.kind = kind;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *map = it->second; return 0; }
  }
  int rc = encode_tmap(map, kind, base, rows, cols, ld, box_rows, batch1, batch2, s1, s2);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() >= 16384) cache.clear();
  cache.emplace(key, *map);
  return 0;
}

// What the launcher decides per problem: tile width, CTAs per SM (kernel build) and split-K factor.
struct Plan { int bn, ctas, splits; };

template <int BN, bool A_MN, bool B_MN, int EPI, int CTAS>
int launch(const csts_gemm_args& a, const Plan& plan, cudaStream_t stream) {
  using C = Cfg<BN, CTAS>;
  static std::atomic<bool> attr_set{false};      // set from the forward thread and from autograd's backward thread
  if (!attr_set.load(std::memory_order_acquire)) {
    CSTS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN, EPI, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set.store(true, std::memory_order_release);
  }
  CUtensorMap ta, tb;
  // K-major operand X[mn][k]: tensor [mn rows, K cols], box [tile rows, 64 k].
  // MN-major operand X[k][mn]: tensor [K rows, mn cols], box [64 k rows, 64 mn].
  int rc = A_MN ? make_tmap(&ta, 0, a.A, a.K, a.M, a.lda, BK, a.batch1, a.batch2, a.sA1, a.sA2)
                : make_tmap(&ta, 0, a.A, a.M, a.K, a.lda, BM, a.batch1, a.batch2, a.sA1, a.sA2);
  if (rc) return rc;
  rc = B_MN ? make_tmap(&tb, 0, a.B, a.K, a.N, a.ldb, BK, a.batch1, a.batch2, a.sB1, a.sB2)
            : make_tmap(&tb, 0, a.B, a.N, a.K, a.ldb, BN, a.batch1, a.batch2, a.sB1, a.sB2);
  if (rc) return rc;
  // output tiles (32 x 32) leave through TMA stores / f32 reduce-adds; the whole-row softmax epilogues store directly
  CUtensorMap tc, tz;
  if (EPI == EPI_LSE) {
    tc = ta;                                     // the logsumexp pass writes one float per row with plain stores
  } else {
    rc = make_tmap(&tc, a.c_dtype == 0 ? 2 : 1, a.C, a.M, a.N, a.ldc, 32, a.batch1, a.batch2, a.sC1, a.sC2);
    if (rc) return rc;
  }
  tz = tc;
  if (EPI == EPI_BF16_GELU && a.Z) {
    rc = make_tmap(&tz, 1, a.Z, a.M, a.N, a.ldz, 32, 1, 1, 0, 0);
    if (rc) return rc;
  }
  TcParams p;
  p.C = a.C; p.Z = a.Z; p.bias = a.bias; p.residual = a.residual; p.rowsum = a.rowsum; p.rowvec = a.rowvec;
  p.row_scale = a.row_scale; p.rows_per_scale = a.rows_per_scale > 0 ? a.rows_per_scale : 1;
  p.ldc = a.ldc; p.ldz = a.ldz; p.ldr = a.ldr;
  p.sC1 = a.sC1; p.sC2 = a.sC2; p.batch = a.batch1 * a.batch2; p.batch2 = a.batch2;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.c_dtype = a.c_dtype; p.act = a.act; p.accumulate = a.accumulate; p.res_mod = a.res_mod; p.alpha = a.alpha;
  p.a_bf16 = a.a_dtype == CSTS_F16 ? 0 : 1; p.b_bf16 = a.b_dtype == CSTS_F16 ? 0 : 1;
  p.c_half = a.c_dtype == CSTS_F16; p.z_half = a.z_dtype == CSTS_F16;
  const int kblocks = ceil_div(a.K, BK);
  p.kblocks_per_split = ceil_div(kblocks, plan.splits);
  p.splits = ceil_div(kblocks, p.kblocks_per_split);
  if (p.splits > 1 && !a.accumulate) {
    CSTS_REQUIRE(p.batch == 1, "gemm_tc: split-K without accumulate supports a single batch");
    CSTS_CUDA(cudaMemset2DAsync(a.C, a.ldc * sizeof(float), 0, (size_t)a.N * sizeof(float), a.M, stream));
  }
  const int items = ceil_div(a.M, BM) * (EPI == EPI_LSE ? 1 : ceil_div(a.N, BN)) * p.splits * p.batch;
  const int slots = csts_num_sms() * CTAS;
  const int grid = items < slots ? items : slots;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSTS_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, A_MN, B_MN, EPI, CTAS>, ta, tb, tc, tz, p));
  return csts_check_launch("gemm_tc_kernel");
}

int normalized_splits(int K, int splits) {
  const int kblocks = ceil_div(K, BK);
  if (splits < 1) splits = 1;
  if (splits > kblocks) splits = kblocks;
  return ceil_div(kblocks, ceil_div(kblocks, splits));
}

// Launch-time model used to rank (tile width, CTAs per SM, split-K) candidates: a work item costs a fixed part
// (pipeline fill, first TMA round trip, accumulator drain), its k-blocks and its epilogue; items run in waves over the
// CTA slots.  Nanoseconds, calibrated on the per-shape event timings in profiles/ (tuning/gemm_tune.py re-measures it).
double model_ns(const csts_gemm_args& a, const Plan& pl) {
  const int sms = csts_num_sms();
  const int kblocks = ceil_div(a.K, BK);
  const int kb = ceil_div(kblocks, pl.splits);
  const long items = (long)ceil_div(a.M, BM) * ceil_div(a.N, pl.bn) * pl.splits * a.batch1 * a.batch2;
  const long slots = (long)sms * pl.ctas;
  const long waves = (items + slots - 1) / slots;
  const bool shared_pipe = pl.ctas == 2 && items > sms;          // two resident CTAs share the SM's tensor pipe and LSU
  const bool f32_out = a.c_dtype == 0;
  const double per_col = 1.75;                                   // 128 x 64 x 2 flop per column and k-block at 9.4 TF/s per SM
  double epi = pl.bn * (f32_out ? 14.0 : 8.0) * (pl.splits > 1 ? 1.5 : 1.0);
  if (a.act == 1 || a.act == 2) epi *= 1.6;
  const bool side_input = a.act == 2 || a.residual != nullptr || a.accumulate;
  if (pl.ctas == 2 && !side_input) epi *= 2.6;                   // 4 epilogue warps instead of 12; epilogues that wait for a
                                                                 // side tile (Z, residual, old C) are latency-bound either way
  const double fixed = pl.ctas == 2 ? 1400.0 : 2200.0;
  const double tile = fixed + kb * (pl.bn * per_col * (shared_pipe ? 1.7 : 1.0) + 45.0) + epi;
  const double memset_ns = (pl.splits > 1 && !a.accumulate) ? 2500.0 : 0.0;
  return waves * tile + memset_ns;
}

bool small_variant_exists(int bn, int act) { return bn != 256 && act != 3 && act != 4 && act != 5; }

// Measured plans (benchmarks/tune_gemm.py on a B200) for the problems of the benchmarked training step
struct TunedPlan { int M, N, K, batch, a_kmajor, b_kmajor, act, c16, rowsum, auto_split, bn, ctas, splits; };
const TunedPlan kTuned[] = {
#include "gemm_tune.inc"
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};

Plan plan_for(const csts_gemm_args& a) {
  static const bool use_table = getenv("CSTS_GEMM_NO_TABLE") == nullptr;
  if (use_table && a.tile_n == 0 && a.ctas == 0 && (a.split_k < 0 || a.split_k == 0 || a.split_k == 1)) {
    const int nb = a.batch1 * a.batch2;
    for (const TunedPlan* t = kTuned; t->M != 0; ++t) {
      if (t->M == a.M && t->N == a.N && t->K == a.K && t->batch == nb && t->a_kmajor == (a.a_kmajor != 0) && t->b_kmajor == (a.b_kmajor != 0) &&
          t->act == a.act && t->c16 == (a.c_dtype != 0) && t->rowsum == (a.rowsum != nullptr) && t->auto_split == (a.split_k < 0)) {
        if (t->ctas == 2 && !small_variant_exists(t->bn, a.act)) break;
        return Plan{t->bn, t->ctas, normalized_splits(a.K, a.split_k < 0 ? t->splits : 1)};
      }
    }
  }
  static const int force_bn = getenv("CSTS_FORCE_BN") ? atoi(getenv("CSTS_FORCE_BN")) : 0;        // tuning experiments only
  static const int force_ctas = getenv("CSTS_FORCE_CTAS") ? atoi(getenv("CSTS_FORCE_CTAS")) : 0;
  const int cands[4] = {256, 192, 128, 96};
  const int kblocks = ceil_div(a.K, BK);
  int bn_lo = 0, bn_hi = 3;
  int fixed_bn = a.tile_n ? a.tile_n : force_bn;
  if (a.act == 3 || a.act == 4) fixed_bn = a.N <= 96 ? 96 : (a.N <= 128 ? 128 : (a.N <= 192 ? 192 : 256));   // one tile holds all keys
  else if (a.act == 5) fixed_bn = 256;           // logsumexp pass: two 256-column accumulators alternate
  else if (a.rowsum) fixed_bn = a.N <= 96 ? 96 : 192;            // tile widths that leave TMEM columns for the fused row sums
  int best_pad = 1 << 30;
  for (int i = 0; i < 4; ++i) best_pad = std::min(best_pad, ceil_div(a.N, cands[i]) * cands[i]);
  const int split_opts[12] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64};
  Plan best = {96, 1, 1};
  double best_ns = 1e30;
  for (int i = bn_lo; i <= bn_hi; ++i) {
    const int bn = cands[i];
    if (fixed_bn ? bn != fixed_bn : ceil_div(a.N, bn) * bn != best_pad) continue;
    for (int ctas = 1; ctas <= 2; ++ctas) {
      if (ctas == 2 && !small_variant_exists(bn, a.act)) continue;
      const int want_ctas = a.ctas ? a.ctas : force_ctas;
      if (want_ctas && ctas != want_ctas && !(want_ctas == 2 && !small_variant_exists(bn, a.act))) continue;
      for (int si = 0; si < 12; ++si) {
        int sp = split_opts[si];
        if (a.split_k >= 0) { if (si > 0) break; sp = a.split_k > 1 ? a.split_k : 1; }      // caller's choice
        else if (sp > 1 && ceil_div(kblocks, sp) < 4) break;                                 // auto: >= 4 k-blocks per split
        Plan pl = {bn, ctas, normalized_splits(a.K, sp)};
        const double ns = model_ns(a, pl);
        if (ns < best_ns) { best_ns = ns; best = pl; }
      }
    }
  }
  return best;
}

template <bool A_MN, bool B_MN, int EPI>
int dispatch(const csts_gemm_args& a, const Plan& pl, cudaStream_t stream) {
  constexpr bool ROWS = EPI == EPI_SOFTMAX || EPI == EPI_DSOFTMAX || EPI == EPI_LSE;
  if (pl.ctas == 2) {
    if constexpr (!ROWS) {
      switch (pl.bn) {
        case 192: return launch<192, A_MN, B_MN, EPI, 2>(a, pl, stream);
        case 128: return launch<128, A_MN, B_MN, EPI, 2>(a, pl, stream);
        case 96: return launch<96, A_MN, B_MN, EPI, 2>(a, pl, stream);
      }
    }
    csts_set_error("gemm_tc: no two-CTA build for tile width %d / epilogue %d", pl.bn, EPI);
    return 2;
  }
  switch (pl.bn) {
    case 256: return launch<256, A_MN, B_MN, EPI, 1>(a, pl, stream);
    case 192: return launch<192, A_MN, B_MN, EPI, 1>(a, pl, stream);
    case 128: return launch<128, A_MN, B_MN, EPI, 1>(a, pl, stream);
    case 96: return launch<96, A_MN, B_MN, EPI, 1>(a, pl, stream);
  }
  csts_set_error("gemm_tc: no tile width divides N=%d", a.N);
  return 2;
}

}  // namespace

bool csts_gemm_tc_supported(const csts_gemm_args& a) {
  if (!a.a_kmajor && a.b_kmajor) return false;         // (MN-major A, K-major B) never occurs on the path
  if (a.a_dtype != a.b_dtype) return false;            // one kind::f16 MMA cannot mix f16 and bf16 operands (faults on sm_100a)
  const int nb = a.batch1 * a.batch2;
  if (a.lda % 8 != 0 || a.ldb % 8 != 0) return false;
  if (((uintptr_t)a.A & 15) || ((uintptr_t)a.B & 15)) return false;
  if (nb > 1 && ((a.sA1 | a.sA2 | a.sB1 | a.sB2) % 8 != 0)) return false;
  const int cbytes = a.c_dtype == 0 ? 4 : 2;
  // coalesced epilogue: 16-byte aligned rows, whole 16-byte chunks
  if (((uintptr_t)a.C & 15) || (a.ldc * cbytes) % 16 || ((int64_t)a.N * cbytes) % 16) return false;
  if (nb > 1 && (((a.sC1 | a.sC2) * cbytes) % 16 != 0)) return false;
  if (a.Z && a.act != 4 && a.act != 7 && (nb > 1 || ((uintptr_t)a.Z & 15) || (a.ldz * 2) % 16)) return false;
  if (a.residual && (nb > 1 || ((uintptr_t)a.residual & 15) || (a.ldr * 4) % 16)) return false;
  if (a.res_mod > 0 && a.res_mod % 32 != 0) return false;
  if (a.bias && (((uintptr_t)a.bias & 15) || a.N % 4 != 0)) return false;
  if ((a.split_k > 1 || a.split_k < 0) && (a.c_dtype != 0 || a.act != 0 || a.row_scale)) return false;
  if ((a.split_k > 1 || a.split_k < 0) && !a.accumulate && a.batch1 * a.batch2 != 1) return false;
  if (a.act == 3 || a.act == 4) {                      // fused softmax / softmax-backward rows
    if (!a.a_kmajor || !a.b_kmajor || a.c_dtype == 0 || a.N > 256 || a.bias || a.residual || a.accumulate || a.row_scale ||
        a.split_k > 1 || a.ldc % 8 != 0 || a.ldc < a.N || a.M < 64)
      return false;
    if (a.act == 4 && (!a.Z || a.ldz != a.ldc || ((uintptr_t)a.Z & 15))) return false;
    return true;
  }
  if (a.act >= 5 && a.act <= 7) {                      // two-pass attention epilogues (more keys than one tile holds)
    if (!a.a_kmajor || !a.b_kmajor || a.bias || a.residual || a.accumulate || a.row_scale || a.split_k > 1 || a.split_k < 0 || a.M < 64)
      return false;
    if (a.act == 5) return a.c_dtype == 0 && !a.Z;                                   // C = logsumexp rows, f32 [batch][M]
    if (!a.rowvec || a.c_dtype == 0 || a.ldc % 8 != 0) return false;
    if (a.act == 7 && (!a.Z || a.ldz != a.ldc || ((uintptr_t)a.Z & 15))) return false;
    return true;
  }
  if (a.c_dtype == 0 && a.act != 0) return false;      // activations pair with 16-bit outputs only
  if (a.act != 0 && (a.accumulate || a.residual)) return false;
  if (a.c_dtype != 0 && a.residual) return false;
  // skinny problems go to the generic kernel — unless they stream a large operand (the (1,8,8) frame pools: 32 rows against
  // a 75 MB weight): those are bandwidth problems, and TMA + a deep ring move bytes faster than cp.async, whatever the tile's
  // fill (rows beyond M are zero-filled by the tensor map and clipped by the TMA store)
  if (a.M < 64 && (int64_t)a.N * a.K * nb < (1 << 24)) return false;
  return true;
}

int csts_gemm_tc_plan(const csts_gemm_args& a, int* bn, int* ctas, int* splits) {
  Plan pl = plan_for(a);
  *bn = pl.bn; *ctas = pl.ctas; *splits = pl.splits;
  return 0;
}

int csts_gemm_tc_launch(const csts_gemm_args& a, cudaStream_t stream) {
  CSTS_REQUIRE(csts_gemm_tc_supported(a), "gemm_tc: unsupported problem (M=%d N=%d K=%d)", a.M, a.N, a.K);
  if (a.act == 2) CSTS_REQUIRE(a.Z != nullptr, "gemm: act==2 needs Z");
  const Plan pl = plan_for(a);
  const bool atomic = pl.splits > 1;
  if (a.act == 3) return dispatch<false, false, EPI_SOFTMAX>(a, pl, stream);
  if (a.act == 4) return dispatch<false, false, EPI_DSOFTMAX>(a, pl, stream);
  if (a.act == 5) return dispatch<false, false, EPI_LSE>(a, pl, stream);
  if (a.act == 6) return dispatch<false, false, EPI_EXPSUB>(a, pl, stream);
  if (a.act == 7) return dispatch<false, false, EPI_DSOFTMAX_D>(a, pl, stream);
  if (!a.a_kmajor) {                                    // (MN, MN): weight gradients, dV / dK of attention
    if (a.c_dtype != 0) return dispatch<true, true, EPI_BF16>(a, pl, stream);
    return atomic ? dispatch<true, true, EPI_F32_ATOMIC>(a, pl, stream) : dispatch<true, true, EPI_F32>(a, pl, stream);
  }
  if (!a.b_kmajor) {                                    // (K, MN): P.V and dS.K of attention; dX = dY . W of every Linear
    CSTS_REQUIRE((a.act == 0 || a.act == 2) && !atomic, "gemm_tc: (K-major, MN-major) products have no GELU / split-K epilogue");
    if (a.act == 2) return dispatch<false, true, EPI_BF16_DGELU>(a, pl, stream);
    return a.c_dtype != 0 ? dispatch<false, true, EPI_BF16>(a, pl, stream) : dispatch<false, true, EPI_F32>(a, pl, stream);
  }
  if (a.c_dtype == 0) return atomic ? dispatch<false, false, EPI_F32_ATOMIC>(a, pl, stream) : dispatch<false, false, EPI_F32>(a, pl, stream);
  if (a.act == 1) return dispatch<false, false, EPI_BF16_GELU>(a, pl, stream);
  if (a.act == 2) return dispatch<false, false, EPI_BF16_DGELU>(a, pl, stream);
  return dispatch<false, false, EPI_BF16>(a, pl, stream);
}
