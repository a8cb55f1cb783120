// Shared pieces of the fused row-tile kernels (fused_post.cu, fused_cpe.cu, fused_qkv.cu): PTX wrappers for
// mbarrier / TMA (2-D tensor maps, gather4, bulk stores) / tcgen05, the fp16 hi-lo split, the fast GELU, and the
// host-side tensor-map encoder.  sm_100a only.
//
// Every fused kernel works on 128-row tiles with THREAD == ROW == TMEM LANE for the 128 "row threads" (warps 0-3):
// row-local math (bias, GELU, residual, LayerNorm) needs no cross-thread traffic, accumulators are read with
// tcgen05.ld (32x32b), turned into the next GEMM's fp16 hi/lo A operand in registers and written back to tensor
// memory with tcgen05.st, so chained skinny GEMMs never leave the SM.  Global memory is only touched by TMA.
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#include "common.cuh"

namespace fz {

constexpr int BM = 128;                // rows per tile
constexpr int KC = 32;                 // fp32 elements per k-chunk (one 128-byte swizzle span per row)
constexpr int NT = 128;                // packed-weight N tile (cdseg_gemm_pack_b layout)
constexpr int BLK = NT * KC;           // fp16 elements of one packed hi (or lo) block
constexpr int IN_STAGE = BM * KC * 4;  // 16 KB: one [128 x 32] fp32 box
constexpr int B_STAGE = 2 * BLK * 2;   // 16 KB: hi block + lo block
constexpr long long WAIT_TIMEOUT_CYCLES = 4000000000ll;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) __trap();       // a protocol bug must trap, never hang
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// four rows r0..r3 (any order, negative / out-of-range rows are zero-filled) x the box width at column c0
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no-swizzle shared-memory operand descriptor (8x8 core matrices of 128 B): the cdseg_gemm_pack_b block layout
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// kind::f16 instruction descriptor: D fp32, A = B = fp16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
// D[tmem] (+)= A[tmem] . B[smem]   (A: lane = row, two fp16 K elements per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) { tmem_ld16(taddr, r); tmem_ld16(taddr + 16, r + 16); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
               "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 fp32 values of one row -> the [hi 16 | lo 16] 32-bit columns of a K = 32 A-operand chunk
// (x = hi + lo, hi = fp16(x), lo = fp16(x - hi): 22 significant bits, see gemm_tc.cu)
__device__ __forceinline__ void split_store(const float* v, uint32_t taddr) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
    lo[j] = *reinterpret_cast<const uint32_t*>(&l);
  }
  tmem_st16(taddr, hi);
  tmem_st16(taddr + 16, lo);
}

// one row (128 B) of a SWIZZLE_128B [rows x 32] fp32 box whose base is 1024-byte aligned: 16-byte chunk j of row r
// lives at chunk j ^ (r & 7)  (conflict-free for thread == row)
__device__ __forceinline__ void lds_row(const uint8_t* box, int r, float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(box + r * 128 + ((j ^ (r & 7)) << 4));
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ void sts_row(uint8_t* box, int r, const float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(box + r * 128 + ((j ^ (r & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// GELU(x) = x Phi(x) (nn.GELU() default, exact erf form) with Phi(-u) = 2^-(1 + u q(u)), u = min(|x|, 8.5), q a degree-7
// fit: |x Phi(x) - exact| <= 1.2e-8 in exact arithmetic and <= 4e-7 evaluated in fp32 (= rounding of the result; erff-based
// evaluation measures 3.3e-7 on the same grid, profiles/r01b_microbench_aload.txt).  Branch-free: 9 FMA + 1 MUFU.EX2,
// about half the issue slots of the erff form, which is what bounds a GELU epilogue (same profile).
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fminf(fabsf(x), 8.5f);
  float q = 2.0539439447020413e-06f;
  q = fmaf(q, u, -3.0070181310293265e-05f);
  q = fmaf(q, u, 0.0001422710920451209f);
  q = fmaf(q, u, 0.00024190108524635434f);
  q = fmaf(q, u, -0.007198362145572901f);
  q = fmaf(q, u, 0.052587080746889114f);
  q = fmaf(q, u, 0.45917975902557373f);
  q = fmaf(q, u, 1.1511081457138062f);
  const float h = fmaf(q, u, 1.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-h));
  return x * (x > 0.f ? 1.0f - e : e);
}

}  // namespace fz

// fp32 row-major [rows, cols] (row pitch ld floats) -> 2-D tensor map with a [box_rows x 32 cols] SWIZZLE_128B box
// (box_rows = 1 for tile::gather4).  Returns 0 on success.  Host only.
int cdseg_make_tmap_f32(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
