// Serialized patch attention, second generation (sm_100a): same math and numerics contract as attn_tc.cu
// (fp16 q/k/v, fp32 accumulation, fp16 probabilities, fp16-rounded output; replaces flash_attn varlen at
// pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:282-289, 1038-1047 + ":290 feat[inverse]"),
// re-organised around what the round-1 profiles showed (profiles/r01_ncu_attn_tc_stage0.md: MUFU 50 %, 9 warps/SM):
//   * ONE 128-row q tile per CTA and K / V streamed in 64-key chunks through a 4-stage TMA ring instead of a
//     whole-patch K/V image: 35 KB of shared memory per CTA instead of 105 KB -> 3 CTAs (12 softmax warps) per SM;
//   * V is packed 32 wide, [v | 1 | 0..], so ONE tcgen05.mma per 16 keys yields P.V and the row sum P.1
//     (5 MMAs per chunk instead of 9: a thread issues only one tcgen05.mma every ~50 cycles,
//     profiles/r01_microbench_mma_issue_latency.txt);
//   * S, P and O are single-buffered: S is copied to registers at once (so the next QK^T can start), the previous
//     chunk's O is folded into the register accumulator before P is overwritten -> TMEM 128 columns per CTA.
#include "common.cuh"
#include <cstdlib>

namespace tc2 {

constexpr int NC = 64;                 // keys per chunk
constexpr int NTHREADS = 160;          // 4 softmax warps + 1 control warp
constexpr int TMEM_COLS = 128;
constexpr int COL_S = 0;               // S : 64 fp32 columns
constexpr int COL_O = 64;              // O : 32 columns (0..15 = P.V, 16 = P.1)
constexpr int R = 4;                   // K/V ring stages
constexpr int KBYTES = NC * 32;        // K chunk: 64 keys x 16 d fp16
constexpr int VBYTES = NC * 64;        // V chunk: 64 keys x 32 (v | 1 | 0) fp16
constexpr int STAGE = KBYTES + VBYTES;
constexpr int SQ_BYTES = 128 * 32;
constexpr int SP_BYTES = 128 * NC * 2;
constexpr long long WAIT_TIMEOUT_CYCLES = 4000000000ll;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(done)
               : "r"(bar), "r"(parity)
               : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) __trap();       // a protocol bug must trap, never hang
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) {   // one FMNMX3 instead of two FMNMX (sm_100+)
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a degree-4 polynomial
// (|rel err| <= 3.1e-6, far below the 4.9e-4 fp16 rounding of P), 2^n by adding n to the exponent field.  The softmax is bound by the 16
// MUFU.EX2 per clock and SM (profiles/r01d_ncu_attn_tc2_kernel.md: XU pipe 65 %, issue slots 42 %), so moving a fraction of the
// exponentials onto the idle FMA slots raises the ceiling (the trick FlashAttention-4 uses on the same hardware).
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -120.f);
  const float t = x + 12582912.f;                     // 1.5 * 2^23: the low mantissa bits of t now hold round(x)
  const float f = x - (t - 12582912.f);
  float p = 0.0096004f;                               // relative-error least-squares fit of 2^f on Chebyshev nodes: max 3.1e-6 in fp32
  p = fmaf(p, f, 0.05591689f);
  p = fmaf(p, f, 0.24023718f);
  p = fmaf(p, f, 0.69312199f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

constexpr uint32_t IDESC_S = (1u << 4) | ((uint32_t)(NC >> 3) << 17) | ((128u >> 4) << 24);              // M128 N64, A,B K-major
constexpr uint32_t IDESC_O = (1u << 4) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);          // M128 N32, B MN-major

struct Bars {
  uint64_t q, kv_full[R], kv_empty[R], s_full, s_free, p_full, o_full;
  uint32_t tmem_slot, pad;
};

template <int OCC, int POLY>       // POLY of every 8 exponentials go to the FMA pipe (0 = all on MUFU)
__global__ void __launch_bounds__(NTHREADS, OCC)
attn_tc2_kernel(const __half* __restrict__ Qp, const __half* __restrict__ Kpk, const __half* __restrict__ Vp,
                const int32_t* __restrict__ patch_len, const int32_t* __restrict__ slot_dst, int H, int T, int Kp, float sl2,
                float* __restrict__ out, int64_t out_ld) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int qt = blockIdx.x, t = blockIdx.y, h = blockIdx.z;
  const int len = patch_len[t];
  if (qt * 128 >= len) return;                                   // CTA-uniform: no valid query row in this tile
  const int nc = (len + NC - 1) / NC;                            // key chunks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* sKV = smem;                                           // R stages of [K chunk | V chunk]
  uint8_t* sQ = sKV + R * STAGE;
  uint8_t* sP = sQ + SQ_BYTES;
  Bars* bars = reinterpret_cast<Bars*>(sP + SP_BYTES);

  if (threadIdx.x == 128) {
    mbar_init(smem_u32(&bars->q), 1);
    for (int s = 0; s < R; ++s) {
      mbar_init(smem_u32(&bars->kv_full[s]), 1);
      mbar_init(smem_u32(&bars->kv_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->s_full), 1);
    mbar_init(smem_u32(&bars->s_free), 128);
    mbar_init(smem_u32(&bars->p_full), 128);
    mbar_init(smem_u32(&bars->o_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  const int64_t blk16 = ((int64_t)h * T + t) * Kp * 16;          // element offsets of this (head, patch)
  const int64_t blk32 = ((int64_t)h * T + t) * Kp * 32;

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------ control lane: TMA + MMA issue ------------------------------
      mbar_expect_tx(smem_u32(&bars->q), SQ_BYTES);
      tma_load_1d(smem_u32(sQ), Qp + blk16 + (int64_t)qt * 128 * 16, SQ_BYTES, smem_u32(&bars->q));
      auto load_kv = [&](int c) {
        const int s = c % R, u = c / R;
        if (u > 0) mbar_wait(smem_u32(&bars->kv_empty[s]), (uint32_t)((u - 1) & 1));
        uint8_t* st = sKV + s * STAGE;
        mbar_expect_tx(smem_u32(&bars->kv_full[s]), STAGE);
        tma_load_1d(smem_u32(st), Kpk + blk16 + (int64_t)c * NC * 16, KBYTES, smem_u32(&bars->kv_full[s]));
        tma_load_1d(smem_u32(st + KBYTES), Vp + blk32 + (int64_t)c * NC * 32, VBYTES, smem_u32(&bars->kv_full[s]));
      };
      constexpr int D = R - 2;                                   // chunks requested ahead (see ring-reuse argument in DESIGN.md)
      for (int c = 0; c < min(D, nc); ++c) load_kv(c);
      mbar_wait(smem_u32(&bars->q), 0);
      const uint64_t qd = make_desc(smem_u32(sQ), 128, 256);
      for (int g = 0; g <= nc; ++g) {
        if (g + D < nc) load_kv(g + D);
        if (g < nc) {                                            // MMA1(g): S = Q . K_g^T
          const int s = g % R;
          mbar_wait(smem_u32(&bars->kv_full[s]), (uint32_t)((g / R) & 1));
          if (g >= 1) mbar_wait(smem_u32(&bars->s_free), (uint32_t)((g - 1) & 1));
          tc_fence_after();
          umma_f16(tmem + COL_S, qd, make_desc(smem_u32(sKV + s * STAGE), 128, 256), IDESC_S, 0);
          umma_commit(smem_u32(&bars->s_full));
        }
        if (g >= 1) {                                            // MMA2(g-1): [O | L] = P . [V | 1]
          const int gp = g - 1, s = gp % R;
          mbar_wait(smem_u32(&bars->p_full), (uint32_t)(gp & 1));
          tc_fence_after();
          const uint32_t vb = smem_u32(sKV + s * STAGE + KBYTES);
#pragma unroll
          for (int kk = 0; kk < NC / 16; ++kk)
            umma_f16(tmem + COL_O, make_desc(smem_u32(sP + kk * 2 * 2048), 2048, 128), make_desc(vb + kk * 2 * 512, 512, 128),
                     IDESC_O, kk > 0);
          umma_commit(smem_u32(&bars->o_full));
          umma_commit(smem_u32(&bars->kv_empty[s]));
        }
      }
    }
  } else {
    // ------------------------------------- softmax threads -------------------------------------
    const int r = threadIdx.x;                                   // row of the q tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float acc[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[d] = 0.f;
    float l = 0.f, m = -INFINITY, a_prev = 0.f;
    uint8_t* prow = sP + (uint32_t)((r >> 3) * 128 + (r & 7) * 16);

    auto fold = [&](int gp, float a) {                           // acc = acc * a + O_gp ; l likewise
      mbar_wait(smem_u32(&bars->o_full), (uint32_t)(gp & 1));
      tc_fence_after();
      uint32_t o[17];
      tmem_ld16(tmem + lane_base + COL_O, o);
      tmem_ld1(tmem + lane_base + COL_O + 16, o + 16);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 16; ++d) acc[d] = fmaf(acc[d], a, __uint_as_float(o[d]));
      l = fmaf(l, a, __uint_as_float(o[16]));
    };

    for (int g = 0; g < nc; ++g) {
      mbar_wait(smem_u32(&bars->s_full), (uint32_t)(g & 1));
      tc_fence_after();
      uint32_t s[NC];
      tmem_ld32(tmem + lane_base + COL_S, s);
      tmem_ld32(tmem + lane_base + COL_S + 32, s + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->s_free));                      // S lives in registers now: the next QK^T may start
      const int valid = len - g * NC;                            // keys >= valid are padding
      float mx = -INFINITY;
      if (valid < NC) {
#pragma unroll
        for (int j = 0; j < NC; ++j)
          if (j >= valid) s[j] = 0xff800000u;                    // -inf
      }
#pragma unroll
      for (int j = 0; j < NC; j += 2) mx = max3(mx, __uint_as_float(s[j]), __uint_as_float(s[j + 1]));
      const float m_new = fmaxf(m, mx);
      const float msc = m_new * sl2;
      const float a_g = ex2(m * sl2 - msc);                      // first chunk: m = -inf -> 0
      m = m_new;
      if (g > 0) fold(g - 1, a_prev);                            // also guarantees MMA2(g-1) is done reading P
      a_prev = a_g;
#pragma unroll
      for (int kg = 0; kg < NC / 8; ++kg) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x0 = fmaf(__uint_as_float(s[kg * 8 + 2 * j]), sl2, -msc), x1 = fmaf(__uint_as_float(s[kg * 8 + 2 * j + 1]), sl2, -msc);
          const float p0 = (2 * j < POLY) ? ex2_fma(x0) : ex2(x0);
          const float p1 = (2 * j + 1 < POLY) ? ex2_fma(x1) : ex2(x1);
          __half2 hh = __floats2half2_rn(p0, p1);
          pk[j] = *reinterpret_cast<uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(prow + kg * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_async_smem();                                        // generic-proxy writes -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->p_full));
    }
    fold(nc - 1, a_prev);
    const int32_t dst = slot_dst[(int64_t)t * Kp + qt * 128 + r];
    if (dst >= 0) {
      const float inv = 1.f / l;
      float4* op = reinterpret_cast<float4*>(out + (int64_t)dst * out_ld + h * 16);
      // flash_attn returns fp16 and the reference widens it again (ptv3.py:289): same rounding point here
#pragma unroll
      for (int j = 0; j < 4; ++j)
        op[j] = make_float4(__half2float(__float2half_rn(acc[4 * j] * inv)), __half2float(__float2half_rn(acc[4 * j + 1] * inv)),
                            __half2float(__float2half_rn(acc[4 * j + 2] * inv)), __half2float(__float2half_rn(acc[4 * j + 3] * inv)));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

}  // namespace tc2

// Q, K: fp16 packed [H][T][Kp][16]; V: fp16 packed 32 wide [H][T][Kp][32] with the ones column (cdseg_attn_pack_f16v, v_ones=1).
// out: fp32 [n, out_ld]; head h -> columns h*16 .. h*16+15 of row slot_dst[slot].
// exponentials per group of 8 computed on the FMA pipe instead of MUFU (0..3); env CDSEG_ATTN_POLY or cdseg_attn_set_poly
extern int g_cdseg_attn_poly;      // attn_tc3.cu (cdseg_attn_set_poly)
#define g_attn_poly g_cdseg_attn_poly

CDSEG_API int cdseg_attn_tc2(const void* Q, const void* K, const void* V32, const int32_t* patch_len,
                             const int32_t* slot_dst, int H, int T, int Kp, float scale, float* out, int64_t out_ld,
                             void* stream) {
  if (H <= 0 || T < 0 || (Kp % 128) || (out_ld & 3)) return CDSEG_EINVAL;
  if (T == 0) return CDSEG_OK;
  const size_t smem = (size_t)tc2::R * tc2::STAGE + tc2::SQ_BYTES + tc2::SP_BYTES + sizeof(tc2::Bars) + 1024;
  static int occ = 0;
  if (!occ) {
    const char* e = getenv("CDSEG_ATTN_OCC");                    // 3 (113 regs, no spill) or 4 (96 regs, 84 B spill) CTAs per SM
    occ = (e && e[0] == '4') ? 4 : 3;                            // measured equal (100.4 vs 102.4 us at stage 0): default 3
    cudaError_t e1 = cudaFuncSetAttribute(tc2::attn_tc2_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e2 = cudaFuncSetAttribute(tc2::attn_tc2_kernel<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e3 = cudaFuncSetAttribute(tc2::attn_tc2_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e4 = cudaFuncSetAttribute(tc2::attn_tc2_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e5 = cudaFuncSetAttribute(tc2::attn_tc2_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 != cudaSuccess) return (int)e1;
    if (e2 != cudaSuccess) return (int)e2;
    if (e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess) return (int)(e3 != cudaSuccess ? e3 : e4 != cudaSuccess ? e4 : e5);
  }
  dim3 g(Kp / 128, T, H);
  const float sl2 = scale * 1.4426950408889634f;
#define CDSEG_ATTN_LAUNCH(O, P)                                                                                              \
  tc2::attn_tc2_kernel<O, P><<<g, tc2::NTHREADS, smem, (cudaStream_t)stream>>>((const __half*)Q, (const __half*)K, (const __half*)V32, \
                                                                               patch_len, slot_dst, H, T, Kp, sl2, out, out_ld)
  if (occ == 4) CDSEG_ATTN_LAUNCH(4, 0);
  else if (g_attn_poly == 1) CDSEG_ATTN_LAUNCH(3, 1);
  else if (g_attn_poly == 2) CDSEG_ATTN_LAUNCH(3, 2);
  else if (g_attn_poly == 3) CDSEG_ATTN_LAUNCH(3, 3);
  else CDSEG_ATTN_LAUNCH(3, 0);
#undef CDSEG_ATTN_LAUNCH
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
