// Serialized patch attention, third generation (sm_100a).  Replaces flash_attn varlen at
// pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:282-289, 1038-1047 (+ ":290 feat[inverse]") and, in its
// fp32-faithful mode, the dense branch at :264-280.
//
// What changed against attn_tc2.cu (profiles/r01d_ncu_attn_tc2_kernel.md: MUFU 65 %, issue slots 42 %, every softmax warp idle
// ~half of each chunk): the per-chunk dependency chain  S ready -> softmax -> P ready -> P.V -> fold O  is cut.
//   * P (shared memory) and O (tensor memory) are DOUBLE-buffered and the fold of chunk g-1's O into the register accumulator
//     is deferred until after chunk g's probabilities are written, so a softmax warp never waits for a P.V product it has just
//     requested; S(g+1) = Q.K^T is issued as soon as every row thread has copied S(g) to registers.  The warps then run
//     ld S -> max -> 64 x ex2 -> store P -> fold O(g-1) back to back and the MUFU pipe becomes the limiter it should be.
//   * the TMA producer and the MMA issuer are separate warps (a ring-slot wait no longer delays an MMA issue).
//   * MODE 1 ("tc32"): fp32-class numerics on the tensor cores.  q, k, v are split x = hi + lo (fp16 each, 22 significant bits),
//     S = q_hi.k_hi + q_lo.k_hi + q_hi.k_lo, the probabilities are split the same way and
//     O = P_hi.[v_hi | 1 | v_lo] + P_lo.[v_hi | 1]  (11 MMAs per 64-key chunk instead of 5), output stays fp32.
//     End to end this is within 2e-5 of the fp32 dense branch where the fp16 flash numerics (MODE 0) are at 3e-3.
#include "common.cuh"
#include <cstdlib>

namespace tc3 {

constexpr int NC = 64;                 // keys per chunk
constexpr int NTHREADS = 192;          // 4 softmax warps + TMA producer warp + MMA issuer warp
constexpr uint32_t WAIT_TIMEOUT_POLLS = 1u << 26;

template <int MODE> struct Cfg;
template <> struct Cfg<0> {            // fp16 operands (flash-branch numerics); P / O double-buffered, O folded one chunk late
  static constexpr bool DEFER = true;
  static constexpr int R = 4;                          // K/V ring stages
  static constexpr int KBYTES = NC * 32;               // 64 keys x 16 d fp16
  static constexpr int VW = 32;                        // V operand width: [v | 1 | 0 x 15]
  static constexpr int VBYTES = NC * VW * 2;
  static constexpr int STAGE = KBYTES + VBYTES;
  static constexpr int SQ_BYTES = 128 * 32;
  static constexpr int SP_ONE = 128 * NC * 2;          // one P buffer
  static constexpr int SP_BYTES = 2 * SP_ONE;          // double-buffered
  static constexpr int TMEM_COLS = 128;
  static constexpr int COL_S = 0, COL_O0 = 64, COL_O1 = 96;
};
#ifdef CDSEG_ATTN_TC32_DEFER           // A/B build (profiles/): mode 1 with double-buffered P / O at 2 CTAs per SM
template <> struct Cfg<1> {
  static constexpr bool DEFER = true;
  static constexpr int R = 3;
  static constexpr int KBYTES = 2 * NC * 32;
  static constexpr int VW = 48;
  static constexpr int VBYTES = NC * VW * 2;
  static constexpr int STAGE = KBYTES + VBYTES;
  static constexpr int SQ_BYTES = 2 * 128 * 32;
  static constexpr int SP_ONE = 2 * 128 * NC * 2;
  static constexpr int SP_BYTES = 2 * SP_ONE;
  static constexpr int TMEM_COLS = 256;
  static constexpr int COL_S = 0, COL_O0 = 64, COL_O1 = 128;
};
#else
template <> struct Cfg<1> {            // hi/lo split operands (fp32-class numerics); P / O single-buffered, O folded in order
  static constexpr bool DEFER = false;
  static constexpr int R = 3;
  static constexpr int KBYTES = 2 * NC * 32;           // k_hi chunk | k_lo chunk
  static constexpr int VW = 48;                        // [v_hi | 1 | 0 x 15 | v_lo]
  static constexpr int VBYTES = NC * VW * 2;
  static constexpr int STAGE = KBYTES + VBYTES;
  static constexpr int SQ_BYTES = 2 * 128 * 32;        // q_hi | q_lo
  static constexpr int SP_ONE = 2 * 128 * NC * 2;      // P_hi | P_lo
  static constexpr int SP_BYTES = SP_ONE;
  static constexpr int TMEM_COLS = 128;
  static constexpr int COL_S = 0, COL_O0 = 64, COL_O1 = 64;
};
#endif

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
  // try_wait suspends the thread for a hardware-defined interval per call; counting polls instead of reading clock64 keeps the
  // spin loop at 4 instructions (the control warps share their scheduler with softmax warps: the round-1 loop with two clock reads
  // per poll was 20 % of all issued instructions, profiles/r02_attn_tc3_f16 hot lines)
  uint32_t spins = 0;
  while (!mbar_try(bar, parity))
    if (++spins > WAIT_TIMEOUT_POLLS) __trap();               // a protocol bug must trap, never hang
}
// control warps (TMA producer, MMA issuer) wait most of the time: back off between polls so that the spin does not take issue
// slots from the softmax warps on the same scheduler
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (sleep_ns) __nanosleep(sleep_ns);
    if (++spins > WAIT_TIMEOUT_POLLS) __trap();
  }
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
// (|rel err| <= 3.1e-6, far below the 4.9e-4 fp16 rounding of P), 2^n by adding n to the exponent field.
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -120.f);
  const float t = x + 12582912.f;                     // 1.5 * 2^23: the low mantissa bits of t now hold round(x)
  const float f = x - (t - 12582912.f);
  float p = 0.0096004f;
  p = fmaf(p, f, 0.05591689f);
  p = fmaf(p, f, 0.24023718f);
  p = fmaf(p, f, 0.69312199f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__host__ __device__ constexpr uint32_t idesc(int N, bool b_mn) {                      // M128, fp16 operands, fp32 accumulate
  return (1u << 4) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

template <int R> struct Bars {
  uint64_t q, kv_full[R], kv_empty[R], s_full, s_free, p_full[2], o_full[2];
  uint32_t tmem_slot, pad;
};

// DBG != 0: what-if variants for profiling ONLY (wrong results, selected by cdseg_attn_set_debug, never by the product path):
// 1 no exponentials, 2 S read from tensor memory only for the first chunk, 3 no P stores, 4 no row max, 5 no P.V MMAs, 6 no O fold loads
template <int MODE, int POLY, int DBG = 0>      // POLY of every 8 exponentials go to the FMA pipe (0 = all on MUFU); MODE 1 always uses MUFU
#if defined(CDSEG_ATTN_TC32_DEFER)
__global__ void __launch_bounds__(NTHREADS, MODE == 0 ? 3 : 2)
#elif defined(CDSEG_ATTN_MAXNREG)
__global__ void __maxnreg__(112)   // A/B build: measured 14 % slower than the 96-register build below (profiles/r02_attention_ab.txt)
#else
__global__ void __launch_bounds__(NTHREADS, 3)   // 96 registers, 3 CTAs (18 warps) per SM in both modes
#endif
attn_tc3_kernel(const __half* __restrict__ Qp, const __half* __restrict__ Kpk, const __half* __restrict__ Vp,
                const int32_t* __restrict__ patch_len, const int32_t* __restrict__ slot_dst, int H, int T, int Kp, float sl2,
                float* __restrict__ out, int64_t out_ld, uint32_t sleep_ns) {
  using C = Cfg<MODE>;
  constexpr int R = C::R;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int qt = blockIdx.x, t = blockIdx.y, h = blockIdx.z;
  PDL_TRIGGER_EARLY();
  pdl_wait();                                                    // before the first global read (patch_len may come from the kernel before the pack)
  const int len = patch_len[t];
  if (qt * 128 >= len) return;                                   // CTA-uniform: no valid query row in this tile
  const int nc = (len + NC - 1) / NC;                            // key chunks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* sKV = smem;                                           // R stages of [K chunk(s) | V chunk]
  uint8_t* sQ = sKV + R * C::STAGE;
  uint8_t* sP = sQ + C::SQ_BYTES;
  Bars<R>* bars = reinterpret_cast<Bars<R>*>(sP + C::SP_BYTES);

  if (threadIdx.x == 128) {
    mbar_init(smem_u32(&bars->q), 1);
    for (int s = 0; s < R; ++s) {
      mbar_init(smem_u32(&bars->kv_full[s]), 1);
      mbar_init(smem_u32(&bars->kv_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->s_full), 1);
    mbar_init(smem_u32(&bars->s_free), 128);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars->p_full[b]), 128);
      mbar_init(smem_u32(&bars->o_full[b]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  const int64_t HT = (int64_t)H * T;
  const int64_t blk16 = ((int64_t)h * T + t) * Kp * 16;          // element offsets of this (head, patch)
  const int64_t blkV = ((int64_t)h * T + t) * Kp * C::VW;
  const int64_t lo_off = HT * Kp * 16;                           // MODE 1: the lo halves of q / k follow all the hi halves

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------------ TMA producer ------------------------------------
      mbar_expect_tx(smem_u32(&bars->q), C::SQ_BYTES);
      tma_load_1d(smem_u32(sQ), Qp + blk16 + (int64_t)qt * 128 * 16, 128 * 32, smem_u32(&bars->q));
      if (MODE == 1) tma_load_1d(smem_u32(sQ + 128 * 32), Qp + lo_off + blk16 + (int64_t)qt * 128 * 16, 128 * 32, smem_u32(&bars->q));
      for (int c = 0; c < nc; ++c) {
        const int s = c % R, u = c / R;
        if (u > 0) mbar_wait_idle(smem_u32(&bars->kv_empty[s]), (uint32_t)((u - 1) & 1), sleep_ns);
        uint8_t* st = sKV + s * C::STAGE;
        const uint32_t bar = smem_u32(&bars->kv_full[s]);
        mbar_expect_tx(bar, C::STAGE);
        tma_load_1d(smem_u32(st), Kpk + blk16 + (int64_t)c * NC * 16, NC * 32, bar);
        if (MODE == 1) tma_load_1d(smem_u32(st + NC * 32), Kpk + lo_off + blk16 + (int64_t)c * NC * 16, NC * 32, bar);
        tma_load_1d(smem_u32(st + C::KBYTES), Vp + blkV + (int64_t)c * NC * C::VW, C::VBYTES, bar);
      }
    }
  } else if (warp == 5) {
    {
      // ------------------------------------- MMA issuer -------------------------------------
      // whole warp, one elected lane issues (a lane-0 branch around this loop costs an R2UR + ELECT + branch sequence per tcgen05
      // instruction; see gemm_tc.cu)
      const uint32_t tmem = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);
      constexpr uint32_t IDESC_S = idesc(NC, false);             // M128 N64, A and B K-major
      constexpr uint32_t IDESC_O = idesc(C::VW, true);           // M128 N32|48, B MN-major
      constexpr uint32_t IDESC_OL = idesc(32, true);             // MODE 1: P_lo . [v_hi | 1]
      constexpr uint32_t VKG = C::VW * 16;                       // bytes between 8-key groups of the V operand
      mbar_wait_idle(smem_u32(&bars->q), 0, sleep_ns);
      const uint64_t qd = make_desc(smem_u32(sQ), 128, 256);
      const uint64_t qd_lo = make_desc(smem_u32(sQ + 128 * 32), 128, 256);
      for (int g = 0; g <= nc; ++g) {
        if (g < nc) {                                            // S(g) = Q . K_g^T
          const int s = g % R;
          mbar_wait_idle(smem_u32(&bars->kv_full[s]), (uint32_t)((g / R) & 1), sleep_ns);
          if (g >= 1) mbar_wait_idle(smem_u32(&bars->s_free), (uint32_t)((g - 1) & 1), sleep_ns);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t kb = smem_u32(sKV + s * C::STAGE);
            umma_f16(tmem + C::COL_S, qd, make_desc(kb, 128, 256), IDESC_S, 0);
            if (MODE == 1) {
              umma_f16(tmem + C::COL_S, qd_lo, make_desc(kb, 128, 256), IDESC_S, 1);
              umma_f16(tmem + C::COL_S, qd, make_desc(kb + NC * 32, 128, 256), IDESC_S, 1);
            }
            umma_commit(smem_u32(&bars->s_full));
          }
          __syncwarp();
        }
        if (g >= 1) {                                            // [O | L](g-1) = P . [V | 1]
          const int gp = g - 1, s = gp % R, b = C::DEFER ? (gp & 1) : 0;
          mbar_wait_idle(smem_u32(&bars->p_full[b]), (uint32_t)((C::DEFER ? (gp >> 1) : gp) & 1), sleep_ns);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t vb = smem_u32(sKV + s * C::STAGE + C::KBYTES);
            const uint32_t pb = smem_u32(sP + b * C::SP_ONE);
            const uint32_t od = tmem + (b ? C::COL_O1 : C::COL_O0);
#pragma unroll
            for (int kk = 0; kk < NC / 16; ++kk)
              if (DBG != 5 || gp == 0) umma_f16(od, make_desc(pb + kk * 2 * 2048, 2048, 128), make_desc(vb + kk * 2 * VKG, VKG, 128), IDESC_O, kk > 0);
            if (MODE == 1) {
#pragma unroll
              for (int kk = 0; kk < NC / 16; ++kk)
                umma_f16(od, make_desc(pb + 128 * NC * 2 + kk * 2 * 2048, 2048, 128), make_desc(vb + kk * 2 * VKG, VKG, 128), IDESC_OL, 1);
            }
            umma_commit(smem_u32(&bars->o_full[b]));
            umma_commit(smem_u32(&bars->kv_empty[s]));
          }
          __syncwarp();
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
      const int b = C::DEFER ? (gp & 1) : 0;
      mbar_wait(smem_u32(&bars->o_full[b]), (uint32_t)((C::DEFER ? (gp >> 1) : gp) & 1));
      tc_fence_after();
      const uint32_t oc = tmem + lane_base + (b ? C::COL_O1 : C::COL_O0);
      if (MODE == 0) {
        uint32_t o[17];
        if (DBG == 6 && gp > 0) {
#pragma unroll
          for (int d = 0; d < 17; ++d) o[d] = 0x3f800000u;
        } else {
          tmem_ld16(oc, o);
          tmem_ld1(oc + 16, o + 16);
          tmem_ld_wait();
        }
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d] = fmaf(acc[d], a, __uint_as_float(o[d]));
        l = fmaf(l, a, __uint_as_float(o[16]));
      } else {
        uint32_t o[16], ol[16], o1[1];
        tmem_ld16(oc, o);
        tmem_ld1(oc + 16, o1);
        tmem_ld16(oc + 32, ol);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d] = fmaf(acc[d], a, __uint_as_float(o[d]) + __uint_as_float(ol[d]));
        l = fmaf(l, a, __uint_as_float(o1[0]));
      }
    };

    for (int g = 0; g < nc; ++g) {
      // in-order variant: fold O(g-1) first (before S occupies 64 registers) -- this also waits for P.V(g-1), the last reader of the
      // single P buffer and the last writer of the single O buffer
      if (!C::DEFER && g > 0) fold(g - 1, a_prev);
      mbar_wait(smem_u32(&bars->s_full), (uint32_t)(g & 1));
      tc_fence_after();
      uint32_t s[NC];
      if (DBG == 2 && g > 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j) s[j] = __float_as_uint((float)(j + g) * 0.01f);
      } else {
        tmem_ld32(tmem + lane_base + C::COL_S, s);
        tmem_ld32(tmem + lane_base + C::COL_S + 32, s + 32);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->s_free));                      // S lives in registers now: the next QK^T may start
      const int valid = len - g * NC;                            // keys >= valid are padding
      float mx = -INFINITY;
      if (valid < NC) {
#pragma unroll
        for (int j = 0; j < NC; ++j)
          if (j >= valid) s[j] = 0xff800000u;                    // -inf
      }
      if (DBG == 4) mx = __uint_as_float(s[0]);
      else {                                                     // four independent chains: 8 dependent FMNMX3 instead of 32
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int j = 0; j < NC; j += 8) {
          m0 = max3(m0, __uint_as_float(s[j]), __uint_as_float(s[j + 1]));
          m1 = max3(m1, __uint_as_float(s[j + 2]), __uint_as_float(s[j + 3]));
          m2 = max3(m2, __uint_as_float(s[j + 4]), __uint_as_float(s[j + 5]));
          m3 = max3(m3, __uint_as_float(s[j + 6]), __uint_as_float(s[j + 7]));
        }
        mx = fmaxf(max3(m0, m1, m2), m3);
      }
      const float m_new = fmaxf(m, mx);
      const float msc = m_new * sl2;
      const float a_g = ex2(m * sl2 - msc);                      // first chunk: m = -inf -> 0
      m = m_new;
      // deferred variant: P buffer g&1 was last read by P.V(g-2); fold(g-2) (previous iteration) waited for that product
      uint8_t* pw = prow + (C::DEFER ? (g & 1) : 0) * C::SP_ONE;
#pragma unroll
      for (int kg = 0; kg < NC / 8; ++kg) {
        uint32_t pk[4], pl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x0 = fmaf(__uint_as_float(s[kg * 8 + 2 * j]), sl2, -msc), x1 = fmaf(__uint_as_float(s[kg * 8 + 2 * j + 1]), sl2, -msc);
          const float p0 = DBG == 1 ? x0 : ((MODE == 0 && 2 * j < POLY) ? ex2_fma(x0) : ex2(x0));
          const float p1 = DBG == 1 ? x1 : ((MODE == 0 && 2 * j + 1 < POLY) ? ex2_fma(x1) : ex2(x1));
          __half2 hh = __floats2half2_rn(p0, p1);
          pk[j] = *reinterpret_cast<uint32_t*>(&hh);
          if (MODE == 1) {
            const float2 back = __half22float2(hh);
            __half2 ll = __floats2half2_rn(p0 - back.x, p1 - back.y);
            pl[j] = *reinterpret_cast<uint32_t*>(&ll);
          }
        }
        if (DBG != 3) *reinterpret_cast<uint4*>(pw + kg * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        else if (kg == 0) *reinterpret_cast<uint4*>(pw) = make_uint4(pk[0] ^ pk[1], pk[2] ^ pk[3], 0, 0);
        if (MODE == 1) *reinterpret_cast<uint4*>(pw + 128 * NC * 2 + kg * 2048) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
      fence_async_smem();                                        // generic-proxy writes -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->p_full[C::DEFER ? (g & 1) : 0]));
      if (C::DEFER && g > 0) fold(g - 1, a_prev);                // P.V(g-1) was requested a whole chunk ago: no stall in steady state
      a_prev = a_g;
    }
    PDL_TRIGGER_LATE();                                          // last fold + output row left
    fold(nc - 1, a_prev);
    const int32_t dst = slot_dst[(int64_t)t * Kp + qt * 128 + r];
    if (dst >= 0) {
      const float inv = 1.f / l;
      float4* op = reinterpret_cast<float4*>(out + (int64_t)dst * out_ld + h * 16);
      if (MODE == 0) {
        // flash_attn returns fp16 and the reference widens it again (ptv3.py:289): same rounding point here
#pragma unroll
        for (int j = 0; j < 4; ++j)
          op[j] = make_float4(__half2float(__float2half_rn(acc[4 * j] * inv)), __half2float(__float2half_rn(acc[4 * j + 1] * inv)),
                              __half2float(__float2half_rn(acc[4 * j + 2] * inv)), __half2float(__float2half_rn(acc[4 * j + 3] * inv)));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) op[j] = make_float4(acc[4 * j] * inv, acc[4 * j + 1] * inv, acc[4 * j + 2] * inv, acc[4 * j + 3] * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

template <int MODE> constexpr size_t smem_bytes() {
  return (size_t)Cfg<MODE>::R * Cfg<MODE>::STAGE + Cfg<MODE>::SQ_BYTES + Cfg<MODE>::SP_BYTES + sizeof(Bars<Cfg<MODE>::R>) + 1024;
}

}  // namespace tc3

// exponentials per group of 8 computed on the FMA pipe instead of MUFU (0..3), MODE 0 only; env CDSEG_ATTN_POLY or cdseg_attn_set_poly
static int g_attn3_debug = 0;
static uint32_t g_attn3_sleep = [] { const char* e = getenv("CDSEG_ATTN_SLEEP"); return e ? (uint32_t)atoi(e) : 0u; }();   // ns between polls of the control warps
CDSEG_API void cdseg_attn_set_debug(int variant) { g_attn3_debug = variant; }
int g_cdseg_attn_poly = [] { const char* e = getenv("CDSEG_ATTN_POLY"); return e ? atoi(e) : 0; }();
CDSEG_API void cdseg_attn_set_poly(int per8) { g_cdseg_attn_poly = per8 < 0 ? 0 : (per8 > 3 ? 3 : per8); }

// mode 0: Q, K fp16 packed [H][T][Kp][16], V fp16 packed 32 wide with the ones column (cdseg_attn_pack_f16v, v_ones = 1); fp16
//         probabilities, fp16-rounded output (flash-branch numerics).
// mode 1: operands from cdseg_attn_pack_split (q / k as hi | lo halves, V 48 wide [v_hi | 1 | v_lo]); fp32-class results.
// out: fp32 [n, out_ld]; head h -> columns h*16 .. h*16+15 of row slot_dst[slot].
CDSEG_API int cdseg_attn_tc3(const void* Q, const void* K, const void* V, const int32_t* patch_len, const int32_t* slot_dst, int H,
                             int T, int Kp, float scale, int mode, float* out, int64_t out_ld, void* stream) {
  if (H <= 0 || T < 0 || (Kp % 128) || (out_ld & 3) || (mode != 0 && mode != 1)) return CDSEG_EINVAL;
  if (T == 0) return CDSEG_OK;
  static bool init = false;
  if (!init) {
    cudaError_t e;
#define CDSEG_SET(MODE, POLY)                                                                                              \
  e = cudaFuncSetAttribute(tc3::attn_tc3_kernel<MODE, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc3::smem_bytes<MODE>()); \
  if (e != cudaSuccess) return (int)e;
    CDSEG_SET(0, 0) CDSEG_SET(0, 1) CDSEG_SET(0, 2) CDSEG_SET(0, 3) CDSEG_SET(1, 0)
#undef CDSEG_SET
#define CDSEG_SETD(D)                                                                                                      \
  e = cudaFuncSetAttribute(tc3::attn_tc3_kernel<0, 0, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc3::smem_bytes<0>()); \
  if (e != cudaSuccess) return (int)e;
    CDSEG_SETD(1) CDSEG_SETD(2) CDSEG_SETD(3) CDSEG_SETD(4) CDSEG_SETD(5) CDSEG_SETD(6)
#undef CDSEG_SETD
    init = true;
  }
  dim3 g(Kp / 128, T, H);
  const float sl2 = scale * 1.4426950408889634f;
#define CDSEG_ATTN_LAUNCH(MODE, POLY)                                                                                      \
  cdseg_launch_pdl(tc3::attn_tc3_kernel<MODE, POLY, 0>, g, dim3(tc3::NTHREADS), tc3::smem_bytes<MODE>(), (cudaStream_t)stream, \
      (const __half*)Q, (const __half*)K, (const __half*)V, patch_len, slot_dst, H, T, Kp, sl2, out, (int64_t)out_ld, (uint32_t)g_attn3_sleep)
#define CDSEG_ATTN_LAUNCH_D(D)                                                                                             \
  tc3::attn_tc3_kernel<0, 0, D><<<g, tc3::NTHREADS, tc3::smem_bytes<0>(), (cudaStream_t)stream>>>(                         \
      (const __half*)Q, (const __half*)K, (const __half*)V, patch_len, slot_dst, H, T, Kp, sl2, out, out_ld, g_attn3_sleep)
  if (mode == 0 && g_attn3_debug) {
    switch (g_attn3_debug) {
      case 1: CDSEG_ATTN_LAUNCH_D(1); break;
      case 2: CDSEG_ATTN_LAUNCH_D(2); break;
      case 3: CDSEG_ATTN_LAUNCH_D(3); break;
      case 4: CDSEG_ATTN_LAUNCH_D(4); break;
      case 5: CDSEG_ATTN_LAUNCH_D(5); break;
      default: CDSEG_ATTN_LAUNCH_D(6); break;
    }
  } else
  if (mode == 1) CDSEG_ATTN_LAUNCH(1, 0);
  else if (g_cdseg_attn_poly == 1) CDSEG_ATTN_LAUNCH(0, 1);
  else if (g_cdseg_attn_poly == 2) CDSEG_ATTN_LAUNCH(0, 2);
  else if (g_cdseg_attn_poly == 3) CDSEG_ATTN_LAUNCH(0, 3);
  else CDSEG_ATTN_LAUNCH(0, 0);
#undef CDSEG_ATTN_LAUNCH
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
