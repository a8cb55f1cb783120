// Serialized patch attention on the 5th-gen tensor cores (sm_100a): per (patch, head)
// S = Q K^T and O = P V run as tcgen05.mma with accumulators in TMEM; operands are staged
// in shared memory by 1-D TMA bulk copies straight from the pre-tiled ("core matrix")
// packed buffers written by cdseg_attn_pack_f16; softmax runs in registers, one thread
// per query row (TMEM lane == row), with the row sum produced by the tensor core as well
// (P . 1).  Output rows are scattered to the original point order (fuses "feat[inverse]").
//
// Replaces flash_attn.flash_attn_varlen_{qkv,kv}packed_func at
//   pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:282-289, 1038-1047
// with the same numerics contract: fp16 q/k/v, fp32 accumulation, fp16 probabilities.
// head_dim is 16 at every stage of every shipped config (configs/scannet/CDSegNet.py:66-82).
//
// Pipeline (one CTA = 4 softmax warps + 1 control warp; 2 CTAs per SM; TMEM 256 columns):
//   control lane: TMA(K,V) ; TMA(Q_i) ; for chunk g: MMA1(g): S[g&1] = Q_i K_c^T  (N = 64 keys)
//                                                    MMA2(g-1): O[(g-1)&1] = P_{g-1} [V_c | 1]
//   softmax thr : wait S_g -> regs ; free S ; m,alpha ; P_g = exp2(.) -> fp16 smem (UMMA layout)
//                 signal P ; then fold O_{g-1} (TMEM) into the fp32 register accumulator
// All hand-offs are mbarriers; no __syncthreads in the main loop.
#include "common.cuh"

namespace tc {

constexpr int NC = 64;                 // keys per chunk
constexpr int NTHREADS = 160;
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0;               // S[b] : COL_S + 64 b
constexpr int COL_O = 128;             // O[b] : COL_O + 32 b   (0..15 = P.V, 16..31 = P.1)
constexpr int SQ_BYTES = 128 * 32;     // one Q tile
constexpr int SP_BYTES = 128 * NC * 2; // one P tile
constexpr long long WAIT_TIMEOUT_CYCLES = 4000000000ll;   // ~2 s: a protocol bug must trap, never hang

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!done && clock64() - t0 > WAIT_TIMEOUT_CYCLES) __trap();
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

// shared-memory matrix descriptor, no swizzle, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// D[tmem] (+)= A[smem] . B[smem], fp16 inputs, fp32 accumulate
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// instruction descriptors (kind::f16): D fp32, A/B fp16, M = 128
constexpr uint32_t IDESC_S = (1u << 4) | ((uint32_t)(NC >> 3) << 17) | ((128u >> 4) << 24);              // N=64, A,B K-major
constexpr uint32_t IDESC_O = (1u << 4) | (1u << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24);          // N=16, B MN-major

struct Bars {
  uint64_t kv, q[2], s[2], sfree[2], p[2], o[2];
  uint32_t tmem_slot, pad;
};

__global__ void __launch_bounds__(NTHREADS, 2)
attn_tc_kernel(const __half* __restrict__ Qp, const __half* __restrict__ Kpk, const __half* __restrict__ Vp,
               const int32_t* __restrict__ patch_len, const int32_t* __restrict__ slot_dst, int H, int T, int Kp,
               int tiles_per_cta, float sl2, float* __restrict__ out, int64_t out_ld) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int t = blockIdx.y, h = blockIdx.z;
  const int len = patch_len[t];
  const int nq_valid = (len + 127) >> 7;
  const int q0 = blockIdx.x * tiles_per_cta;
  if (q0 >= nq_valid) return;                                   // CTA-uniform
  const int nqt = min(tiles_per_cta, nq_valid - q0);
  const int nc = (len + NC - 1) / NC;                           // key chunks
  const int G = nqt * nc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* sK = smem;
  uint8_t* sV = sK + (size_t)Kp * 32;
  uint8_t* sQ = sV + (size_t)Kp * 32;
  uint8_t* sP = sQ + 2 * SQ_BYTES;
  uint8_t* sOnes = sP + 2 * SP_BYTES;
  Bars* bars = reinterpret_cast<Bars*>(sOnes + 512);

  if (threadIdx.x == 128) {
    mbar_init(smem_u32(&bars->kv), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars->q[b]), 1);
      mbar_init(smem_u32(&bars->s[b]), 1);
      mbar_init(smem_u32(&bars->sfree[b]), 128);
      mbar_init(smem_u32(&bars->p[b]), 128);
      mbar_init(smem_u32(&bars->o[b]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 128) {                                      // B operand of the row-sum MMA: 16x16 ones
    reinterpret_cast<uint32_t*>(sOnes)[threadIdx.x] = 0x3C003C00u;
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  const int64_t blk = ((int64_t)h * T + t) * Kp * 16;           // element offset of this (head, patch)

  if (warp == 4) {
    if (lane == 0) {
      // ------------------------------ control lane ------------------------------
      const uint32_t kvbytes = (uint32_t)nc * NC * 32;
      mbar_expect_tx(smem_u32(&bars->kv), 2 * kvbytes);
      tma_load_1d(smem_u32(sK), Kpk + blk, kvbytes, smem_u32(&bars->kv));
      tma_load_1d(smem_u32(sV), Vp + blk, kvbytes, smem_u32(&bars->kv));
      for (int i = 0; i < min(2, nqt); ++i) {
        mbar_expect_tx(smem_u32(&bars->q[i]), SQ_BYTES);
        tma_load_1d(smem_u32(sQ + i * SQ_BYTES), Qp + blk + (int64_t)(q0 + i) * 128 * 16, SQ_BYTES, smem_u32(&bars->q[i]));
      }
      mbar_wait(smem_u32(&bars->kv), 0);
      for (int g = 0; g <= G; ++g) {
        if (g < G) {                                            // MMA1(g): S[g&1] = Q_i . K_c^T
          const int i = g / nc, c = g - i * nc, sb = g & 1;
          if (g >= 2) mbar_wait(smem_u32(&bars->sfree[sb]), (uint32_t)(((g >> 1) - 1) & 1));
          if (c == 0) mbar_wait(smem_u32(&bars->q[i & 1]), (uint32_t)((i >> 1) & 1));
          tc_fence_after();
          const uint64_t ad = make_desc(smem_u32(sQ + (i & 1) * SQ_BYTES), 128, 256);
          const uint64_t bd = make_desc(smem_u32(sK + (size_t)c * (NC / 8) * 256), 128, 256);
          umma_f16(tmem + COL_S + sb * NC, ad, bd, IDESC_S, 0);
          umma_commit(smem_u32(&bars->s[sb]));
        }
        if (g >= 1) {                                           // MMA2(g-1): O = P . V , L = P . 1
          const int gp = g - 1, ip = gp / nc, cp = gp - ip * nc, pb = gp & 1;
          mbar_wait(smem_u32(&bars->p[pb]), (uint32_t)((gp >> 1) & 1));
          tc_fence_after();
          if (cp == nc - 1 && ip + 2 < nqt) {                   // Q buffer of tile ip is free: prefetch tile ip+2
            mbar_expect_tx(smem_u32(&bars->q[ip & 1]), SQ_BYTES);
            tma_load_1d(smem_u32(sQ + (ip & 1) * SQ_BYTES), Qp + blk + (int64_t)(q0 + ip + 2) * 128 * 16, SQ_BYTES,
                        smem_u32(&bars->q[ip & 1]));
          }
          const uint32_t d_o = tmem + COL_O + pb * 32;
          const uint64_t ones = make_desc(smem_u32(sOnes), 256, 128);
#pragma unroll
          for (int kk = 0; kk < NC / 16; ++kk) {
            const uint64_t ad = make_desc(smem_u32(sP + pb * SP_BYTES + kk * 2 * 2048), 2048, 128);
            const uint64_t bd = make_desc(smem_u32(sV + ((size_t)cp * (NC / 8) + kk * 2) * 256), 256, 128);
            umma_f16(d_o, ad, bd, IDESC_O, kk > 0);
            umma_f16(d_o + 16, ad, ones, IDESC_O, kk > 0);
          }
          umma_commit(smem_u32(&bars->o[pb]));
        }
      }
    }
  } else {
    // -------------------------------- softmax threads --------------------------------
    const int r = threadIdx.x;                                   // row inside the 128-row q tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float acc[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[d] = 0.f;
    float l = 0.f, m = -INFINITY, a_prev = 0.f;
    const uint32_t p_row = (uint32_t)((r >> 3) * 128 + (r & 7) * 16);

    auto fold = [&](int gp, float a) {                           // acc = acc * a + O_gp ; write out at tile end
      const int pb = gp & 1;
      mbar_wait(smem_u32(&bars->o[pb]), (uint32_t)((gp >> 1) & 1));
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(tmem + lane_base + COL_O + pb * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 16; ++d) acc[d] = fmaf(acc[d], a, __uint_as_float(o[d]));
      l = fmaf(l, a, __uint_as_float(o[16]));
      const int ip = gp / nc;
      if (gp - ip * nc == nc - 1) {                              // last chunk of q tile ip
        const int32_t dst = slot_dst[(int64_t)t * Kp + (q0 + ip) * 128 + r];
        if (dst >= 0) {
          const float inv = 1.f / l;
          float4* op = reinterpret_cast<float4*>(out + (int64_t)dst * out_ld + h * 16);
          // flash_attn returns fp16 and the reference widens it again (ptv3.py:289): same rounding point here
#pragma unroll
          for (int j = 0; j < 4; ++j)
            op[j] = make_float4(__half2float(__float2half_rn(acc[4 * j] * inv)), __half2float(__float2half_rn(acc[4 * j + 1] * inv)),
                                __half2float(__float2half_rn(acc[4 * j + 2] * inv)), __half2float(__float2half_rn(acc[4 * j + 3] * inv)));
        }
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d] = 0.f;
        l = 0.f;
      }
    };

    for (int g = 0; g < G; ++g) {
      const int i = g / nc, c = g - i * nc, sb = g & 1;
      if (c == 0) m = -INFINITY;
      mbar_wait(smem_u32(&bars->s[sb]), (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      uint32_t s[NC];
      tmem_ld32(tmem + lane_base + COL_S + sb * NC, s);
      tmem_ld32(tmem + lane_base + COL_S + sb * NC + 32, s + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->sfree[sb]));
      const int valid = len - c * NC;                            // keys >= valid are padding
      float mx = -INFINITY;
      if (valid >= NC) {
#pragma unroll
        for (int j = 0; j < NC; ++j) mx = fmaxf(mx, __uint_as_float(s[j]));
      } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          if (j >= valid) s[j] = 0xff800000u;                    // -inf
          mx = fmaxf(mx, __uint_as_float(s[j]));
        }
      }
      const float m_new = fmaxf(m, mx);
      const float msc = m_new * sl2;
      const float a_g = ex2(m * sl2 - msc);                      // first chunk: m = -inf -> 0
      m = m_new;
      uint8_t* prow = sP + sb * SP_BYTES + p_row;
#pragma unroll
      for (int kg = 0; kg < NC / 8; ++kg) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p0 = ex2(fmaf(__uint_as_float(s[kg * 8 + 2 * j]), sl2, -msc));
          const float p1 = ex2(fmaf(__uint_as_float(s[kg * 8 + 2 * j + 1]), sl2, -msc));
          __half2 hh = __floats2half2_rn(p0, p1);
          pk[j] = *reinterpret_cast<uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(prow + kg * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_async_smem();                                        // generic-proxy writes -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->p[sb]));
      if (g > 0) fold(g - 1, a_prev);
      a_prev = a_g;
    }
    fold(G - 1, a_prev);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

}  // namespace tc

CDSEG_API size_t cdseg_attn_tc_smem_bytes(int Kp) {
  return (size_t)Kp * 64 + 2 * tc::SQ_BYTES + 2 * tc::SP_BYTES + 512 + sizeof(tc::Bars) + 1024;
}

// Q,K,V: fp16 packed [H][T][Kp][16] (cdseg_attn_pack_f16).  out: fp32 [n, out_ld]; head h -> columns h*16 .. h*16+15
// of row slot_dst[slot].  scale = softmax scale (head_dim^-0.5).
CDSEG_API int cdseg_attn_tc(const void* Q, const void* K, const void* V, const int32_t* patch_len,
                            const int32_t* slot_dst, int H, int T, int Kp, float scale, float* out, int64_t out_ld,
                            void* stream) {
  if (H <= 0 || T < 0 || (Kp % 128) || Kp > 1024 || (out_ld & 3)) return CDSEG_EINVAL;
  if (T == 0) return CDSEG_OK;
  const size_t smem = cdseg_attn_tc_smem_bytes(Kp);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(tc::attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  const int nq = Kp / 128;
  // split the q tiles of a (patch, head) over several CTAs until the grid covers ~2 waves of 2 CTAs/SM
  int tiles_per_cta = nq;
  while (tiles_per_cta > 1 && (int64_t)T * H * ((nq + tiles_per_cta - 1) / tiles_per_cta) < 148 * 2 * 2) tiles_per_cta >>= 1;
  dim3 g((nq + tiles_per_cta - 1) / tiles_per_cta, T, H);
  const float sl2 = scale * 1.4426950408889634f;
  tc::attn_tc_kernel<<<g, tc::NTHREADS, smem, (cudaStream_t)stream>>>((const __half*)Q, (const __half*)K, (const __half*)V,
                                                                     patch_len, slot_dst, H, T, Kp, tiles_per_cta, sl2,
                                                                     out, out_ld);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
