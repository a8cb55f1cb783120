// Post-attention half of a PTv3 Block as ONE kernel (C = 32 / 64 / 128):
//     x2  = x1 + proj(o)                  attn.proj + residual        (ptv3.py:290-296, 416-417)
//     h   = LayerNorm(x2)                 norm2                       (ptv3.py:420)
//     out = x2 + fc2(GELU(fc1(h)))        mlp + residual              (ptv3.py:311-321, 421-424)
// (ptv3.py = pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py).  Before: 4 launches (3 GEMMs + 1
// residual/LayerNorm kernel) that wrote and re-read the [n, C] sum, its LayerNorm and the [n, 4C] hidden activation
// (at stage 0: 61 MB written + 61 MB read per block for the hidden tensor alone).  Here a persistent CTA walks
// 128-row tiles; per tile the three GEMMs are chained through TENSOR MEMORY:
//
//   TMA(o tile) -> split -> A operand in TMEM -> GEMM0 (proj) -> ACC0
//   ACC0 + b + x1 (TMA) -> x2 written back to ACC0, LayerNorm in registers (thread == row) -> A operand -> GEMM1 (fc1)
//   per 128 hidden columns: ACC1 + b1 -> GELU -> A operand IN PLACE -> GEMM2 (fc2) accumulates ONTO x2 in ACC0
//   ACC0 + b2 -> swizzled staging -> TMA store
//
// fp32-faithful numerics as in gemm_tc.cu (3-term fp16 hi/lo split, fp32 accumulation).  The weights are the blocks
// cdseg_gemm_pack_b already caches per Linear, streamed through a TMA ring (they stay L2 resident).
//
// Warps: 0-3 row threads (operand conversion + all epilogues), 4 input loader, 5 weight loader, 6 MMA issuer / TMEM owner.
// TMEM columns: ACC1 [0,128) fc1 chunk accumulator -> GELU'd operand | ACC0 [128,128+C) | H [128+C,128+2C) o / h operand.
#include "fused_common.cuh"
#include <cstdlib>

namespace fz {

constexpr int P_THREADS = 224;
constexpr int P_SI = 2, P_SB = 2;               // input / weight ring depth
constexpr int P_STG = 4 * 2 * 4096;             // output staging: per warp two [32 x 32] fp32 boxes

struct PostParams {
  int M, C, ntiles, tmem_cols;
  float eps;
  const __half *Bp_proj, *Bp_fc1, *Bp_fc2;
  const float *b_proj, *ln_g, *ln_b, *b_fc1, *b_fc2;
  int single;                                   // fp16 x fp16 products only (cdseg_set_gemm_precision)
};

struct PostBars {
  uint64_t in_full[P_SI], in_empty[P_SI], b_full[P_SB], b_empty[P_SB], a_rdy[4], acc_done, g_free;
  uint32_t tmem_slot, pad;
};

__global__ void __launch_bounds__(P_THREADS, 2)
post_kernel(const PostParams p, const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX,
            const __grid_constant__ CUtensorMap tmY) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_in = smem;
  uint8_t* s_b = s_in + P_SI * IN_STAGE;
  uint8_t* s_stg = s_b + P_SB * B_STAGE;
  float* s_par = reinterpret_cast<float*>(s_stg + P_STG);       // b_proj[C] g[C] b[C] b2[C] b1[4C]
  PostBars* bars = reinterpret_cast<PostBars*>(s_par + 8 * 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, nc = C / KC, J = nc;                        // J hidden chunks of 128 = 4C / 128
  float* s_bp = s_par; float* s_g = s_par + C; float* s_be = s_par + 2 * C; float* s_b2 = s_par + 3 * C; float* s_b1 = s_par + 4 * C;

  PDL_TRIGGER_EARLY();
  if (threadIdx.x == 0) {
    for (int s = 0; s < P_SI; ++s) { mbar_init(smem_u32(&bars->in_full[s]), 1); mbar_init(smem_u32(&bars->in_empty[s]), 128); }
    for (int s = 0; s < P_SB; ++s) { mbar_init(smem_u32(&bars->b_full[s]), 1); mbar_init(smem_u32(&bars->b_empty[s]), 1); }
    for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&bars->a_rdy[k]), 128);
    mbar_init(smem_u32(&bars->acc_done), 1);
    mbar_init(smem_u32(&bars->g_free), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < C; i += P_THREADS) { s_bp[i] = p.b_proj[i]; s_g[i] = p.ln_g[i]; s_be[i] = p.ln_b[i]; s_b2[i] = p.b_fc2[i]; }
  for (int i = threadIdx.x; i < 4 * C; i += P_THREADS) s_b1[i] = p.b_fc1[i];
  if (warp == 6) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)),
                 "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();                 // everything above read parameters only (never written during a forward); every thread waits here
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  const uint32_t ACC1 = tmem, ACC0 = tmem + 128, H = tmem + 128 + C;

  if (warp < 4) {
    // =========================================== row threads ===========================================
    const int r = threadIdx.x;
    const uint32_t lb = (uint32_t)(warp * 32) << 16;
    uint32_t in_cnt = 0, acc_cnt = 0;
    uint8_t* stg = s_stg + warp * 8192;
    auto in_wait = [&]() -> const uint8_t* {
      const int s = in_cnt % P_SI;
      mbar_wait(smem_u32(&bars->in_full[s]), (in_cnt / P_SI) & 1);
      return s_in + s * IN_STAGE;
    };
    // The box was read through the generic proxy (LDS) and will be overwritten through the async proxy (the next TMA load): the
    // proxy fence orders this thread's reads before the loader's next bulk copy.  Without it the copy could land while a row's LDS
    // were still in flight when a co-resident CTA of another stream saturated the shared-memory pipe -- 8-row groups of an output
    // tile then picked up 16-byte chunks of the NEXT tile's input (profiles/r02_two_stream_race.md).
    auto in_release = [&]() { fence_async_smem(); mbar_arrive(smem_u32(&bars->in_empty[in_cnt % P_SI])); ++in_cnt; };
    auto acc_wait = [&]() { mbar_wait(smem_u32(&bars->acc_done), acc_cnt & 1); ++acc_cnt; tc_fence_after(); };
    auto operand_ready = [&](int k) { tmem_st_wait(); tc_fence_before(); mbar_arrive(smem_u32(&bars->a_rdy[k])); };

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      if (tile + (int)gridDim.x >= p.ntiles) PDL_TRIGGER_LATE();      // this CTA's last tile
      float v[32];
      // ---- o tile -> A operand of GEMM0 (proj)
      for (int kc = 0; kc < nc; ++kc) {
        const uint8_t* box = in_wait();
        lds_row(box, r, v);
        in_release();
        split_store(v, H + lb + kc * 32);
        operand_ready(kc);
      }
      // ---- x2 = proj + b + x1 (kept in ACC0), LayerNorm statistics (shifted single sweep)
      acc_wait();
      float v0 = 0.f, s1 = 0.f, s2 = 0.f;
      for (int c = 0; c < nc; ++c) {
        uint32_t a[32];
        tmem_ld32(ACC0 + lb + c * 32, a);
        const uint8_t* box = in_wait();
        lds_row(box, r, v);
        in_release();
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(a[i]) + s_bp[c * 32 + i];
        if (c == 0) v0 = v[0];
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float d = v[i] - v0; s1 += d; s2 = fmaf(d, d, s2); }
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = __float_as_uint(v[i]);
        tmem_st16(ACC0 + lb + c * 32, a);
        tmem_st16(ACC0 + lb + c * 32 + 16, a + 16);
      }
      tmem_st_wait();
      const float inv_c = 1.0f / (float)C;
      const float dm = s1 * inv_c, mean = v0 + dm;
      const float rstd = rsqrtf(fmaxf(s2 * inv_c - dm * dm, 0.f) + p.eps);
      // ---- h = LayerNorm(x2) -> A operand of GEMM1 (fc1)
      for (int c = 0; c < nc; ++c) {
        uint32_t a[32];
        tmem_ld32(ACC0 + lb + c * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf((__uint_as_float(a[i]) - mean) * rstd, s_g[c * 32 + i], s_be[c * 32 + i]);
        split_store(v, H + lb + c * 32);
        operand_ready(c);
      }
      // ---- hidden chunks: GELU(fc1 + b1) becomes the A operand of GEMM2 (fc2) in place
      for (int j = 0; j < J; ++j) {
        acc_wait();
        for (int k = 0; k < 4; ++k) {
          uint32_t a[32];
          tmem_ld32(ACC1 + lb + k * 32, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_fast(__uint_as_float(a[i]) + s_b1[j * 128 + k * 32 + i]);
          split_store(v, ACC1 + lb + k * 32);
          operand_ready(k);
        }
      }
      // ---- out = ACC0 (= x2 + fc2 sums) + b2 -> staging -> TMA store
      acc_wait();
      for (int c = 0; c < nc; ++c) {
        uint32_t a[32];
        tmem_ld32(ACC0 + lb + c * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a[i]) + s_b2[c * 32 + i];
        uint8_t* buf = stg + (c & 1) * 4096;
        if (c >= 2) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }
        sts_row(buf, lane, v);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { tma_store_2d(&tmY, c * 32, tile * BM + warp * 32, smem_u32(buf)); bulk_commit(); }
      }
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      tc_fence_before();          // this tile's TMEM reads are ordered before the next tile's operand_ready arrivals
    }
  } else if (warp == 4) {
    // =========================================== input loader ===========================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
        for (int i = 0; i < 2 * nc; ++i, ++it) {
          const int s = it % P_SI;
          if (it >= P_SI) mbar_wait(smem_u32(&bars->in_empty[s]), ((it / P_SI) - 1) & 1);
          const uint32_t bar = smem_u32(&bars->in_full[s]);
          mbar_expect_tx(bar, IN_STAGE);
          tma_load_2d(smem_u32(s_in + s * IN_STAGE), i < nc ? &tmO : &tmX, (i % nc) * KC, tile * BM, bar);
        }
    }
  } else if (warp == 5) {
    // =========================================== weight loader ===========================================
    if (lane == 0) {
      uint32_t it = 0;
      auto load = [&](const __half* blk, int un) {
        const int s = it % P_SB;
        if (it >= P_SB) mbar_wait(smem_u32(&bars->b_empty[s]), ((it / P_SB) - 1) & 1);
        const uint32_t bar = smem_u32(&bars->b_full[s]), bytes = (uint32_t)un * KC * 2;
        mbar_expect_tx(bar, 2 * bytes);
        tma_load_1d(smem_u32(s_b + s * B_STAGE), blk, bytes, bar);
        tma_load_1d(smem_u32(s_b + s * B_STAGE + BLK * 2), blk + BLK, bytes, bar);
        ++it;
      };
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nc; ++kc) load(p.Bp_proj + (size_t)kc * 2 * BLK, C);
        for (int j = 0; j < J; ++j) {
          for (int kc = 0; kc < nc; ++kc) load(p.Bp_fc1 + ((size_t)kc * J + j) * 2 * BLK, 128);
          for (int k = 0; k < 4; ++k) load(p.Bp_fc2 + (size_t)(4 * j + k) * 2 * BLK, C);
        }
      }
    }
  } else {
    // =========================================== MMA issuer ===========================================
    // whole warp, one elected lane issues (see gemm_tc.cu / fused_pre.cu)
    {
      uint32_t b_it = 0, ar_use[4] = {0, 0, 0, 0}, gf_cnt = 0;
      const uint32_t idC = idesc_f16(C), id128 = idesc_f16(128);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t uACC1 = tm, uACC0 = tm + 128, uH = tm + 128 + C;
      const bool single = p.single != 0;
      auto a_wait = [&](int k) { mbar_wait(smem_u32(&bars->a_rdy[k]), ar_use[k] & 1); ++ar_use[k]; };
      // one K = 32 chunk: 2 K-steps x (lo.hi + hi.lo + hi.hi), small terms first
      auto chunk = [&](uint32_t d, uint32_t a, uint32_t idesc, bool first) {
        const int s = b_it % P_SB;
        mbar_wait(smem_u32(&bars->b_full[s]), (b_it / P_SB) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t bh = smem_u32(s_b + s * B_STAGE), bl = bh + BLK * 2;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t dh = make_desc(bh + ks * 256, 128, 512), dl = make_desc(bl + ks * 256, 128, 512);
            if (single) {
              umma_f16_ts(d, a + ks * 8, dh, idesc, (first && ks == 0) ? 0u : 1u);
            } else {
              umma_f16_ts(d, a + 16 + ks * 8, dh, idesc, (first && ks == 0) ? 0u : 1u);
              umma_f16_ts(d, a + ks * 8, dl, idesc, 1u);
              umma_f16_ts(d, a + ks * 8, dh, idesc, 1u);
            }
          }
          umma_commit(smem_u32(&bars->b_empty[s]));
        }
        __syncwarp();
        ++b_it;
      };
      auto commit_one = [&](uint32_t bar) { if (elect_one()) umma_commit(bar); __syncwarp(); };
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nc; ++kc) { a_wait(kc); chunk(uACC0, uH + kc * 32, idC, kc == 0); }        // GEMM0: proj
        commit_one(smem_u32(&bars->acc_done));
        for (int j = 0; j < J; ++j) {
          if (j > 0) { mbar_wait(smem_u32(&bars->g_free), gf_cnt & 1); ++gf_cnt; }
          for (int kc = 0; kc < nc; ++kc) {                                                             // GEMM1: fc1 chunk j
            if (j == 0) a_wait(kc);
            chunk(uACC1, uH + kc * 32, id128, kc == 0);
          }
          commit_one(smem_u32(&bars->acc_done));
          for (int k = 0; k < 4; ++k) { a_wait(k); chunk(uACC0, uACC1 + k * 32, idC, false); }            // GEMM2: fc2, onto x2
          commit_one(j + 1 < J ? smem_u32(&bars->g_free) : smem_u32(&bars->acc_done));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

}  // namespace fz

// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int cdseg_make_tmap_f32(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  static EncodeTiledFn enc = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return (EncodeTiledFn)fn;
  }();
  if (!enc) return CDSEG_EINVAL;
  cuuint64_t gdim[2] = {cols, rows}, gstr[1] = {ld * 4};
  cuuint32_t box[2] = {32, box_rows}, es[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CDSEG_OK : CDSEG_EINVAL;
}

extern int g_cdseg_gemm_single;                  // gemm_tc.cu
extern int g_cdseg_fused_per_sm;                 // block_exec.cu

static int sm_count() {
  static int n = [] { int d = 0, v = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
  return n;
}

// out = x2 + fc2(GELU(fc1(LN(x2)))),  x2 = x1 + proj(o) + b_proj.   o, x1, out: fp32 [n, C] contiguous, 16-byte aligned;
// *_Bp: cdseg_gemm_pack_b blocks of W^T ([1][C][C], [1][C][4C], [1][4C][C]).  C in {32, 64, 128}.  See include/cdseg_b200.h.
CDSEG_API int cdseg_post_attn(const float* o, const float* x1, int64_t n, int C, const float* proj_Bp, const float* proj_b,
                              const float* ln_g, const float* ln_b, float eps, const float* fc1_Bp, const float* fc1_b,
                              const float* fc2_Bp, const float* fc2_b, float* out, void* stream) {
  if (n < 0 || (C != 32 && C != 64 && C != 128) || !o || !x1 || !out) return CDSEG_EINVAL;
  if (((uintptr_t)o | (uintptr_t)x1 | (uintptr_t)out) & 15) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  CUtensorMap tmO, tmX, tmY;
  if (cdseg_make_tmap_f32(&tmO, o, (uint64_t)n, C, C, fz::BM) || cdseg_make_tmap_f32(&tmX, x1, (uint64_t)n, C, C, fz::BM) ||
      cdseg_make_tmap_f32(&tmY, out, (uint64_t)n, C, C, 32))
    return CDSEG_EINVAL;
  fz::PostParams p;
  p.M = (int)n; p.C = C; p.ntiles = cdseg_div_up(n, fz::BM); p.eps = eps;
  static const bool excl = [] { const char* e = getenv("CDSEG_TMEM_EXCL"); return e && atoi(e) != 0; }();   // diagnostic: whole TMEM per CTA
  p.tmem_cols = (C <= 64 && !excl) ? 256 : 512;
  p.single = g_cdseg_gemm_single;
  p.Bp_proj = reinterpret_cast<const __half*>(proj_Bp); p.Bp_fc1 = reinterpret_cast<const __half*>(fc1_Bp);
  p.Bp_fc2 = reinterpret_cast<const __half*>(fc2_Bp);
  p.b_proj = proj_b; p.ln_g = ln_g; p.ln_b = ln_b; p.b_fc1 = fc1_b; p.b_fc2 = fc2_b;
  const int per_sm = (C <= 64 && !excl) ? 2 : 1;                               // TMEM: 256 columns per CTA up to C = 64, all 512 at C = 128
  size_t smem = (size_t)fz::P_SI * fz::IN_STAGE + (size_t)fz::P_SB * fz::B_STAGE + fz::P_STG + 8 * 128 * 4 + sizeof(fz::PostBars) + 1024;
  if (per_sm == 1) smem = smem > 120 * 1024 ? smem : 120 * 1024;    // keep a second CTA (blocked in tcgen05.alloc) off the SM
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(fz::post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  const int per_sm_run = (g_cdseg_fused_per_sm > 0 && g_cdseg_fused_per_sm < per_sm) ? g_cdseg_fused_per_sm : per_sm;
  const int grid = p.ntiles < per_sm_run * sm_count() ? p.ntiles : per_sm_run * sm_count();
  cdseg_launch_pdl(fz::post_kernel, dim3(grid), dim3(fz::P_THREADS), smem, (cudaStream_t)stream, p, tmO, tmX, tmY);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
