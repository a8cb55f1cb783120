// Pre-attention half of a PTv3 Block as ONE kernel (C = 32 / 64 / 128):
//     y   = LayerNorm_cpe(Linear(SubMConv3d_k3(conv_in)))          cpe            (ptv3.py:355-362, 400-402)
//     x1  = x + y (+ t_mlp(t_emb)[batch])                          residual + NN timestep add  (ptv3.py:402-411)
//     qkv = Linear_qkv(LayerNorm_1(x1))                            norm1 + attn.qkv            (ptv3.py:413, 258)
// (ptv3.py = pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py).  Before: conv GEMM, Linear GEMM, two
// residual/LayerNorm kernels and the qkv GEMM = 5 launches with four [n, C] round trips through HBM.  Here a
// persistent CTA walks 128-row tiles and chains the three GEMMs through TENSOR MEMORY (see fused_common.cuh):
//
//   conv operand: the UNIQUE neighbour rows of the tile (about 190 of its 3456 (row, tap) slots on ScanNet-like surfaces;
//   cdseg_conv_tile_plan builds the list once per level) are fetched ONCE per 32-channel chunk with TMA tile::gather4 into a
//   shared-memory cache and split to fp16 hi/lo in place; per tap every row thread copies its neighbour's 128 B (or
//   zeros) from the cache into the A ring in TMEM (taps no row of the tile has are skipped)
//     -> GEMM conv -> ACCc ; + b -> A operand in place -> GEMM lin -> ACCl
//   ACCl + b -> LayerNorm_cpe -> + x (TMA) + t -> x1 : TMA store, kept in ACCl -> LayerNorm_1 -> A operand in place
//     -> GEMM qkv in 128-column chunks -> ACCq (aliases the idle A ring) -> + b -> TMA store
//
// (The first version fetched a [128 x 32] box per (tap, chunk) with gather4: 27 x 16 KB per tile, two thirds of it zero
//  fill, one TMA round trip per tap on the critical path: 198 us at stage 0 and 116 us for a single C = 128 tile.)
// A tile whose neighbourhood does not fit the cache (ucount > Q_UCAP: unsorted or volumetric inputs) takes the
// fallback: the row threads read their neighbours straight from global memory, tap by tap.
// Warps: 0-3 row threads, 4 input loader, 5 weight loader, 6 MMA issuer / TMEM owner.
// TMEM columns: RING / ACCq [0,128) | ACCc [128,128+C) | ACCl [128+C,128+2C).
#include "fused_common.cuh"
#include <cstdlib>
#include "../../include/cdseg_b200.h"

namespace fz {

constexpr int Q_THREADS = 224;
constexpr int Q_SB = 2, Q_AT = 4, Q_SX = 2;
constexpr int Q_UCAP = CDSEG_CONV_PLAN_UCAP;     // rows of the neighbour cache (128 B each)
constexpr int Q_CACHE = Q_UCAP * 128;            // 48 KB; after the conv the same bytes hold the x boxes and the store staging
constexpr int Q_STG_OFF = Q_SX * IN_STAGE;       // 4 x 4 KB staging boxes (one [32 x 32] fp32 box per warp) behind the x boxes
constexpr int Q_LIDX = BM * 27 * 2;              // the tile's local neighbour indices (int16 [128][27])
constexpr int Q_REC = 16 + Q_UCAP * 4 + Q_LIDX;  // plan record of one tile: ucount, pad[3], uniq[Q_UCAP], lidx[128][27]
constexpr int Q_PAR = 1280;                      // floats of per-channel parameters kept in shared memory (9 C)
static_assert(Q_STG_OFF + 4 * 4096 <= Q_CACHE, "x boxes + staging must fit the cache bytes");
static_assert(Q_REC % 16 == 0 && (16 + Q_UCAP * 4) % 16 == 0, "bulk copies need 16-byte alignment");

struct PreParams {
  int M, C, ntiles, tmem_cols, nq;              // nq = number of 128-column chunks of the qkv GEMM
  float eps;
  const int32_t* nbr; const uint32_t* tile_mask;
  const uint8_t* plan;                          // cdseg_conv_tile_plan records
  const float* conv_in;                         // read directly only by the overflow fallback
  const __half *Bp_conv, *Bp_lin, *Bp_qkv;
  const float *b_conv, *b_lin, *cpe_g, *cpe_b, *n1_g, *n1_b, *b_qkv;
  const float* tproj; const int32_t* batch;     // [B, C] per-scene timestep projection + scene id per row, or NULL
  long long* trace; int trace_cta;              // profiling hook (cdseg_pre_attn_set_trace): clock64 stamps of one CTA, null in production
  int single;                                   // fp16 x fp16 products only (cdseg_set_gemm_precision)
};

struct PreBars {
  uint64_t fill_full, cache_free, x_full[Q_SX], x_empty[Q_SX], b_full[Q_SB], b_empty[Q_SB], a_full[Q_AT], a_empty[Q_AT], a_rdy[4],
      acc_done, q_free;
  uint32_t tmem_slot, pad;
};

__global__ void __launch_bounds__(Q_THREADS, 2)
pre_kernel(const PreParams p, const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
           const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmQ) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_cache = smem;
  uint8_t* s_b = s_cache + Q_CACHE;
  int16_t* s_lidx = reinterpret_cast<int16_t*>(s_b + Q_SB * B_STAGE);
  float* s_par = reinterpret_cast<float*>(s_b + Q_SB * B_STAGE + Q_LIDX);
  PreBars* bars = reinterpret_cast<PreBars*>(s_par + Q_PAR);
  // 128 zero bytes: the operand row of an absent neighbour, so that the tap loop reads "some row" without a divergent branch
  const uint8_t* s_zero = reinterpret_cast<const uint8_t*>((reinterpret_cast<uintptr_t>(bars + 1) + 15) & ~(uintptr_t)15);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, nc = C / KC, nq = p.nq;
  float *s_bc = s_par, *s_bl = s_par + C, *s_cg = s_par + 2 * C, *s_cb = s_par + 3 * C, *s_g1 = s_par + 4 * C, *s_b1 = s_par + 5 * C,
        *s_bq = s_par + 6 * C;

  PDL_TRIGGER_EARLY();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->fill_full), 1);
    mbar_init(smem_u32(&bars->cache_free), 128);
    for (int s = 0; s < Q_SX; ++s) { mbar_init(smem_u32(&bars->x_full[s]), 1); mbar_init(smem_u32(&bars->x_empty[s]), 128); }
    for (int s = 0; s < Q_SB; ++s) { mbar_init(smem_u32(&bars->b_full[s]), 1); mbar_init(smem_u32(&bars->b_empty[s]), 1); }
    for (int s = 0; s < Q_AT; ++s) { mbar_init(smem_u32(&bars->a_full[s]), 128); mbar_init(smem_u32(&bars->a_empty[s]), 1); }
    for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&bars->a_rdy[k]), 128);
    mbar_init(smem_u32(&bars->acc_done), 1);
    mbar_init(smem_u32(&bars->q_free), 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < C; i += Q_THREADS) {
    s_bc[i] = p.b_conv[i]; s_bl[i] = p.b_lin[i]; s_cg[i] = p.cpe_g[i]; s_cb[i] = p.cpe_b[i]; s_g1[i] = p.n1_g[i]; s_b1[i] = p.n1_b[i];
  }
  for (int i = threadIdx.x; i < 3 * C; i += Q_THREADS) s_bq[i] = p.b_qkv[i];
  if (threadIdx.x < 32) reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(s_zero))[threadIdx.x] = 0u;
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
  const uint32_t RING = tmem, ACCQ = tmem, ACCC = tmem + 128, ACCL = tmem + 128 + C;

  if (warp < 4) {
    // =========================================== row threads ===========================================
    const int r = threadIdx.x;
    const uint32_t lb = (uint32_t)(warp * 32) << 16;
    uint32_t fill_cnt = 0, x_cnt = 0, acc_cnt = 0, a_it = 0;
    uint8_t* stg = s_cache + Q_STG_OFF + warp * 4096;
    auto x_wait = [&]() -> const uint8_t* {
      const int s = x_cnt % Q_SX;
      mbar_wait(smem_u32(&bars->x_full[s]), (x_cnt / Q_SX) & 1);
      return s_cache + s * IN_STAGE;
    };
    // generic-proxy reads (LDS) of the box before the loader's next async-proxy write into it: proxy fence, see fused_post.cu
    auto x_release = [&]() { fence_async_smem(); mbar_arrive(smem_u32(&bars->x_empty[x_cnt % Q_SX])); ++x_cnt; };
    auto acc_wait = [&]() { mbar_wait(smem_u32(&bars->acc_done), acc_cnt & 1); ++acc_cnt; tc_fence_after(); };
    auto operand_ready = [&](int k) { tmem_st_wait(); tc_fence_before(); mbar_arrive(smem_u32(&bars->a_rdy[k])); };
    // one [32 rows x 32 cols] box of this warp: registers -> swizzled staging -> TMA store
    auto store_box = [&](const CUtensorMap* tm, int col, int row0, const float* v) {
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      sts_row(stg, lane, v);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_2d(tm, col, row0, smem_u32(stg)); bulk_commit(); }
    };

    const bool tr = p.trace && (int)blockIdx.x == p.trace_cta && threadIdx.x == 0;
    int tn = 0;                                                    // 12 stamps per tile: see profiles/trace_pre.py
    long long w_empty = 0, w_gather = 0, w_st = 0;                 // tap-loop split of the traced thread: a_empty waits / gather + store issue / wait::st + arrive
#define STAMP(k) do { if (tr && tn < 6) p.trace[tn * 12 + (k)] = clock64(); } while (0)
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tn) {
      STAMP(0);
      if (tile + (int)gridDim.x >= p.ntiles) PDL_TRIGGER_LATE();      // this CTA's last tile
      const long long m = (long long)tile * BM + r;
      const int row0 = tile * BM + warp * 32;
      const int U = __ldg(reinterpret_cast<const int*>(p.plan + (size_t)tile * Q_REC));
      const bool cached = U <= Q_UCAP;
      const uint32_t tmask = p.tile_mask[tile] & 0x7ffffffu;
      float v[32];
      // ---- conv as an implicit GEMM over (32-channel chunk, tap this tile has): neighbour rows -> A ring
      for (int kc = 0; kc < nc; ++kc) {
        if (cached || kc == 0) {                                 // chunk 0 always shakes hands with the loader (see there)
          mbar_wait(smem_u32(&bars->fill_full), fill_cnt & 1);
          ++fill_cnt;
        }
        if (kc == 0) STAMP(1);
        if (cached) {
          for (int u = r; u < U; u += BM) {                      // raw fp32 row -> [hi 16 words | lo 16 words], same swizzled chunks
            uint8_t* row = s_cache + u * 128;
            lds_row(s_cache, u, v);
            uint32_t w[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
              const float2 f = __half22float2(h);
              const __half2 l = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
              w[j] = *reinterpret_cast<const uint32_t*>(&h);
              w[16 + j] = *reinterpret_cast<const uint32_t*>(&l);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(row + ((j ^ (u & 7)) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (kc == 0) STAMP(2);
        for (uint32_t mk = tmask; mk; mk &= mk - 1, ++a_it) {
          const long long c1 = tr ? clock64() : 0;
          const int t = __ffs(mk) - 1;
          uint32_t w[32];
          if (cached) {
            // absent neighbour (li < 0): read the zero row -- one uniform path for the warp (about half of a tile's rows miss any
            // given tap on a surface scan, so the branchy version executed both sides nearly every time)
            const int li = s_lidx[r * 27 + t];
            const uint8_t* row = li >= 0 ? s_cache + li * 128 : s_zero;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 q4 = *reinterpret_cast<const uint4*>(row + ((j ^ (li & 7)) << 4));
              w[4 * j] = q4.x; w[4 * j + 1] = q4.y; w[4 * j + 2] = q4.z; w[4 * j + 3] = q4.w;
            }
          } else {
            const int g = m < p.M ? __ldg(p.nbr + m * 27 + t) : -1;
            if (g >= 0) {
              const float4* src = reinterpret_cast<const float4*>(p.conv_in + (size_t)g * C + kc * KC);
#pragma unroll
              for (int j = 0; j < 8; ++j) { const float4 f4 = __ldg(src + j); v[4 * j] = f4.x; v[4 * j + 1] = f4.y; v[4 * j + 2] = f4.z; v[4 * j + 3] = f4.w; }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
                const float2 f = __half22float2(h);
                const __half2 l = __floats2half2_rn(v[2 * j] - f.x, v[2 * j + 1] - f.y);
                w[j] = *reinterpret_cast<const uint32_t*>(&h);
                w[16 + j] = *reinterpret_cast<const uint32_t*>(&l);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) w[j] = 0u;
            }
          }
          const int q = a_it % Q_AT;
          const long long c0 = tr ? clock64() : 0;
          if (a_it >= Q_AT) { mbar_wait(smem_u32(&bars->a_empty[q]), ((a_it / Q_AT) - 1) & 1); tc_fence_after(); }
          const long long c2 = tr ? clock64() : 0;
          tmem_st16(RING + lb + q * 32, w);
          tmem_st16(RING + lb + q * 32 + 16, w + 16);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(&bars->a_full[q]));
          if (tr) { w_gather += c0 - c1; w_empty += c2 - c0; w_st += clock64() - c2; }
        }
        if (cached || kc == nc - 1) {
          fence_async_smem();                                    // the next fill / x box (async proxy) overwrites what was read and written here
          mbar_arrive(smem_u32(&bars->cache_free));
        }
      }
      STAMP(3);
      // ---- conv + bias -> A operand of the cpe Linear, in place
      acc_wait();
      STAMP(4);
      for (int c = 0; c < nc; ++c) {
        uint32_t a[32];
        tmem_ld32(ACCC + lb + c * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a[i]) + s_bc[c * 32 + i];
        split_store(v, ACCC + lb + c * 32);
        operand_ready(c);
      }
      STAMP(5);
      // ---- u = lin + b: LayerNorm_cpe statistics
      acc_wait();
      STAMP(6);
      float u0 = 0.f, s1 = 0.f, s2 = 0.f;
      for (int c = 0; c < nc; ++c) {
        uint32_t a[32];
        tmem_ld32(ACCL + lb + c * 32, a);
        tmem_ld_wait();
        if (c == 0) u0 = __uint_as_float(a[0]) + s_bl[0];
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float d = __uint_as_float(a[i]) + s_bl[c * 32 + i] - u0; s1 += d; s2 = fmaf(d, d, s2); }
      }
      const float inv_c = 1.0f / (float)C;
      const float dm1 = s1 * inv_c, mean1 = u0 + dm1;
      const float rstd1 = rsqrtf(fmaxf(s2 * inv_c - dm1 * dm1, 0.f) + p.eps);
      // ---- x1 = x + LayerNorm_cpe(u) (+ t): stored, kept in ACCl, LayerNorm_1 statistics
      const float* trow = (p.tproj && m < p.M) ? p.tproj + (long long)__ldg(p.batch + m) * C : nullptr;
      float w0 = 0.f, t1 = 0.f, t2 = 0.f;
      for (int c = 0; c < nc; ++c) {
        uint32_t a[32];
        tmem_ld32(ACCL + lb + c * 32, a);
        const uint8_t* box = x_wait();
        lds_row(box, r, v);
        x_release();
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float u = __uint_as_float(a[i]) + s_bl[c * 32 + i];
          v[i] += fmaf((u - mean1) * rstd1, s_cg[c * 32 + i], s_cb[c * 32 + i]);
        }
        if (trow) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(trow + c * 32) + j);
            v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
          }
        }
        if (c == 0) w0 = v[0];
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float d = v[i] - w0; t1 += d; t2 = fmaf(d, d, t2); a[i] = __float_as_uint(v[i]); }
        tmem_st16(ACCL + lb + c * 32, a);
        tmem_st16(ACCL + lb + c * 32 + 16, a + 16);
        // thread == row, but the staging box is written row by row of the WARP's 32 rows: lane == row inside the box
        store_box(&tmX1, c * 32, row0, v);
      }
      tmem_st_wait();
      STAMP(7);
      const float dm2 = t1 * inv_c, mean2 = w0 + dm2;
      const float rstd2 = rsqrtf(fmaxf(t2 * inv_c - dm2 * dm2, 0.f) + p.eps);
      // ---- h = LayerNorm_1(x1) -> A operand of the qkv GEMM, in place
      for (int c = 0; c < nc; ++c) {
        uint32_t a[32];
        tmem_ld32(ACCL + lb + c * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf((__uint_as_float(a[i]) - mean2) * rstd2, s_g1[c * 32 + i], s_b1[c * 32 + i]);
        split_store(v, ACCL + lb + c * 32);
        operand_ready(c);
      }
      STAMP(8);
      // ---- qkv chunks: + bias -> TMA store
      for (int j = 0; j < nq; ++j) {
        acc_wait();
        if (j == 0) STAMP(9);
        const int un = min(128, 3 * C - 128 * j);
        for (int c = 0; c < un / 32; ++c) {
          uint32_t a[32];
          tmem_ld32(ACCQ + lb + c * 32, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a[i]) + s_bq[j * 128 + c * 32 + i];
          store_box(&tmQ, j * 128 + c * 32, row0, v);
        }
        if (j + 1 < nq) { tc_fence_before(); mbar_arrive(smem_u32(&bars->q_free)); }
      }
      tc_fence_before();
      // ---- the staging boxes live in the cache bytes: the next tile's fill may start once the stores have read them
      STAMP(10);
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      mbar_arrive(smem_u32(&bars->cache_free));
      STAMP(11);
      if (tr && tn < 6) { p.trace[72 + tn * 3] = w_empty; p.trace[73 + tn * 3] = w_gather; p.trace[74 + tn * 3] = w_st; w_empty = w_gather = w_st = 0; }
    }
#undef STAMP
  } else if (warp == 4) {
    // =========================================== input loader ===========================================
    uint32_t free_cnt = 0, x_it = 0;
    bool first = true;
    auto free_wait = [&]() { mbar_wait(smem_u32(&bars->cache_free), free_cnt & 1); ++free_cnt; };
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const uint8_t* rec = p.plan + (size_t)tile * Q_REC;
      const int U = __ldg(reinterpret_cast<const int*>(rec));
      const bool cached = U <= Q_UCAP;
      const int ng = (U + 3) >> 2;
      // cache_free phases per tile: one per chunk whose taps are done (cached tiles: every chunk, others: the last one only)
      // + one when the tile's stores have left the staging boxes.  Every wait below is matched by a row-thread wait on
      // fill_full / x_full before their next arrival, so the two sides never drift by more than one phase.
      for (int kc = 0; kc < nc; ++kc) {
        if (kc == 0) { if (!first) free_wait(); first = false; }
        else if (cached) free_wait();
        const uint32_t bar = smem_u32(&bars->fill_full);
        if (!cached) {
          if (kc == 0 && lane == 0) mbar_arrive(bar);            // nothing to fetch: release the row threads
          __syncwarp();
        } else {
          if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t)ng * 512u + (kc == 0 ? (uint32_t)Q_LIDX : 0u));
            if (kc == 0) tma_load_1d(smem_u32(s_lidx), rec + 16 + Q_UCAP * 4, Q_LIDX, bar);
          }
          __syncwarp();
          for (int g = lane; g < ng; g += 32) {
            const int4 rr = __ldg(reinterpret_cast<const int4*>(rec + 16) + g);
            tma_gather4(smem_u32(s_cache) + g * 512, &tmG, kc * KC, rr.x, rr.y, rr.z, rr.w, bar);
          }
        }
      }
      free_wait();                                               // conv operand consumed: the cache bytes become x boxes + staging
      for (int c = 0; c < nc; ++c, ++x_it) {                     // residual tile
        const int s = x_it % Q_SX;
        if (x_it >= Q_SX) mbar_wait(smem_u32(&bars->x_empty[s]), ((x_it / Q_SX) - 1) & 1);
        if (lane == 0) {
          const uint32_t bar = smem_u32(&bars->x_full[s]);
          mbar_expect_tx(bar, IN_STAGE);
          tma_load_2d(smem_u32(s_cache + s * IN_STAGE), &tmX, c * KC, tile * BM, bar);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // =========================================== weight loader ===========================================
    if (lane == 0) {
      uint32_t it = 0;
      auto load = [&](const __half* blk, int un) {
        const int s = it % Q_SB;
        if (it >= Q_SB) mbar_wait(smem_u32(&bars->b_empty[s]), ((it / Q_SB) - 1) & 1);
        const uint32_t bar = smem_u32(&bars->b_full[s]), bytes = (uint32_t)un * KC * 2;
        mbar_expect_tx(bar, 2 * bytes);
        tma_load_1d(smem_u32(s_b + s * B_STAGE), blk, bytes, bar);
        tma_load_1d(smem_u32(s_b + s * B_STAGE + BLK * 2), blk + BLK, bytes, bar);
        ++it;
      };
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const uint32_t tmask = p.tile_mask[tile] & 0x7ffffffu;
        for (int kc = 0; kc < nc; ++kc)
          for (uint32_t mk = tmask; mk; mk &= mk - 1) load(p.Bp_conv + ((size_t)(__ffs(mk) - 1) * nc + kc) * 2 * BLK, C);
        for (int kc = 0; kc < nc; ++kc) load(p.Bp_lin + (size_t)kc * 2 * BLK, C);
        for (int j = 0; j < nq; ++j)
          for (int kc = 0; kc < nc; ++kc) load(p.Bp_qkv + ((size_t)kc * nq + j) * 2 * BLK, min(128, 3 * C - 128 * j));
      }
    }
  } else {
    // =========================================== MMA issuer ===========================================
    // the whole warp walks the schedule and one elected lane issues (see gemm_tc.cu: a lane-0 branch around the loop costs an
    // R2UR + ELECT + branch sequence per tcgen05 instruction, ~900 cycles per 6-MMA chunk on this one thread)
    {
      uint32_t b_it = 0, a_it = 0, ar_use[4] = {0, 0, 0, 0}, qf_cnt = 0;
      const uint32_t idC = idesc_f16(C);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t uRING = tm, uACCQ = tm, uACCC = tm + 128, uACCL = tm + 128 + C;
      const bool single = p.single != 0;
      auto a_wait = [&](int k) { mbar_wait(smem_u32(&bars->a_rdy[k]), ar_use[k] & 1); ++ar_use[k]; };
      auto chunk = [&](uint32_t d, uint32_t a, uint32_t idesc, bool first, uint32_t extra_commit) {
        const int s = b_it % Q_SB;
        mbar_wait(smem_u32(&bars->b_full[s]), (b_it / Q_SB) & 1);
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
          if (extra_commit) umma_commit(extra_commit);
        }
        __syncwarp();
        ++b_it;
      };
      auto commit_one = [&](uint32_t bar) { if (elect_one()) umma_commit(bar); __syncwarp(); };
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int n_conv = __popc(__shfl_sync(0xffffffffu, p.tile_mask[tile], 0) & 0x7ffffffu) * nc;
        for (int it = 0; it < n_conv; ++it, ++a_it) {                                     // conv
          const int q = a_it % Q_AT;
          mbar_wait(smem_u32(&bars->a_full[q]), (a_it / Q_AT) & 1);
          chunk(uACCC, uRING + q * 32, idC, it == 0, smem_u32(&bars->a_empty[q]));
        }
        commit_one(smem_u32(&bars->acc_done));
        for (int kc = 0; kc < nc; ++kc) { a_wait(kc); chunk(uACCL, uACCC + kc * 32, idC, kc == 0, 0u); }   // cpe Linear
        commit_one(smem_u32(&bars->acc_done));
        for (int j = 0; j < nq; ++j) {                                                                 // qkv, 128 columns at a time
          if (j > 0) { mbar_wait(smem_u32(&bars->q_free), qf_cnt & 1); ++qf_cnt; tc_fence_after(); }
          const uint32_t idq = idesc_f16(min(128, 3 * C - 128 * j));
          for (int kc = 0; kc < nc; ++kc) {
            if (j == 0) a_wait(kc);
            chunk(uACCQ, uACCL + kc * 32, idq, kc == 0, 0u);
          }
          commit_one(smem_u32(&bars->acc_done));
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

// ---- the tile plan: per 128-row tile the sorted list of distinct neighbour rows and every (row, tap)'s index into it ----
constexpr int PL_HASH = 4096;
__device__ __forceinline__ uint32_t pl_hash(int k) { return ((uint32_t)k * 2654435761u) >> 20; }

__global__ void __launch_bounds__(BM) conv_tile_plan_kernel(const int32_t* __restrict__ nbr, int M, uint8_t* __restrict__ plan) {
  __shared__ int s_key[PL_HASH];
  __shared__ int16_t s_val[PL_HASH];
  __shared__ int s_list[Q_UCAP], s_sorted[Q_UCAP];
  __shared__ int s_cnt, s_cnt2;
  const int r = threadIdx.x;
  uint8_t* rec = plan + (size_t)blockIdx.x * Q_REC;
  for (int i = r; i < PL_HASH; i += BM) s_key[i] = -1;
  if (r == 0) { s_cnt = 0; s_cnt2 = 0; }
  __syncthreads();
  const long long m = (long long)blockIdx.x * BM + r;
  int nb[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    const int k = m < M ? __ldg(nbr + m * 27 + t) : -1;
    nb[t] = k;
    if (k >= 0) {
      uint32_t h = pl_hash(k);
      while (true) {
        const int prev = atomicCAS(&s_key[h], -1, k);
        if (prev == -1) { atomicAdd(&s_cnt, 1); break; }
        if (prev == k) break;
        h = (h + 1) & (PL_HASH - 1);
      }
    }
  }
  __syncthreads();
  const int U = s_cnt;
  if (U > Q_UCAP) {                                      // does not fit the cache: the consumer takes its fallback path
    if (r == 0) *reinterpret_cast<int*>(rec) = U;
    return;
  }
  for (int i = r; i < PL_HASH; i += BM)
    if (s_key[i] >= 0) s_list[atomicAdd(&s_cnt2, 1)] = s_key[i];
  __syncthreads();
  auto find = [&](int k) {
    uint32_t h = pl_hash(k);
    while (s_key[h] != k) h = (h + 1) & (PL_HASH - 1);
    return h;
  };
  for (int i = r; i < U; i += BM) {                      // rank sort (U <= 384): ascending row ids, so gather4 quads share cache lines
    const int k = s_list[i];
    int rank = 0;
    for (int j = 0; j < U; ++j) rank += s_list[j] < k;
    s_sorted[rank] = k;
    s_val[find(k)] = (int16_t)rank;
  }
  __syncthreads();
  int* hdr = reinterpret_cast<int*>(rec);
  if (r < 4) hdr[r] = r == 0 ? U : 0;
  for (int i = r; i < Q_UCAP; i += BM) hdr[4 + i] = i < U ? s_sorted[i] : -1;
  int16_t* lidx = reinterpret_cast<int16_t*>(rec + 16 + Q_UCAP * 4) + r * 27;
#pragma unroll
  for (int t = 0; t < 27; ++t) lidx[t] = nb[t] >= 0 ? s_val[find(nb[t])] : (int16_t)-1;
}

}  // namespace fz

static long long* g_pre_trace = nullptr;
static int g_pre_trace_cta = 0;
extern int g_cdseg_gemm_single;                  // gemm_tc.cu
extern int g_cdseg_fused_per_sm;                 // block_exec.cu
CDSEG_API void cdseg_pre_attn_set_trace(long long* buf, int cta) { g_pre_trace = buf; g_pre_trace_cta = cta; }

static int sm_count_pre() {
  static int n = [] { int d = 0, v = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
  return n;
}

CDSEG_API size_t cdseg_conv_plan_bytes(int64_t n) { return (size_t)cdseg_div_up(n > 0 ? n : 1, fz::BM) * fz::Q_REC; }

CDSEG_API int cdseg_conv_tile_plan(const int32_t* nbr, int64_t n, void* plan, void* stream) {
  if (n < 0 || !nbr || !plan || ((uintptr_t)plan & 15)) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  fz::conv_tile_plan_kernel<<<(unsigned)cdseg_div_up(n, fz::BM), fz::BM, 0, (cudaStream_t)stream>>>(nbr, (int)n, (uint8_t*)plan);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// x1 = x + LN_cpe(lin(conv3(conv_in))) (+ tproj[batch]) ; qkv = qkv_lin(LN_1(x1)).  See include/cdseg_b200.h.
CDSEG_API int cdseg_pre_attn(const float* conv_in, const float* x, int64_t n, int C, const int32_t* nbr, const uint32_t* tile_mask,
                             const void* plan, const float* conv_Bp, const float* conv_b, const float* lin_Bp, const float* lin_b,
                             const float* cpe_g, const float* cpe_b, const float* tproj, const int32_t* batch,
                             const float* n1_g, const float* n1_b, float eps, const float* qkv_Bp, const float* qkv_b,
                             float* x1, float* qkv, void* stream) {
  if (n < 0 || (C != 32 && C != 64 && C != 128) || !conv_in || !x || !x1 || !qkv || !nbr || !tile_mask || !plan || (tproj && !batch))
    return CDSEG_EINVAL;
  if (((uintptr_t)conv_in | (uintptr_t)x | (uintptr_t)x1 | (uintptr_t)qkv) & 15) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  CUtensorMap tmG, tmX, tmX1, tmQ;
  if (cdseg_make_tmap_f32(&tmG, conv_in, (uint64_t)n, C, C, 1) || cdseg_make_tmap_f32(&tmX, x, (uint64_t)n, C, C, fz::BM) ||
      cdseg_make_tmap_f32(&tmX1, x1, (uint64_t)n, C, C, 32) || cdseg_make_tmap_f32(&tmQ, qkv, (uint64_t)n, 3 * C, 3 * C, 32))
    return CDSEG_EINVAL;
  fz::PreParams p;
  p.M = (int)n; p.C = C; p.ntiles = cdseg_div_up(n, fz::BM); p.eps = eps; p.nq = (3 * C + 127) / 128;
  static const bool excl = [] { const char* e = getenv("CDSEG_TMEM_EXCL"); return e && atoi(e) != 0; }();   // diagnostic: whole TMEM per CTA
  p.tmem_cols = (C <= 64 && !excl) ? 256 : 512;
  p.nbr = nbr; p.tile_mask = tile_mask; p.plan = (const uint8_t*)plan; p.conv_in = conv_in;
  p.Bp_conv = reinterpret_cast<const __half*>(conv_Bp); p.Bp_lin = reinterpret_cast<const __half*>(lin_Bp);
  p.Bp_qkv = reinterpret_cast<const __half*>(qkv_Bp);
  p.b_conv = conv_b; p.b_lin = lin_b; p.cpe_g = cpe_g; p.cpe_b = cpe_b; p.n1_g = n1_g; p.n1_b = n1_b; p.b_qkv = qkv_b;
  p.tproj = tproj; p.batch = batch;
  p.trace = g_pre_trace; p.trace_cta = g_pre_trace_cta;
  p.single = g_cdseg_gemm_single;
  const int per_sm = (C <= 64 && !excl) ? 2 : 1;
  size_t smem = (size_t)fz::Q_CACHE + (size_t)fz::Q_SB * fz::B_STAGE + fz::Q_LIDX + fz::Q_PAR * 4 + sizeof(fz::PreBars) + 144 + 1024;
  if (per_sm == 1) smem = smem > 120 * 1024 ? smem : 120 * 1024;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(fz::pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  const int per_sm_run = (g_cdseg_fused_per_sm > 0 && g_cdseg_fused_per_sm < per_sm) ? g_cdseg_fused_per_sm : per_sm;
  const int grid = p.ntiles < per_sm_run * sm_count_pre() ? p.ntiles : per_sm_run * sm_count_pre();
  cdseg_launch_pdl(fz::pre_kernel, dim3(grid), dim3(fz::Q_THREADS), smem, (cudaStream_t)stream, p, tmG, tmX, tmX1, tmQ);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
