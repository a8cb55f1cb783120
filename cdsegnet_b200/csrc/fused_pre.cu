// Pre-attention half of a PTv3 Block as ONE kernel (C = 32 / 64 / 128):
//     y   = LayerNorm_cpe(Linear(SubMConv3d_k3(conv_in)))          cpe            (ptv3.py:355-362, 400-402)
//     x1  = x + y (+ t_mlp(t_emb)[batch])                          residual + NN timestep add  (ptv3.py:402-411)
//     qkv = Linear_qkv(LayerNorm_1(x1))                            norm1 + attn.qkv            (ptv3.py:413, 258)
// (ptv3.py = pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py).  Before: conv GEMM, Linear GEMM, two
// residual/LayerNorm kernels and the qkv GEMM = 5 launches with four [n, C] round trips through HBM.  Here a
// persistent CTA walks 128-row tiles and chains the three GEMMs through TENSOR MEMORY (see fused_common.cuh):
//
//   neighbour rows (TMA tile::gather4, absent neighbours zero-filled; taps no row of the tile has are skipped)
//     -> split -> A ring in TMEM -> GEMM conv -> ACCc ; + b -> A operand in place -> GEMM lin -> ACCl
//   ACCl + b -> LayerNorm_cpe -> + x (TMA) + t -> x1 : TMA store, kept in ACCl -> LayerNorm_1 -> A operand in place
//     -> GEMM qkv in 128-column chunks -> ACCq (aliases the idle A ring) -> + b -> TMA store
//
// Warps: 0-3 row threads, 4 input loader (32 lanes: one gather4 each = 128 rows), 5 weight loader, 6 MMA issuer / TMEM owner.
// TMEM columns: RING / ACCq [0,128) | ACCc [128,128+C) | ACCl [128+C,128+2C).
#include "fused_common.cuh"

namespace fz {

constexpr int Q_THREADS = 224;
constexpr int Q_SI = 3, Q_SB = 2, Q_AT = 4;
constexpr int Q_STG = 4 * 4096;                 // output staging: one [32 x 32] fp32 box per warp
constexpr int Q_PAR = 1280;                     // floats of per-channel parameters kept in shared memory (9 C)

struct PreParams {
  int M, C, ntiles, tmem_cols, nq;              // nq = number of 128-column chunks of the qkv GEMM
  float eps;
  const int32_t* nbr; const uint32_t* tile_mask;
  const __half *Bp_conv, *Bp_lin, *Bp_qkv;
  const float *b_conv, *b_lin, *cpe_g, *cpe_b, *n1_g, *n1_b, *b_qkv;
  const float* tproj; const int32_t* batch;     // [B, C] per-scene timestep projection + scene id per row, or NULL
};

struct PreBars {
  uint64_t in_full[Q_SI], in_empty[Q_SI], b_full[Q_SB], b_empty[Q_SB], a_full[Q_AT], a_empty[Q_AT], a_rdy[4], acc_done, q_free;
  uint32_t tmem_slot, pad;
};

__global__ void __launch_bounds__(Q_THREADS, 2)
pre_kernel(const PreParams p, const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
           const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmQ) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_in = smem;
  uint8_t* s_b = s_in + Q_SI * IN_STAGE;
  uint8_t* s_stg = s_b + Q_SB * B_STAGE;
  float* s_par = reinterpret_cast<float*>(s_stg + Q_STG);
  PreBars* bars = reinterpret_cast<PreBars*>(s_par + Q_PAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, nc = C / KC, nq = p.nq;
  float *s_bc = s_par, *s_bl = s_par + C, *s_cg = s_par + 2 * C, *s_cb = s_par + 3 * C, *s_g1 = s_par + 4 * C, *s_b1 = s_par + 5 * C,
        *s_bq = s_par + 6 * C;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Q_SI; ++s) { mbar_init(smem_u32(&bars->in_full[s]), 1); mbar_init(smem_u32(&bars->in_empty[s]), 128); }
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
  if (warp == 6) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)),
                 "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  const uint32_t RING = tmem, ACCQ = tmem, ACCC = tmem + 128, ACCL = tmem + 128 + C;

  if (warp < 4) {
    // =========================================== row threads ===========================================
    const int r = threadIdx.x;
    const uint32_t lb = (uint32_t)(warp * 32) << 16;
    uint32_t in_cnt = 0, acc_cnt = 0, a_it = 0;
    uint8_t* stg = s_stg + warp * 4096;
    auto in_wait = [&]() -> const uint8_t* {
      const int s = in_cnt % Q_SI;
      mbar_wait(smem_u32(&bars->in_full[s]), (in_cnt / Q_SI) & 1);
      return s_in + s * IN_STAGE;
    };
    auto in_release = [&]() { mbar_arrive(smem_u32(&bars->in_empty[in_cnt % Q_SI])); ++in_cnt; };
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

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const long long m = (long long)tile * BM + r;
      const int row0 = tile * BM + warp * 32;
      float v[32];
      // ---- gathered neighbour rows -> A ring (conv as an implicit GEMM over the taps this tile has)
      const int n_conv = __popc(p.tile_mask[tile] & 0x7ffffffu) * nc;
      for (int it = 0; it < n_conv; ++it, ++a_it) {
        const uint8_t* box = in_wait();
        lds_row(box, r, v);
        in_release();
        const int q = a_it % Q_AT;
        if (a_it >= Q_AT) { mbar_wait(smem_u32(&bars->a_empty[q]), ((a_it / Q_AT) - 1) & 1); tc_fence_after(); }
        split_store(v, RING + lb + q * 32);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->a_full[q]));
      }
      // ---- conv + bias -> A operand of the cpe Linear, in place
      acc_wait();
      for (int c = 0; c < nc; ++c) {
        uint32_t a[32];
        tmem_ld32(ACCC + lb + c * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a[i]) + s_bc[c * 32 + i];
        split_store(v, ACCC + lb + c * 32);
        operand_ready(c);
      }
      // ---- u = lin + b: LayerNorm_cpe statistics
      acc_wait();
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
        const uint8_t* box = in_wait();
        lds_row(box, r, v);
        in_release();
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
      // ---- qkv chunks: + bias -> TMA store
      for (int j = 0; j < nq; ++j) {
        acc_wait();
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
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
  } else if (warp == 4) {
    // =========================================== input loader ===========================================
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      uint32_t mask = p.tile_mask[tile] & 0x7ffffffu;
      const long long m4 = (long long)tile * BM + lane * 4;
      int rr[4];
      auto load_idx = [&](int t) {
#pragma unroll
        for (int q = 0; q < 4; ++q) rr[q] = (m4 + q < p.M) ? __ldg(p.nbr + (m4 + q) * 27 + t) : -1;
      };
      if (mask) load_idx(__ffs(mask) - 1);
      while (mask) {
        mask &= mask - 1;
        const int r0 = rr[0], r1 = rr[1], r2 = rr[2], r3 = rr[3];
        if (mask) load_idx(__ffs(mask) - 1);                    // next tap's indices while this tap's copies are issued
        for (int kc = 0; kc < nc; ++kc, ++it) {
          const int s = it % Q_SI;
          if (it >= Q_SI) mbar_wait(smem_u32(&bars->in_empty[s]), ((it / Q_SI) - 1) & 1);
          const uint32_t bar = smem_u32(&bars->in_full[s]);
          if (lane == 0) mbar_expect_tx(bar, IN_STAGE);
          __syncwarp();
          tma_gather4(smem_u32(s_in + s * IN_STAGE) + lane * 512, &tmG, kc * KC, r0, r1, r2, r3, bar);
        }
      }
      for (int c = 0; c < nc; ++c, ++it) {                      // residual tile
        const int s = it % Q_SI;
        if (it >= Q_SI) mbar_wait(smem_u32(&bars->in_empty[s]), ((it / Q_SI) - 1) & 1);
        if (lane == 0) {
          const uint32_t bar = smem_u32(&bars->in_full[s]);
          mbar_expect_tx(bar, IN_STAGE);
          tma_load_2d(smem_u32(s_in + s * IN_STAGE), &tmX, c * KC, tile * BM, bar);
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
        for (uint32_t mask = p.tile_mask[tile] & 0x7ffffffu; mask; mask &= mask - 1) {
          const int t = __ffs(mask) - 1;
          for (int kc = 0; kc < nc; ++kc) load(p.Bp_conv + ((size_t)t * nc + kc) * 2 * BLK, C);
        }
        for (int kc = 0; kc < nc; ++kc) load(p.Bp_lin + (size_t)kc * 2 * BLK, C);
        for (int j = 0; j < nq; ++j)
          for (int kc = 0; kc < nc; ++kc) load(p.Bp_qkv + ((size_t)kc * nq + j) * 2 * BLK, min(128, 3 * C - 128 * j));
      }
    }
  } else {
    // =========================================== MMA issuer ===========================================
    if (lane == 0) {
      uint32_t b_it = 0, a_it = 0, ar_use[4] = {0, 0, 0, 0}, qf_cnt = 0;
      const uint32_t idC = idesc_f16(C);
      auto a_wait = [&](int k) { mbar_wait(smem_u32(&bars->a_rdy[k]), ar_use[k] & 1); ++ar_use[k]; };
      auto chunk = [&](uint32_t d, uint32_t a, uint32_t idesc, bool first) {
        const int s = b_it % Q_SB;
        mbar_wait(smem_u32(&bars->b_full[s]), (b_it / Q_SB) & 1);
        tc_fence_after();
        const uint32_t bh = smem_u32(s_b + s * B_STAGE), bl = bh + BLK * 2;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t dh = make_desc(bh + ks * 256, 128, 512), dl = make_desc(bl + ks * 256, 128, 512);
          umma_f16_ts(d, a + 16 + ks * 8, dh, idesc, (first && ks == 0) ? 0u : 1u);
          umma_f16_ts(d, a + ks * 8, dl, idesc, 1u);
          umma_f16_ts(d, a + ks * 8, dh, idesc, 1u);
        }
        umma_commit(smem_u32(&bars->b_empty[s]));
        ++b_it;
      };
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int n_conv = __popc(p.tile_mask[tile] & 0x7ffffffu) * nc;
        for (int it = 0; it < n_conv; ++it, ++a_it) {                                     // conv
          const int q = a_it % Q_AT;
          mbar_wait(smem_u32(&bars->a_full[q]), (a_it / Q_AT) & 1);
          chunk(ACCC, RING + q * 32, idC, it == 0);
          umma_commit(smem_u32(&bars->a_empty[q]));
        }
        umma_commit(smem_u32(&bars->acc_done));
        for (int kc = 0; kc < nc; ++kc) { a_wait(kc); chunk(ACCL, ACCC + kc * 32, idC, kc == 0); }   // cpe Linear
        umma_commit(smem_u32(&bars->acc_done));
        for (int j = 0; j < nq; ++j) {                                                                 // qkv, 128 columns at a time
          if (j > 0) { mbar_wait(smem_u32(&bars->q_free), qf_cnt & 1); ++qf_cnt; tc_fence_after(); }
          const uint32_t idq = idesc_f16(min(128, 3 * C - 128 * j));
          for (int kc = 0; kc < nc; ++kc) {
            if (j == 0) a_wait(kc);
            chunk(ACCQ, ACCL + kc * 32, idq, kc == 0);
          }
          umma_commit(smem_u32(&bars->acc_done));
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

static int sm_count_pre() {
  static int n = [] { int d = 0, v = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
  return n;
}

// x1 = x + LN_cpe(lin(conv3(conv_in))) (+ tproj[batch]) ; qkv = qkv_lin(LN_1(x1)).  See include/cdseg_b200.h.
CDSEG_API int cdseg_pre_attn(const float* conv_in, const float* x, int64_t n, int C, const int32_t* nbr, const uint32_t* tile_mask,
                             const float* conv_Bp, const float* conv_b, const float* lin_Bp, const float* lin_b,
                             const float* cpe_g, const float* cpe_b, const float* tproj, const int32_t* batch,
                             const float* n1_g, const float* n1_b, float eps, const float* qkv_Bp, const float* qkv_b,
                             float* x1, float* qkv, void* stream) {
  if (n < 0 || (C != 32 && C != 64 && C != 128) || !conv_in || !x || !x1 || !qkv || !nbr || !tile_mask || (tproj && !batch))
    return CDSEG_EINVAL;
  if (((uintptr_t)conv_in | (uintptr_t)x | (uintptr_t)x1 | (uintptr_t)qkv) & 15) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  CUtensorMap tmG, tmX, tmX1, tmQ;
  if (cdseg_make_tmap_f32(&tmG, conv_in, (uint64_t)n, C, C, 1) || cdseg_make_tmap_f32(&tmX, x, (uint64_t)n, C, C, fz::BM) ||
      cdseg_make_tmap_f32(&tmX1, x1, (uint64_t)n, C, C, 32) || cdseg_make_tmap_f32(&tmQ, qkv, (uint64_t)n, 3 * C, 3 * C, 32))
    return CDSEG_EINVAL;
  fz::PreParams p;
  p.M = (int)n; p.C = C; p.ntiles = cdseg_div_up(n, fz::BM); p.eps = eps; p.nq = (3 * C + 127) / 128;
  p.tmem_cols = C <= 64 ? 256 : 512;
  p.nbr = nbr; p.tile_mask = tile_mask;
  p.Bp_conv = reinterpret_cast<const __half*>(conv_Bp); p.Bp_lin = reinterpret_cast<const __half*>(lin_Bp);
  p.Bp_qkv = reinterpret_cast<const __half*>(qkv_Bp);
  p.b_conv = conv_b; p.b_lin = lin_b; p.cpe_g = cpe_g; p.cpe_b = cpe_b; p.n1_g = n1_g; p.n1_b = n1_b; p.b_qkv = qkv_b;
  p.tproj = tproj; p.batch = batch;
  const int per_sm = C <= 64 ? 2 : 1;
  size_t smem = (size_t)fz::Q_SI * fz::IN_STAGE + (size_t)fz::Q_SB * fz::B_STAGE + fz::Q_STG + fz::Q_PAR * 4 + sizeof(fz::PreBars) + 1024;
  if (per_sm == 1) smem = smem > 120 * 1024 ? smem : 120 * 1024;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(fz::pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  const int grid = p.ntiles < per_sm * sm_count_pre() ? p.ntiles : per_sm * sm_count_pre();
  fz::pre_kernel<<<grid, fz::Q_THREADS, smem, (cudaStream_t)stream>>>(p, tmG, tmX, tmX1, tmQ);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
