// fp32-faithful GEMM on the 5th-gen tensor cores (tcgen05, kind::f16, 3-term fp16 hi/lo split) with
// (a) optionally GATHERED A rows -- the submanifold convolution as an implicit GEMM over taps --
// and (b) a fused epilogue (bias, GELU, residual).  One kernel serves
//   * spconv.SubMConv3d k=3 (ptv3.py:356-362, 1106-1123):  out[m] = b + sum_t in[nbr[m,t]] . W_t
//   * every nn.Linear on the path (cpe.1, attn.qkv/proj, mlp.fc1/fc2, pooling/unpooling proj,
//     cross-attention q/kv/proj: ptv3.py:185-186, 311-313, 359, 458, 575-581, 911-913)
// (ptv3.py = pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py)
//
// Numerics: every fp32 operand x is split into hi = fp16(x) and lo = fp16(x - hi) (22 significant bits);
// acc += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi with fp32 accumulation in TMEM: relative error ~2^-21 per product,
// i.e. fp32-class results (the reference runs these layers in fp32 at inference).  fp16 rather than TF32 halves
// because ONE tcgen05.mma covers K=16 fp16 but only K=8 tf32, and the single issuing thread -- ~50 cycles per
// tcgen05.mma regardless of N <= 64, measured in profiles/r01_microbench_mma_issue_latency.txt -- is what bounds
// these skinny (N = 32..128) GEMMs.  Valid for |x| < 65504 (activations / weights of this network are O(1..100)).
//
// CTA = 128 output rows x one N tile (<= 128 columns), 192 threads:
//   warps 0-3  A producers (thread == row == TMEM lane): gather 64 B of the row per k-chunk, split hi/lo, store
//              both into tensor memory (the A operand never touches shared memory); afterwards the epilogue
//              (TMEM lane == row): + bias, GELU, + residual, fp32 rows to HBM
//   warp 4     B loader: one TMA bulk copy per k-chunk of the pre-split, pre-tiled weight block
//   warp 5     MMA issuer: 6 x tcgen05.mma (M128 x N x K8) per k-chunk, commit -> frees the stage
// A ring in TENSOR MEMORY (tcgen05.st by the producers, TS-form MMA), TMA ring for B in shared memory; taps that no row of the tile has are skipped via a per-tile tap mask;
// optional split over taps (grid.z) for levels with few rows (partials reduced by a second kernel).
#include "common.cuh"
#include <cstdlib>

namespace gt {

constexpr int BM = 128;
constexpr int NT = 128;                // max N tile (TMEM columns); NT = 64 (4 CTAs per SM) measured no faster: 19.36 vs 19.08 ms/step
constexpr int KC = 32;                 // fp32 elements per k-chunk (128 B per row) = 2 MMA K-steps of 16
constexpr int NTHREADS = 192;
// a stage holds A_hi | A_lo | B_hi | B_lo ; the B blocks are sized for this launch's widest N tile
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
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Params {
  const float* A; long long lda;
  const int32_t* idx; int T;              // idx [M, T] (row of A feeding tap t); null => row m, tap t = columns [tK, (t+1)K)
  int sub, taps_ld;                       // sub = 4: im2col mode, iteration t = taps 4t..4t+3 x 8 channels of idx [M, taps_ld]
  const uint32_t* tile_mask;              // per 128-row tile: bit t set <=> some row has tap t ; null => all taps
  const float* Bp;                        // packed weights [T][K/KC][ntiles][2][NT*KC]
  int M, N, K;
  const float* bias; const float* res; long long ldr; int act;
  float* out; long long ldo;
  float* part; int nsplit;                // nsplit > 1: raw partial sums to part[z][M][N]
  int nw;                                 // output-tile width in columns (32 / 64 / 128): a CTA owns columns [blockIdx.y * nw, + nw)
  int vec_ok;                             // output / residual rows are 16-byte aligned
  int fast_epi;                           // aligned full-width tiles: the compact epilogue (see the kernel)
  int a256;                               // A rows are 32-byte aligned: 256-bit loads (half the LSU wavefronts of the row gather)
  int AT, SB, b_bytes, acc_cols, tmem_cols;   // A ring slots (TMEM), B ring stages (smem), bytes of one B block, TMEM layout
  long long* trace; int trace_cta;        // profiling hook: clock64 stamps of one CTA (null in production)
  int single;                             // 1: fp16 x fp16 products only (hi.hi; the lo terms are skipped): the reduced-precision mode
};

constexpr int MAX_RING = 4;             // A ring slots (TMEM)
constexpr int MAX_RING_B = 8;           // B ring stages (shared memory): weights depend on nothing, so the loader may run far ahead
constexpr int STG_BYTES = 4 * 32 * 36 * 4;   // epilogue staging: 4 warps x 32 rows x 36 floats

struct Bars {
  uint64_t full_a[MAX_RING], empty_a[MAX_RING], full_b[MAX_RING_B], empty_b[MAX_RING_B], acc;
  uint32_t tmem_slot, pad;
  uint8_t taps[32];                       // present taps of this CTA's split, ascending
};

// D[tmem] (+)= A[tmem] . B[smem]   (A: lane = row, two fp16 K elements per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// The A operand never touches shared memory: each producer thread gathers its row (prefetched in registers), splits it
// into fp16 hi / lo and stores both straight into TENSOR MEMORY (lane == row), from where tcgen05.mma reads it (TS form).
// Variants measured on B200 for the stage-0 conv (120k x 27 taps x 32->32), see profiles/r01_gemm_tc_history.md:
// smem-A tf32 143 us, +deep cp.async ring 270 us, TMEM-A tf32 152 us, TMEM-A fp16 (this) 127 us, coalesced smem-A 230 us.
template <int MINB, int AMODE>
__global__ void __launch_bounds__(NTHREADS, MINB) gemm_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int AT = p.AT, SB = p.SB, B_BYTES = p.b_bytes;
  uint8_t* s_stg = smem;
  uint8_t* s_b = smem + STG_BYTES;
  Bars* bars = reinterpret_cast<Bars*>(s_b + (size_t)SB * 2 * B_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y, z = blockIdx.z;
  const int ntiles = (p.N + NT - 1) / NT;
  const int n0 = tile_n * p.nw;
  const int wn = min(p.nw, p.N - n0);               // valid output columns of this tile
  const int un = (wn + 15) & ~15;                   // UMMA N (multiple of 16)
  const int kch = (p.K + KC - 1) / KC;
  const int t_begin = (int)((long long)p.T * z / p.nsplit), t_end = (int)((long long)p.T * (z + 1) / p.nsplit);

  PDL_TRIGGER_EARLY();
  if (threadIdx.x < 2 * MAX_RING + 2 * MAX_RING_B + 1) {   // one barrier per thread (Bars: full_a, empty_a, full_b, empty_b, acc are contiguous)
    const int b = threadIdx.x;
    mbar_init(smem_u32(&bars->full_a[0]) + 8u * b, b < MAX_RING ? 128u : 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Dense mode: the tap list is the identity and the B loader (warp 4) reads weights only, so it skips the wait and streams the first
  // weight blocks while the preceding kernel of the stream is still finishing; every other thread waits before its first global access.
  if (AMODE != 0 || warp != 4) pdl_wait();        // nothing above touches global memory
  const uint32_t mask = (AMODE != 0 && p.tile_mask) ? p.tile_mask[tile_m] : 0xffffffffu;
  int ntap = t_end - t_begin;
  if (AMODE != 0 && p.T <= 32) {
    const uint32_t width = (uint32_t)(t_end - t_begin);
    const uint32_t present = mask & ((width >= 32 ? 0xffffffffu : ((1u << width) - 1u)) << t_begin);
    ntap = __popc(present);
    if (threadIdx.x == 32) {
      int n = 0;
      for (uint32_t m = present; m; m &= m - 1) bars->taps[n++] = (uint8_t)(__ffs(m) - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  const bool tr = p.trace && blockIdx.x == p.trace_cta && blockIdx.y == 0 && blockIdx.z == 0;
  if (tr && threadIdx.x == 0) p.trace[0] = clock64();
  const uint32_t tmem_a = tmem + (uint32_t)p.acc_cols;          // A ring: slot q at +32q : [hi 16 cols | lo 16 cols]

  // present taps of this CTA's split (listed in shared memory by thread 32 before the barrier above)
  // -> iteration it = (tap taps[it / kch], chunk it % kch)
  const int n_iter = ntap * kch;
  auto tap_of_slot = [&](int j) { return (AMODE != 0 && p.T <= 32) ? (int)bars->taps[j] : t_begin + j; };
  auto tap_of = [&](int it) { return tap_of_slot(it / kch); };

  if (warp < 4) {
    // ------------------------------- A producers -------------------------------
    // AMODE (compile time, so that each launch runs only its own address arithmetic -- the generic version of this loop was ~480
    // instructions per k-iteration, and with one producer warp per scheduler at the deep levels that IS the iteration time):
    //   0 dense rows (Linear, split-K slices)   1 gathered rows, tile's indices staged in shared memory (T <= 32)
    //   2 im2col (4 taps x 8 channels per chunk)   3 gathered rows, indices read from global memory (T > 32)
    const int r = threadIdx.x;
    const long long m = (long long)tile_m * BM + r;
    const bool row_ok = m < p.M;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int32_t* s_idx = reinterpret_cast<int32_t*>(s_b + (size_t)SB * 2 * B_BYTES + sizeof(Bars) + 64);
    if (AMODE == 1) {
      // this tile's neighbour indices, staged once (one coalesced block read) so that the per-tap row address no longer hangs off a
      // dependent global load
      const long long base = (long long)tile_m * BM * p.T, total = (long long)p.M * p.T;
      for (int j = r; j < BM * p.T; j += 128) s_idx[j] = base + j < total ? __ldg(p.idx + base + j) : -1;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const bool full_k = (p.K % KC) == 0;
    // fetches are requested in iteration order: (tap slot fj, chunk fkc) advance incrementally instead of it / kch, it % kch
    int fj = 0, fkc = 0;
    auto fetch = [&](float4* v) {
      const int t = tap_of_slot(fj);
      const int kc = fkc;
      if (++fkc == kch) { fkc = 0; ++fj; }
      if (AMODE == 2) {
        // im2col chunk (tiny C_in, padded to 8): K = 32 = 4 taps x 8 channels, each tap a 32-byte row of A
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int tap = t * 4 + q;
          const int s8 = (row_ok && tap < p.taps_ld) ? __ldg(p.idx + m * p.taps_ld + tap) : -1;
          if (s8 >= 0) {
            const float4* row = reinterpret_cast<const float4*>(p.A + (long long)s8 * 8);
            if (p.a256) ldg256(p.A + (long long)s8 * 8, v[2 * q], v[2 * q + 1]);
            else { v[2 * q] = __ldg(row); v[2 * q + 1] = __ldg(row + 1); }
          } else {
            v[2 * q] = v[2 * q + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        return;
      }
      long long src = -1;
      if (AMODE == 1) src = s_idx[r * p.T + t];
      else if (row_ok) src = AMODE == 3 ? (long long)__ldg(p.idx + m * p.T + t) : m;
      if (src >= 0) {
        const float4* row = reinterpret_cast<const float4*>(p.A + src * p.lda + (AMODE == 0 ? (long long)t * p.K : 0)) + kc * (KC / 4);
        if (full_k && p.a256) {
#pragma unroll
          for (int j = 0; j < KC / 4; j += 2) ldg256(reinterpret_cast<const float*>(row + j), v[j], v[j + 1]);
        } else if (full_k) {
#pragma unroll
          for (int j = 0; j < KC / 4; ++j) v[j] = __ldg(row + j);
        } else {
          const int kleft = p.K - kc * KC;                                // valid fp32 elements of this chunk (K may be 16 mod 32)
#pragma unroll
          for (int j = 0; j < KC / 4; ++j) v[j] = j * 4 < kleft ? __ldg(row + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
#pragma unroll
        for (int j = 0; j < KC / 4; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // two gathers alternate in registers: iteration it+2 is requested right after iteration it has been stored
    int pq = 0, pu = 0;                                                  // ring slot / round of the iteration being stored
    auto proc = [&](int it, float4* v) {
      const int q = pq, u = pu;
      if (++pq == AT) { pq = 0; ++pu; }
      uint32_t hi[KC / 2], lo[KC / 2];
#pragma unroll
      for (int j = 0; j < KC / 4; ++j) {
        const __half2 h0 = __floats2half2_rn(v[j].x, v[j].y), h1 = __floats2half2_rn(v[j].z, v[j].w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(v[j].x - f0.x, v[j].y - f0.y), l1 = __floats2half2_rn(v[j].z - f1.x, v[j].w - f1.y);
        hi[2 * j] = *reinterpret_cast<const uint32_t*>(&h0); hi[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h1);
        lo[2 * j] = *reinterpret_cast<const uint32_t*>(&l0); lo[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&l1);
      }
      if (tr && threadIdx.x == 0 && it == 0) p.trace[2] = clock64();     // first chunk arrived + converted
      if (it + 2 < n_iter) fetch(v);
      if (u > 0) {
        mbar_wait(smem_u32(&bars->empty_a[q]), (uint32_t)((u - 1) & 1));
        tc_fence_after();
      }
      tmem_st16(tmem_a + lane_base + q * 32, hi);
      tmem_st16(tmem_a + lane_base + q * 32 + 16, lo);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->full_a[q]));
      if (tr && threadIdx.x == 0 && it == 0) p.trace[3] = clock64();     // first chunk stored to TMEM + signalled
    };
    float4 v0[KC / 4], v1[KC / 4];
    if (n_iter > 0) fetch(v0);
    if (n_iter > 1) fetch(v1);
    if (tr && threadIdx.x == 0) p.trace[1] = clock64();                  // first gathers issued
    for (int it = 0; it < n_iter; it += 2) {
      proc(it, v0);
      if (it + 1 < n_iter) proc(it + 1, v1);
    }
    // --------------------------------- epilogue ---------------------------------
    // TMEM (lane == row) -> registers -> bias/GELU -> shared-memory staging (each warp only touches its own
    // 32 rows, so __syncwarp suffices) -> coalesced 128-byte row segments to HBM (+ residual, read coalesced).
    if (tr && threadIdx.x == 0) p.trace[4] = clock64();                  // producer loop done
    PDL_TRIGGER_LATE();                                                  // only the last MMAs and the epilogue are left
    if (n_iter > 0) {
      mbar_wait(smem_u32(&bars->acc), 0);
      tc_fence_after();
    }
    if (tr && threadIdx.x == 0) p.trace[5] = clock64();                  // accumulator complete
    constexpr int SLD = 36;                         // staging row stride in floats (conflict-free float4 access)
    float* stg = reinterpret_cast<float*>(s_stg) + (size_t)warp * 32 * SLD;
    const bool final_out = p.nsplit == 1;
    float* obase = final_out ? p.out : p.part + (long long)z * p.M * p.N;
    const long long old = final_out ? p.ldo : (long long)p.N;
    if (p.fast_epi && (wn & 31) == 0 && n_iter > 0) {
      // Compact epilogue for the common case (aligned rows, tile width a multiple of 32, so every pass is 32 full columns).  The general
      // loop below compiles to ~2 300 instructions per pass (scalar fallbacks, 64-bit index arithmetic per element); with one producer
      // warp per scheduler at the deep levels that is 2 500 cycles per pass and 10 000 per 128-column tile
      // (profiles/r02_trace_gemm_deep.txt).  Same arithmetic in the same order: bit-identical results.
      const int rsub = lane >> 3, cc = (lane & 7) * 4;
      const long long m0 = (long long)tile_m * BM + warp * 32 + rsub;
      const long long left = (long long)p.M - m0;
      const int nvalid = left <= 0 ? 0 : (left >= 29 ? 8 : (int)((left + 3) >> 2));      // rows m0 + 4 i < M
      const bool has_res = final_out && p.res != nullptr;
      const bool do_gelu = final_out && p.act == 1;
      const bool has_bias = final_out && p.bias != nullptr;
      float* op0 = obase + m0 * old + n0 + cc;
      const float* rp0 = has_res ? p.res + m0 * p.ldr + n0 + cc : nullptr;
      const long long ostep = 4 * old, rstep = 4 * p.ldr;
      const float* srow = stg + rsub * SLD + cc;
      for (int c0 = 0; c0 < wn; c0 += 32) {
        uint32_t acc[32];
        tmem_ld16(tmem + lane_base + c0, acc);
        tmem_ld16(tmem + lane_base + c0 + 16, acc + 16);
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + cc));
        float4 q[8];
        if (has_res) {
          const float* rp = rp0 + c0;
#pragma unroll
          for (int i = 0; i < 8; ++i, rp += rstep) q[i] = i < nvalid ? __ldg(reinterpret_cast<const float4*>(rp)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<uint4*>(stg + lane * SLD + j4 * 4) = make_uint4(acc[j4 * 4], acc[j4 * 4 + 1], acc[j4 * 4 + 2], acc[j4 * 4 + 3]);
        __syncwarp();
        float* op = op0 + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i, op += ostep) {
          if (i < nvalid) {
            float4 v = *reinterpret_cast<const float4*>(srow + i * 4 * SLD);
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            if (do_gelu) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
            if (has_res) { v.x += q[i].x; v.y += q[i].y; v.z += q[i].z; v.w += q[i].w; }
            *reinterpret_cast<float4*>(op) = v;
          }
        }
        __syncwarp();
        if (tr && threadIdx.x == 0 && c0 / 32 < 4) p.trace[6 + c0 / 32] = clock64();   // epilogue pass done
      }
    } else
    for (int c0 = 0; c0 < un; c0 += 32) {
      const int cw = min(32, un - c0);              // 32 or 16 accumulator columns in this pass
      uint32_t acc[32];
      if (n_iter > 0) {
        tmem_ld16(tmem + lane_base + c0, acc);
        if (cw == 32) tmem_ld16(tmem + lane_base + c0 + 16, acc + 16);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0u;
      }
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        if (j4 * 4 < cw)
          *reinterpret_cast<uint4*>(stg + lane * SLD + j4 * 4) = make_uint4(acc[j4 * 4], acc[j4 * 4 + 1], acc[j4 * 4 + 2], acc[j4 * 4 + 3]);
      __syncwarp();
      // write-out: lane owns columns cc..cc+3 of rows rr = 4i + lane/8, so bias is one 4-vector per pass and the
      // residual reads / output writes are full 128-byte row segments
      const int cc = (lane & 7) * 4, c = c0 + cc;
      const bool col_ok = cc < cw && c < wn;
      const bool vec = p.vec_ok && c + 3 < wn;
      float b4[4] = {0.f, 0.f, 0.f, 0.f};
      if (final_out && p.bias && col_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < wn) b4[j] = __ldg(p.bias + n0 + c + j);
      }
      const bool has_res = final_out && p.res != nullptr;
      const bool do_gelu = final_out && p.act == 1;
      const long long m_base = (long long)tile_m * BM + warp * 32 + (lane >> 3);
      float4 q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long mm = m_base + i * 4;
        q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_res && col_ok && mm < p.M) {
          const float* rp = p.res + mm * p.ldr + n0 + c;
          if (vec) q[i] = *reinterpret_cast<const float4*>(rp);
          else {
            q[i].x = rp[0];
            if (c + 1 < wn) q[i].y = rp[1];
            if (c + 2 < wn) q[i].z = rp[2];
            if (c + 3 < wn) q[i].w = rp[3];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long mm = m_base + i * 4;
        if (mm < p.M && col_ok) {
          float4 v = *reinterpret_cast<const float4*>(stg + (i * 4 + (lane >> 3)) * SLD + cc);
          v.x += b4[0]; v.y += b4[1]; v.z += b4[2]; v.w += b4[3];
          if (do_gelu) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
          v.x += q[i].x; v.y += q[i].y; v.z += q[i].z; v.w += q[i].w;
          float* op = obase + mm * old + n0 + c;
          if (vec) *reinterpret_cast<float4*>(op) = v;
          else {
            op[0] = v.x;
            if (c + 1 < wn) op[1] = v.y;
            if (c + 2 < wn) op[2] = v.z;
            if (c + 3 < wn) op[3] = v.w;
          }
        }
      }
      __syncwarp();
      if (tr && threadIdx.x == 0 && c0 / 32 < 4) p.trace[6 + c0 / 32] = clock64();   // epilogue pass done
    }
  } else if (warp == 4) {
    // -------------------------------- B loader --------------------------------
    // (this loop and the MMA issuer's run on ONE thread each: every integer division by a runtime ring depth or chunk count costs that
    // thread ~100 cycles, so stage / phase / tap / chunk advance incrementally -- profiles/r02_microbench_mma_commit.txt)
    if (lane == 0) {
      const uint32_t bbytes = (uint32_t)un * KC * 2;            // prefix of the fp16 hi / lo block (n-groups are outermost)
      const long long blk_step = (long long)ntiles * 2 * (NT * KC);       // halves between consecutive (tap, chunk) blocks
      const __half* blk0 = reinterpret_cast<const __half*>(p.Bp) + (long long)(n0 / NT) * 2 * (NT * KC) + (n0 % NT) * KC;
      int s = 0, kc = 0, j = 0;
      uint32_t ph = 1;                                           // parity to wait for on empty_b: the first round passes
      uint32_t b_hi = smem_u32(s_b);
      const __half* blk = blk0 + (long long)tap_of_slot(0) * kch * blk_step;
      for (int it = 0; it < n_iter; ++it) {
        if (it >= SB) mbar_wait(smem_u32(&bars->empty_b[s]), ph);
        // packed block of the 128-column group this tile lies in; a narrower tile starts (n0 % NT) / 8 row-groups of 8 n x 32 k into it
        const uint32_t full = smem_u32(&bars->full_b[s]);
        mbar_expect_tx(full, 2 * bbytes);
        tma_load_1d(b_hi, blk, bbytes, full);
        tma_load_1d(b_hi + B_BYTES, blk + NT * KC, bbytes, full);
        b_hi += 2 * B_BYTES;
        if (++s == SB) { s = 0; b_hi = smem_u32(s_b); ph ^= 1; }
        blk += blk_step;
        if (++kc == kch) { kc = 0; ++j; if (it + 1 < n_iter) blk = blk0 + (long long)tap_of_slot(j) * kch * blk_step; }
      }
    }
  } else {
    // ------------------------------- MMA issuer -------------------------------
    // The WHOLE warp runs this loop and one elected lane issues: with `if (lane == 0)` around the loop the operands live in per-lane
    // registers and the compiler wraps every tcgen05 instruction (a uniform-datapath instruction) in an R2UR + ELECT + BRA.U.ANY
    // sequence -- ~150 instructions and ~900 cycles per k-iteration for 6 MMAs, slower than the four producer warps deliver A.
    if (n_iter > 0) {
      const uint32_t tmem = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);          // warp-uniform copies
      const uint32_t tmem_a = tmem + (uint32_t)p.acc_cols;
      // kind::f16: D fp32 (1<<4), A = B = F16 (format 0), both K-major, N, M=128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(un >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t dconst = make_desc(0, 128, (KC / 8) * 128);          // + (shared address >> 4) in the low 14 bits
      int q = 0, s = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t a_hi = tmem_a, b_hi = smem_u32(s_b) >> 4;
      const uint32_t b_lo_off = (uint32_t)B_BYTES >> 4;
      const bool single = p.single != 0;
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(smem_u32(&bars->full_a[q]), pa);
        if (tr && it == 0 && lane == 0) p.trace[10] = clock64();         // MMA: A ready
        mbar_wait(smem_u32(&bars->full_b[s]), pb);
        if (tr && it == 0 && lane == 0) p.trace[11] = clock64();         // MMA: B ready
        tc_fence_after();
        const uint32_t a_lo = a_hi + 16;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {                 // K = 16 per MMA: 8 TMEM columns of A, 2 core matrices (256 B) of B
            const uint64_t bh = dconst | (uint64_t)(b_hi + ks * 16), bl = dconst | (uint64_t)(b_hi + b_lo_off + ks * 16);
            if (single) {
              umma_f16_ts(tmem, a_hi + ks * 8, bh, idesc, (it | ks) != 0);
            } else {
              umma_f16_ts(tmem, a_lo + ks * 8, bh, idesc, (it | ks) != 0);   // small terms first
              umma_f16_ts(tmem, a_hi + ks * 8, bl, idesc, 1);
              umma_f16_ts(tmem, a_hi + ks * 8, bh, idesc, 1);
            }
          }
          umma_commit(smem_u32(&bars->empty_a[q]));              // each commit tracks every MMA issued so far
          umma_commit(smem_u32(&bars->empty_b[s]));
        }
        __syncwarp();
        a_hi += 32; b_hi += (uint32_t)(2 * B_BYTES) >> 4;
        if (++q == AT) { q = 0; a_hi = tmem_a; pa ^= 1; }
        if (++s == SB) { s = 0; b_hi = smem_u32(s_b) >> 4; pb ^= 1; }
      }
      if (elect_one()) umma_commit(smem_u32(&bars->acc));
      if (tr && lane == 0) p.trace[12] = clock64();                      // all MMAs issued
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) p.trace[13] = clock64();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// out[m][n] = act(bias[n] + sum_z part[z][m][n]) + res[m][n]
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int nsplit, long long M, int N,
                                     const float* __restrict__ bias, const float* __restrict__ res, long long ldr,
                                     int act, float* __restrict__ out, long long ldo) {
  pdl_trigger();
  pdl_wait();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const long long m = i / N;
  const int n = (int)(i % N);
  float s = 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(long long)z * M * N + i];
  if (bias) s += bias[n];
  if (act == 1) s = gelu_erf(s);
  if (res) s += res[m * ldr + n];
  out[m * ldo + n] = s;
}

// the same for N % 4 == 0 and 16-byte aligned rows: four columns per thread (the partials are summed in the same order z = 0, 1, ...
// per element, so the result is bit-identical to the scalar kernel's)
__global__ void splitk_reduce4_kernel(const float4* __restrict__ part, int nsplit, long long M, int N4,
                                      const float4* __restrict__ bias, const float* __restrict__ res, long long ldr,
                                      int act, float* __restrict__ out, long long ldo) {
  pdl_trigger();
  pdl_wait();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long MN4 = M * N4;
  if (i >= MN4) return;
  const long long m = i / N4;
  const int n4 = (int)(i % N4);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int z = 0; z < nsplit; ++z) {
    const float4 v = part[(long long)z * MN4 + i];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  if (bias) { const float4 b = __ldg(bias + n4); s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w; }
  if (act == 1) { s.x = gelu_erf(s.x); s.y = gelu_erf(s.y); s.z = gelu_erf(s.z); s.w = gelu_erf(s.w); }
  if (res) { const float4 r = *reinterpret_cast<const float4*>(res + m * ldr + 4 * n4); s.x += r.x; s.y += r.y; s.z += r.z; s.w += r.w; }
  *reinterpret_cast<float4*>(out + m * ldo + 4 * n4) = s;
}

// per 128-row tile: OR over rows of (nbr[row][t] >= 0) << t     (T <= 32)
__global__ void tile_mask_kernel(const int32_t* __restrict__ nbr, long long M, int T, uint32_t* __restrict__ mask) {
  const int tile = blockIdx.x;
  const long long m = (long long)tile * BM + threadIdx.x;
  uint32_t b = 0;
  if (m < M)
    for (int t = 0; t < T; ++t) b |= (nbr[m * T + t] >= 0 ? 1u : 0u) << t;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b |= __shfl_xor_sync(0xffffffffu, b, o);
  __shared__ uint32_t w[BM / 32];
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) mask[tile] = w[0] | w[1] | w[2] | w[3];
}

// pack W [T][K][N] (fp32, row-major: the tap-major transposed conv weight, or weight^T of a Linear) into
// Bp [T][ceil(K/KC)][ntiles][2][NT*KC] fp16: K-major core-matrix tiles (8 n x 8 k), hi block then lo block; K padded with 0
__global__ void pack_b_kernel(const float* __restrict__ W, int T, int K, int N, __half* __restrict__ Bp) {
  const int ntiles = (N + NT - 1) / NT, kch = (K + KC - 1) / KC;
  const long long total = (long long)T * kch * ntiles * NT * KC;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  // destination element order inside a block: (n/8, k/8, n%8, k%8)
  const int e = (int)(i % (NT * KC));
  const long long blk = i / (NT * KC);
  const int tn = (int)(blk % ntiles);
  const int kc = (int)((blk / ntiles) % kch);
  const int t = (int)(blk / ((long long)ntiles * kch));
  const int k8 = e % 8, n8 = (e / 8) % 8, kg = (e / 64) % (KC / 8), ng = e / (64 * (KC / 8));
  const int n = tn * NT + ng * 8 + n8, k = kc * KC + kg * 8 + k8;
  const float x = (n < N && k < K) ? W[((long long)t * K + k) * N + n] : 0.f;
  const __half h = __float2half_rn(x);
  __half* dst = Bp + blk * 2 * (NT * KC);
  dst[e] = h;
  dst[NT * KC + e] = __float2half_rn(x - __half2float(h));
}

// ---- skinny Linear on the FP32 pipes: out[M, N] = act(bias + A[M, K] . W) + res for K <= 64 and many rows --------------------------
// The level-0 linears around the blocks (grid-pool / unpool projections, proj_cat, heads: K = 32 / 64, N = 6 .. 64 (200), 120 000+ rows)
// move 45 MB for 0.5 GFLOP: pure HBM work.  Through the tcgen05 tile kernel they took 80 us each (12 % of the HBM roofline: one
// 128-row tile per CTA, prologue + TMEM staging + epilogue per 8 KB of input); this kernel streams them at HBM speed with exact
// fp32 FMAs.  W is decoded from the same packed (hi | lo) operand blocks the tensor-core path uses, so callers do not change.
// CTA = 256 threads = SK_ROWS rows x 64 columns; thread = SK_ROWS / 16 rows x 4 columns.
constexpr int SK_COLS = 64;
template <int K, int SK_ROWS>        // SK_ROWS = 128 (K = 32) or 64 (K = 64): x tile + W chunk stay under 48 KB of static shared memory
__global__ void __launch_bounds__(256) skinny_linear_kernel(const float* __restrict__ A, long long lda, const __half* __restrict__ Bp, int M,
                                                            int N, const float* __restrict__ bias, const float* __restrict__ res,
                                                            long long ldr, int act, float* __restrict__ out, long long ldo) {
  __shared__ float xs[SK_ROWS][K + 1];
  __shared__ __align__(16) float ws[K][SK_COLS];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * SK_ROWS;
  const int n0 = blockIdx.y * SK_COLS;
  // W chunk: packed blocks [k / KC][n / NT][hi | lo][(n%NT)/8][(k%KC)/8][n%8][k%8]
  const int ntiles = (N + NT - 1) / NT;
  for (int i = tid; i < K * SK_COLS; i += 256) {
    const int k = i / SK_COLS, c = i % SK_COLS, n = n0 + c;
    float w = 0.f;
    if (n < N) {
      const __half* blk = Bp + ((long long)(k / KC) * ntiles + n / NT) * 2 * (NT * KC);
      const int e = ((n % NT) / 8) * (64 * (KC / 8)) + ((k % KC) / 8) * 64 + (n % 8) * 8 + (k % 8);
      w = __half2float(blk[e]) + __half2float(blk[NT * KC + e]);
    }
    ws[k][c] = w;
  }
  pdl_trigger();
  pdl_wait();                      // the weight decode above reads parameters only; A / res / out come from the stream's earlier kernels
  for (int i = tid; i < SK_ROWS * (K / 4); i += 256) {
    const int r = i / (K / 4), q = i % (K / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + r < M) v = __ldcs(reinterpret_cast<const float4*>(A + (m0 + r) * lda) + q);
    xs[r][4 * q] = v.x; xs[r][4 * q + 1] = v.y; xs[r][4 * q + 2] = v.z; xs[r][4 * q + 3] = v.w;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  constexpr int RT = SK_ROWS / 16;
  float acc[RT][4];
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    const float4 b = *reinterpret_cast<const float4*>(&ws[k][tx * 4]);
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      const float a = xs[ty * RT + i][k];
      acc[i][0] = fmaf(a, b.x, acc[i][0]); acc[i][1] = fmaf(a, b.y, acc[i][1]);
      acc[i][2] = fmaf(a, b.z, acc[i][2]); acc[i][3] = fmaf(a, b.w, acc[i][3]);
    }
  }
  const int c = n0 + tx * 4;
  if (c >= N) return;
  float b4[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < N) b4[j] = bias[c + j];
  const bool vec = c + 3 < N && !(ldo & 3) && !((uintptr_t)out & 15) && (!res || (!(ldr & 3) && !((uintptr_t)res & 15)));
#pragma unroll
  for (int i = 0; i < RT; ++i) {
    const long long m = m0 + ty * RT + i;
    if (m >= M) break;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = acc[i][j] + b4[j]; if (act == 1) v[j] = gelu_erf(v[j]); }
    if (vec) {
      if (res) { const float4 r4 = *reinterpret_cast<const float4*>(res + m * ldr + c); v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w; }
      __stcs(reinterpret_cast<float4*>(out + m * ldo + c), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < N) out[m * ldo + c + j] = v[j] + (res ? res[m * ldr + c + j] : 0.f);
    }
  }
}

}  // namespace gt

CDSEG_API size_t cdseg_gemm_packed_b_floats(int T, int K, int N) {   // size in 4-byte units (the blocks hold fp16 pairs)
  return (size_t)T * ((K + gt::KC - 1) / gt::KC) * ((N + gt::NT - 1) / gt::NT) * gt::NT * gt::KC;
}

// W: fp32 [T][K][N] -> Bp (cdseg_gemm_packed_b_floats 4-byte units).  K % 16 == 0.
CDSEG_API int cdseg_gemm_pack_b(const float* W, int T, int K, int N, float* Bp, void* stream) {
  if (T <= 0 || K <= 0 || (K % 16) || N <= 0) return CDSEG_EINVAL;
  const long long total = (long long)T * ((K + gt::KC - 1) / gt::KC) * ((N + gt::NT - 1) / gt::NT) * gt::NT * gt::KC;
  gt::pack_b_kernel<<<cdseg_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(W, T, K, N, reinterpret_cast<__half*>(Bp));
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// mask: uint32 [ceil(M/128)]
CDSEG_API int cdseg_tile_tap_mask(const int32_t* nbr, int64_t M, int T, uint32_t* mask, void* stream) {
  if (T <= 0 || T > 32) return CDSEG_EINVAL;
  if (M == 0) return CDSEG_OK;
  gt::tile_mask_kernel<<<cdseg_div_up(M, gt::BM), gt::BM, 0, (cudaStream_t)stream>>>(nbr, M, T, mask);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// A/B switch, OFF by default: measured on the full step (profiles/r01f_narrow_tiles.md) narrow tiles lose to split-K + reduce
// (12.01 vs 11.76 ms): every CTA re-reads the whole A row block and N = 32 MMAs run at a third of the N = 128 rate
static bool g_narrow = [] { const char* e = getenv("CDSEG_GEMM_NARROW"); return e && atoi(e) != 0; }();
CDSEG_API void cdseg_gemm_tc_set_narrow(int on) { g_narrow = on != 0; }
static long long* g_trace = nullptr;
static int g_trace_cta = 0;
// Dense-layer precision of the whole library (gemm_tc, fused pre / post kernels, stem): 0 = fp32-faithful (x = hi + lo fp16 halves,
// three MMAs per product term), 1 = fp16 operands with fp32 accumulation (one MMA; the numerics of the reference's autocast training
// and of BASELINE.json's reduced-precision configs).
int g_cdseg_gemm_single = [] { const char* e = getenv("CDSEG_GEMM_FP16"); return e && atoi(e) != 0 ? 1 : 0; }();
CDSEG_API void cdseg_set_gemm_precision(int fp16_single) { g_cdseg_gemm_single = fp16_single ? 1 : 0; }
CDSEG_API int cdseg_get_gemm_precision(void) { return g_cdseg_gemm_single; }
// profiling hook: clock64 stamps of CTA (cta,0,0) of subsequent launches into a device buffer of >= 16 int64; NULL disables
CDSEG_API void cdseg_gemm_tc_set_trace(long long* buf, int cta) { g_trace = buf; g_trace_cta = cta; }

CDSEG_API size_t cdseg_gemm_tc_workspace_bytes(int64_t M, int N, int nsplit) {
  return nsplit > 1 ? (size_t)nsplit * M * N * sizeof(float) : 0;
}

// out[M,N] = act(bias + sum_t A[idx[:,t]] @ W_t) + res   (see include/cdseg_b200.h)
static int gemm_tc_launch(const float* A, int64_t lda, const int32_t* idx, int T, const uint32_t* tile_mask,
                          const float* Bp, int64_t M, int N, int K, const float* bias, const float* res, int64_t ldr,
                          int act, float* out, int64_t ldo, int nsplit, void* workspace, size_t workspace_bytes,
                          void* stream, int sub, int taps_ld) {
  cudaStream_t st = (cudaStream_t)stream;
  if (M < 0 || N <= 0 || K <= 0 || (K % 16) || (lda & 3) || T <= 0 ||
      nsplit < 1 || nsplit > T || (tile_mask && T > 32))
    return CDSEG_EINVAL;
  if (M == 0) return CDSEG_OK;
  if (workspace_bytes < cdseg_gemm_tc_workspace_bytes(M, N, nsplit)) return CDSEG_ENOSPC;
  // skinny Linear (K = 32 / 64, >= 16 384 rows): HBM-bound, served by the FP32-pipe streaming kernel (exact fp32 products)
  static const bool no_skinny = [] { const char* e = getenv("CDSEG_NO_SKINNY"); return e && atoi(e) != 0; }();
  if (!idx && T == 1 && !sub && nsplit == 1 && out && (K == 32 || K == 64) && M >= 16384 && !no_skinny && !g_cdseg_gemm_single) {
    dim3 g(cdseg_div_up(M, K == 32 ? 128 : 64), cdseg_div_up(N, gt::SK_COLS));
    if (K == 32) cdseg_launch_pdl(gt::skinny_linear_kernel<32, 128>, g, dim3(256), 0, st, A, (long long)lda, reinterpret_cast<const __half*>(Bp), (int)M, N, bias, res, (long long)ldr, act, out, (long long)ldo);
    else cdseg_launch_pdl(gt::skinny_linear_kernel<64, 64>, g, dim3(256), 0, st, A, (long long)lda, reinterpret_cast<const __half*>(Bp), (int)M, N, bias, res, (long long)ldr, act, out, (long long)ldo);
    CDSEG_COUNT_LAUNCH(1);
    CDSEG_LAUNCH_CHECK();
    return CDSEG_OK;
  }
  // Experiment (CDSEG_GEMM_NARROW=1): a dense Linear whose caller asked for a K split (few row tiles, K >= 256) served by
  // narrower output tiles instead -- the same number of CTAs, no partial sums, no reduce launch.
  // (T, K) chunks of a contiguous row are the same operand as (1, T * K): the packed weight order [t][kc][tile] is unchanged.
  int nw = gt::NT;
  if (!idx && !sub && nsplit > 1 && (long long)T * K <= 1024 && (N % 32) == 0 && g_narrow) {
    K *= T; T = 1; nsplit = 1;
    const long long tm = (M + gt::BM - 1) / gt::BM;
    while (nw > 32 && tm * ((N + nw - 1) / nw) < 120) nw >>= 1;
  }
  const int un_max = N >= nw ? nw : ((N + 15) & ~15);
  const int b_bytes = un_max * gt::KC * 2;
  const int iters = (int)((long long)T * ((K + gt::KC - 1) / gt::KC) / nsplit);      // upper bound of k-iterations per CTA
  // TMEM: accumulator columns + A ring (32 columns per slot); allocation must be a power of two >= 32
  const int acc_cols = un_max <= 32 ? 32 : (un_max <= 64 ? 64 : 128);
  const int AT = acc_cols == 128 ? 4 : ((iters >= 4 && acc_cols <= 32) ? 3 : 2);   // 32+96=128, 64+64=128, 128+128=256 columns
  int tmem_cols = 32;
  while (tmem_cols < acc_cols + AT * 32) tmem_cols <<= 1;
  // B ring depth: a stage is 2 * b_bytes (4 KB at 32 columns, 16 KB at 128).  Narrow tiles take the whole k-loop in flight: with 4 stages
  // the MMA thread waited ~1 200 cycles per iteration for weights while the producers delivered A every 700 (r02_trace_gemm_deep_amode.txt)
  const int SB = iters >= 4 ? (un_max <= 32 ? (iters >= 8 ? 8 : 4) : (un_max <= 64 ? 4 : (iters >= 8 ? 4 : 3))) : 2;
  const size_t smem = (size_t)gt::STG_BYTES + (size_t)SB * 2 * b_bytes + sizeof(gt::Bars) + 64 +
                      ((idx && T <= 32 && !sub) ? (size_t)gt::BM * T * 4 : 0) + 1024;
  // two register budgets: 3+ CTAs per SM (<= 112 registers, a few spills) when TMEM and shared memory allow that many,
  // otherwise the spill-free 2-per-SM build.  CDSEG_GEMM_MINB=2|3 forces one for experiments.
  static const int forced = [] { const char* e = getenv("CDSEG_GEMM_MINB"); return e ? atoi(e) : 0; }();
  const bool dense3 = forced ? forced == 3 : (tmem_cols <= 128 && smem * 3 <= 220 * 1024);
  if (!idx && tile_mask) return CDSEG_EINVAL;
  const int amode = sub ? 2 : (!idx ? 0 : (T <= 32 ? 1 : 3));      // see the kernel's producer section
  typedef void (*KernelFn)(const gt::Params);
  static const KernelFn kernels[2][4] = {
      {gt::gemm_tc_kernel<2, 0>, gt::gemm_tc_kernel<2, 1>, gt::gemm_tc_kernel<2, 2>, gt::gemm_tc_kernel<2, 3>},
      {gt::gemm_tc_kernel<3, 0>, gt::gemm_tc_kernel<3, 1>, gt::gemm_tc_kernel<3, 2>, gt::gemm_tc_kernel<3, 3>}};
  const KernelFn kernel = kernels[dense3][amode];
  static size_t configured[2][4] = {};
  if (smem > configured[dense3][amode]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured[dense3][amode] = smem;
  }
  gt::Params p;
  p.A = A; p.lda = lda; p.idx = idx; p.T = T; p.tile_mask = tile_mask; p.Bp = Bp; p.sub = sub; p.taps_ld = taps_ld;
  p.M = (int)M; p.N = N; p.K = K; p.bias = bias; p.res = res; p.ldr = ldr; p.act = act; p.out = out; p.ldo = ldo;
  p.part = (float*)workspace; p.nsplit = nsplit; p.nw = nw;
  p.AT = AT; p.SB = SB; p.b_bytes = b_bytes; p.acc_cols = acc_cols; p.tmem_cols = tmem_cols;
  p.trace = g_trace; p.trace_cta = g_trace_cta;
  p.single = g_cdseg_gemm_single;
  static const bool no256 = [] { const char* e = getenv("CDSEG_NO_LDG256"); return e && atoi(e) != 0; }();
  p.a256 = (!no256 && ((uintptr_t)A & 31) == 0 && (lda & 7) == 0 && (K & 7) == 0) ? 1 : 0;
  static const bool no_fast = [] { const char* e = getenv("CDSEG_NO_FAST_EPI"); return e && atoi(e) != 0; }();
  p.fast_epi = 0;                                  // set below, once vec_ok is known
  p.vec_ok = ((ldo & 3) == 0 && (!res || (ldr & 3) == 0) && (N & 3) == 0 && ((uintptr_t)out & 15) == 0 &&
              (!res || ((uintptr_t)res & 15) == 0)) ? 1 : 0;
  p.fast_epi = (!no_fast && p.vec_ok && (nw & 31) == 0 && !(((uintptr_t)bias | (uintptr_t)workspace) & 15)) ? 1 : 0;
  dim3 g(cdseg_div_up(M, gt::BM), (N + nw - 1) / nw, nsplit);
  cdseg_launch_pdl(kernel, g, dim3(gt::NTHREADS), smem, st, p);
  CDSEG_COUNT_LAUNCH(1);
  if (nsplit > 1 && out) {                   // out == NULL: the caller consumes the raw partials part[z][M][N] itself (cdseg_reduce_ln)
    const bool v4 = (N & 3) == 0 && (ldo & 3) == 0 && (!res || (ldr & 3) == 0) &&
                    !(((uintptr_t)workspace | (uintptr_t)bias | (uintptr_t)res | (uintptr_t)out) & 15);
    if (v4)
      cdseg_launch_pdl(gt::splitk_reduce4_kernel, dim3(cdseg_div_up(M * (N / 4), 256)), dim3(256), 0, st, (const float4*)workspace, nsplit,
                       (long long)M, N / 4, (const float4*)bias, res, (long long)ldr, act, out, (long long)ldo);
    else
      cdseg_launch_pdl(gt::splitk_reduce_kernel, dim3(cdseg_div_up(M * N, 256)), dim3(256), 0, st, (const float*)workspace, nsplit,
                       (long long)M, N, bias, res, (long long)ldr, act, out, (long long)ldo);
    CDSEG_COUNT_LAUNCH(1);
  }
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_gemm_tc(const float* A, int64_t lda, const int32_t* idx, int T, const uint32_t* tile_mask,
                            const float* Bp, int64_t M, int N, int K, const float* bias, const float* res, int64_t ldr,
                            int act, float* out, int64_t ldo, int nsplit, void* workspace, size_t workspace_bytes,
                            void* stream) {
  return gemm_tc_launch(A, lda, idx, T, tile_mask, Bp, M, N, K, bias, res, ldr, act, out, ldo, nsplit, workspace,
                        workspace_bytes, stream, 0, 0);
}

// Sparse conv with a tiny C_in (the k=5 stem, 6 input channels) as an im2col GEMM on the tensor cores:
// A8 [rows, 8] = input padded to 8 channels, nbr [M, taps], Bp = cdseg_gemm_pack_b of [ceil(taps/4)][32][N] where row
// 8*q + c of block t holds the weight of tap 4t+q, channel c (zero for c >= C_in or tap >= taps).
// out = act(bias + conv) -- eval-BatchNorm folded into Bp / bias by the caller (see include/cdseg_b200.h).
CDSEG_API int cdseg_conv_im2col_tc(const float* A8, const int32_t* nbr, int taps, const float* Bp, int64_t M, int N,
                                   const float* bias, int act, float* out, int64_t ldo, void* stream) {
  if (!nbr || taps <= 0) return CDSEG_EINVAL;
  return gemm_tc_launch(A8, 8, nbr, (taps + 3) / 4, nullptr, Bp, M, N, 32, bias, nullptr, 0, act, out, ldo, 1, nullptr, 0,
                        stream, 4, taps);
}
