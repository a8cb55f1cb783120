// Native executor for one PTv3 Block (ptv3.py:399-428 of the reference: cpe conv -> Linear -> LayerNorm -> residual
// [-> + t_mlp(t_emb)] -> norm1 -> serialized attention -> residual -> norm2 -> MLP -> residual).
//
// The per-block schedule is a fixed sequence of 12-13 launches of kernels that already live in this library; issuing
// them from Python costs ~14 us per launch (ctypes marshalling, torch allocations, stream queries) and left the forward
// host-bound (profiles/r01_host_overhead.txt: 14.4 ms of enqueue per 19.5 ms step).  This entry point takes one struct of
// raw pointers and a caller-provided scratch arena and enqueues the whole block from C++.
#include "common.cuh"
#include "../../include/cdseg_b200.h"

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static int pick_split(int64_t tiles, int T) {
  if (tiles >= 120 || T == 1) return 1;
  if (T == 27) {
    // 27-tap convs at the deep levels: CTAs per launch = tiles * s on 2 x 148 resident slots.  The divisor rule below picked s = 9 for
    // 48 tiles (432 CTAs = 1.46 waves of 3 taps); an uneven split that fills ONE wave (s = 6: 288 CTAs of 4-5 taps) is shorter.
    // cost(s) = waves * (taps per CTA + 1 tap-time of prologue / epilogue)
    int best = 1;
    long long best_cost = -1;
    for (int s = 1; s <= T; ++s) {
      const long long waves = (tiles * s + 295) / 296, cost = waves * ((T + s - 1) / s + 1);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
    }
    return best;
  }
  const int want = (int)((200 + tiles - 1) / tiles);
  for (int s = 1; s <= T; ++s)
    if (T % s == 0 && s >= want) return s;
  return T;
}

CDSEG_API int cdseg_debug_pick_split_block(int64_t tiles, int T) { return pick_split(tiles, T); }   // see net_exec.cu

struct Gemm { const float* Bp; const float* bias; };

static int g_fused_mask = 0x7fffffff;
CDSEG_API void cdseg_set_fused_mask(int mask) { g_fused_mask = mask; }
// resident CTAs per SM of the persistent fused kernels launched from now on (0 = as many as fit).  cdseg_net_forward sets 1 around the
// low-priority branch: its persistent CTAs then leave half of every SM to the kernels of the critical stream.
int g_cdseg_fused_per_sm = 0;
CDSEG_API void cdseg_set_fused_ctas_per_sm(int n) { g_cdseg_fused_per_sm = n; }

// one Linear through cdseg_gemm_tc with the same split heuristic as the Python path (cdsegnet_b200/ptv3.py::linear)
static int run_linear(const float* x, int64_t n, int K, int N, const float* Bp, const float* bias, const float* res, int act,
                      float* out, void* ws, size_t ws_bytes, void* stream) {
  const int64_t tiles = ((n + 127) / 128) * ((N + 127) / 128);
  int T = 1;
  if (tiles < 120 && K >= 256) T = K / 64;
  const int ns = pick_split(tiles, T);
  return cdseg_gemm_tc(x, K, nullptr, T, nullptr, Bp, n, N, K / T, bias, res, N, act, out, N, ns, ws, ws_bytes, stream);
}

// a Linear whose split-K partials (if any) are left for cdseg_reduce_ln: returns the status, *ns_out = number of partials in ws
// (1: `out` holds the finished result, bias included)
static int run_linear_raw(const float* x, int64_t n, int K, int N, const float* Bp, const float* bias, float* out, void* ws,
                          size_t ws_bytes, void* stream, int* ns_out) {
  const int64_t tiles = ((n + 127) / 128) * ((N + 127) / 128);
  int T = 1;
  if (tiles < 120 && K >= 256) T = K / 64;
  const int ns = pick_split(tiles, T);
  *ns_out = ns;
  if (ns > 1 && cdseg_gemm_tc_workspace_bytes(n, N, ns) > ws_bytes) return CDSEG_ENOSPC;
  return cdseg_gemm_tc(x, K, nullptr, T, nullptr, Bp, n, N, K / T, bias, nullptr, N, 0, ns > 1 ? nullptr : out, N, ns, ws, ws_bytes, stream);
}

CDSEG_API size_t cdseg_block_scratch_bytes(int64_t n, int C, int H, int T, int Kp, int B) {
  const size_t row = align_up((size_t)n * C * 4);
  size_t s = 0;
  s += 6 * row;                                   // y1, y2, x1, h, o, a  (a doubles as x2-input)
  s += align_up((size_t)n * 3 * C * 4);           // qkv
  s += align_up((size_t)n * 4 * C * 4);           // hidden
  s += 7 * align_up((size_t)H * T * Kp * 16 * 2); // packed q, k (16 wide) and v (32 wide: [v | 1 | 0..]) in fp16; CDSEG_ATTN_EXACT packs fp32 q, k, v
                                                  // (3 x 2 units); CDSEG_ATTN_TC32 packs q / k as fp16 hi | lo (2 + 2) and v 48 wide (3): 7 units
  s += align_up((size_t)B * C * 4);               // t projection
  s += (size_t)32 << 20;                          // split-K partials: only launches with < 120 output tiles split, so
                                                  // nsplit * M * N * 4 B stays below ~26 MB (see pick_split)
  return s + 4096;
}

CDSEG_API int cdseg_block_forward(const CdsegBlockArgs* a, void* stream) {
  if (!a || a->n <= 0 || a->C <= 0 || (a->C % 16) || a->H * 16 != a->C) return CDSEG_EINVAL;
  const int64_t n = a->n;
  const int C = a->C;
  char* p = (char*)a->scratch;
  char* const end = p + a->scratch_bytes;
  auto take = [&](size_t bytes) { char* q = p; p += align_up(bytes); return q; };
  float* y1 = (float*)take((size_t)n * C * 4);
  float* y2 = (float*)take((size_t)n * C * 4);
  float* x1 = (float*)take((size_t)n * C * 4);
  float* h = (float*)take((size_t)n * C * 4);
  float* o = (float*)take((size_t)n * C * 4);
  float* att = (float*)take((size_t)n * C * 4);
  float* qkv = (float*)take((size_t)n * 3 * C * 4);
  float* hid = (float*)take((size_t)n * 4 * C * 4);
  const size_t pk = (size_t)a->H * a->T * a->Kp * 16 * 2;
  const int am = a->attn_mode;
  if (am != CDSEG_ATTN_F16 && am != CDSEG_ATTN_EXACT && am != CDSEG_ATTN_TC32) return CDSEG_EINVAL;
  const size_t qk_units = am == CDSEG_ATTN_F16 ? 1 : 2, v_units = am == CDSEG_ATTN_TC32 ? 3 : 2;
  void* qp = take(qk_units * pk); void* kp = take(qk_units * pk); void* vp = take(v_units * pk);
  float* tproj = (float*)take((size_t)a->B * C * 4);
  if (p > end) return CDSEG_ENOSPC;
  void* ws = p;
  const size_t ws_bytes = (size_t)(end - p);
  int st;
#define RUN(call) do { st = (call); if (st != CDSEG_OK) return st; } while (0)
  const bool fused_c = C == 32 || C == 64 || C == 128;
  if ((g_fused_mask & 2) && fused_c && a->conv_plan) {
    // cpe conv + Linear + LayerNorm + residual (+ t) + norm1 + qkv: one kernel (fused_pre.cu)
    const float* tp = nullptr;
    if (a->t_scene) {
      RUN(cdseg_small_linear(a->t_scene, a->t_W, a->t_b, 0, a->B, a->T_dim, C, tproj, stream));
      tp = tproj;
    }
    if (a->ev[4]) cudaEventRecord((cudaEvent_t)a->ev[4], (cudaStream_t)stream);
    RUN(cdseg_pre_attn(a->conv_in ? a->conv_in : a->x, a->x, n, C, a->nbr, a->tile_mask, a->conv_plan, a->conv_Bp, a->conv_b, a->lin_Bp,
                       a->lin_b, a->cpe_g, a->cpe_b, tp, tp ? a->batch : nullptr, a->n1_g, a->n1_b, a->ln_eps, a->qkv_Bp, a->qkv_b, x1,
                       qkv, stream));
    if (a->ev[5]) cudaEventRecord((cudaEvent_t)a->ev[5], (cudaStream_t)stream);
  } else {
  // cpe: conv (implicit GEMM over 27 taps) -> Linear -> LayerNorm
  {
    const int64_t tiles = ((n + 127) / 128) * ((C + 127) / 128);
    const int ns = pick_split(tiles, 27);
    if (cdseg_gemm_tc_workspace_bytes(n, C, ns) > ws_bytes) return CDSEG_ENOSPC;
    if (a->ev[4]) cudaEventRecord((cudaEvent_t)a->ev[4], (cudaStream_t)stream);
    RUN(cdseg_gemm_tc(a->conv_in ? a->conv_in : a->x, C, a->nbr, 27, a->tile_mask, a->conv_Bp, n, C, C, a->conv_b, nullptr, 0, 0, y1, C,
                      ns, ws, ws_bytes, stream));
    if (a->ev[5]) cudaEventRecord((cudaEvent_t)a->ev[5], (cudaStream_t)stream);
  }
  const float* tp = nullptr;
  if (a->t_scene) {
    RUN(cdseg_small_linear(a->t_scene, a->t_W, a->t_b, 0, a->B, a->T_dim, C, tproj, stream));
    tp = tproj;
  }
  if (g_fused_mask & 4) {
    // Linear (partials left unreduced) -> ONE kernel: reduce + bias + LayerNorm_cpe + residual (+ t) + norm1
    int ns = 1;
    RUN(run_linear_raw(y1, n, C, C, a->lin_Bp, a->lin_b, y2, ws, ws_bytes, stream, &ns));
    RUN(cdseg_reduce_ln(ns > 1 ? (const float*)ws : y2, ns, ns > 1 ? a->lin_b : nullptr, a->cpe_g, a->cpe_b, a->x, tp, tp ? a->batch : nullptr,
                        a->n1_g, a->n1_b, a->ln_eps, n, C, x1, h, stream));
  } else {
  RUN(run_linear(y1, n, C, C, a->lin_Bp, a->lin_b, nullptr, 0, y2, ws, ws_bytes, stream));
  RUN(cdseg_add_layernorm(y2, nullptr, nullptr, nullptr, a->cpe_g, a->cpe_b, a->ln_eps, n, C, nullptr, y1, stream));
  // residual (+ per-scene timestep projection) + norm1
  RUN(cdseg_add_layernorm(a->x, y1, tp, tp ? a->batch : nullptr, a->n1_g, a->n1_b, a->ln_eps, n, C, x1, h, stream));
  }
  // attention
  RUN(run_linear(h, n, C, 3 * C, a->qkv_Bp, a->qkv_b, nullptr, 0, qkv, ws, ws_bytes, stream));
  }
  if (am == CDSEG_ATTN_EXACT) {       // dense-branch numerics (ptv3.py:264-280): fp32 operands, SIMT kernel
    RUN(cdseg_attn_pack_f32(qkv, 3 * C, 0, C, 3, a->slot_src, a->H, a->T, a->Kp, (float*)qp, (float*)kp, (float*)vp, stream));
    if (a->ev[0]) cudaEventRecord((cudaEvent_t)a->ev[0], (cudaStream_t)stream);
    RUN(cdseg_attn_exact((const float*)qp, (const float*)kp, (const float*)vp, a->patch_len, a->slot_dst, a->H, a->T, a->Kp, a->scale, o, C,
                         stream));
  } else if (am == CDSEG_ATTN_TC32) { // tcgen05 with hi/lo-split operands: fp32-class results at tensor-core speed
    RUN(cdseg_attn_pack_split(qkv, 3 * C, 0, C, 3, a->slot_src, a->H, a->T, a->Kp, qp, kp, vp, 1, stream));
    if (a->ev[0]) cudaEventRecord((cudaEvent_t)a->ev[0], (cudaStream_t)stream);
    RUN(cdseg_attn_tc3(qp, kp, vp, a->patch_len, a->slot_dst, a->H, a->T, a->Kp, a->scale, 1, o, C, stream));
  } else {                            // flash-branch numerics (ptv3.py:282-289): fp16 operands / probabilities / output
    RUN(cdseg_attn_pack_f16v(qkv, 3 * C, 0, C, 3, a->slot_src, a->H, a->T, a->Kp, qp, kp, vp, 1, stream));
    if (a->ev[0]) cudaEventRecord((cudaEvent_t)a->ev[0], (cudaStream_t)stream);
    RUN(cdseg_attn_tc3(qp, kp, vp, a->patch_len, a->slot_dst, a->H, a->T, a->Kp, a->scale, 0, o, C, stream));
  }
  if (a->ev[1]) cudaEventRecord((cudaEvent_t)a->ev[1], (cudaStream_t)stream);
  if ((g_fused_mask & 1) && fused_c) {
    // proj + residual + norm2 + MLP + residual: one kernel, intermediates in tensor memory (fused_post.cu)
    if (a->ev[2]) cudaEventRecord((cudaEvent_t)a->ev[2], (cudaStream_t)stream);
    RUN(cdseg_post_attn(o, x1, n, C, a->proj_Bp, a->proj_b, a->n2_g, a->n2_b, a->ln_eps, a->fc1_Bp, a->fc1_b, a->fc2_Bp,
                        a->fc2_b, a->out, stream));
    if (a->ev[3]) cudaEventRecord((cudaEvent_t)a->ev[3], (cudaStream_t)stream);
    return CDSEG_OK;
  }
  if (g_fused_mask & 4) {
    // proj (partials left unreduced) -> ONE kernel: reduce + bias + residual + norm2
    int ns = 1;
    RUN(run_linear_raw(o, n, C, C, a->proj_Bp, a->proj_b, att, ws, ws_bytes, stream, &ns));
    RUN(cdseg_reduce_ln(ns > 1 ? (const float*)ws : att, ns, ns > 1 ? a->proj_b : nullptr, nullptr, nullptr, x1, nullptr, nullptr, a->n2_g,
                        a->n2_b, a->ln_eps, n, C, y2, h, stream));
  } else {
  RUN(run_linear(o, n, C, C, a->proj_Bp, a->proj_b, nullptr, 0, att, ws, ws_bytes, stream));
  // residual + norm2 + MLP (fc1+GELU, fc2+residual fused in the GEMM epilogues)
  RUN(cdseg_add_layernorm(x1, att, nullptr, nullptr, a->n2_g, a->n2_b, a->ln_eps, n, C, y2, h, stream));
  }
  if (a->ev[2]) cudaEventRecord((cudaEvent_t)a->ev[2], (cudaStream_t)stream);
  RUN(run_linear(h, n, C, 4 * C, a->fc1_Bp, a->fc1_b, nullptr, 1, hid, ws, ws_bytes, stream));
  if (a->ev[3]) cudaEventRecord((cudaEvent_t)a->ev[3], (cudaStream_t)stream);
  RUN(run_linear(hid, n, 4 * C, C, a->fc2_Bp, a->fc2_b, y2, 0, a->out, ws, ws_bytes, stream));
#undef RUN
  return CDSEG_OK;
}

// CUDA events for bench.py: kernel durations measured live, on the launching stream, inside the timed region
CDSEG_API void* cdseg_event_create(void) {
  cudaEvent_t e = nullptr;
  return cudaEventCreate(&e) == cudaSuccess ? (void*)e : nullptr;
}
CDSEG_API void cdseg_event_destroy(void* e) { if (e) cudaEventDestroy((cudaEvent_t)e); }
CDSEG_API int cdseg_event_elapsed_ms(void* e0, void* e1, float* ms) { return (int)cudaEventElapsedTime(ms, (cudaEvent_t)e0, (cudaEvent_t)e1); }
