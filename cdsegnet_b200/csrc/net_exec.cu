// Native executor of the feature phase of one CDSegNet forward (PointTransformerV3.forward, point_transformer_v3m1_base.py:1757-1845,
// everything after serialization): Embedding stems, interleaved encoder stages of the Conditional Network (code prefix _n_) and the
// Noise Network (_c_, on a second stream), TransferModule (CrossBlock, :1179-1223), decoders (SerializedUnpooling :601-630 + Blocks),
// heads, and the gathers between the caller's numbering and the internal curve-order numbering.
//
// Why: the forward is ~450 kernel launches; issued through ~120 Python -> ctypes transitions it cost 6.6 ms of host time per
// 11.6 ms step and the deep levels (15 us kernels) ran host-bound (profiles/r02_host_overhead.txt).  This file enqueues the same
// kernels from C++ out of two caller-provided bump arenas (no allocator calls, no Python between launches).
// A dry run of the same code path (launch = false) sizes the arenas, so the sizing can never drift from the execution.
#include "common.cuh"
#include "../../include/cdseg_b200.h"
#include <cstring>
#include <cstdlib>
#include <cstdio>

namespace {

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

int pick_split(int64_t tiles, int T) {
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

int g_net_block_limit = [] { const char* e = getenv("CDSEG_NET_BLOCK_LIMIT"); return e ? atoi(e) : 0; }();
int g_side_per_sm = [] { const char* e = getenv("CDSEG_SIDE_CTAS"); return e ? atoi(e) : 0; }();
int g_net_debug = [] { const char* e = getenv("CDSEG_NET_DEBUG"); return e ? atoi(e) : 0; }();

struct Arena {
  char* base; size_t cap; size_t off; size_t high; bool dry;
  void* take(size_t bytes) {
    const size_t o = off;
    if (!dry && (g_net_debug & 4)) fprintf(stderr, "[net] arena %p alloc off=%zu bytes=%zu\n", (void*)base, o, bytes);
    off += al256(bytes);
    if (off > high) high = off;
    return dry ? (void*)(uintptr_t)(0x1000 + o) : (void*)(base + o);     // dry: fake non-NULL addresses, never dereferenced
  }
  bool ok() const { return dry || off <= cap; }
};

struct Ctx {
  const CdsegForwardArgs* a;
  bool dry;
  int status;
  cudaStream_t st;        // stream of the branch being enqueued
  Arena* ar;              // its arena
  void* scratch; size_t scratch_bytes;   // block scratch of this branch (sized for its largest block)
  void* ws; size_t ws_bytes;             // split-K workspace for the linears outside blocks
  int ev_index;
  int nblocks; bool skip;                // debug: stop enqueueing after g_net_block_limit blocks of this branch
};

#define RUN(call) do { if (!c.dry && !c.skip && c.status == CDSEG_OK) { c.status = (call); } } while (0)

// plan built with an aux stream: the indice tables of a level are complete once its `ready` event has fired (include/cdseg_b200.h)
inline void need(Ctx& c, void* ev) { if (!c.dry && ev) cudaStreamWaitEvent(c.st, (cudaEvent_t)ev, 0); }

// one Linear through cdseg_gemm_tc with the split heuristic of cdsegnet_b200/ptv3.py::linear / block_exec.cu::run_linear
void linear(Ctx& c, const float* x, int64_t n, const CdsegLinW& w, const float* res, int act, float* out) {
  const int64_t tiles = ((n + 127) / 128) * ((w.N + 127) / 128);
  int T = 1;
  if (tiles < 120 && w.K >= 256) T = w.K / 64;
  const int ns = pick_split(tiles, T);
  if (c.dry) {
    const size_t need = cdseg_gemm_tc_workspace_bytes(n, w.N, ns);
    if (need > c.ws_bytes) c.ws_bytes = need;
    return;
  }
  RUN(cdseg_gemm_tc(x, w.K, nullptr, T, nullptr, w.Bp, n, w.N, w.K / T, w.bias, res, w.N, act, out, w.N, ns, c.ws, c.ws_bytes, c.st));
}

void conv3(Ctx& c, const float* x, const CdsegPlanLevel& L, int C, const float* Bp, const float* bias, float* out) {
  const int64_t tiles = ((L.n + 127) / 128) * ((C + 127) / 128);
  const int ns = pick_split(tiles, 27);
  if (c.dry) {
    const size_t need = cdseg_gemm_tc_workspace_bytes(L.n, C, ns);
    if (need > c.ws_bytes) c.ws_bytes = need;
    return;
  }
  RUN(cdseg_gemm_tc(x, C, L.nbr3, 27, L.tile_mask3, Bp, L.n, C, C, bias, nullptr, 0, 0, out, C, ns, c.ws, c.ws_bytes, c.st));
}

float* f32(Ctx& c, int64_t n, int C) { return (float*)c.ar->take((size_t)n * C * 4); }

// Block (ptv3.py:399-428) through cdseg_block_forward
float* block(Ctx& c, const CdsegBlockW& w, const CdsegPlanLevel& L, const float* x, const float* conv_in, const float* t_scene) {
  const CdsegPatchMap& pm = L.pm[w.order_index];
  float* out = f32(c, L.n, w.C);
  if (c.dry) {
    const size_t need = cdseg_block_scratch_bytes(L.n, w.C, w.H, pm.T, pm.Kp, c.a->B);
    if (need > c.scratch_bytes) c.scratch_bytes = need;
    ++c.ev_index;
    return out;
  }
  if (pm.T <= 0 || !pm.slot_src) { if (c.status == CDSEG_OK) c.status = CDSEG_EINVAL; return out; }
  if (g_net_block_limit && ++c.nblocks > g_net_block_limit) c.skip = true;
  CdsegBlockArgs b;
  memset(&b, 0, sizeof(b));
  b.n = L.n; b.C = w.C; b.H = w.H; b.T_dim = w.T_dim > 0 ? w.T_dim : 0; b.B = c.a->B;
  b.x = x; b.conv_in = conv_in; b.nbr = L.nbr3; b.tile_mask = L.tile_mask3; b.batch = L.batch;
  b.conv_plan = (w.C == 32 || w.C == 64 || w.C == 128) ? L.conv_plan3 : nullptr;
  b.t_scene = (w.T_dim > 0) ? t_scene : nullptr;
  b.slot_src = pm.slot_src; b.slot_dst = pm.slot_dst; b.patch_len = pm.patch_len; b.T = pm.T; b.Kp = pm.Kp; b.scale = w.scale;
  b.conv_Bp = w.conv_Bp; b.conv_b = w.conv_b; b.lin_Bp = w.lin.Bp; b.lin_b = w.lin.bias; b.cpe_g = w.cpe_ln.g; b.cpe_b = w.cpe_ln.b;
  b.t_W = w.t_W; b.t_b = w.t_b; b.n1_g = w.n1.g; b.n1_b = w.n1.b; b.qkv_Bp = w.qkv.Bp; b.qkv_b = w.qkv.bias;
  b.proj_Bp = w.proj.Bp; b.proj_b = w.proj.bias; b.n2_g = w.n2.g; b.n2_b = w.n2.b; b.fc1_Bp = w.fc1.Bp; b.fc1_b = w.fc1.bias;
  b.fc2_Bp = w.fc2.Bp; b.fc2_b = w.fc2.bias; b.ln_eps = w.ln_eps; b.attn_mode = c.a->attn_mode;
  if ((g_net_debug & 8) && c.st == (cudaStream_t)c.a->stream_side) b.attn_mode = CDSEG_ATTN_EXACT;       // debug: SIMT attention on the side stream
  if ((g_net_debug & 16) && c.st == (cudaStream_t)c.a->stream_main) b.attn_mode = CDSEG_ATTN_EXACT;      // debug: SIMT attention on the main stream
  b.out = out; b.scratch = c.scratch; b.scratch_bytes = c.scratch_bytes;
  if (c.a->block_events)
    for (int i = 0; i < 6; ++i) b.ev[i] = c.a->block_events[(size_t)c.ev_index * 6 + i];
  ++c.ev_index;
  // low-priority branch next to the critical one: one persistent CTA per SM, so the critical stream's kernels always find room
  const bool yield = c.a->stream_side && c.st == (cudaStream_t)c.a->stream_side && c.st != (cudaStream_t)c.a->stream_main && g_side_per_sm > 0;
  if (yield) cdseg_set_fused_ctas_per_sm(g_side_per_sm);
  RUN(cdseg_block_forward(&b, c.st));
  if (yield) cdseg_set_fused_ctas_per_sm(0);
  return out;
}

// encoder stage: [SerializedPooling] + blocks.  x: features of the parent level (stage > 0) or of this level (stage 0)
float* enc_stage(Ctx& c, const CdsegStageW& s, const CdsegPlanLevel* lv, int base, const float* x, const float* t_scene) {
  const CdsegPlanLevel& L = lv[base + s.level];
  if (L.parent >= 0 && !c.dry) {                          // first pooled stage: the plan's deferred launches go out now (include/cdseg_b200.h)
    const int r = cdseg_plan_finish();
    if (r != CDSEG_OK && c.status == CDSEG_OK) c.status = r;
  }
  if (s.has_pool) {
    const CdsegPlanLevel& P = lv[L.parent];
    float* p = f32(c, P.n, s.pool.proj.N);
    linear(c, x, P.n, s.pool.proj, nullptr, 0, p);
    float* f = f32(c, L.n, s.pool.proj.N);
    // members = the parent's points in the order of the curve that defined the clusters (physical row c0)
    RUN(cdseg_pool_reduce(p, nullptr, P.order + (size_t)L.c0 * P.cap, L.idx_ptr, L.n, s.pool.proj.N, s.pool.bn_scale, s.pool.bn_shift, 1, f,
                          nullptr, c.st));
    x = f;
  }
  need(c, L.ready);
  for (int i = 0; i < s.n_blocks; ++i) x = block(c, s.blocks[i], L, x, nullptr, t_scene);
  return (float*)x;
}

// decoder stage: SerializedUnpooling (coarse x_up at level+1, skip x_skip at `level`) + blocks
float* dec_stage(Ctx& c, const CdsegStageW& s, const CdsegPlanLevel* lv, int base, const float* x_up, const float* x_skip,
                 const float* t_scene) {
  const CdsegPlanLevel& L = lv[base + s.level];          // fine level
  const CdsegPlanLevel& Cc = lv[base + s.level + 1];     // coarse level
  const CdsegUnpoolW& u = s.up;
  const int co = u.proj.N;
  float* up = f32(c, Cc.n, co);
  linear(c, x_up, Cc.n, u.proj, nullptr, 1, up);          // Linear + folded BN + GELU
  float* skip = f32(c, L.n, co);
  linear(c, x_skip, L.n, u.proj_skip, nullptr, 1, skip);
  float* feat = f32(c, L.n, co);
  if (!u.cat) {
    RUN(cdseg_unpool_add(skip, up, Cc.cluster, L.n, co, u.alpha, feat, c.st));
  } else {
    float* a = f32(c, L.n, co);
    float* b = f32(c, Cc.n, co);
    linear(c, skip, L.n, u.cat_a, nullptr, 0, a);
    linear(c, up, Cc.n, u.cat_b, nullptr, 0, b);
    RUN(cdseg_unpool_add(a, b, Cc.cluster, L.n, co, u.alpha, feat, c.st));
  }
  // reference quirk (ptv3.py:608-625): parent.sparse_conv_feat keeps the unscaled proj_skip output, so the FIRST block's CPE
  // convolves `skip`, not the fused features
  const float* x = feat;
  need(c, L.ready);
  for (int i = 0; i < s.n_blocks; ++i) x = block(c, s.blocks[i], L, x, i == 0 ? skip : nullptr, t_scene);
  return (float*)x;
}

float* stem(Ctx& c, const CdsegStemW& w, const CdsegPlanLevel& L0, const float* feat_caller) {
  float* x8 = f32(c, L0.n, 8);
  RUN(cdseg_gather_rows_pad(feat_caller, L0.perm, L0.n, w.cin, 8, x8, c.st));
  float* out = f32(c, L0.n, w.cout);
  const int taps = w.ksize * w.ksize * w.ksize;
  need(c, L0.ready_stem);
  RUN(cdseg_conv_im2col_tc(x8, L0.nbr_stem, taps, w.Bp, L0.n, w.cout, w.shift, 1, out, w.cout, c.st));
  return out;
}

void head(Ctx& c, const CdsegLinW& w, const CdsegPlanLevel& L0, const float* x, float* out_caller) {
  float* y = f32(c, L0.n, w.N);
  linear(c, x, L0.n, w, nullptr, 0, y);
  RUN(cdseg_gather_rows(y, L0.inv_perm, L0.n, w.N * 4, out_caller, c.st));
}

// y = LN_cpe(lin(conv3(x)))  (PointSequential(SubMConv3d, Linear, LayerNorm), ptv3.py:1105-1123)
float* cpe(Ctx& c, const float* x, const CdsegPlanLevel& L, int C, const float* conv_Bp, const float* conv_b, const CdsegLinW& lin,
           const CdsegLnW& ln, float eps) {
  float* y1 = f32(c, L.n, C);
  conv3(c, x, L, C, conv_Bp, conv_b, y1);
  float* y2 = f32(c, L.n, C);
  linear(c, y1, L.n, lin, nullptr, 0, y2);
  RUN(cdseg_add_layernorm(y2, nullptr, nullptr, nullptr, ln.g, ln.b, eps, L.n, C, nullptr, y1, c.st));
  return y1;
}

// CrossBlock (ptv3.py:1179-1223): q = features of the 5-stage network at its last level, kv = features of the 3-stage network at its
// last level.  The kv half -- kv_point.feat = LN(kv + cpe(kv)), which the reference leaves in kv_point.feat (ptv3.py:1190-1192) and the
// 3-stage decoder then consumes -- depends on the 3-stage encoder only, so it runs on that network's stream and its decoder never waits
// for the 5-stage encoder (profiles/r02_timeline.md: that wait idled the side stream for 3 ms while the main stream ran its
// latency-bound deep levels on a mostly empty GPU).
float* cross_kv(Ctx& c, const CdsegCrossW& w, const CdsegPlanLevel& Lk, const float* xkv) {
  const int Ck = w.Ckv;
  need(c, Lk.ready);
  float* ck = cpe(c, xkv, Lk, Ck, w.kv_conv_Bp, w.kv_conv_b, w.kv_lin, w.kv_cpe_ln, w.ln_eps);
  float* hkv = f32(c, Lk.n, Ck);
  RUN(cdseg_add_layernorm(xkv, ck, nullptr, nullptr, w.kv_norm1.g, w.kv_norm1.b, w.ln_eps, Lk.n, Ck, nullptr, hkv, c.st));
  return hkv;
}

// the q half: returns the new q features.  hkv = cross_kv's result (ordered by the caller's event when it comes from another stream)
float* cross_block(Ctx& c, const CdsegCrossW& w, const CdsegPlanLevel& Lq, const CdsegPlanLevel& Lk, const float* xq, const float* hkv) {
  if (!c.dry && Lq.n != Lk.n) { if (c.status == CDSEG_OK) c.status = CDSEG_EINVAL; }     // ptv3.py:1008-1010 index kv with q's pad map
  const int64_t n = Lq.n;
  const int Cq = w.Cq, H = w.H;
  need(c, Lq.ready); need(c, Lk.ready);
  float* cq = cpe(c, xq, Lq, Cq, w.q_conv_Bp, w.q_conv_b, w.q_lin, w.q_cpe_ln, w.ln_eps);
  float* q1 = f32(c, n, Cq);
  float* hq = f32(c, n, Cq);
  RUN(cdseg_add_layernorm(xq, cq, nullptr, nullptr, w.q_norm1.g, w.q_norm1.b, w.ln_eps, n, Cq, q1, hq, c.st));
  float* Q = f32(c, n, Cq);
  linear(c, hq, n, w.q, nullptr, 0, Q);
  float* KV = f32(c, Lk.n, 2 * Cq);
  linear(c, hkv, Lk.n, w.kv, nullptr, 0, KV);
  // slot maps: q along logical curve 0 of its level; kv rows by the NN's logical curve 0 with q's scene counts / patch size
  const CdsegPatchMap& pm = Lq.pm[0];
  const int T = pm.T, Kp = pm.Kp;
  int32_t* kv_src = (int32_t*)c.ar->take((size_t)T * Kp * 4);
  int32_t* kv_dst = (int32_t*)c.ar->take((size_t)T * Kp * 4);
  int32_t* kv_ps = (int32_t*)c.ar->take((size_t)n * 4);
  int32_t* kv_len = (int32_t*)c.ar->take((size_t)(T + 1) * 4);
  const size_t unit = (size_t)H * T * Kp * 16 * 2;
  const int am = c.a->attn_mode;
  const size_t qk_units = am == CDSEG_ATTN_F16 ? 1 : 2, v_units = am == CDSEG_ATTN_TC32 ? 3 : 2;
  void* qp = c.ar->take(qk_units * unit);
  void* kp = c.ar->take(qk_units * unit);
  void* vp = c.ar->take(v_units * unit);
  float* o = f32(c, n, Cq);
  if (!c.dry && !c.skip && c.status == CDSEG_OK) {
    int64_t cnt[CDSEG_MAX_SCENES];
    for (int b = 0; b < Lq.B; ++b) cnt[b] = Lq.offset_host[b] - (b ? Lq.offset_host[b - 1] : 0);
    RUN(cdseg_patch_maps(Lk.order + (size_t)Lk.rowmap[0] * Lk.cap, cnt, Lq.B, pm.K, Kp, kv_src, kv_dst, kv_ps, kv_len, c.st));
    if (am == CDSEG_ATTN_EXACT) {
      RUN(cdseg_attn_pack_f32(Q, Cq, 0, Cq, 1, pm.slot_src, H, T, Kp, (float*)qp, nullptr, nullptr, c.st));
      RUN(cdseg_attn_pack_f32(KV, 2 * Cq, 0, Cq, 2, kv_src, H, T, Kp, (float*)kp, (float*)vp, nullptr, c.st));
      RUN(cdseg_attn_exact((const float*)qp, (const float*)kp, (const float*)vp, pm.patch_len, pm.slot_dst, H, T, Kp, w.scale, o, Cq, c.st));
    } else if (am == CDSEG_ATTN_TC32) {
      RUN(cdseg_attn_pack_split(Q, Cq, 0, Cq, 1, pm.slot_src, H, T, Kp, qp, nullptr, nullptr, 0, c.st));
      RUN(cdseg_attn_pack_split(KV, 2 * Cq, 0, Cq, 2, kv_src, H, T, Kp, kp, vp, nullptr, 1, c.st));
      RUN(cdseg_attn_tc3(qp, kp, vp, pm.patch_len, pm.slot_dst, H, T, Kp, w.scale, 1, o, Cq, c.st));
    } else {
      RUN(cdseg_attn_pack_f16v(Q, Cq, 0, Cq, 1, pm.slot_src, H, T, Kp, qp, nullptr, nullptr, 0, c.st));
      RUN(cdseg_attn_pack_f16v(KV, 2 * Cq, 0, Cq, 2, kv_src, H, T, Kp, kp, vp, nullptr, 1, c.st));
      RUN(cdseg_attn_tc3(qp, kp, vp, pm.patch_len, pm.slot_dst, H, T, Kp, w.scale, 0, o, Cq, c.st));
    }
  }
  float* a = f32(c, n, Cq);
  linear(c, o, n, w.proj, nullptr, 0, a);
  if (w.tm_feat != 1.0f) RUN(cdseg_axpy_scale(a, a, 0.f, w.tm_feat, n * Cq, c.st));
  float* q2 = f32(c, n, Cq);
  float* h = f32(c, n, Cq);
  RUN(cdseg_add_layernorm(q1, a, nullptr, nullptr, w.q_norm2.g, w.q_norm2.b, w.ln_eps, n, Cq, q2, h, c.st));
  float* hid = f32(c, n, w.fc1.N);
  linear(c, h, n, w.fc1, nullptr, 1, hid);
  float* out = f32(c, n, Cq);
  linear(c, hid, n, w.fc2, q2, 0, out);
  return out;
}

cudaEvent_t g_ev_all[16][4] = {};                 // fork / join events, one set per device of the process

struct Sizes { size_t act_m, act_s, scratch_m, ws_m, scratch_s, ws_s; };

// One walk over the network.  dry: nothing is launched, *sz receives the activation / scratch / workspace bytes of each branch.
// !dry: *sz (from a dry walk of the same arguments) carves the reusable regions, then activations are bump-allocated behind them.
int walk(const CdsegForwardArgs* a, bool dry, Sizes* sz) {
  if (!a || !a->w || !a->levels || a->n_lv_n < 1 || a->N <= 0 || a->B <= 0 || a->B > CDSEG_MAX_SCENES) return CDSEG_EINVAL;
  const CdsegNetW& w = *a->w;
  if (w.n_enc != a->n_lv_n || w.n_dec != w.n_enc - 1 || w.n_enc > CDSEG_MAX_STAGES) return CDSEG_EINVAL;
  if (w.condition && (w.c_enc != a->n_lv_c || w.c_dec != w.c_enc - 1 || w.c_enc < 1 || w.c_enc > CDSEG_MAX_STAGES)) return CDSEG_EINVAL;
  if (!dry && (!a->n_feat || !a->n_out || !a->arena_main || (w.condition && (!a->c_feat || !a->c_out)))) return CDSEG_EINVAL;
  const bool two = w.condition && a->stream_side && a->stream_side != a->stream_main;
  if (two && !dry && !a->arena_side) return CDSEG_EINVAL;
  cudaStream_t sm = (cudaStream_t)a->stream_main, ss = two ? (cudaStream_t)a->stream_side : sm;
  int dev_id = 0;
  if (two && !dry && (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 16)) return (int)cudaErrorUnknown;
  cudaEvent_t* g_ev = g_ev_all[dev_id];
  if (two && !dry && !g_ev[0])
    for (int i = 0; i < 4; ++i)
      if (cudaEventCreateWithFlags(&g_ev[i], cudaEventDisableTiming) != cudaSuccess) return (int)cudaErrorUnknown;

  Arena am{(char*)a->arena_main, a->arena_main_bytes, 0, 0, dry};
  Arena as_{(char*)a->arena_side, a->arena_side_bytes, 0, 0, dry};
  Ctx cm{a, dry, CDSEG_OK, sm, &am, nullptr, 0, nullptr, 0, 0, 0, false};
  Ctx cs{a, dry, CDSEG_OK, ss, two ? &as_ : &am, nullptr, 0, nullptr, 0, 0, 0, false};
  if (!dry) {
    cm.scratch_bytes = sz->scratch_m; cm.ws_bytes = sz->ws_m; cs.scratch_bytes = sz->scratch_s; cs.ws_bytes = sz->ws_s;
    cm.scratch = am.take(cm.scratch_bytes); cm.ws = am.take(cm.ws_bytes);
    if (two) { cs.scratch = as_.take(cs.scratch_bytes); cs.ws = as_.take(cs.ws_bytes); }
    else { cs.scratch = cm.scratch; cs.ws = cm.ws; }
    if (!am.ok() || !as_.ok()) return CDSEG_ENOSPC;
  }

  const CdsegPlanLevel* lv = a->levels;
  const int nb = 0, cb = a->n_lv_n;                       // index of each network's level 0
  int n_blocks_n = 0;
  for (int s = 0; s < w.n_enc; ++s) n_blocks_n += w.n_enc_st[s].n_blocks;
  for (int s = 0; s < w.n_dec; ++s) n_blocks_n += w.n_dec_st[s].n_blocks;
  cs.ev_index = n_blocks_n;

  if (two && !dry) { cudaEventRecord(g_ev[0], sm); cudaStreamWaitEvent(ss, g_ev[0], 0); }     // plan tables were built on the main stream

  // Enqueue order.  The 5-stage network is the critical path, so with two streams its kernels are enqueued first: the host needs
  // ~0.25 ms to enqueue the 3-stage encoder, which used to delay the first kernel of the critical stream by as much
  // (profiles/r02_timeline.md).  One stream (and the debug serialisations) keep the historical order.
  const bool main_first = two && !(g_net_debug & 3);
  float* cx = nullptr;
  float* cx_kv = nullptr;
  float* t_scene = nullptr;
  float* c_skip[CDSEG_MAX_STAGES] = {nullptr};
  float* n_skip[CDSEG_MAX_STAGES] = {nullptr};
  float* nx = nullptr;
  auto side_encoder = [&]() {                             // ---- 3-stage network: timestep MLP, stem, encoder (side stream) ----
    Ctx& c = cs;
    if (w.T_dim > 0 && a->t_emb) {                        // timestep MLP once per scene (ptv3.py:1772-1778): fc_t1 -> swish -> fc_t2 -> swish
      float* t1 = f32(c, a->B, 4 * w.T_dim);
      t_scene = f32(c, a->B, w.T_dim);
      RUN(cdseg_small_linear(a->t_emb, w.fc_t1_W, w.fc_t1_b, 2, a->B, w.T_dim, 4 * w.T_dim, t1, c.st));
      RUN(cdseg_small_linear(t1, w.fc_t2_W, w.fc_t2_b, 2, a->B, 4 * w.T_dim, w.T_dim, t_scene, c.st));
    }
    cx = stem(c, w.c_stem, lv[cb], a->c_feat);
    for (int s = 0; s < w.c_enc; ++s) { cx = enc_stage(c, w.c_enc_st[s], lv, cb, cx, t_scene); c_skip[s] = cx; }
    cx = cx_kv = cross_kv(c, w.tm, lv[cb + w.c_enc - 1], cx);     // TransferModule, kv half: the decoder below starts from it
    if (two && !dry) cudaEventRecord(g_ev[1], ss);
    if (two && !dry && (g_net_debug & 1)) cudaStreamWaitEvent(sm, g_ev[1], 0);          // debug: serialise the two encoders
  };
  auto main_encoder = [&]() {                             // ---- 5-stage network: stem, encoder (main stream) ----
    nx = stem(cm, w.n_stem, lv[nb], a->n_feat);
    for (int s = 0; s < w.n_enc; ++s) { nx = enc_stage(cm, w.n_enc_st[s], lv, nb, nx, nullptr); n_skip[s] = nx; }
  };
  auto side_decoder = [&]() {                             // ---- 3-stage network: decoder + head (side stream) ----
    Ctx& c = cs;
    for (int j = 0; j < w.c_dec; ++j) cx = dec_stage(c, w.c_dec_st[j], lv, cb, cx, c_skip[w.c_dec_st[j].level], t_scene);
    head(c, w.c_head, lv[cb], cx, a->c_out);
    if (two && !dry) cudaEventRecord(g_ev[3], ss);
    if (two && !dry && (g_net_debug & 2)) cudaStreamWaitEvent(sm, g_ev[3], 0);          // debug: serialise the two decoders
  };
  auto main_decoder = [&]() {                             // ---- 5-stage network: decoder + head (main stream) ----
    for (int j = 0; j < w.n_dec; ++j) nx = dec_stage(cm, w.n_dec_st[j], lv, nb, nx, n_skip[w.n_dec_st[j].level], nullptr);
    head(cm, w.n_head, lv[nb], nx, a->n_out);
  };
  if (!w.condition) {
    main_encoder();
    main_decoder();
  } else {
    // the 3-stage network never waits for the 5-stage one: encoder, kv half of the TransferModule and decoder are one chain
    if (main_first) { main_encoder(); side_encoder(); side_decoder(); } else { side_encoder(); side_decoder(); main_encoder(); }
    // ---- TransferModule, q half (main stream) ----
    if (two && !dry) cudaStreamWaitEvent(sm, g_ev[1], 0);
    nx = cross_block(cm, w.tm, lv[nb + w.n_enc - 1], lv[cb + w.c_enc - 1], nx, cx_kv);
    main_decoder();
    if (two && !dry) cudaStreamWaitEvent(sm, g_ev[3], 0);
  }

  if (dry) {
    sz->act_m = am.high; sz->act_s = as_.high;
    sz->scratch_m = cm.scratch_bytes; sz->ws_m = cm.ws_bytes; sz->scratch_s = cs.scratch_bytes; sz->ws_s = cs.ws_bytes;
    if (!two) {                                           // one stream: one scratch / workspace serves both networks
      sz->scratch_m = sz->scratch_s = cm.scratch_bytes > cs.scratch_bytes ? cm.scratch_bytes : cs.scratch_bytes;
      sz->ws_m = sz->ws_s = cm.ws_bytes > cs.ws_bytes ? cm.ws_bytes : cs.ws_bytes;
    }
    return CDSEG_OK;
  }
  if (cm.status != CDSEG_OK) return cm.status;
  if (cs.status != CDSEG_OK) return cs.status;
  if (!am.ok() || !as_.ok()) return CDSEG_ENOSPC;
  return CDSEG_OK;
}

void totals(const CdsegForwardArgs* a, const Sizes& z, size_t* m, size_t* s) {
  const bool two = a->w->condition && a->stream_side && a->stream_side != a->stream_main;
  *m = z.act_m + al256(z.scratch_m) + al256(z.ws_m) + 4096;
  *s = two ? z.act_s + al256(z.scratch_s) + al256(z.ws_s) + 4096 : 0;
}

}  // namespace

// debug / profiling switches (also env CDSEG_NET_DEBUG): bit 0 serialise the two encoders, bit 1 serialise the two decoders,
// bit 2 log every arena allocation to stderr
CDSEG_API void cdseg_net_set_debug(int flags) { g_net_debug = flags; }

// the split heuristic of this file (block_exec.cu and cdsegnet_b200/ops.py::pick_split carry copies: the three launch paths must make the
// same choice to stay bit-identical; tests/test_cpu_oracle.py compares them through this export)
CDSEG_API int cdseg_debug_pick_split(int64_t tiles, int T) { return pick_split(tiles, T); }

CDSEG_API int cdseg_net_arena_bytes(const CdsegForwardArgs* args, size_t* main_bytes, size_t* side_bytes) {
  Sizes z{};
  const int st = walk(args, true, &z);
  if (st != CDSEG_OK) return st;
  size_t m, s;
  totals(args, z, &m, &s);
  if (main_bytes) *main_bytes = m;
  if (side_bytes) *side_bytes = s;
  return CDSEG_OK;
}

CDSEG_API int cdseg_net_forward(const CdsegForwardArgs* args) {
  Sizes z{};
  int st = walk(args, true, &z);                          // sizing pass: host arithmetic only
  if (st != CDSEG_OK) return st;
  size_t m, s;
  totals(args, z, &m, &s);
  if (m > args->arena_main_bytes || s > args->arena_side_bytes) return CDSEG_ENOSPC;
  return walk(args, false, &z);
}

// sizeof of every struct of the C ABI, in the order documented in include/cdseg_b200.h: lets a binding (ctypes, cgo, JNI ...) verify its
// mirror of the layouts without a GPU (tests/test_cpu_oracle.py)
CDSEG_API int cdseg_struct_sizes(size_t* out, int n) {
  const size_t v[] = {sizeof(CdsegBlockArgs), sizeof(CdsegPatchMap), sizeof(CdsegPlanLevel), sizeof(CdsegLinW), sizeof(CdsegLnW),
                      sizeof(CdsegBlockW), sizeof(CdsegPoolW), sizeof(CdsegUnpoolW), sizeof(CdsegStageW), sizeof(CdsegStemW),
                      sizeof(CdsegCrossW), sizeof(CdsegNetW), sizeof(CdsegForwardArgs)};
  const int m = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < m; ++i) out[i] = v[i];
  return m;
}
