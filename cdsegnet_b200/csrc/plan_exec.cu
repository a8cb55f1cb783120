// Native plan phase: everything of a CDSegNet forward that depends only on grid_coord / offset, in ONE C-ABI call.
//
// Replaces (reference): Point.serialization (pointcept/models/utils/structure.py:47-102), the structural half of every
// SerializedPooling of both networks (point_transformer_v3m1_base.py:464-505), Point.sparsify's indice tables (structure.py:104-140,
// spconv indice_key caches) and SerializedAttention.get_padding_and_inverse (point_transformer_v3m1_base.py:188-244).
//
// Round 1 issued these ~75 launches from Python (cdsegnet_b200/structure.py): 4.3 ms of host time per forward during which the GPU
// had nothing to run (profiles/r02_host_overhead.txt).  Here the same kernels are enqueued from C++ with two stream syncs:
//   sync #1  max(grid) -> depth / key bits; offsets            (the reference syncs here too, structure.py:66)
//   sync #2  point counts + per-scene offsets of the 6 pooled levels (+ caller flags)
// Caller allocates: one device arena (cdseg_plan_arena_bytes), the level descriptors are filled with pointers into it.
#include "common.cuh"
#include "../../include/cdseg_b200.h"
#include <cstring>
#include <functional>
#include <vector>

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// n_levels: pooled levels of both networks; n_level0: level-0 entries (1 or 2); K_min / K_max: smallest / largest patch size in use
CDSEG_API size_t cdseg_plan_arena_bytes(int64_t N, int B, int k, int n_levels, int n_level0, int stem_ksize, int K_min, int K_max) {
  if (N <= 0 || B <= 0 || k <= 0 || n_levels < 0 || K_min <= 0 || K_max < K_min) return 0;
  const size_t cap = (size_t)N;
  size_t s = 0;
  // level-0 originals (caller numbering) + internal copies
  s += al256(cap * 4) + 2 * (al256(cap * k * 8) + 2 * al256(cap * k * 4)) + al256(cap * 12) + al256(cap * 4);
  // pooled levels: cluster, idx_ptr, head, code, order, inverse, grid, batch
  s += (size_t)n_levels * (al256(cap * 4) + al256((cap + 1) * 4) + al256(cap * 4) + al256(cap * k * 8) + 2 * al256(cap * k * 4) +
                           al256(cap * 12) + al256(cap * 4));
  s += al256((size_t)n_levels * 4) + al256((size_t)n_levels * B * 8) + 4096;                     // counts, offsets
  // tables: nbr3 + tile mask + conv plan per level, nbr(stem) at level 0, patch maps (<= k per level; slots <= n + B * Kp, Kp <= 1024)
  s += (size_t)n_levels * (al256(cap * 27 * 4) + al256((cap / 128 + 1) * 4) + al256(cdseg_conv_plan_bytes(N)));
  s += al256(cap * (size_t)stem_ksize * stem_ksize * stem_ksize * 4);
  // patch slots of one map: T * Kp with T <= n / K + B, Kp = round_up(K, 128)
  const size_t slots_a = (cap / K_min + B + 1) * (size_t)((K_min + 127) / 128 * 128);
  const size_t slots_b = 2 * cap + (size_t)(B + 1) * ((K_max + 127) / 128 * 128);
  const size_t slots = slots_a > slots_b ? slots_a : slots_b;
  s += (size_t)(n_levels + n_level0) * k * (2 * al256(slots * 4) + al256(cap * 4) + al256((cap / K_min + B + 8) * 4));
  // workspaces (argsort / pool plan / hash): reused
  size_t ws = cdseg_argsort_workspace_bytes(k, N);
  const size_t w2 = cdseg_pool_plan_workspace_bytes(k, N), w3 = cdseg_nbr_workspace_bytes(N);
  ws = ws > w2 ? ws : w2;
  ws = ws > w3 ? ws : w3;
  return s + al256(ws) + al256(w3) + (1 << 16);                     // + the aux stream's hash workspace
}

// host staging for the two device->host copies (pinned, grown on demand)
static void* g_stage = nullptr;
static size_t g_stage_bytes = 0;
static void* stage(size_t bytes) {
  if (bytes > g_stage_bytes) {
    if (g_stage) cudaFreeHost(g_stage);
    g_stage_bytes = bytes < 65536 ? 65536 : bytes * 2;
    if (cudaHostAlloc(&g_stage, g_stage_bytes, cudaHostAllocDefault) != cudaSuccess) { g_stage = nullptr; g_stage_bytes = 0; }
  }
  return g_stage;
}

// events of the aux-stream schedule: one per level (tables complete), one for the stem table, two forks.  Static: a later plan's
// records land later on the same in-order aux stream, so a consumer that waits on a re-recorded event only over-waits.
enum { MAX_LV = 16 };
enum { MAX_DEV = 16 };
static cudaEvent_t g_pev_all[MAX_DEV][MAX_LV + 3] = {};         // events belong to a device: one set per device of the process
static cudaEvent_t* plan_events() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
  for (auto& e : g_pev_all[dev])
    if (!e && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  return g_pev_all[dev];
}

// Launches of the pooled levels' tables and slot maps that an aux-stream build leaves for cdseg_plan_finish(): after the second host sync
// the GPU is waiting for the host, and these ~60 launches (0.25 ms of enqueue time) feed nothing before the second stage of either
// network -- cdseg_net_forward enqueues the stems and the level-0 stage first and calls cdseg_plan_finish() before its first pooled
// stage (profiles/r02_timeline.md).  One plan at a time per process (the list is static); a new build flushes what is left.
static std::vector<std::function<int()>> g_pending;
static int g_pending_dev = -1;                                   // device the pending launches belong to
CDSEG_API int cdseg_plan_finish(void) {
  if (g_pending.empty()) return CDSEG_OK;
  int cur = 0, rc = CDSEG_OK;
  cudaGetDevice(&cur);
  if (g_pending_dev >= 0 && g_pending_dev != cur) cudaSetDevice(g_pending_dev);
  for (auto& f : g_pending) {
    const int r = f();
    if (rc == CDSEG_OK) rc = r;
  }
  g_pending.clear();
  if (g_pending_dev >= 0 && g_pending_dev != cur) cudaSetDevice(cur);
  return rc;
}

// levels[0 .. n_lv): every level of both networks.  Inputs per level (host): parent (index of the level it is pooled from, -1 for a
// level-0 entry), stride, rowmap (logical curve row -> physical row), K + pm_mask (patch maps wanted, by logical curve), want_*.
// Level-0 entries share one set of arrays (the two networks serialize the same points).  A parent must precede its children.
CDSEG_API int cdseg_plan_build(const int32_t* grid, const int64_t* offset, int64_t N, int B, const int* order_ids, int k,
                               CdsegPlanLevel* lv, int n_lv, const int32_t* extra_flags, int n_flags, int32_t* flags_host,
                               void* arena, size_t arena_bytes, void* stream, void* aux_stream) {
  if (!grid || !offset || N <= 0 || B <= 0 || B > CDSEG_MAX_SCENES || k <= 0 || k > 4 || !lv || n_lv <= 0 || n_lv > MAX_LV || !arena)
    return CDSEG_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  // aux stream: the indice tables (60 % of the plan phase's device time) only feed the feature phase, so they are built beside the
  // pooling hierarchy and beside the first kernels of the forward; consumers wait on the per-level `ready` events.
  const bool two = aux_stream && aux_stream != stream;
  cudaStream_t sa = two ? (cudaStream_t)aux_stream : st;
  cudaEvent_t* g_pev = two ? plan_events() : nullptr;
  if (two && !g_pev) return (int)cudaErrorUnknown;
  if (!g_pending.empty()) { const int r = cdseg_plan_finish(); if (r != CDSEG_OK) return r; }
  bool defer = false;                                              // set after the second sync (pooled levels only)
  char* p = (char*)arena;
  char* const end = p + arena_bytes;
  bool oom = false;
  auto take = [&](size_t bytes) -> void* {
    char* q = p;
    p += al256(bytes);
    if (p > end) { oom = true; return arena; }
    return q;
  };
  int st_;
#define RUN(call) do { st_ = (call); if (st_ != CDSEG_OK) return st_; } while (0)
  const size_t cap = (size_t)N;
  int n_pooled = 0;
  for (int i = 0; i < n_lv; ++i) n_pooled += lv[i].parent >= 0;

  // ---------------- sync #1: depth + offsets ----------------
  int32_t* gm = (int32_t*)take(256);
  RUN(cdseg_grid_max(grid, N * 3, gm, st));
  char* hs = (char*)stage(8 + (size_t)B * 8 + (size_t)(n_pooled + 1) * (4 + (size_t)B * 8) + (size_t)n_flags * 4 + 64);
  if (!hs) return (int)cudaErrorMemoryAllocation;
  cudaMemcpyAsync(hs, gm, 4, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(hs + 8, offset, (size_t)B * 8, cudaMemcpyDeviceToHost, st);
  cudaError_t ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) return (int)ce;
  const int32_t gmax = *(int32_t*)hs;
  int depth = 0;
  while ((1ll << depth) <= (long long)gmax) ++depth;                 // bit_length(max)
  if (depth > 16) return CDSEG_EINVAL;                               // structure.py:74
  int64_t off_host[CDSEG_MAX_SCENES];
  memcpy(off_host, hs + 8, (size_t)B * 8);
  if (off_host[B - 1] != N) return CDSEG_EINVAL;
  int bbits = 0;
  while ((1 << bbits) < B) ++bbits;                                  // bit_length(B - 1)
  const int nbits = 3 * depth + bbits;

  // ---------------- level 0: encode, sort, renumber along curve 0 ----------------
  int32_t* batch0 = (int32_t*)take(cap * 4);
  int64_t* o_code = (int64_t*)take(cap * k * 8);
  int32_t* o_order = (int32_t*)take(cap * k * 4);
  int32_t* o_inverse = (int32_t*)take(cap * k * 4);
  int32_t* i_grid = (int32_t*)take(cap * 12);
  int32_t* i_batch = (int32_t*)take(cap * 4);
  int64_t* i_code = (int64_t*)take(cap * k * 8);
  int32_t* i_order = (int32_t*)take(cap * k * 4);
  int32_t* i_inverse = (int32_t*)take(cap * k * 4);
  int32_t* cnt_buf = (int32_t*)take((size_t)(n_pooled + 1) * 4);
  int64_t* off_buf = (int64_t*)take((size_t)(n_pooled + 1) * B * 8);
  size_t ws_bytes = cdseg_argsort_workspace_bytes(k, N);
  {
    const size_t w2 = cdseg_pool_plan_workspace_bytes(k, N), w3 = cdseg_nbr_workspace_bytes(N);
    ws_bytes = ws_bytes > w2 ? ws_bytes : w2;
    ws_bytes = ws_bytes > w3 ? ws_bytes : w3;
  }
  void* ws = take(ws_bytes);
  if (oom) return CDSEG_ENOSPC;
  RUN(cdseg_offset2batch(offset, B, N, batch0, st));
  RUN(cdseg_encode_codes(grid, batch0, N, depth, order_ids, k, o_code, st));
  RUN(cdseg_argsort_rows(o_code, k, N, nbits, o_order, o_inverse, ws, ws_bytes, st));
  RUN(cdseg_renumber(o_order, o_inverse, grid, batch0, o_code, o_order, o_inverse, k, N, i_grid, i_batch, i_code, i_order, i_inverse, st));
  cudaMemsetAsync(cnt_buf, 0, (size_t)(n_pooled + 1) * 4, st);
  cudaMemsetAsync(off_buf, 0, (size_t)(n_pooled + 1) * B * 8, st);

  for (int i = 0; i < n_lv; ++i) {
    CdsegPlanLevel& L = lv[i];
    L.B = B;
    L.cap = N;
    L.ready = L.ready_stem = nullptr;
    if (L.parent < 0) {
      L.n = N; L.depth = depth; L.c0 = -1; L.pooling_depth = 0;
      L.grid = i_grid; L.batch = i_batch; L.code = i_code; L.order = i_order; L.inverse = i_inverse;
      L.cluster = nullptr; L.idx_ptr = nullptr; L.head = nullptr; L.m_dev = nullptr; L.offset_dev = (int64_t*)offset;
      L.perm = o_order; L.inv_perm = o_inverse; L.o_code = o_code; L.o_order = o_order; L.o_inverse = o_inverse;
      memcpy(L.offset_host, off_host, (size_t)B * 8);
    }
  }
  // ---------------- tables: neighbour indices, tap masks, conv tile plans (aux stream when given) ----------------
  const size_t ws_a_bytes = two ? cdseg_nbr_workspace_bytes(N) : ws_bytes;
  void* ws_a = two ? take(ws_a_bytes) : ws;
  if (oom) return CDSEG_ENOSPC;
  int first0 = -1;
  auto build_tables = [&](int i) -> int {
    CdsegPlanLevel& L = lv[i];
    if (L.parent < 0 && first0 >= 0) {                               // the second network's level 0: same points, same tables
      const CdsegPlanLevel& F = lv[first0];
      L.nbr3 = F.nbr3; L.tile_mask3 = F.tile_mask3; L.conv_plan3 = F.conv_plan3; L.nbr_stem = F.nbr_stem;
      L.ready = F.ready; L.ready_stem = F.ready_stem;
      return CDSEG_OK;
    }
    if (L.parent < 0) first0 = i;
    const int64_t n = L.n;
    L.nbr3 = (int32_t*)take((size_t)n * 27 * 4);
    L.tile_mask3 = (uint32_t*)take(((size_t)n / 128 + 1) * 4);
    L.conv_plan3 = nullptr;
    L.nbr_stem = nullptr;
    bool want_plan = L.want_conv_plan != 0, want_stem = L.stem_ksize > 0;
    if (L.parent < 0)
      for (int j = i + 1; j < n_lv; ++j)
        if (lv[j].parent < 0) { want_plan |= lv[j].want_conv_plan != 0; want_stem |= lv[j].stem_ksize > 0; }
    if (want_plan) L.conv_plan3 = take(cdseg_conv_plan_bytes(n));
    const int sk = L.stem_ksize > 0 ? L.stem_ksize : 5;
    if (want_stem) L.nbr_stem = (int32_t*)take((size_t)n * sk * sk * sk * 4);
    if (oom) return CDSEG_ENOSPC;
    if (two) { L.ready = g_pev[i]; if (L.nbr_stem) L.ready_stem = g_pev[MAX_LV]; }
    const CdsegPlanLevel D = L;                                      // by value: the closure may run after this call has returned
    cudaEvent_t ev = two ? g_pev[i] : nullptr, ev_stem = two ? g_pev[MAX_LV] : nullptr;
    auto launches = [D, n, sk, ws_a, ws_a_bytes, sa, ev, ev_stem]() -> int {
      int st_;
      if (n > 0) {
        if (D.nbr_stem) {                                            // first: the stems are the first consumers
          RUN(cdseg_nbr_build(D.grid, D.batch, n, sk, D.nbr_stem, ws_a, ws_a_bytes, sa));
          if (ev_stem) cudaEventRecord(ev_stem, sa);
        }
        RUN(cdseg_nbr_build(D.grid, D.batch, n, 3, D.nbr3, ws_a, ws_a_bytes, sa));
        RUN(cdseg_tile_tap_mask(D.nbr3, n, 27, D.tile_mask3, sa));
        if (D.conv_plan3) RUN(cdseg_conv_tile_plan(D.nbr3, n, D.conv_plan3, sa));
      }
      if (ev) cudaEventRecord(ev, sa);
      return CDSEG_OK;
    };
    if (defer) { g_pending.push_back(launches); return CDSEG_OK; }
    return launches();
  };
  if (two) { cudaEventRecord(g_pev[MAX_LV + 1], st); cudaStreamWaitEvent(sa, g_pev[MAX_LV + 1], 0); }
  for (int i = 0; i < n_lv; ++i)                                    // level 0 needs nothing from the pooling hierarchy: before it
    if (lv[i].parent < 0) RUN(build_tables(i));
  // ---------------- pooled levels (sync-free: child counts stay on the device) ----------------
  int slot = 0;
  for (int i = 0; i < n_lv; ++i) {
    CdsegPlanLevel& L = lv[i];
    if (L.parent < 0) continue;
    if (L.parent >= i) return CDSEG_EINVAL;
    const CdsegPlanLevel& P = lv[L.parent];
    int pd = 0;
    for (int s = L.stride - 1; s > 0; s >>= 1) ++pd;                 // (stride - 1).bit_length()
    if (pd > P.depth) pd = 0;
    L.pooling_depth = pd;
    L.depth = P.depth - pd;
    L.c0 = P.rowmap[0];
    L.cluster = (int32_t*)take(cap * 4);
    L.idx_ptr = (int32_t*)take((cap + 1) * 4);
    L.head = (int32_t*)take(cap * 4);
    L.code = (int64_t*)take(cap * k * 8);
    L.order = (int32_t*)take(cap * k * 4);
    L.inverse = (int32_t*)take(cap * k * 4);
    L.grid = (int32_t*)take(cap * 12);
    L.batch = (int32_t*)take(cap * 4);
    L.m_dev = cnt_buf + slot;
    L.offset_dev = off_buf + (size_t)slot * B;
    L.slot = slot++;
    L.perm = L.inv_perm = nullptr; L.o_code = nullptr; L.o_order = L.o_inverse = nullptr;
    if (oom) return CDSEG_ENOSPC;
    RUN(cdseg_pool_plan(P.code, P.order, k, N, P.m_dev, P.m_dev ? 0 : P.n, L.c0, pd, P.grid, P.batch, L.cluster, L.idx_ptr, L.head, L.code,
                        L.order, L.inverse, N, L.grid, L.batch, L.m_dev, L.offset_dev, ws, ws_bytes, st));
  }
  // ---------------- sync #2: pooled sizes + offsets (+ caller flags) ----------------
  if (n_pooled || n_flags) {
    char* h = hs;
    if (n_pooled) {
      cudaMemcpyAsync(h, cnt_buf, (size_t)n_pooled * 4, cudaMemcpyDeviceToHost, st);
      cudaMemcpyAsync(h + al256((size_t)n_pooled * 4), off_buf, (size_t)n_pooled * B * 8, cudaMemcpyDeviceToHost, st);
    }
    char* hf = h + al256((size_t)n_pooled * 4) + al256((size_t)n_pooled * B * 8);
    if (n_flags) cudaMemcpyAsync(hf, extra_flags, (size_t)n_flags * 4, cudaMemcpyDeviceToHost, st);
    ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return (int)ce;
    for (int i = 0; i < n_lv; ++i) {
      CdsegPlanLevel& L = lv[i];
      if (L.parent < 0) continue;
      L.n = ((int32_t*)h)[L.slot];
      memcpy(L.offset_host, h + al256((size_t)n_pooled * 4) + (size_t)L.slot * B * 8, (size_t)B * 8);
    }
    if (n_flags && flags_host) memcpy(flags_host, hf, (size_t)n_flags * 4);
  }
  // ---------------- tables of the pooled levels + patch slot maps ----------------
  if (two) { cudaEventRecord(g_pev[MAX_LV + 2], st); cudaStreamWaitEvent(sa, g_pev[MAX_LV + 2], 0); }     // the pooled grids come from `st`
  defer = two;
  if (defer) cudaGetDevice(&g_pending_dev);
  for (int i = 0; i < n_lv; ++i) {
    CdsegPlanLevel& L = lv[i];
    if (L.parent >= 0) RUN(build_tables(i));
    // patch maps: level 0 now (its blocks come first), pooled levels with the deferred launches -- on the aux stream then, so that
    // `ready` orders them too
    const bool later = defer && L.parent >= 0;
    cudaStream_t sp = later ? sa : st;
    // patch maps, one per distinct PHYSICAL row among the wanted logical curves
    int64_t cnt[CDSEG_MAX_SCENES];
    for (int b = 0; b < B; ++b) cnt[b] = L.offset_host[b] - (b ? L.offset_host[b - 1] : 0);
    for (int r = 0; r < 4; ++r) memset(&L.pm[r], 0, sizeof(CdsegPatchMap));
    for (int r = 0; r < k; ++r) {
      if (!(L.pm_mask & (1u << r)) || L.K <= 0 || L.n <= 0) continue;
      const int prow = L.rowmap[r];
      int same = -1;
      for (int q = 0; q < r; ++q)
        if ((L.pm_mask & (1u << q)) && L.rowmap[q] == prow) same = q;
      if (same >= 0) { L.pm[r] = L.pm[same]; continue; }
      int T = 0;
      RUN(cdseg_patch_count(cnt, B, L.K, &T));
      const int Kp = (L.K + 127) / 128 * 128;
      CdsegPatchMap& M = L.pm[r];
      M.T = T; M.Kp = Kp; M.K = L.K;
      M.slot_src = (int32_t*)take((size_t)T * Kp * 4);
      M.slot_dst = (int32_t*)take((size_t)T * Kp * 4);
      M.point_slot = (int32_t*)take((size_t)L.n * 4);
      M.patch_len = (int32_t*)take((size_t)(T + 1) * 4);
      int64_t pairs = 0;
      for (int b = 0; b < B; ++b) pairs += cnt[b] * (cnt[b] > L.K ? L.K : cnt[b]);
      M.pairs = pairs;
      if (oom) return CDSEG_ENOSPC;
      if (later) {
        const int32_t* ord = L.order + (size_t)prow * N;
        const int K_ = L.K;
        std::vector<int64_t> cv(cnt, cnt + B);
        const CdsegPatchMap Mv = M;
        cudaEvent_t ev = g_pev[i];
        g_pending.push_back([ord, cv, B, K_, Kp, Mv, sp, ev]() -> int {
          const int r = cdseg_patch_maps(ord, cv.data(), B, K_, Kp, Mv.slot_src, Mv.slot_dst, Mv.point_slot, Mv.patch_len, sp);
          cudaEventRecord(ev, sp);                                   // re-record: `ready` now covers this map as well
          return r;
        });
      } else {
        RUN(cdseg_patch_maps(L.order + (size_t)prow * N, cnt, B, L.K, Kp, M.slot_src, M.slot_dst, M.point_slot, M.patch_len, sp));
      }
    }
  }
#undef RUN
  return oom ? CDSEG_ENOSPC : CDSEG_OK;
}
