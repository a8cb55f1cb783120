// Test-time fragment pipeline: the step on both sides of the forward in tools/test_CDSegNet_*.py.
//
//   GridSample(mode="test")  pointcept/datasets/transform.py:796-905   voxelise a raw scene, hash the voxels (FNV64-1A
//                            or ravel, :907-933), sort by hash, cut the scene into count.max() fragments that each hold one
//                            point of every voxel (the (i % count)-th of its run)
//   vote accumulation        pointcept/engines/test.py:198-267          pred[idx_part] += softmax(logits); argmax at the end
//
// The reference does this with numpy on the host (argsort + unique + a python loop over fragments) and moves every fragment
// to the GPU; here the raw scene is uploaded once and everything is a coalesced HBM kernel: floor/min/max reduce, key encode,
// the 64-bit radix argsort of serialize.cu, run flags + scan (voxel ids, starts, counts), one gather for ALL fragment index
// rows, and a fused softmax + scatter-add for the votes.
#include <limits.h>
#include "common.cuh"
#include "../../include/cdseg_b200.h"

namespace fr {

constexpr int TH = 256;
constexpr int ITEMS = 4;
constexpr int TILE = TH * ITEMS;

// coordinate element e of a float32 (f64 == 0) or float64 array
__device__ __forceinline__ double ld_coord(const void* coord, int64_t e, int f64) {
  return f64 ? reinterpret_cast<const double*>(coord)[e] : (double)reinterpret_cast<const float*>(coord)[e];
}
__device__ __forceinline__ int floor_div(double c, double gs, int f32) {
  // NumPy >= 2 divides float32 coordinates by np.array(grid_size) in float64; NumPy 1.x value-based casting kept float32
  return f32 ? (int)floorf((float)c / (float)gs) : (int)floor(c / gs);
}

// mm[0..2] = per-axis min, mm[3..5] = per-axis max of floor(coord / grid_size)
__global__ void __launch_bounds__(TH) voxel_minmax_kernel(const void* __restrict__ coord, int f64, int64_t n, double gs, int f32, int* __restrict__ mm) {
  __shared__ int s[6][TH / 32];
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int64_t i = (int64_t)blockIdx.x * TH + threadIdx.x; i < n; i += (int64_t)gridDim.x * TH)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int g = floor_div(ld_coord(coord, 3 * i + a, f64), gs, f32);
      lo[a] = min(lo[a], g); hi[a] = max(hi[a], g);
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) { lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
    if (lane == 0) { s[a][warp] = lo[a]; s[3 + a][warp] = hi[a]; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    int v = s[threadIdx.x][0];
    for (int w = 1; w < TH / 32; ++w) v = threadIdx.x < 3 ? min(v, s[threadIdx.x][w]) : max(v, s[threadIdx.x][w]);
    if (threadIdx.x < 3) atomicMin(mm + threadIdx.x, v); else atomicMax(mm + threadIdx.x, v);
  }
}

// grid = floor(coord / gs) - min ; key = FNV64-1A (multiply, then xor, per axis: transform.py:918-933) or the ravel hash (:907-916)
__global__ void voxel_key_kernel(const void* __restrict__ coord, int f64, int64_t n, double gs, int f32, const int* __restrict__ mm, int hash_fnv,
                                 int32_t* __restrict__ grid, int64_t* __restrict__ key) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int v = floor_div(ld_coord(coord, 3 * i + a, f64), gs, f32) - mm[a];
    grid[3 * i + a] = v;
    g[a] = (uint64_t)v;
  }
  uint64_t h;
  if (hash_fnv) {
    h = 14695981039346656037ull;
#pragma unroll
    for (int a = 0; a < 3; ++a) { h *= 1099511628211ull; h ^= g[a]; }
  } else {
    const uint64_t m1 = (uint64_t)(mm[4] - mm[1]) + 1, m2 = (uint64_t)(mm[5] - mm[2]) + 1;
    h = (g[0] * m1 + g[1]) * m2 + g[2];
  }
  key[i] = (int64_t)h;
}

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[TH / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < TH / 32 ? wsum[lane] : 0, wi = w;
#pragma unroll
    for (int o = 1; o < TH / 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += x; }
    if (lane < TH / 32) wsum[lane] = wi - w;
    if (lane == TH / 32 - 1) *total = wi;
  }
  __syncthreads();
  return wsum[warp] + inc - v;
}

// run heads of the sorted key sequence: seg[i] = tile-local exclusive rank | flag << 31, blk[tile] = heads in the tile
__global__ void __launch_bounds__(TH) run_flag_kernel(const int64_t* __restrict__ key, const int32_t* __restrict__ order, int64_t n,
                                                      uint32_t* __restrict__ seg, uint32_t* __restrict__ blk) {
  __shared__ uint32_t total;
  const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
  uint32_t f[ITEMS], cnt = 0;
  int64_t prev = (base > 0 && base - 1 < n) ? key[order[base - 1]] : 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + j;
    if (i < n) {
      const int64_t k = key[order[i]];
      f[j] = (i == 0 || k != prev) ? 1u : 0u;
      prev = k;
    } else f[j] = 0;
    cnt += f[j];
  }
  uint32_t ex = block_excl_scan(cnt, &total);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = base + j;
    if (i < n) seg[i] = ex | (f[j] << 31);
    ex += f[j];
  }
  if (threadIdx.x == 0) blk[blockIdx.x] = total;
}

// exclusive scan of the tile totals (one CTA); stats[0] = number of voxels
__global__ void __launch_bounds__(TH) run_blkscan_kernel(uint32_t* __restrict__ blk, int ntiles, int32_t* __restrict__ stats) {
  __shared__ uint32_t total, carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < ntiles; b0 += TH) {
    const int t = b0 + threadIdx.x;
    const uint32_t v = t < ntiles ? blk[t] : 0;
    const uint32_t ex = block_excl_scan(v, &total);
    if (t < ntiles) blk[t] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) stats[0] = (int32_t)carry;
}

// voxel id of every point (GridSample's `inverse`), start of every run in the sorted order (start[V] = n)
__global__ void run_write_kernel(const int32_t* __restrict__ order, int64_t n, const uint32_t* __restrict__ seg, const uint32_t* __restrict__ blk,
                                 const int32_t* __restrict__ stats, int32_t* __restrict__ voxel_of_point, int32_t* __restrict__ start) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) start[stats[0]] = (int32_t)n;
  if (i >= n) return;
  const uint32_t s = seg[i];
  const bool flag = s >> 31;
  const uint32_t id = (s & 0x7fffffffu) + blk[i / TILE] - (flag ? 0u : 1u);
  voxel_of_point[order[i]] = (int32_t)id;
  if (flag) start[id] = (int32_t)i;
}

// count[v] = start[v + 1] - start[v]; stats[1] = max count (= number of fragments)
__global__ void run_count_kernel(const int32_t* __restrict__ start, int32_t* __restrict__ count, int32_t* stats) {
  const int V = stats[0];
  int mx = 0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
    const int c = start[v + 1] - start[v];
    count[v] = c;
    mx = max(mx, c);
  }
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(stats + 1, mx);
}

// index[f][v] = order[start[v] + f % count[v]]   (transform.py:868-870)
__global__ void fragment_index_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ start, int V, int F,
                                      int32_t* __restrict__ index) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int s = start[v], c = start[v + 1] - s;
  for (int f = 0; f < F; ++f) index[(int64_t)f * V + v] = order[s + f % c];
}

// pred[index[r], :] += softmax(logits[r, :])   (test.py:252-257); one warp per row
__global__ void __launch_bounds__(256) vote_softmax_add_kernel(const float* __restrict__ logits, const int32_t* __restrict__ index, int64_t n,
                                                               int C, float* __restrict__ pred) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= n) return;
  const float* row = logits + r * C;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
  s = warp_sum(s);
  float* dst = pred + (int64_t)index[r] * C;
  for (int c = lane; c < C; c += 32) atomicAdd(dst + c, expf(row[c] - m) / s);
}

// out[r] = first index of the row maximum
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ x, int64_t n, int C, int64_t* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= n) return;
  float best = -INFINITY;
  int bi = INT_MAX;
  for (int c = lane; c < C; c += 32) {
    const float v = x[r * C + c];
    if (v > best || (v == best && c < bi)) { best = v; bi = c; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) out[r] = bi == INT_MAX ? 0 : bi;
}

}  // namespace fr

CDSEG_API size_t cdseg_grid_sample_workspace_bytes(int64_t n) {
  const int64_t ntiles = (n + fr::TILE - 1) / fr::TILE;
  return (size_t)n * 4 * 2 + (size_t)(ntiles + 64) * 4 + cdseg_argsort_workspace_bytes(1, n) + 1024;
}

// see include/cdseg_b200.h
CDSEG_API int cdseg_grid_sample_plan(const void* coord, int coord_f64, int64_t n, double grid_size, int hash_fnv, int legacy_f32, int32_t* grid,
                                     int64_t* key, int32_t* order, int32_t* voxel_of_point, int32_t* start, int32_t* count,
                                     int32_t* stats, void* workspace, size_t workspace_bytes, void* stream) {
  if (n <= 0 || n >= (1ll << 31) || !coord || !(grid_size > 0) || !grid || !key || !order || !voxel_of_point || !start || !count || !stats ||
      !workspace)
    return CDSEG_EINVAL;
  if (workspace_bytes < cdseg_grid_sample_workspace_bytes(n)) return CDSEG_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  const int ntiles = (int)((n + fr::TILE - 1) / fr::TILE);
  char* p = (char*)workspace;
  int32_t* inverse = (int32_t*)p; p += (size_t)n * 4;
  uint32_t* seg = (uint32_t*)p; p += (size_t)n * 4;
  uint32_t* blk = (uint32_t*)p; p += (size_t)(ntiles + 64) * 4;
  p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  const size_t sort_ws = (size_t)((char*)workspace + workspace_bytes - p);
  // stats: [0] voxels, [1] max count, [2..4] min, [5..7] max of floor(coord / grid_size)
  const int init[8] = {0, 0, INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  cudaMemcpyAsync(stats, init, sizeof(init), cudaMemcpyHostToDevice, st);
  const int blocks = (int)((n + fr::TH - 1) / fr::TH < 1184 ? (n + fr::TH - 1) / fr::TH : 1184);
  fr::voxel_minmax_kernel<<<blocks, fr::TH, 0, st>>>(coord, coord_f64, n, grid_size, legacy_f32 && !coord_f64, stats + 2);
  fr::voxel_key_kernel<<<cdseg_div_up(n, 256), 256, 0, st>>>(coord, coord_f64, n, grid_size, legacy_f32 && !coord_f64, stats + 2, hash_fnv, grid, key);
  CDSEG_COUNT_LAUNCH(2);
  int s = cdseg_argsort_rows(key, 1, n, 64, order, inverse, p, sort_ws, stream);
  if (s != CDSEG_OK) return s;
  fr::run_flag_kernel<<<ntiles, fr::TH, 0, st>>>(key, order, n, seg, blk);
  fr::run_blkscan_kernel<<<1, fr::TH, 0, st>>>(blk, ntiles, stats);
  fr::run_write_kernel<<<cdseg_div_up(n, 256), 256, 0, st>>>(order, n, seg, blk, stats, voxel_of_point, start);
  fr::run_count_kernel<<<blocks, fr::TH, 0, st>>>(start, count, stats);
  CDSEG_COUNT_LAUNCH(4);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_fragment_index(const int32_t* order, const int32_t* start, int V, int F, int32_t* index, void* stream) {
  if (V < 0 || F < 0 || !order || !start || !index) return CDSEG_EINVAL;
  if (V == 0 || F == 0) return CDSEG_OK;
  fr::fragment_index_kernel<<<cdseg_div_up(V, 256), 256, 0, (cudaStream_t)stream>>>(order, start, V, F, index);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_vote_softmax_add(const float* logits, const int32_t* index, int64_t n, int C, float* pred, void* stream) {
  if (n < 0 || C <= 0 || !logits || !index || !pred) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  fr::vote_softmax_add_kernel<<<cdseg_div_up(n, 8), 256, 0, (cudaStream_t)stream>>>(logits, index, n, C, pred);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_argmax_rows(const float* x, int64_t n, int C, int64_t* out, void* stream) {
  if (n < 0 || C <= 0 || !x || !out) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  fr::argmax_rows_kernel<<<cdseg_div_up(n, 8), 256, 0, (cudaStream_t)stream>>>(x, n, C, out);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
