// Grid pooling / unpooling on B200.
//
// Replaces (reference file:line, ptv3.py = pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py):
//   ptv3.py:464-505  SerializedPooling plan: code >> 3d, torch.unique, torch.sort(cluster),
//                    cumsum(counts), 4x argsort of the pooled codes, scatter_ inverse
//   ptv3.py:507-531  torch_scatter.segment_csr(proj(feat)[indices], idx_ptr, "max") / coord "mean"
//   ptv3.py:623      parent.feat + point.feat[inverse]  (unpool gather-add)
//
// B200-first restatement: the parent is already sorted along every curve, and a
// right shift keeps a sorted sequence sorted, so clusters are RUNS of equal shifted
// keys in each curve's sorted order.  One flag + scan + compact per curve yields
// cluster ids, idx_ptr, head indices AND the pooled order/inverse of all curves --
// no second sort, no unique, no host sync (the pooled count stays in device memory
// and later levels read it from there).  Pooled point j == j-th smallest shifted
// code of the clustering curve, exactly as torch.unique(sorted=True) numbers them.
#include "common.cuh"

constexpr int PL_THREADS = 256;
constexpr int PL_ITEMS = 4;
constexpr int PL_TILE = PL_THREADS * PL_ITEMS;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[PL_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < PL_THREADS / 32 ? wsum[lane] : 0, wi = w;
#pragma unroll
    for (int o = 1; o < PL_THREADS / 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += n; }
    if (lane < PL_THREADS / 32) wsum[lane] = wi - w;
    if (lane == PL_THREADS / 32 - 1) *total = wi;
  }
  __syncthreads();
  return wsum[warp] + inc - v;
}

// A: per (tile, curve): run-head flags in sorted order + tile-local exclusive scan.
// seg[c][i] = local exclusive rank | flag << 31 ; blk[c][tile] = #heads in tile
__global__ void pool_flag_kernel(const int64_t* __restrict__ code, const int32_t* __restrict__ order, int64_t ld,
                                 const int32_t* __restrict__ n_dev, int64_t n_host, int shift, int ntiles,
                                 uint32_t* __restrict__ seg, uint32_t* __restrict__ blk) {
  __shared__ uint32_t total;
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
  const int c = blockIdx.y, tile = blockIdx.x;
  const int64_t base = (int64_t)tile * PL_TILE + (int64_t)threadIdx.x * PL_ITEMS;
  const int64_t* cd = code + (int64_t)c * ld;
  const int32_t* od = order + (int64_t)c * ld;
  uint32_t f[PL_ITEMS], cntv = 0;
  int64_t prev = (base > 0 && base - 1 < n) ? (cd[od[base - 1]] >> shift) : -1;
#pragma unroll
  for (int j = 0; j < PL_ITEMS; ++j) {
    const int64_t i = base + j;
    if (i < n) {
      const int64_t key = cd[od[i]] >> shift;
      f[j] = (i == 0 || key != prev) ? 1u : 0u;
      prev = key;
    } else f[j] = 0;
    cntv += f[j];
  }
  uint32_t ex = block_excl_scan(cntv, &total);
#pragma unroll
  for (int j = 0; j < PL_ITEMS; ++j) {
    const int64_t i = base + j;
    if (i < n) seg[(int64_t)c * ld + i] = ex | (f[j] << 31);
    ex += f[j];
  }
  if (threadIdx.x == 0) blk[c * ntiles + tile] = total;
}

// B: exclusive scan of tile totals per curve; pooled count of curve c0 -> *m_dev
__global__ void pool_blkscan_kernel(uint32_t* __restrict__ blk, int ntiles, int c0, int32_t* __restrict__ m_dev) {
  __shared__ uint32_t total;
  __shared__ uint32_t carry;
  const int c = blockIdx.x;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < ntiles; b0 += PL_THREADS) {
    const int t = b0 + threadIdx.x;
    const uint32_t v = t < ntiles ? blk[c * ntiles + t] : 0;
    const uint32_t ex = block_excl_scan(v, &total);
    if (t < ntiles) blk[c * ntiles + t] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && c == c0) *m_dev = (int32_t)carry;
}

// C0: clustering curve c0: cluster ids, idx_ptr, head, pooled attributes, pooled offsets
__global__ void pool_write0_kernel(const int64_t* __restrict__ code, const int32_t* __restrict__ order, int64_t ld,
                                   const int32_t* __restrict__ n_dev, int64_t n_host, int k, int c0, int shift, int pd,
                                   int ntiles, const uint32_t* __restrict__ seg, const uint32_t* __restrict__ blk,
                                   const int32_t* __restrict__ grid, const int32_t* __restrict__ batch,
                                   int32_t* __restrict__ cluster, int32_t* __restrict__ idx_ptr,
                                   int32_t* __restrict__ head, int64_t* __restrict__ c_code,
                                   int32_t* __restrict__ c_order, int32_t* __restrict__ c_inverse, int64_t ld_c,
                                   int32_t* __restrict__ c_grid, int32_t* __restrict__ c_batch,
                                   int64_t* __restrict__ c_offset) {
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int tile = (int)(i / PL_TILE);
  const uint32_t s = seg[(int64_t)c0 * ld + i];
  const bool flag = s >> 31;
  const uint32_t id = (s & 0x7fffffffu) + blk[c0 * ntiles + tile] - (flag ? 0u : 1u);   // run id of position i
  const int32_t p = order[(int64_t)c0 * ld + i];
  cluster[p] = (int32_t)id;
  const int32_t b = batch[p];
  if (flag) {
    idx_ptr[id] = (int32_t)i;
    head[id] = p;
    for (int c = 0; c < k; ++c) c_code[(int64_t)c * ld_c + id] = code[(int64_t)c * ld + p] >> shift;
    c_order[(int64_t)c0 * ld_c + id] = (int32_t)id;
    c_inverse[(int64_t)c0 * ld_c + id] = (int32_t)id;
    c_grid[3 * (int64_t)id + 0] = grid[3 * (int64_t)p + 0] >> pd;
    c_grid[3 * (int64_t)id + 1] = grid[3 * (int64_t)p + 1] >> pd;
    c_grid[3 * (int64_t)id + 2] = grid[3 * (int64_t)p + 2] >> pd;
    c_batch[id] = b;
  }
  // batch ids live in the top key bits, so scenes are contiguous in every curve's order
  if (i == n - 1) {
    idx_ptr[id + 1] = (int32_t)n;
    c_offset[b] = (int64_t)id + 1;
  } else {
    const int32_t bn = batch[order[(int64_t)c0 * ld + i + 1]];
    if (bn != b) c_offset[b] = (int64_t)id + 1;
  }
}

// C1: other curves: pooled order = cluster ids in first-visit order along that curve
__global__ void pool_write1_kernel(const int32_t* __restrict__ order, int64_t ld, const int32_t* __restrict__ n_dev,
                                   int64_t n_host, int c0, int ntiles, const uint32_t* __restrict__ seg,
                                   const uint32_t* __restrict__ blk, const int32_t* __restrict__ cluster,
                                   int32_t* __restrict__ c_order, int32_t* __restrict__ c_inverse, int64_t ld_c) {
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;
  const int c = blockIdx.y;
  if (c == c0) return;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = seg[(int64_t)c * ld + i];
  if (!(s >> 31)) return;
  const uint32_t rank = (s & 0x7fffffffu) + blk[c * ntiles + (int)(i / PL_TILE)];
  const int32_t cid = cluster[order[(int64_t)c * ld + i]];
  c_order[(int64_t)c * ld_c + rank] = cid;
  c_inverse[(int64_t)c * ld_c + cid] = (int32_t)rank;
}

CDSEG_API size_t cdseg_pool_plan_workspace_bytes(int k, int64_t ld) {
  const int64_t ntiles = (ld + PL_TILE - 1) / PL_TILE;
  return (size_t)k * ld * 4 + (size_t)k * ntiles * 4 + 256;
}

// See include/cdseg_b200.h for the contract.
CDSEG_API int cdseg_pool_plan(const int64_t* code, const int32_t* order, int k, int64_t ld, const int32_t* n_dev,
                              int64_t n_host, int c0, int pooling_depth, const int32_t* grid, const int32_t* batch,
                              int32_t* cluster, int32_t* idx_ptr, int32_t* head, int64_t* c_code, int32_t* c_order,
                              int32_t* c_inverse, int64_t ld_c, int32_t* c_grid, int32_t* c_batch, int32_t* m_dev,
                              int64_t* c_offset, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 0 || k > 8 || c0 < 0 || c0 >= k || pooling_depth < 0 || ld <= 0) return CDSEG_EINVAL;
  const int64_t cap = n_dev ? ld : n_host;          // launch bound; the true count may live on the device
  if (cap > ld || cap <= 0) return CDSEG_EINVAL;
  if (workspace_bytes < cdseg_pool_plan_workspace_bytes(k, ld)) return CDSEG_ENOSPC;
  const int ntiles = (int)((cap + PL_TILE - 1) / PL_TILE);
  uint32_t* seg = (uint32_t*)workspace;             // [k][ld]
  uint32_t* blk = (uint32_t*)((char*)workspace + (size_t)k * ld * 4);
  const int shift = 3 * pooling_depth;
  dim3 g(ntiles, k);
  pool_flag_kernel<<<g, PL_THREADS, 0, st>>>(code, order, ld, n_dev, n_host, shift, ntiles, seg, blk);
  pool_blkscan_kernel<<<k, PL_THREADS, 0, st>>>(blk, ntiles, c0, m_dev);
  pool_write0_kernel<<<cdseg_div_up(cap, 256), 256, 0, st>>>(code, order, ld, n_dev, n_host, k, c0, shift,
                                                             pooling_depth, ntiles, seg, blk, grid, batch, cluster,
                                                             idx_ptr, head, c_code, c_order, c_inverse, ld_c, c_grid,
                                                             c_batch, c_offset);
  CDSEG_COUNT_LAUNCH(3);
  if (k > 1) {
    dim3 g1(cdseg_div_up(cap, 256), k);
    pool_write1_kernel<<<g1, 256, 0, st>>>(order, ld, n_dev, n_host, c0, ntiles, seg, blk, cluster, c_order,
                                           c_inverse, ld_c);
    CDSEG_COUNT_LAUNCH(1);
  }
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// ---------------------------------------------------------------------------------
// feature pooling: out[m] = act(bn(max_{i in cluster m} x[i])), coord mean
// one warp per pooled point; lanes stride the channel dimension (coalesced rows)
// ---------------------------------------------------------------------------------
__global__ void pool_reduce_kernel(const float* __restrict__ x, const float* __restrict__ coord,
                                   const int32_t* __restrict__ members, const int32_t* __restrict__ idx_ptr,
                                   int64_t m, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                                   int gelu, float* __restrict__ out, float* __restrict__ out_coord) {
  pdl_trigger();
  pdl_wait();
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= m) return;
  const int beg = idx_ptr[w], end = idx_ptr[w + 1];
  for (int c = lane; c < C; c += 32) {
    float v = -INFINITY;
    for (int j = beg; j < end; ++j) v = fmaxf(v, x[(int64_t)members[j] * C + c]);
    if (scale) v = v * scale[c] + shift[c];
    if (gelu) v = gelu_erf(v);
    out[w * C + c] = v;
  }
  if (out_coord && lane < 3) {
    float s = 0.f;
    for (int j = beg; j < end; ++j) s += coord[(int64_t)members[j] * 3 + lane];
    out_coord[w * 3 + lane] = s / (float)(end - beg);
  }
}

CDSEG_API int cdseg_pool_reduce(const float* x, const float* coord, const int32_t* members, const int32_t* idx_ptr,
                                int64_t m, int C, const float* bn_scale, const float* bn_shift, int gelu, float* out,
                                float* out_coord, void* stream) {
  if (C <= 0) return CDSEG_EINVAL;
  if (m == 0) return CDSEG_OK;
  cdseg_launch_pdl(pool_reduce_kernel, dim3(cdseg_div_up(m * 32, 256)), dim3(256), 0, (cudaStream_t)stream, x, coord, members, idx_ptr, m, C,
                   bn_scale, bn_shift, gelu, out, out_coord);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// ---------------------------------------------------------------------------------
// unpool: out[i] = a[i] * alpha + b[cluster[i]]     (float4 rows)
// ---------------------------------------------------------------------------------
__global__ void unpool_add_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                  const int32_t* __restrict__ cluster, int64_t n, int C4, float alpha,
                                  float4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n * C4) return;
  const int64_t i = t / C4;
  const int c = (int)(t % C4);
  const float4 u = a[t], v = b[(int64_t)cluster[i] * C4 + c];
  out[t] = make_float4(u.x * alpha + v.x, u.y * alpha + v.y, u.z * alpha + v.z, u.w * alpha + v.w);
}

CDSEG_API int cdseg_unpool_add(const float* a, const float* b, const int32_t* cluster, int64_t n, int C, float alpha,
                               float* out, void* stream) {
  if (C <= 0 || (C & 3)) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  cdseg_launch_pdl(unpool_add_kernel, dim3(cdseg_div_up(n * (C / 4), 256)), dim3(256), 0, (cudaStream_t)stream, (const float4*)a,
                   (const float4*)b, cluster, n, C / 4, alpha, (float4*)out);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
