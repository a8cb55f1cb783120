// pointops.knn_query on B200: the one other native op the CDSegNet evaluator reaches
// (pointcept/engines/hooks/evaluator.py:132-141 maps voxel predictions back to the original points with k = 1).
//
// Reference: libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:60-104 -- one thread per query scans EVERY point of
// its batch (O(m n)) keeping a max-heap of the nsample best squared distances, heap-sorts them ascending.
//
// Here: exact k-NN over a uniform cell grid.  (1) bounding box + cell size chosen on the device (no host sync);
// (2) cell key (batch, cx, cy, cz) per point, 64-bit radix argsort (serialize.cu), points copied into cell order so a
// cell is one contiguous, coalesced run; (3) open-addressing hash cell -> [start, end); (4) one thread per query walks
// Chebyshev shells of cells around its own (clamped) cell and stops once the k-th best distance is closer than
// anything an unvisited shell can hold.  Distances use the reference's expression, ties go to the lower point index
// (what the reference's strict `<` yields for k = 1), results ascending like its heap sort, missing neighbours
// (batch smaller than nsample) are idx = -1, dist2 = 1e10 like its initial values.
#include <float.h>
#include "common.cuh"
#include "../../include/cdseg_b200.h"

namespace kn {

constexpr uint64_t EMPTY = ~0ull;

struct Grid {          // device-resident description written by grid_setup_kernel
  float lo[3];
  float h, inv_h;
  int dim[3];
};

__device__ __forceinline__ uint64_t cell_key(int b, int x, int y, int z) {
  return ((uint64_t)(uint32_t)b << 48) | ((uint64_t)(uint32_t)x << 32) | ((uint64_t)(uint32_t)y << 16) | (uint64_t)(uint32_t)z;
}
__device__ __forceinline__ uint32_t hash64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k;
}
__device__ __forceinline__ unsigned f2ord(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// bb[0..2] = min, bb[3..5] = max as order-preserving unsigned ints
__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ xyz, int64_t n, unsigned* __restrict__ bb) {
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const unsigned u = f2ord(xyz[3 * i + a]);
      lo[a] = min(lo[a], u); hi[a] = max(hi[a], u);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) { lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(bb + a, lo[a]); atomicMax(bb + 3 + a, hi[a]); }
  }
}

// cell size: about `target` points per occupied cell for points on surfaces (occupied cells ~ area / h^2), never more
// than 60000 cells per axis
__global__ void grid_setup_kernel(const unsigned* __restrict__ bb, int64_t n, int B, float target, Grid* __restrict__ g) {
  float ext = 0.f;
  for (int a = 0; a < 3; ++a) { g->lo[a] = ord2f(bb[a]); ext = fmaxf(ext, ord2f(bb[3 + a]) - ord2f(bb[a])); }
  const float per_batch = fmaxf((float)n / (float)max(B, 1), 1.0f);
  float h = ext * sqrtf(target / per_batch);
  h = fmaxf(h, ext / 60000.0f);
  if (!(h > 0.f)) h = 1.0f;                                  // all points coincide
  g->h = h; g->inv_h = 1.0f / h;
  for (int a = 0; a < 3; ++a) g->dim[a] = min(65535, (int)((ord2f(bb[3 + a]) - g->lo[a]) * g->inv_h) + 1);
}

__device__ __forceinline__ int cell_of(float p, float lo, float inv_h, int dim) {
  const int c = (int)floorf((p - lo) * inv_h);
  return min(max(c, 0), dim - 1);
}

__global__ void cell_key_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ offset, int B, int64_t n, const Grid* __restrict__ g,
                                int64_t* __restrict__ key) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b = 0;
  while (b < B - 1 && i >= offset[b]) ++b;
  key[i] = (int64_t)cell_key(b, cell_of(xyz[3 * i], g->lo[0], g->inv_h, g->dim[0]), cell_of(xyz[3 * i + 1], g->lo[1], g->inv_h, g->dim[1]),
                             cell_of(xyz[3 * i + 2], g->lo[2], g->inv_h, g->dim[2]));
}

// points in cell order (x, y, z, original index); run heads insert cell -> [start, end)
__global__ void cell_table_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ key, const int32_t* __restrict__ order, int64_t n,
                                  float4* __restrict__ sorted, uint64_t* __restrict__ hkeys, int2* __restrict__ hvals, uint32_t cap_mask) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int i = order[j];
  sorted[j] = make_float4(xyz[3 * (int64_t)i], xyz[3 * (int64_t)i + 1], xyz[3 * (int64_t)i + 2], __int_as_float(i));
  const uint64_t k = (uint64_t)key[i];
  if (j > 0 && (uint64_t)key[order[j - 1]] == k) return;
  int64_t e = j + 1;
  while (e < n && (uint64_t)key[order[e]] == k) ++e;
  uint32_t slot = hash64(k) & cap_mask;
  while (true) {
    const uint64_t prev = atomicCAS((unsigned long long*)&hkeys[slot], (unsigned long long)EMPTY, (unsigned long long)k);
    if (prev == EMPTY) { hvals[slot] = make_int2((int)j, (int)e); return; }
    slot = (slot + 1) & cap_mask;
  }
}

// (d2, idx) lexicographic "worse than"
__device__ __forceinline__ bool worse(float d1, int i1, float d2, int i2) { return d1 > d2 || (d1 == d2 && i1 > i2); }

template <int KMAX>
__global__ void __launch_bounds__(128) knn_kernel(const float* __restrict__ new_xyz, const int32_t* __restrict__ offset,
                                                  const int32_t* __restrict__ new_offset, int B, int64_t m, int k, const Grid* __restrict__ gp,
                                                  const float4* __restrict__ sorted, const uint64_t* __restrict__ hkeys,
                                                  const int2* __restrict__ hvals, uint32_t cap_mask, int32_t* __restrict__ idx,
                                                  float* __restrict__ dist2) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const Grid g = *gp;
  int b = 0;
  while (b < B - 1 && q >= new_offset[b]) ++b;
  const int n_b = offset[b] - (b ? offset[b - 1] : 0);
  const float qx = new_xyz[3 * q], qy = new_xyz[3 * q + 1], qz = new_xyz[3 * q + 2];
  const int cx = cell_of(qx, g.lo[0], g.inv_h, g.dim[0]), cy = cell_of(qy, g.lo[1], g.inv_h, g.dim[1]), cz = cell_of(qz, g.lo[2], g.inv_h, g.dim[2]);
  float bd[KMAX];                      // max-heap on (d2, idx): bd[0] / bi[0] is the worst kept candidate
  int bi[KMAX];
  for (int j = 0; j < k; ++j) { bd[j] = 1e10f; bi[j] = INT_MAX; }
  const int want = min(k, n_b);
  int found = 0;
  const int rmax = max(g.dim[0], max(g.dim[1], g.dim[2]));
  for (int r = 0; r <= rmax; ++r) {
    if (found >= want) {
      if (want < k) break;                                                       // the whole batch is already in the heap
      // shells 0..r-1 are done: an unvisited cell is >= r cells away, its points >= (r - 1) h (minus rounding slack)
      const float reach = (float)(r - 1) * g.h * 0.99999f;
      if (r > 0 && bd[0] < reach * reach) break;
    }
    for (int dz = -r; dz <= r; ++dz) {
      const int z = cz + dz;
      if (z < 0 || z >= g.dim[2]) continue;
      for (int dy = -r; dy <= r; ++dy) {
        const int y = cy + dy;
        if (y < 0 || y >= g.dim[1]) continue;
        const bool face = (dz == -r || dz == r || dy == -r || dy == r);
        const int step = (face || r == 0) ? 1 : 2 * r;                          // interior rows of the shell: only the two end cells
        for (int dx = -r; dx <= r; dx += step) {
          const int x = cx + dx;
          if (x < 0 || x >= g.dim[0]) continue;
          const uint64_t key = cell_key(b, x, y, z);
          uint32_t slot = hash64(key) & cap_mask;
          int2 run = make_int2(0, 0);
          while (true) {
            const uint64_t cur = hkeys[slot];
            if (cur == key) { run = hvals[slot]; break; }
            if (cur == EMPTY) break;
            slot = (slot + 1) & cap_mask;
          }
          for (int j = run.x; j < run.y; ++j) {
            const float4 p = sorted[j];
            const float x0 = p.x, y0 = p.y, z0 = p.z;
            const float d2 = (qx - x0) * (qx - x0) + (qy - y0) * (qy - y0) + (qz - z0) * (qz - z0);   // the reference's expression
            const int pi = __float_as_int(p.w);
            if (KMAX == 1) {
              if (worse(bd[0], bi[0], d2, pi)) { bd[0] = d2; bi[0] = pi; found = 1; }
            } else if (worse(bd[0], bi[0], d2, pi)) {
              bd[0] = d2; bi[0] = pi;
              if (found < k) ++found;
              int root = 0;                                                       // sift down
              while (true) {
                int child = 2 * root + 1;
                if (child >= k) break;
                if (child + 1 < k && worse(bd[child + 1], bi[child + 1], bd[child], bi[child])) ++child;
                if (!worse(bd[child], bi[child], bd[root], bi[root])) break;
                const float td = bd[root]; bd[root] = bd[child]; bd[child] = td;
                const int ti = bi[root]; bi[root] = bi[child]; bi[child] = ti;
                root = child;
              }
            }
          }
        }
      }
    }
  }
  // ascending (d2, idx): selection from the heap arrays (k is small)
  for (int a = 0; a < k; ++a) {
    int best = a;
    for (int c = a + 1; c < k; ++c)
      if (worse(bd[best], bi[best], bd[c], bi[c])) best = c;
    const float td = bd[a]; bd[a] = bd[best]; bd[best] = td;
    const int ti = bi[a]; bi[a] = bi[best]; bi[best] = ti;
    idx[q * k + a] = bi[a] == INT_MAX ? -1 : bi[a];
    dist2[q * k + a] = bd[a];
  }
}

}  // namespace kn

static int64_t knn_hash_capacity(int64_t n) {
  int64_t c = 1024;
  while (c < 2 * n) c <<= 1;
  return c;
}

CDSEG_API size_t cdseg_knn_workspace_bytes(int64_t n) {
  return 256 + (size_t)n * 8 + (size_t)n * 4 * 2 + (size_t)n * 16 + (size_t)knn_hash_capacity(n) * 16 + cdseg_argsort_workspace_bytes(1, n) + 1024;
}

// see include/cdseg_b200.h
CDSEG_API int cdseg_knn_query(int m, int nsample, const float* xyz, const float* new_xyz, const int32_t* offset, const int32_t* new_offset,
                              int B, int64_t n, int32_t* idx, float* dist2, void* workspace, size_t workspace_bytes, void* stream) {
  if (m < 0 || nsample <= 0 || nsample > 128 || n < 0 || B <= 0 || B >= 65536 || !xyz || !new_xyz || !offset || !new_offset || !idx || !dist2 ||
      !workspace)
    return CDSEG_EINVAL;
  if (m == 0) return CDSEG_OK;
  if (workspace_bytes < cdseg_knn_workspace_bytes(n)) return CDSEG_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  char* p = (char*)workspace;
  unsigned* bb = (unsigned*)p;
  kn::Grid* g = (kn::Grid*)(p + 64); p += 256;
  int64_t* key = (int64_t*)p; p += (size_t)n * 8;
  int32_t* order = (int32_t*)p; p += (size_t)n * 4;
  int32_t* inverse = (int32_t*)p; p += (size_t)n * 4;
  p = (char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  float4* sorted = (float4*)p; p += (size_t)n * 16;
  const int64_t cap = knn_hash_capacity(n);
  uint64_t* hkeys = (uint64_t*)p; p += (size_t)cap * 8;
  int2* hvals = (int2*)p; p += (size_t)cap * 8;
  p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  const size_t sort_ws = (size_t)((char*)workspace + workspace_bytes - p);
  const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  cudaMemcpyAsync(bb, init, sizeof(init), cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(hkeys, 0xff, (size_t)cap * 8, st);
  if (n > 0) {
    const int blocks = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    kn::bbox_kernel<<<blocks, 256, 0, st>>>(xyz, n, bb);
  }
  kn::grid_setup_kernel<<<1, 1, 0, st>>>(bb, n, B, 4.0f, g);
  CDSEG_COUNT_LAUNCH(2);
  if (n > 0) {
    kn::cell_key_kernel<<<cdseg_div_up(n, 256), 256, 0, st>>>(xyz, offset, B, n, g, key);
    int s = cdseg_argsort_rows(key, 1, n, 64, order, inverse, p, sort_ws, stream);
    if (s != CDSEG_OK) return s;
    kn::cell_table_kernel<<<cdseg_div_up(n, 256), 256, 0, st>>>(xyz, key, order, n, sorted, hkeys, hvals, (uint32_t)(cap - 1));
    CDSEG_COUNT_LAUNCH(2);
  }
  const int grid = cdseg_div_up(m, 128);
  if (nsample == 1)
    kn::knn_kernel<1><<<grid, 128, 0, st>>>(new_xyz, offset, new_offset, B, m, nsample, g, sorted, hkeys, hvals, (uint32_t)(cap - 1), idx, dist2);
  else if (nsample <= 16)
    kn::knn_kernel<16><<<grid, 128, 0, st>>>(new_xyz, offset, new_offset, B, m, nsample, g, sorted, hkeys, hvals, (uint32_t)(cap - 1), idx, dist2);
  else
    kn::knn_kernel<128><<<grid, 128, 0, st>>>(new_xyz, offset, new_offset, B, m, nsample, g, sorted, hkeys, hvals, (uint32_t)(cap - 1), idx, dist2);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
