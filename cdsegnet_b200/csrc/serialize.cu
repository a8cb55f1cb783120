// Space-filling-curve serialization on B200: key encode (Morton / Hilbert, 4 curves
// in one pass), batched LSD radix argsort + inverse, and the patch slot maps.
//
// Replaces (reference file:line):
//   pointcept/models/utils/serialization/z_order.py:66-101     (LUT Morton)
//   pointcept/models/utils/serialization/hilbert.py:91-198     (byte-tensor Skilling)
//   pointcept/models/utils/serialization/default.py:9-24       (encode + batch bits)
//   pointcept/models/utils/structure.py:47-102                 (Point.serialization: argsort + scatter_ inverse)
//   pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:188-244 (pad / unpad maps)
//
// All kernels are HBM/L2-bound integer work: coalesced 32/64-bit accesses, one
// thread per element, keys staged through shared memory in the sort.
#include "common.cuh"
#include <cstdlib>

unsigned long long g_cdseg_launches = 0;
int g_cdseg_pdl = [] { const char* e = getenv("CDSEG_PDL"); return e ? atoi(e) : 1; }();        // common.cuh: programmatic dependent launch
CDSEG_API void cdseg_set_pdl(int on) { g_cdseg_pdl = on; }
CDSEG_API int cdseg_get_pdl(void) { return g_cdseg_pdl; }

CDSEG_API unsigned long long cdseg_launch_count(void) { return g_cdseg_launches; }
CDSEG_API void cdseg_launch_count_reset(void) { g_cdseg_launches = 0; }
CDSEG_API int cdseg_abi_version(void) { return 3; }

// ---------------------------------------------------------------------------------
// grid max (for serialized_depth = bit_length(max), structure.py:66)
// ---------------------------------------------------------------------------------
__global__ void grid_max_kernel(const int32_t* __restrict__ g, int64_t n, int32_t* __restrict__ out) {
  int m = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, g[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

CDSEG_API int cdseg_grid_max(const int32_t* grid, int64_t n_elems, int32_t* out_max, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out_max, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  if (n_elems > 0) {
    int64_t want = (n_elems + 1023) / 1024; int blocks = (int)(want < 148 * 4 ? want : 148 * 4);
    grid_max_kernel<<<blocks, 256, 0, st>>>(grid, n_elems, out_max);
    CDSEG_COUNT_LAUNCH(1);
  }
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// ---------------------------------------------------------------------------------
// offset -> batch (utils/misc.py:19-24)
// ---------------------------------------------------------------------------------
__global__ void offset2batch_kernel(const int64_t* __restrict__ offset, int B, int64_t N, int32_t* __restrict__ batch) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  int lo = 0, hi = B - 1;            // first b with offset[b] > i
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (offset[mid] > i) hi = mid; else lo = mid + 1;
  }
  batch[i] = lo;
}

CDSEG_API int cdseg_offset2batch(const int64_t* offset, int B, int64_t N, int32_t* batch, void* stream) {
  if (B <= 0) return CDSEG_EINVAL;
  if (N == 0) return CDSEG_OK;
  offset2batch_kernel<<<cdseg_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(offset, B, N, batch);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// ---------------------------------------------------------------------------------
// key encode
// ---------------------------------------------------------------------------------
// spread the low 16 bits of v so that bit i lands at bit 3i
__host__ __device__ __forceinline__ uint64_t spread3(uint32_t v) {
  uint64_t x = v & 0xffffu;
  x = (x | (x << 16)) & 0x0000ff0000ffull;   // magic-number bit spread for 16 -> 48 bits
  x = (x | (x << 8)) & 0x00f00f00f00full;
  x = (x | (x << 4)) & 0x0c30c30c30c3ull;
  x = (x | (x << 2)) & 0x249249249249ull;
  return x;
}
// x -> bit 3i+2, y -> 3i+1, z -> 3i (z_order.py:40-50)
__host__ __device__ __forceinline__ uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
  return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
}
// Skilling transpose -> Hilbert index, bit-exact with hilbert.py:91-198
__host__ __device__ __forceinline__ uint64_t hilbert3(uint32_t x, uint32_t y, uint32_t z, int depth) {
  uint32_t X[3] = {x, y, z};
  for (int b = depth - 1; b >= 0; --b) {       // MSB -> LSB of the depth-bit window
    const uint32_t q = 1u << b, low = q - 1u;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (X[d] & q) {
        X[0] ^= low;                           // bit set: invert lower bits of dim 0
      } else {
        const uint32_t t = (X[0] ^ X[d]) & low; // bit clear: exchange differing lower bits
        X[0] ^= t; X[d] ^= t;
      }
    }
  }
  uint64_t g = morton3(X[0], X[1], X[2]);       // bit-major interleave
  g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8; g ^= g >> 16; g ^= g >> 32;   // Gray -> binary
  return g;
}

// order ids: 0 = "z", 1 = "z-trans", 2 = "hilbert", 3 = "hilbert-trans"
struct OrderIds { int id[8]; };

__global__ void encode_kernel(const int32_t* __restrict__ grid, const int32_t* __restrict__ batch, int64_t N,
                              int depth, OrderIds ids, int k, int64_t* __restrict__ codes) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t mask = (depth >= 32) ? 0xffffffffu : ((1u << depth) - 1u);
  const uint32_t x = (uint32_t)grid[3 * i + 0] & mask, y = (uint32_t)grid[3 * i + 1] & mask,
                 z = (uint32_t)grid[3 * i + 2] & mask;
  const uint64_t hi = batch ? ((uint64_t)(uint32_t)batch[i] << (3 * depth)) : 0ull;
  for (int r = 0; r < k; ++r) {
    uint64_t c;
    switch (ids.id[r]) {
      case 0: c = morton3(x, y, z); break;
      case 1: c = morton3(y, x, z); break;
      case 2: c = hilbert3(x, y, z, depth); break;
      default: c = hilbert3(y, x, z, depth); break;
    }
    codes[(int64_t)r * N + i] = (int64_t)(hi | c);
  }
}

// Host-side evaluation of the very same __host__ __device__ bit routines, for the CPU
// ("not gpu") test-suite only: lets the bit twiddling be checked against the golden
// vectors without a GPU.  Never called by the product path.
CDSEG_API int cdseg_debug_encode_host(const int32_t* grid, const int32_t* batch, int64_t N, int depth, int order_id,
                                      int64_t* codes) {
  if (depth < 0 || depth > 16 || order_id < 0 || order_id > 3) return CDSEG_EINVAL;
  const uint32_t mask = (1u << depth) - 1u;
  for (int64_t i = 0; i < N; ++i) {
    const uint32_t x = (uint32_t)grid[3 * i] & mask, y = (uint32_t)grid[3 * i + 1] & mask, z = (uint32_t)grid[3 * i + 2] & mask;
    uint64_t c = order_id == 0 ? morton3(x, y, z) : order_id == 1 ? morton3(y, x, z)
               : order_id == 2 ? hilbert3(x, y, z, depth) : hilbert3(y, x, z, depth);
    codes[i] = (int64_t)((batch ? ((uint64_t)(uint32_t)batch[i] << (3 * depth)) : 0ull) | c);
  }
  return CDSEG_OK;
}

CDSEG_API int cdseg_encode_codes(const int32_t* grid, const int32_t* batch, int64_t N, int depth,
                                 const int* order_ids, int k, int64_t* codes, void* stream) {
  if (k <= 0 || k > 8 || depth < 0 || depth > 16) return CDSEG_EINVAL;
  OrderIds ids;
  for (int r = 0; r < k; ++r) {
    if (order_ids[r] < 0 || order_ids[r] > 3) return CDSEG_EINVAL;
    ids.id[r] = order_ids[r];
  }
  if (N == 0) return CDSEG_OK;
  encode_kernel<<<cdseg_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(grid, batch, N, depth, ids, k, codes);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// ---------------------------------------------------------------------------------
// batched LSD radix argsort (8-bit digits, stable), rows = curves
// ---------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;                 // items per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;

// per-tile digit histogram: hist[row][bin][tile]
__global__ void rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t N, int64_t row_stride, int shift,
                               int ntiles, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  const int row = blockIdx.y, tile = blockIdx.x;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t* k = keys + (int64_t)row * row_stride;
  const int64_t base = (int64_t)tile * RS_TILE;
#pragma unroll 4
  for (int j = 0; j < RS_ITEMS; ++j) {
    int64_t i = base + j * RS_THREADS + threadIdx.x;
    if (i < N) atomicAdd(&h[(k[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[((int64_t)row * 256 + threadIdx.x) * ntiles + tile] = h[threadIdx.x];
}

// exclusive scan over (bin-major, tile-minor) for each row; one block per row
__global__ void rs_scan_kernel(uint32_t* __restrict__ hist, int ntiles) {
  __shared__ uint32_t tot[256];
  const int row = blockIdx.x, bin = threadIdx.x;
  uint32_t* h = hist + ((int64_t)row * 256 + bin) * ntiles;
  uint32_t s = 0;
  for (int t = 0; t < ntiles; ++t) { uint32_t v = h[t]; h[t] = s; s += v; }
  tot[bin] = s;
  __syncthreads();
  // exclusive scan of the 256 bin totals (single warp, 8 per lane)
  if (bin < 32) {
    uint32_t loc[8], run = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { loc[j] = run; run += tot[bin * 8 + j]; }
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (bin >= o) inc += n; }
    uint32_t excl = inc - run;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot[bin * 8 + j] = excl + loc[j];
  }
  __syncthreads();
  const uint32_t b = tot[bin];
  for (int t = 0; t < ntiles; ++t) h[t] += b;
}

// stable scatter.  vals_in == nullptr means "value = element index" (first pass).
// On the last pass also writes inverse[val] = position.
__global__ void rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                  uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                  int32_t* __restrict__ inverse, int64_t N, int64_t in_stride, int64_t out_stride,
                                  int shift, int ntiles, const uint32_t* __restrict__ hist) {
  __shared__ uint32_t cnt[RS_WARPS][256];
  __shared__ uint32_t gbase[256];
  const int row = blockIdx.y, tile = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < RS_WARPS * 256; j += RS_THREADS) (&cnt[0][0])[j] = 0;
  gbase[threadIdx.x] = hist[((int64_t)row * 256 + threadIdx.x) * ntiles + tile];
  __syncthreads();
  const uint64_t* kin = keys_in + (int64_t)row * in_stride;
  const uint32_t* vin = vals_in ? vals_in + (int64_t)row * out_stride : nullptr;
  // warp w owns the contiguous range [base + w*512, base + (w+1)*512), walked in 16 rounds of 32
  const int64_t wbase = (int64_t)tile * RS_TILE + (int64_t)warp * (32 * RS_ITEMS);
  uint64_t key[RS_ITEMS];
  uint32_t rank[RS_ITEMS];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const int64_t i = wbase + j * 32 + lane;
    const bool ok = i < N;
    key[j] = ok ? kin[i] : ~0ull;
    const uint32_t d = ok ? (uint32_t)((key[j] >> shift) & 255u) : 256u + lane;   // invalid lanes never match
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (ok && lane == leader) { old = cnt[warp][d]; cnt[warp][d] = old + __popc(peers); }
    __syncwarp();
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[j] = old + __popc(peers & lt);
  }
  __syncthreads();
  {  // exclusive prefix over warps for each digit, then add the global tile base
    const int d = threadIdx.x;
    uint32_t s = gbase[d];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) { uint32_t v = cnt[w][d]; cnt[w][d] = s; s += v; }
  }
  __syncthreads();
  uint64_t* kout = keys_out + (int64_t)row * out_stride;
  uint32_t* vout = vals_out + (int64_t)row * out_stride;
  int32_t* inv = inverse ? inverse + (int64_t)row * out_stride : nullptr;
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const int64_t i = wbase + j * 32 + lane;
    if (i < N) {
      const uint32_t d = (uint32_t)((key[j] >> shift) & 255u);
      const uint32_t pos = cnt[warp][d] + rank[j];
      const uint32_t v = vin ? vin[i] : (uint32_t)i;
      kout[pos] = key[j];
      vout[pos] = v;
      if (inv) inv[v] = (int32_t)pos;
    }
  }
}

__global__ void iota_rows_kernel(int32_t* __restrict__ order, int32_t* __restrict__ inverse, int64_t total, int64_t N) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < total) { int32_t v = (int32_t)(i % N); order[i] = v; inverse[i] = v; }
}

CDSEG_API size_t cdseg_argsort_workspace_bytes(int k, int64_t N) {
  const int64_t ntiles = (N + RS_TILE - 1) / RS_TILE;
  size_t keys = (size_t)k * N * sizeof(uint64_t) * 2;
  size_t vals = (size_t)k * N * sizeof(uint32_t) * 2;
  size_t hist = (size_t)k * 256 * ntiles * sizeof(uint32_t);
  return keys + vals + hist + 256;
}

// codes [k,N] int64 compared as UNSIGNED 64-bit keys, nbits = number of significant key bits (<= 64).
// order/inverse: int32 [k,N].  codes are left untouched.
CDSEG_API int cdseg_argsort_rows(const int64_t* codes, int k, int64_t N, int nbits, int32_t* order,
                                 int32_t* inverse, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 0 || nbits < 0 || nbits > 64 || N >= (1ll << 31)) return CDSEG_EINVAL;   // keys are compared as uint64
  if (N == 0) return CDSEG_OK;
  if (workspace_bytes < cdseg_argsort_workspace_bytes(k, N)) return CDSEG_ENOSPC;
  const int ntiles = (int)((N + RS_TILE - 1) / RS_TILE);
  uint64_t* kbuf[2];
  uint32_t* vbuf[2];
  char* p = (char*)workspace;
  kbuf[0] = (uint64_t*)p; p += (size_t)k * N * 8;
  kbuf[1] = (uint64_t*)p; p += (size_t)k * N * 8;
  vbuf[0] = (uint32_t*)p; p += (size_t)k * N * 4;
  vbuf[1] = (uint32_t*)p; p += (size_t)k * N * 4;
  uint32_t* hist = (uint32_t*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  const int passes = (nbits + 7) / 8;
  if (passes == 0) {   // all keys equal: identity permutation
    iota_rows_kernel<<<cdseg_div_up((int64_t)k * N, 256), 256, 0, st>>>(order, inverse, (int64_t)k * N, N);
    CDSEG_COUNT_LAUNCH(1);
    CDSEG_LAUNCH_CHECK();
    return CDSEG_OK;
  }
  dim3 grid(ntiles, k);
  const uint64_t* kin = (const uint64_t*)codes;
  const uint32_t* vin = nullptr;
  for (int pass = 0; pass < passes; ++pass) {
    const bool last = pass == passes - 1;
    uint64_t* kout = kbuf[pass & 1];
    uint32_t* vout = last ? (uint32_t*)order : vbuf[pass & 1];
    rs_hist_kernel<<<grid, RS_THREADS, 0, st>>>(kin, N, N, pass * 8, ntiles, hist);
    rs_scan_kernel<<<k, 256, 0, st>>>(hist, ntiles);
    rs_scatter_kernel<<<grid, RS_THREADS, 0, st>>>(kin, vin, kout, vout, last ? inverse : nullptr, N, N, N,
                                                   pass * 8, ntiles, hist);
    CDSEG_COUNT_LAUNCH(3);
    kin = kout;
    vin = vout;
  }
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// ---------------------------------------------------------------------------------
// patch slot maps (ptv3.py:188-244 restated for the packed patch layout)
//
// A scene b with n_b sorted points is cut into patches of K slots (a scene with
// n_b <= K is a single patch of n_b slots).  In the *packed* layout every patch
// owns Kp = round_up(K,128) slots, patch t slot j -> packed slot t*Kp + j.
//   slot_src[t*Kp+j]  = point index feeding that slot (order[pad[.]] in reference
//                       terms: the last patch's filler slots replay the tail of
//                       the previous patch), or -1 for j >= patch_len[t]
//   slot_dst[t*Kp+j]  = slot_src if the slot is that point's OWN slot (the one `inverse`
//                       maps to), else -1: where an attention output row is scattered to
//   point_slot[i]     = packed slot of point i (reference: unpad[inverse[i]])
// ---------------------------------------------------------------------------------
struct SceneTab { int64_t start[64]; int64_t count[64]; int32_t patch0[64]; };

__global__ void patch_maps_kernel(const int32_t* __restrict__ order, int B, SceneTab tab, int K, int Kp, int T,
                                  int32_t* __restrict__ slot_src, int32_t* __restrict__ slot_dst,
                                  int32_t* __restrict__ point_slot, int32_t* __restrict__ patch_len) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= (int64_t)T * Kp) return;
  const int t = (int)(p / Kp), j = (int)(p % Kp);
  int b = 0;
  while (b + 1 < B && tab.patch0[b + 1] <= t) ++b;
  const int64_t n = tab.count[b], s = tab.start[b];
  const int tl = t - tab.patch0[b];                       // patch index inside the scene
  int len;
  int64_t pos = -1;                                       // sorted position feeding the slot
  bool real = false;
  if (n <= K) {
    len = (int)n;
    if (j < len) { pos = j; real = true; }
  } else {
    len = K;
    if (j < K) {
      const int64_t jj = (int64_t)tl * K + j;             // padded slot inside the scene
      real = jj < n;
      pos = real ? jj : jj - K;
    }
  }
  if (j == 0) patch_len[t] = len;
  int32_t src = -1;
  if (pos >= 0) {
    src = order[s + pos];
    if (real) point_slot[src] = (int32_t)p;
  }
  slot_src[p] = src;
  slot_dst[p] = real ? src : -1;
}

// scene_start/scene_count: host arrays [B] (cumulative offsets are host-known at this point).
// returns the number of patches through *T_out when slot_src == nullptr (sizing query).
CDSEG_API int cdseg_patch_count(const int64_t* scene_count, int B, int K, int* T_out) {
  if (B <= 0 || B > 64 || K <= 0) return CDSEG_EINVAL;
  int T = 0;
  for (int b = 0; b < B; ++b) T += scene_count[b] <= K ? 1 : (int)((scene_count[b] + K - 1) / K);
  *T_out = T;
  return CDSEG_OK;
}

CDSEG_API int cdseg_patch_maps(const int32_t* order, const int64_t* scene_count, int B, int K, int Kp,
                               int32_t* slot_src, int32_t* slot_dst, int32_t* point_slot, int32_t* patch_len,
                               void* stream) {
  if (B <= 0 || B > 64 || K <= 0 || Kp < K || (Kp % 128) != 0) return CDSEG_EINVAL;
  SceneTab tab;
  int T = 0;
  int64_t s = 0;
  for (int b = 0; b < B; ++b) {
    tab.start[b] = s; tab.count[b] = scene_count[b]; tab.patch0[b] = T;
    s += scene_count[b];
    T += scene_count[b] <= K ? 1 : (int)((scene_count[b] + K - 1) / K);
  }
  const int64_t total = (int64_t)T * Kp;
  if (total == 0) return CDSEG_OK;
  patch_maps_kernel<<<cdseg_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(order, B, tab, K, Kp, T, slot_src,
                                                                                slot_dst, point_slot, patch_len);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// ---------------------------------------------------------------------------------
// curve-order renumbering of level 0 (structure.py::Plan) and plain row gathers: the "patch gather/scatter" plumbing around the
// network -- feat[perm], coord[perm], logits[inv_perm] -- as coalesced kernels instead of a dozen advanced-indexing launches.
// ---------------------------------------------------------------------------------
// dst[i, :] = src[idx[i], :] for rows of `words` 32-bit words (a thread moves one word; consecutive threads walk a row, then the next row)
__global__ void gather_rows_kernel(const uint32_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t n, int words,
                                   uint32_t* __restrict__ dst) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n * words) return;
  const int64_t i = e / words;
  const int w = (int)(e - i * words);
  dst[e] = src[(int64_t)idx[i] * words + w];
}

CDSEG_API int cdseg_gather_rows(const void* src, const int32_t* idx, int64_t n, int row_bytes, void* dst, void* stream) {
  if (n < 0 || row_bytes <= 0 || (row_bytes & 3) || !src || !idx || !dst) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  const int words = row_bytes / 4;
  gather_rows_kernel<<<cdseg_div_up(n * words, 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)src, idx, n, words, (uint32_t*)dst);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// internal point r = original point perm[r]:  i_grid[r] = grid[perm[r]], i_batch[r] = batch[perm[r]], and for every curve c
// i_code[c][r] = code[c][perm[r]], i_inverse[c][r] = inverse[c][perm[r]], i_order[c][j] = inv_perm[order[c][j]]
__global__ void renumber_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ inv_perm, const int32_t* __restrict__ grid,
                                const int32_t* __restrict__ batch, const int64_t* __restrict__ code, const int32_t* __restrict__ order,
                                const int32_t* __restrict__ inverse, int k, int64_t N, int32_t* __restrict__ i_grid,
                                int32_t* __restrict__ i_batch, int64_t* __restrict__ i_code, int32_t* __restrict__ i_order,
                                int32_t* __restrict__ i_inverse) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int64_t p = perm[r];
  i_grid[3 * r] = grid[3 * p]; i_grid[3 * r + 1] = grid[3 * p + 1]; i_grid[3 * r + 2] = grid[3 * p + 2];
  i_batch[r] = batch[p];
  for (int c = 0; c < k; ++c) {
    i_code[c * N + r] = code[c * N + p];
    i_inverse[c * N + r] = inverse[c * N + p];
    i_order[c * N + r] = inv_perm[order[c * N + r]];
  }
}

// dst[i, 0..C) = src[idx[i], 0..C), dst[i, C..Cp) = 0: the stem's input rows in internal numbering, zero-padded to the 8 channels the
// im2col GEMM consumes (replaces feat[perm] + F.pad in front of cdseg_conv_im2col_tc)
__global__ void gather_rows_pad_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int64_t n, int C, int Cp,
                                       float* __restrict__ dst) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n * Cp) return;
  const int64_t r = t / Cp;
  const int c = (int)(t % Cp);
  dst[t] = c < C ? src[(int64_t)idx[r] * C + c] : 0.f;
}
CDSEG_API int cdseg_gather_rows_pad(const float* src, const int32_t* idx, int64_t n, int C, int Cp, float* dst, void* stream) {
  if (n < 0 || C <= 0 || Cp < C || !src || !idx || !dst) return CDSEG_EINVAL;
  if (n == 0) return CDSEG_OK;
  gather_rows_pad_kernel<<<cdseg_div_up(n * Cp, 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n, C, Cp, dst);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_renumber(const int32_t* perm, const int32_t* inv_perm, const int32_t* grid, const int32_t* batch, const int64_t* code,
                             const int32_t* order, const int32_t* inverse, int k, int64_t N, int32_t* i_grid, int32_t* i_batch,
                             int64_t* i_code, int32_t* i_order, int32_t* i_inverse, void* stream) {
  if (k <= 0 || N < 0 || !perm || !inv_perm || !grid || !batch || !code || !order || !inverse || !i_grid || !i_batch || !i_code || !i_order ||
      !i_inverse)
    return CDSEG_EINVAL;
  if (N == 0) return CDSEG_OK;
  renumber_kernel<<<cdseg_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(perm, inv_perm, grid, batch, code, order, inverse, k, N, i_grid,
                                                                          i_batch, i_code, i_order, i_inverse);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
