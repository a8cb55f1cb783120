// Patch gather ("qkv[order]") for serialized attention, plus the exact fp32 SIMT
// attention used as the high-precision mode.
//
// Reference: pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py
//   :258-262  order = serialized_order[k][pad]; qkv = self.qkv(feat)[order]
//   :264-280  dense branch: softmax((q*scale) @ k^T) @ v per (patch, head)     <- fp32 semantics
//   :282-289  flash branch: qkv.half(), varlen over cu_seqlens                 <- fp16 semantics (attn_tc.cu)
//   :290      feat = feat[inverse]
//   :1003-1047 cross attention: q rows by q_order, kv rows by kv_order[q_pad]
//
// Packed layout (per tensor Q, K, V): [H][T][Kp][16] where the innermost [Kp][16] block of
// one (head, patch) is stored as 8x8 "core matrices": element (r, d) lives at
//   (r/8)*128 + (d/8)*64 + (r%8)*8 + (d%8)          (in elements)
// i.e. exactly the shared-memory image tcgen05.mma's no-swizzle descriptors want
// (K-major for Q and K, MN-major for V), so the attention kernel stages whole tiles
// with plain 1-D bulk copies.  Slots with slot_src < 0 are zero rows.
#include "common.cuh"

CDSEG_API int cdseg_attn_pack_f16v(const float* src, int64_t ld, int col0, int C, int nwhich, const int32_t* slot_src,
                                   int H, int T, int Kp, void* dst0, void* dst1, void* dst2, int v_ones, void* stream);

// src: fp32 rows [n, ld] ; column block for (which, h) starts at col0 + which*C + h*16
// v_ones_which >= 0: that tensor (V) is written 32 wide per key: [v(16) | 1 | 0 x 15], i.e. the MN-major B operand of
// ONE tcgen05.mma that yields P.V and the row sum P.1 together (block layout (k/8)*256 + (n/8)*64 + (k%8)*8 + n%8 elements).
template <typename OutT>
__global__ void pack_heads_kernel(const float* __restrict__ src, int64_t ld, int col0, int C, int nwhich,
                                  const int32_t* __restrict__ slot_src, int H, int T, int Kp,
                                  OutT* __restrict__ dst0, OutT* __restrict__ dst1, OutT* __restrict__ dst2,
                                  int v_ones_which = -1) {
  pdl_trigger();
  pdl_wait();
  // consecutive threads -> consecutive slots (16-byte stores of 8 consecutive rows coalesce)
  const int64_t slots = (int64_t)T * Kp;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (tid >= slots * H * nwhich) return;
  const int64_t p = tid % slots;
  const int h = (int)((tid / slots) % H);
  const int which = (int)(tid / (slots * H));
  const int t = (int)(p / Kp), r = (int)(p % Kp);
  const int32_t s = slot_src[p];
  float v[16];
  if (s >= 0) {
    const float4* row = reinterpret_cast<const float4*>(src + (int64_t)s * ld + col0 + which * C + h * 16);
    float4 q[4];
    // every lane reads a different row: 32-byte loads halve the L1 wavefronts (callers pass 64-byte aligned head columns; the check
    // covers the base pointer and the row stride)
    if ((((uintptr_t)src | (uintptr_t)(ld * 4) | (uintptr_t)(col0 * 4) | (uintptr_t)(C * 4)) & 31) == 0) {
      ldg256(reinterpret_cast<const float*>(row), q[0], q[1]);
      ldg256(reinterpret_cast<const float*>(row + 2), q[2], q[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) q[j] = row[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[4 * j] = q[j].x; v[4 * j + 1] = q[j].y; v[4 * j + 2] = q[j].z; v[4 * j + 3] = q[j].w; }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
  }
  OutT* dst = which == 0 ? dst0 : (which == 1 ? dst1 : dst2);
  if constexpr (sizeof(OutT) == 2) {
    if (which == v_ones_which) {
      OutT* blk32 = dst + ((int64_t)h * T + t) * Kp * 32;
      __half2 hh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) hh[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      uint4* o = reinterpret_cast<uint4*>(blk32 + (r / 8) * 256 + (r % 8) * 8);
      o[0] = *reinterpret_cast<uint4*>(&hh[0]);                                  // n 0..7
      o[8] = *reinterpret_cast<uint4*>(&hh[4]);                                  // n 8..15   (+64 elements)
      o[16] = make_uint4(s >= 0 ? 0x00003C00u : 0u, 0u, 0u, 0u);                 // n 16 = 1.0 (valid keys), n 17..23 = 0
      o[24] = make_uint4(0u, 0u, 0u, 0u);                                        // n 24..31 = 0
      return;
    }
  }
  OutT* blk = dst + ((int64_t)h * T + t) * Kp * 16;
  if constexpr (sizeof(OutT) == 2) {
    __half2 hh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hh[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    uint4* o = reinterpret_cast<uint4*>(blk + (r / 8) * 128 + (r % 8) * 8);
    o[0] = *reinterpret_cast<uint4*>(&hh[0]);       // d 0..7
    o[8] = *reinterpret_cast<uint4*>(&hh[4]);       // d 8..15  (+64 elements = +8 uint4)
  } else {
    float4* o = reinterpret_cast<float4*>(blk + (int64_t)r * 16);   // exact mode: plain [Kp][16] fp32 rows
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}

// fp16 core-matrix packing.  src rows are fp32 [n, ld]; nwhich tensors (1..3) are taken from
// column blocks col0 + w*C.  dst pointers beyond nwhich are ignored.
CDSEG_API int cdseg_attn_pack_f16(const float* src, int64_t ld, int col0, int C, int nwhich, const int32_t* slot_src,
                                  int H, int T, int Kp, void* dst0, void* dst1, void* dst2, void* stream) {
  return cdseg_attn_pack_f16v(src, ld, col0, C, nwhich, slot_src, H, T, Kp, dst0, dst1, dst2, 0, stream);
}

// v_ones != 0: the LAST packed tensor (V) is written 32 wide per key with a ones column (operand of cdseg_attn_tc2)
CDSEG_API int cdseg_attn_pack_f16v(const float* src, int64_t ld, int col0, int C, int nwhich, const int32_t* slot_src,
                                   int H, int T, int Kp, void* dst0, void* dst1, void* dst2, int v_ones, void* stream) {
  if (H <= 0 || C != H * 16 || nwhich < 1 || nwhich > 3 || (Kp % 128) || (ld & 3) || (col0 & 3)) return CDSEG_EINVAL;
  const int64_t total = (int64_t)T * Kp * H * nwhich;
  if (total == 0) return CDSEG_OK;
  cdseg_launch_pdl(pack_heads_kernel<__half>, dim3(cdseg_div_up(total, 256)), dim3(256), 0, (cudaStream_t)stream, src, ld, col0, C, nwhich,
                   slot_src, H, T, Kp, (__half*)dst0, (__half*)dst1, (__half*)dst2, v_ones ? nwhich - 1 : -1);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// hi/lo split packing for cdseg_attn_tc3 mode 1 (fp32-class attention on the tensor cores): x = hi + lo with hi = fp16(x),
// lo = fp16(x - hi).  q / k tensors: the hi halves of all (head, patch) blocks in the [H][T][Kp][16] core-matrix layout, followed
// by the lo halves in the same layout (2 * H*T*Kp*16 halves per tensor).  The V tensor (index v_which, -1: none) is written 48 wide
// per key, [v_hi(16) | 1 | 0 x 15 | v_lo(16)], block layout (k/8)*384 + (n/8)*64 + (k%8)*8 + n%8 (MN-major B operand).
__global__ void pack_split_kernel(const float* __restrict__ src, int64_t ld, int col0, int C, int nwhich,
                                  const int32_t* __restrict__ slot_src, int H, int T, int Kp, __half* __restrict__ dst0,
                                  __half* __restrict__ dst1, __half* __restrict__ dst2, int v_which, int a256) {
  pdl_trigger();
  pdl_wait();
  const int64_t slots = (int64_t)T * Kp;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (tid >= slots * H * nwhich) return;
  const int64_t p = tid % slots;
  const int h = (int)((tid / slots) % H);
  const int which = (int)(tid / (slots * H));
  const int t = (int)(p / Kp), r = (int)(p % Kp);
  const int32_t s = slot_src[p];
  float v[16];
  if (s >= 0) {
    const float4* row = reinterpret_cast<const float4*>(src + (int64_t)s * ld + col0 + which * C + h * 16);
    float4 q[4];
    if (a256) {               // every lane reads a different row: 32-byte loads halve the L1 wavefronts (see gemm_tc.cu)
      ldg256(reinterpret_cast<const float*>(row), q[0], q[1]);
      ldg256(reinterpret_cast<const float*>(row + 2), q[2], q[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) q[j] = row[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[4 * j] = q[j].x; v[4 * j + 1] = q[j].y; v[4 * j + 2] = q[j].z; v[4 * j + 3] = q[j].w; }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
  }
  __half2 hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const float2 b = __half22float2(hi[j]);
    lo[j] = __floats2half2_rn(v[2 * j] - b.x, v[2 * j + 1] - b.y);
  }
  __half* dst = which == 0 ? dst0 : (which == 1 ? dst1 : dst2);
  if (which == v_which) {
    uint4* o = reinterpret_cast<uint4*>(dst + ((int64_t)h * T + t) * Kp * 48 + (r / 8) * 384 + (r % 8) * 8);
    o[0] = *reinterpret_cast<uint4*>(&hi[0]);
    o[8] = *reinterpret_cast<uint4*>(&hi[4]);
    o[16] = make_uint4(s >= 0 ? 0x00003C00u : 0u, 0u, 0u, 0u);                 // n 16 = 1.0 (valid keys)
    o[24] = make_uint4(0u, 0u, 0u, 0u);
    o[32] = *reinterpret_cast<uint4*>(&lo[0]);
    o[40] = *reinterpret_cast<uint4*>(&lo[4]);
    return;
  }
  __half* blk = dst + ((int64_t)h * T + t) * Kp * 16 + (r / 8) * 128 + (r % 8) * 8;
  uint4* oh = reinterpret_cast<uint4*>(blk);
  uint4* ol = reinterpret_cast<uint4*>(blk + (int64_t)H * T * Kp * 16);
  oh[0] = *reinterpret_cast<uint4*>(&hi[0]); oh[8] = *reinterpret_cast<uint4*>(&hi[4]);
  ol[0] = *reinterpret_cast<uint4*>(&lo[0]); ol[8] = *reinterpret_cast<uint4*>(&lo[4]);
}

// has_v != 0: the LAST of the nwhich tensors is V (48 wide); the others are q / k (hi | lo).
CDSEG_API int cdseg_attn_pack_split(const float* src, int64_t ld, int col0, int C, int nwhich, const int32_t* slot_src, int H, int T,
                                    int Kp, void* dst0, void* dst1, void* dst2, int has_v, void* stream) {
  if (H <= 0 || C != H * 16 || nwhich < 1 || nwhich > 3 || (Kp % 128) || (ld & 3) || (col0 & 3)) return CDSEG_EINVAL;
  const int64_t total = (int64_t)T * Kp * H * nwhich;
  if (total == 0) return CDSEG_OK;
  cdseg_launch_pdl(pack_split_kernel, dim3(cdseg_div_up(total, 256)), dim3(256), 0, (cudaStream_t)stream, src, ld, col0, C, nwhich, slot_src,
                   H, T, Kp, (__half*)dst0, (__half*)dst1, (__half*)dst2, has_v ? nwhich - 1 : -1,
                   (((uintptr_t)src & 31) == 0 && (ld & 7) == 0 && (col0 & 7) == 0 && (C & 7) == 0) ? 1 : 0);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

CDSEG_API int cdseg_attn_pack_f32(const float* src, int64_t ld, int col0, int C, int nwhich, const int32_t* slot_src,
                                  int H, int T, int Kp, float* dst0, float* dst1, float* dst2, void* stream) {
  if (H <= 0 || C != H * 16 || nwhich < 1 || nwhich > 3 || (Kp % 128) || (ld & 3) || (col0 & 3)) return CDSEG_EINVAL;
  const int64_t total = (int64_t)T * Kp * H * nwhich;
  if (total == 0) return CDSEG_OK;
  pack_heads_kernel<float><<<cdseg_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(src, ld, col0, C, nwhich, slot_src, H,
                                                                                      T, Kp, dst0, dst1, dst2);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}

// ---------------------------------------------------------------------------------
// exact fp32 attention (SIMT): one thread per query row, keys streamed through shared memory
// in tiles of 32 with one online-softmax rescale per tile.  Output rows are scattered straight
// to the original point order (fuses "feat[inverse]").
// ---------------------------------------------------------------------------------
constexpr int AX_KT = 32;

__global__ void __launch_bounds__(128) attn_exact_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                         const float* __restrict__ V,
                                                         const int32_t* __restrict__ patch_len,
                                                         const int32_t* __restrict__ slot_dst, int H, int T, int Kp,
                                                         float scale, float* __restrict__ out, int64_t out_ld) {
  __shared__ float sK[AX_KT][16];
  __shared__ float sV[AX_KT][16];
  const int qt = blockIdx.x, t = blockIdx.y, h = blockIdx.z;
  const int len = patch_len[t];
  if (qt * 128 >= len) return;
  const int64_t blk = ((int64_t)h * T + t) * Kp * 16;
  const int r = qt * 128 + threadIdx.x;
  float q[16], acc[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 x = reinterpret_cast<const float4*>(Q + blk + (int64_t)r * 16)[j];
    q[4 * j] = x.x * scale; q[4 * j + 1] = x.y * scale; q[4 * j + 2] = x.z * scale; q[4 * j + 3] = x.w * scale;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < len; k0 += AX_KT) {
    __syncthreads();
    for (int j = threadIdx.x; j < AX_KT * 4; j += 128) {
      const int kr = j / 4, kc = j % 4;
      reinterpret_cast<float4*>(&sK[kr][0])[kc] = reinterpret_cast<const float4*>(K + blk + (int64_t)(k0 + kr) * 16)[kc];
      reinterpret_cast<float4*>(&sV[kr][0])[kc] = reinterpret_cast<const float4*>(V + blk + (int64_t)(k0 + kr) * 16)[kc];
    }
    __syncthreads();
    float s[AX_KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < AX_KT; ++j) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) d = fmaf(q[c], sK[j][c], d);
      s[j] = (k0 + j < len) ? d : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float mn = fmaxf(m, tmax);
    const float corr = __expf(m - mn);      // m = -inf on the first tile -> 0
    l *= corr;
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] *= corr;
#pragma unroll
    for (int j = 0; j < AX_KT; ++j) {
      const float p = expf(s[j] - mn);
      l += p;
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = fmaf(p, sV[j][c], acc[c]);
    }
    m = mn;
  }
  const int32_t dst = slot_dst[(int64_t)t * Kp + r];
  if (dst >= 0) {
    const float inv = 1.f / l;
    float4* o = reinterpret_cast<float4*>(out + (int64_t)dst * out_ld + h * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(acc[4 * j] * inv, acc[4 * j + 1] * inv, acc[4 * j + 2] * inv, acc[4 * j + 3] * inv);
  }
}

// Q,K,V: fp32 [H][T][Kp][16] (cdseg_attn_pack_f32).  out: fp32 [n, out_ld], head h -> columns h*16..
CDSEG_API int cdseg_attn_exact(const float* Q, const float* K, const float* V, const int32_t* patch_len,
                               const int32_t* slot_dst, int H, int T, int Kp, float scale, float* out, int64_t out_ld,
                               void* stream) {
  if (H <= 0 || T < 0 || (Kp % 128) || (out_ld & 3)) return CDSEG_EINVAL;
  if (T == 0) return CDSEG_OK;
  dim3 g(Kp / 128, T, H);
  attn_exact_kernel<<<g, 128, 0, (cudaStream_t)stream>>>(Q, K, V, patch_len, slot_dst, H, T, Kp, scale, out, out_ld);
  CDSEG_COUNT_LAUNCH(1);
  CDSEG_LAUNCH_CHECK();
  return CDSEG_OK;
}
