/*
 * cdseg_b200 -- C ABI of the B200 (sm_100a) kernels behind the CDSegNet PTv3 single-step forward.
 *
 * Conventions (mirroring the reference's own native boundary, libs/pointops/src/ * / *_kernel.h,
 * e.g. libs/pointops/src/knn_query/knn_query_cuda_kernel.h:9-17: C linkage, raw device pointers,
 * caller allocates every output) with two additions the reference lacks: an explicit
 * cudaStream_t (passed as void*) and an int status return:
 *     0  ok        >0  cudaError_t of the failed launch
 *    -1  invalid argument (shape / alignment / unsupported size)      -2  workspace too small
 * No entry point allocates, synchronises, or touches the host unless stated.  All pointers are
 * device pointers unless the comment says "host".  "ptv3.py" below abbreviates
 * pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py of the reference.
 */
#ifndef CDSEG_B200_H
#define CDSEG_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int cdseg_abi_version(void);
/* programmatic dependent launch between consecutive kernels of a stream (on by default; CDSEG_PDL=0 in the environment or 0 here
 * turns it off: every launch then waits for the full completion of its predecessor before its first CTA is scheduled) */
void cdseg_set_pdl(int on);
int cdseg_get_pdl(void);
/* number of kernels launched through this library since the last reset (bench.py "gpu_launches") */
unsigned long long cdseg_launch_count(void);
void cdseg_launch_count_reset(void);

/* ---- serialization: replaces Point.serialization, pointcept/models/utils/structure.py:47-102 ------------ */

/* max over grid_coord (-> serialized_depth = bit_length(max), structure.py:66).  out_max: int32[1] */
int cdseg_grid_max(const int32_t* grid, int64_t n_elems, int32_t* out_max, void* stream);
/* batch[i] = scene of point i from cumulative offsets (pointcept/models/utils/misc.py:19-24) */
int cdseg_offset2batch(const int64_t* offset, int B, int64_t N, int32_t* batch, void* stream);
/* codes[r][i] = batch<<3*depth | curve_r(grid[i]) for r < k.  order_ids (host): 0 "z", 1 "z-trans",
 * 2 "hilbert", 3 "hilbert-trans" (serialization/default.py:9-24; z_order.py:66-101; hilbert.py:91-198) */
int cdseg_encode_codes(const int32_t* grid, const int32_t* batch, int64_t N, int depth, const int* order_ids,
                       int k, int64_t* codes, void* stream);
/* host evaluation of the same bit routines (CPU test-suite only, never on the product path) */
int cdseg_debug_encode_host(const int32_t* grid, const int32_t* batch, int64_t N, int depth, int order_id,
                            int64_t* codes);
/* order = argsort(codes) per row, inverse[order[i]] = i (structure.py:83-92).  nbits = significant key bits. */
size_t cdseg_argsort_workspace_bytes(int k, int64_t N);
int cdseg_argsort_rows(const int64_t* codes, int k, int64_t N, int nbits, int32_t* order, int32_t* inverse,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- patch maps: replaces SerializedAttention.get_padding_and_inverse, ptv3.py:188-244 ----------------- */
/* scene_count: host int64[B] points per scene (B <= 64).  K = patch size, Kp = round_up(K,128) slots per patch.
 * slot_src/slot_dst: int32[T*Kp], point_slot: int32[N], patch_len: int32[T]  (T from cdseg_patch_count) */
/* dst[i, :] = src[idx[i], :] for rows of row_bytes bytes (multiple of 4): feat[perm], coord[perm], logits[inv_perm] around the network */
int cdseg_gather_rows(const void* src, const int32_t* idx, int64_t n, int row_bytes, void* dst, void* stream);
/* dst[i, 0..C) = src[idx[i], 0..C), dst[i, C..Cp) = 0 (fp32 rows; the stem's zero-padded input in internal numbering) */
int cdseg_gather_rows_pad(const float* src, const int32_t* idx, int64_t n, int C, int Cp, float* dst, void* stream);
/* level-0 renumbering along the first curve (internal point r = original point perm[r], perm = order[0], inv_perm = inverse[0]):
 * grid / batch / code / inverse gathered by perm, order composed with inv_perm; code int64 [k,N], order / inverse int32 [k,N] */
int cdseg_renumber(const int32_t* perm, const int32_t* inv_perm, const int32_t* grid, const int32_t* batch, const int64_t* code,
                   const int32_t* order, const int32_t* inverse, int k, int64_t N, int32_t* i_grid, int32_t* i_batch,
                   int64_t* i_code, int32_t* i_order, int32_t* i_inverse, void* stream);
int cdseg_patch_count(const int64_t* scene_count, int B, int K, int* T_out);
int cdseg_patch_maps(const int32_t* order, const int64_t* scene_count, int B, int K, int Kp, int32_t* slot_src,
                     int32_t* slot_dst, int32_t* point_slot, int32_t* patch_len, void* stream);

/* ---- grid pooling: replaces SerializedPooling, ptv3.py:464-531 and unpool gather, ptv3.py:623 ---------- */
/* One pooling level, sync-free.  Parent: code int64[k][ld], order int32[k][ld], grid int32[n,3], batch int32[n];
 * n = *n_dev if n_dev != NULL (count produced on the device by the previous level) else n_host.
 * c0 = curve whose codes define the clusters (logical row 0 after the reference's shuffle).
 * Outputs: cluster int32[n] (= pooling_inverse), idx_ptr int32[m+1], head int32[m], child code/order/inverse
 * with row stride ld_c, child grid/batch, *m_dev = m, c_offset int64[B] cumulative child offsets. */
size_t cdseg_pool_plan_workspace_bytes(int k, int64_t ld);
int cdseg_pool_plan(const int64_t* code, const int32_t* order, int k, int64_t ld, const int32_t* n_dev,
                    int64_t n_host, int c0, int pooling_depth, const int32_t* grid, const int32_t* batch,
                    int32_t* cluster, int32_t* idx_ptr, int32_t* head, int64_t* c_code, int32_t* c_order,
                    int32_t* c_inverse, int64_t ld_c, int32_t* c_grid, int32_t* c_batch, int32_t* m_dev,
                    int64_t* c_offset, void* workspace, size_t workspace_bytes, void* stream);
/* out[j] = act(bn(max_{i in cluster j} x[i])) ; out_coord[j] = mean coord (torch_scatter.segment_csr, ptv3.py:507-531)
 * members = parent order of curve c0 (cluster members are contiguous there), idx_ptr from the plan */
int cdseg_pool_reduce(const float* x, const float* coord, const int32_t* members, const int32_t* idx_ptr, int64_t m,
                      int C, const float* bn_scale, const float* bn_shift, int gelu, float* out, float* out_coord,
                      void* stream);
/* out[i] = a[i]*alpha + b[cluster[i]]   (ptv3.py:608-623) */
int cdseg_unpool_add(const float* a, const float* b, const int32_t* cluster, int64_t n, int C, float alpha,
                     float* out, void* stream);

/* ---- submanifold conv: replaces spconv.SubMConv3d at ptv3.py:356-362, 647-654, 1106-1123 ---------------- */
size_t cdseg_nbr_workspace_bytes(int64_t n);
int64_t cdseg_hash_capacity(int64_t n);
/* nbr int32[n, ksize^3]: index of the active voxel at grid[i] + (a-r, b-r, c-r), tap t = (a*ks+b)*ks+c, or -1 */
int cdseg_nbr_build(const int32_t* grid, const int32_t* batch, int64_t n, int ksize, int32_t* nbr, void* workspace,
                    size_t workspace_bytes, void* stream);
/* out[n,Co] = bias + sum_t in[nbr[:,t]] @ wt[t]; wt fp32 [ksize^3][Ci][Co]; optional folded BN + GELU epilogue
 * (only for Ci <= 8, the Embedding stem) */
int cdseg_subm_conv(const float* in, const int32_t* nbr, const float* wt, const float* bias, const float* ep_scale,
                    const float* ep_shift, int ep_gelu, int64_t n, int Ci, int Co, int ksize, float* out,
                    void* stream);

/* ---- serialized patch attention: replaces flash_attn varlen calls at ptv3.py:282-289, 1038-1047 -------- */
/* gather rows by slot_src into per-(head,patch) tiles [H][T][Kp][16]; src fp32 [n, ld]; tensor w in [0,nwhich)
 * takes columns col0 + w*C + h*16 ...  (ptv3.py:258-262 "qkv[order]") */
int cdseg_attn_pack_f16(const float* src, int64_t ld, int col0, int C, int nwhich, const int32_t* slot_src, int H,
                        int T, int Kp, void* dst0, void* dst1, void* dst2, void* stream);
int cdseg_attn_pack_f32(const float* src, int64_t ld, int col0, int C, int nwhich, const int32_t* slot_src, int H,
                        int T, int Kp, float* dst0, float* dst1, float* dst2, void* stream);
/* tcgen05/TMEM kernel: fp16 operands, fp32 accumulate; out fp32 [n, out_ld] rows scattered by slot_dst (":290 feat[inverse]") */
size_t cdseg_attn_tc_smem_bytes(int Kp);
int cdseg_attn_tc(const void* Q, const void* K, const void* V, const int32_t* patch_len, const int32_t* slot_dst,
                  int H, int T, int Kp, float scale, float* out, int64_t out_ld, void* stream);
/* second-generation kernel (default): one q tile per CTA, K/V streamed through a TMA ring, V packed 32 wide with a ones
 * column (cdseg_attn_pack_f16v with v_ones = 1: the LAST packed tensor gets [v(16) | 1 | 0 x 15] per key) */
int cdseg_attn_pack_f16v(const float* src, int64_t ld, int col0, int C, int nwhich, const int32_t* slot_src, int H,
                         int T, int Kp, void* dst0, void* dst1, void* dst2, int v_ones, void* stream);
/* how many of every 8 softmax exponentials cdseg_attn_tc2 evaluates with the FMA-pipe polynomial instead of MUFU.EX2 (0..3) */
void cdseg_attn_set_poly(int per8);
int cdseg_attn_tc2(const void* Q, const void* K, const void* V32, const int32_t* patch_len, const int32_t* slot_dst,
                   int H, int T, int Kp, float scale, float* out, int64_t out_ld, void* stream);
/* exact fp32 SIMT kernel (dense-branch numerics, ptv3.py:264-280) */
/* third-generation tcgen05 kernel (attn_tc3.cu): P / O double-buffered, O folded one chunk late (no softmax warp ever waits for the
 * P.V product it just requested).  mode 0: operands of cdseg_attn_pack_f16v(v_ones = 1), flash-branch numerics (fp16 probabilities,
 * fp16-rounded output).  mode 1: operands of cdseg_attn_pack_split (q / k hi | lo halves, V 48 wide [v_hi | 1 | v_lo]): 22-bit
 * operands and probabilities, fp32 output -- the dense branch (ptv3.py:264-280) to fp32 accuracy on the tensor cores. */
int cdseg_attn_pack_split(const float* src, int64_t ld, int col0, int C, int nwhich, const int32_t* slot_src, int H, int T,
                          int Kp, void* dst0, void* dst1, void* dst2, int has_v, void* stream);
/* profiling only: what-if variants of the mode-0 kernel that drop one piece of work each (results are WRONG): 0 = off */
void cdseg_attn_set_debug(int variant);
int cdseg_attn_tc3(const void* Q, const void* K, const void* V, const int32_t* patch_len, const int32_t* slot_dst, int H,
                   int T, int Kp, float scale, int mode, float* out, int64_t out_ld, void* stream);
int cdseg_attn_exact(const float* Q, const float* K, const float* V, const int32_t* patch_len,
                     const int32_t* slot_dst, int H, int T, int Kp, float scale, float* out, int64_t out_ld,
                     void* stream);

/* ---- row-wise fused ops: ptv3.py:402-424 (residual + t_mlp + LayerNorm), 549-553/575-593 (BN+GELU), 1772-1778 */
/* y = a (+ b) (+ t[batch]) ; y_out = y (nullable) ; ln_out = LayerNorm(y) (nullable) */
int cdseg_add_layernorm(const float* a, const float* b, const float* t, const int32_t* batch, const float* gamma,
                        const float* beta, float eps, int64_t n, int C, float* y_out, float* ln_out, void* stream);
/* out = act(x*scale[c] + shift[c]); act 0 none, 1 GELU(erf) */
/* split-K reduction fused with the row-wise tail of a Linear inside a Block: v = bias + sum_z part[z][n][C] (nsplit = 1: part is a
 * finished [n, C] tensor) -> optional LayerNorm(g1, b1) -> + res (+ t[batch]) -> y_out (nullable) -> LayerNorm(g2, b2) -> ln_out (nullable).
 * part = the workspace of a cdseg_gemm_tc call made with out = NULL and nsplit > 1 (partials left unreduced). */
int cdseg_reduce_ln(const float* part, int nsplit, const float* bias, const float* g1, const float* b1, const float* res,
                    const float* t, const int32_t* batch, const float* g2, const float* b2, float eps, int64_t n, int C,
                    float* y_out, float* ln_out, void* stream);
int cdseg_scale_shift_act(const float* x, const float* scale, const float* shift, int act, int64_t n, int C,
                          float* out, void* stream);
/* out[r][o] = act(bias[o] + x[r].W[o]) for a handful of rows (per-scene timestep MLP); act 0 none, 2 swish */
int cdseg_small_linear(const float* x, const float* W, const float* bias, int act, int R, int K, int O, float* out,
                       void* stream);
/* flag[0] |= 1 if any row of x differs from the first row of its scene */
int cdseg_rows_uniform(const float* x, const int32_t* batch, const int64_t* offset, int64_t n, int C, int32_t* flag,
                       void* stream);

/* ---- fp32-faithful tensor-core GEMM (tcgen05 kind::tf32, 3xTF32 split) with gathered A rows and fused epilogue:
 *      replaces spconv.SubMConv3d k=3 (ptv3.py:356-362, 1106-1123) as an implicit GEMM over taps AND every nn.Linear on
 *      the path (ptv3.py:185-186, 311-313, 359, 458, 575-581, 911-913) ------------------------------------------ */
/* W fp32 [T][K][N] (tap-major transposed conv weight, or weight^T of a Linear with T=1; K % 16 == 0)
 * -> Bp: cdseg_gemm_packed_b_floats(T,K,N) floats of pre-split (hi|lo), pre-tiled UMMA operand blocks */
/* dense-layer precision of every tcgen05 GEMM in the library (gemm_tc, conv, fused pre / post kernels, stem): 0 (default) = fp32-faithful
 * 3-term fp16 hi/lo split; 1 = fp16 operands, fp32 accumulation, one MMA per term (the reference's autocast numerics) */
void cdseg_set_gemm_precision(int fp16_single);
int cdseg_get_gemm_precision(void);
size_t cdseg_gemm_packed_b_floats(int T, int K, int N);
int cdseg_gemm_pack_b(const float* W, int T, int K, int N, float* Bp, void* stream);
/* mask[tile] bit t = some row of the 128-row tile has neighbour t (nbr int32 [M,T], T <= 32) */
int cdseg_tile_tap_mask(const int32_t* nbr, int64_t M, int T, uint32_t* mask, void* stream);
/* out[M,N] = act(bias + sum_t A[idx[m,t]] @ W_t) + res ; idx NULL => row m, tap t = its t-th K-slice (split-K Linear); tile_mask NULL => all
 * taps; act 0 none / 1 GELU(erf); nsplit > 1 splits the taps over grid.z (partials in workspace, reduced by a 2nd kernel) */
size_t cdseg_gemm_tc_workspace_bytes(int64_t M, int N, int nsplit);
/* profiling hook: clock64 stamps of CTA (cta,0,0) of subsequent launches into a device buffer of >= 16 int64; NULL disables */
/* experiment switch (default off, env CDSEG_GEMM_NARROW): serve a dense Linear that was asked to split K with 32 / 64-column
 * output tiles instead of partial sums + a reduce launch (profiles/r01f_narrow_tiles.md) */
void cdseg_gemm_tc_set_narrow(int on);
void cdseg_gemm_tc_set_trace(long long* buf, int cta);
int cdseg_gemm_tc(const float* A, int64_t lda, const int32_t* idx, int T, const uint32_t* tile_mask, const float* Bp,
                  int64_t M, int N, int K, const float* bias, const float* res, int64_t ldr, int act, float* out,
                  int64_t ldo, int nsplit, void* workspace, size_t workspace_bytes, void* stream);
/* Embedding stem (ptv3.py:633-663: spconv.SubMConv3d k=5, C_in=6, bias=False + BatchNorm1d(eval) + GELU) as an im2col
 * GEMM on the tensor cores: A8 fp32 [rows,8] = input zero-padded to 8 channels; nbr int32 [M,taps]; Bp = cdseg_gemm_pack_b
 * of W [ceil(taps/4)][32][N] whose row 8q+c of block t is the (BN-scaled) weight of tap 4t+q, channel c (zero padding);
 * out[M,N] = act(bias + conv) with bias = the folded BN shift */
int cdseg_conv_im2col_tc(const float* A8, const int32_t* nbr, int taps, const float* Bp, int64_t M, int N,
                         const float* bias, int act, float* out, int64_t ldo, void* stream);

/* ---- fused row-tile kernels (C = 32 / 64 / 128): chained skinny GEMMs whose intermediates stay in tensor memory ------
 * post-attention half of a Block, ptv3.py:290-296 + 416-424:  x2 = x1 + proj(o) + b ; out = x2 + fc2(GELU(fc1(LayerNorm(x2))))
 * o, x1, out: fp32 [n, C] contiguous, 16-byte aligned; *_Bp = cdseg_gemm_pack_b of W^T ([1][C][C], [1][C][4C], [1][4C][C]) */
int cdseg_post_attn(const float* o, const float* x1, int64_t n, int C, const float* proj_Bp, const float* proj_b,
                    const float* ln_g, const float* ln_b, float eps, const float* fc1_Bp, const float* fc1_b,
                    const float* fc2_Bp, const float* fc2_b, float* out, void* stream);
/* pre-attention half of a Block, ptv3.py:355-362 + 400-413 + 258:  x1 = x + LayerNorm_cpe(Linear(SubMConv3d_k3(conv_in))) (+ tproj[batch]) ;
 * qkv = Linear_qkv(LayerNorm_1(x1)).  conv_in, x, x1: fp32 [n, C]; qkv: fp32 [n, 3C]; nbr int32 [n, 27] + tile_mask from
 * cdseg_nbr_build / cdseg_tile_tap_mask; conv_Bp = cdseg_gemm_pack_b of the tap-major weight [27][C][C]; lin_Bp / qkv_Bp of W^T;
 * tproj fp32 [B, C] (per-scene t_mlp output) + batch int32 [n], or NULL for the Conditional Network; conv_plan from
 * cdseg_conv_tile_plan (same nbr) */
int cdseg_pre_attn(const float* conv_in, const float* x, int64_t n, int C, const int32_t* nbr, const uint32_t* tile_mask,
                   const void* conv_plan, const float* conv_Bp, const float* conv_b, const float* lin_Bp, const float* lin_b, const float* cpe_g,
                   const float* cpe_b, const float* tproj, const int32_t* batch, const float* n1_g, const float* n1_b,
                   float eps, const float* qkv_Bp, const float* qkv_b, float* x1, float* qkv, void* stream);
/* Per 128-row tile of a k=3 neighbour table (nbr int32 [n, 27], the spconv `indice_key` rulebook analogue, ptv3.py:356):
 * the ascending list of DISTINCT neighbour rows of the tile and, for every (row, tap), its index into that list.  One
 * record of 16 + 4 * CDSEG_CONV_PLAN_UCAP + 128 * 27 * 2 bytes per tile: int32 ucount, int32 pad[3], int32 uniq[UCAP]
 * (-1 padded), int16 lidx[128][27] (-1 = absent).  ucount > UCAP marks a tile whose neighbourhood does not fit (its uniq /
 * lidx are not written; consumers read nbr directly for it).  plan: >= cdseg_conv_plan_bytes(n) bytes, 16-byte aligned. */
#define CDSEG_CONV_PLAN_UCAP 384
/* profiling hook: clock64 stamps (12 per tile, first 6 tiles) of row thread 0 of CTA `cta` of later cdseg_pre_attn launches; NULL = off */
void cdseg_pre_attn_set_trace(long long* buf, int cta);
size_t cdseg_conv_plan_bytes(int64_t n);
int cdseg_conv_tile_plan(const int32_t* nbr, int64_t n, void* plan, void* stream);
/* which fused kernels cdseg_block_forward uses: bit 0 post-attention chain, bit 1 pre-attention chain, bit 2 split-K reduction fused with
 * the LayerNorms that follow it at the wide levels (default: all) */
void cdseg_set_fused_mask(int mask);
/* resident CTAs per SM of the persistent fused kernels launched from now on (0 = as many as fit, the default) */
void cdseg_set_fused_ctas_per_sm(int n);

/* ---- native plan phase (plan_exec.cu): serialization + pooling hierarchy + indice tables + patch maps of BOTH networks in one call --
 * Replaces Point.serialization (structure.py:47-102), the structural half of every SerializedPooling (ptv3.py:464-505), the spconv
 * indice tables of Point.sparsify (structure.py:104-140) and get_padding_and_inverse (ptv3.py:188-244).  Two stream syncs inside
 * (depth + offsets; pooled sizes).  Caller allocates the device arena; every pointer below points into it (or at the inputs). */
#define CDSEG_MAX_SCENES 64
typedef struct CdsegPatchMap {
  int32_t* slot_src; int32_t* slot_dst; int32_t* point_slot; int32_t* patch_len;   /* [T*Kp], [T*Kp], [n], [T] */
  int T, Kp, K, pad_; int64_t pairs;                                                /* pairs: algorithmic (query, key) pairs, padding excluded */
} CdsegPatchMap;
typedef struct CdsegPlanLevel {
  /* in (host) */
  int parent, stride;                 /* index of the level this one is pooled from (-1: a level-0 entry), pooling stride */
  int rowmap[4];                      /* logical curve row -> physical row (the reference's shuffle_orders bookkeeping) */
  int K; unsigned pm_mask;            /* patch size of this level's attention blocks; bit i: build the slot maps of logical curve i */
  int want_conv_plan, stem_ksize;     /* fused pre-attention operand cache plan (C = 32 / 64 / 128 levels); stem conv kernel size (level 0, else 0) */
  /* out (host) */
  int64_t n, cap; int B, depth, c0, pooling_depth, slot, pad_;
  int64_t offset_host[CDSEG_MAX_SCENES];
  /* out (device) */
  int32_t* grid; int32_t* batch; int64_t* code; int32_t* order; int32_t* inverse;      /* code / order / inverse: [k][cap], physical rows */
  int32_t* cluster; int32_t* idx_ptr; int32_t* head; int32_t* m_dev; int64_t* offset_dev;  /* pooled levels: parent point -> this level, CSR, heads */
  int32_t* perm; int32_t* inv_perm; int64_t* o_code; int32_t* o_order; int32_t* o_inverse; /* level 0: internal <-> caller numbering, originals */
  int32_t* nbr3; uint32_t* tile_mask3; void* conv_plan3; int32_t* nbr_stem;
  CdsegPatchMap pm[4];                /* by LOGICAL curve index */
  /* out: NULL, or (plan built with an aux stream) the cudaEvent_t after which nbr3 / tile_mask3 / conv_plan3 (`ready`) and nbr_stem
   * (`ready_stem`) of this level are complete; a consumer makes its stream wait on them before the first kernel that reads those tables.
   * Everything else in the descriptor is ordered by `stream` itself. */
  void* ready; void* ready_stem;
} CdsegPlanLevel;
size_t cdseg_plan_arena_bytes(int64_t N, int B, int k, int n_levels, int n_level0, int stem_ksize, int K_min, int K_max);
int cdseg_plan_build(const int32_t* grid, const int64_t* offset, int64_t N, int B, const int* order_ids, int k,
                     CdsegPlanLevel* levels, int n_levels, const int32_t* extra_flags, int n_flags, int32_t* flags_host,
                     void* arena, size_t arena_bytes, void* stream, void* aux_stream /* NULL: everything on `stream` */);
/* With an aux stream the launches that build the pooled levels' tables and slot maps are NOT enqueued by cdseg_plan_build (the
 * descriptors are complete, the device data is not): cdseg_plan_finish() enqueues them on the aux stream and records the `ready`
 * events.  cdseg_net_forward calls it itself before its first pooled stage; any other consumer calls it before waiting on `ready`.
 * Idempotent; one plan at a time per process. */
int cdseg_plan_finish(void);

/* ---- native executor of one PTv3 Block (ptv3.py:399-428): the 12-13 launches above enqueued from C++ in one call --- */
#define CDSEG_ATTN_F16 0
#define CDSEG_ATTN_EXACT 1
#define CDSEG_ATTN_TC32 2
typedef struct CdsegBlockArgs {
  int64_t n; int C, H, T_dim, B;                 /* points, channels, heads (C = 16 H), timestep width, scenes */
  const float* x; const float* conv_in;           /* block input; conv_in != NULL: tensor the CPE conv reads (stale-feature quirk) */
  const int32_t* nbr; const uint32_t* tile_mask; const int32_t* batch;
  const void* conv_plan;                          /* cdseg_conv_tile_plan(nbr) or NULL (then the unfused conv path runs) */
  const float* t_scene;                           /* [B, T_dim] per-scene timestep features or NULL (CN blocks) */
  const int32_t* slot_src; const int32_t* slot_dst; const int32_t* patch_len; int T, Kp; float scale;
  const float* conv_Bp; const float* conv_b;      /* packed operands come from cdseg_gemm_pack_b */
  const float* lin_Bp; const float* lin_b; const float* cpe_g; const float* cpe_b;
  const float* t_W; const float* t_b;
  const float* n1_g; const float* n1_b; const float* qkv_Bp; const float* qkv_b; const float* proj_Bp; const float* proj_b;
  const float* n2_g; const float* n2_b; const float* fc1_Bp; const float* fc1_b; const float* fc2_Bp; const float* fc2_b;
  float ln_eps;
  int attn_mode;                                  /* CDSEG_ATTN_F16 (flash-branch numerics, tcgen05), CDSEG_ATTN_EXACT (fp32 SIMT, dense-branch
                                                     numerics) or CDSEG_ATTN_TC32 (tcgen05 with hi/lo-split operands: fp32-class results) */
  float* out; void* scratch; size_t scratch_bytes; /* out [n,C]; scratch >= cdseg_block_scratch_bytes(...) */
  void* ev[6];                                    /* optional cudaEvent_t pairs recorded around: [0,1] the attention kernel, [2,3] the post-attention
                                                     kernel (fc1 GEMM on the unfused path), [4,5] the pre-attention kernel (cpe conv GEMM when unfused) */
} CdsegBlockArgs;
size_t cdseg_block_scratch_bytes(int64_t n, int C, int H, int T, int Kp, int B);
int cdseg_block_forward(const CdsegBlockArgs* args, void* stream);
/* ---- native executor of the whole feature phase (net_exec.cu): PointTransformerV3.forward, ptv3.py:1757-1845, after the plan ----
 * Embedding stems, the interleaved encoder stages of both networks (Noise Network on a second stream), TransferModule, decoders,
 * heads and the gathers between the caller's point numbering and the internal (curve-order) one: ~450 launches enqueued from ONE
 * call instead of ~120 Python -> C transitions.  Weights are described once per model by CdsegNetW (packed operands from
 * cdseg_gemm_pack_b, eval-mode BatchNorm folded into the packed weight / bias); levels come from cdseg_plan_build. */
typedef struct CdsegLinW { const float* Bp; const float* bias; int K, N; } CdsegLinW;        /* y = x W^T + b (bias may be NULL) */
typedef struct CdsegLnW { const float* g; const float* b; } CdsegLnW;
typedef struct CdsegBlockW {                       /* Block, ptv3.py:325-428 */
  int C, H, T_dim, order_index; float scale, ln_eps;
  const float* conv_Bp; const float* conv_b; CdsegLinW lin; CdsegLnW cpe_ln; const float* t_W; const float* t_b;
  CdsegLnW n1; CdsegLinW qkv, proj; CdsegLnW n2; CdsegLinW fc1, fc2;
} CdsegBlockW;
typedef struct CdsegPoolW { CdsegLinW proj; const float* bn_scale; const float* bn_shift; } CdsegPoolW;     /* SerializedPooling, ptv3.py:507-555 */
typedef struct CdsegUnpoolW { CdsegLinW proj, proj_skip, cat_a, cat_b; int cat; float alpha; } CdsegUnpoolW;  /* SerializedUnpooling, ptv3.py:601-630 */
typedef struct CdsegStageW { int n_blocks, has_pool, has_up, level; const CdsegBlockW* blocks; CdsegPoolW pool; CdsegUnpoolW up; } CdsegStageW;
typedef struct CdsegStemW { const float* Bp; const float* shift; int cin, cout, ksize, pad_; } CdsegStemW;   /* Embedding, ptv3.py:633-663 */
typedef struct CdsegCrossW {                       /* CrossBlock of the TransferModule, ptv3.py:1058-1223 (pre-norm, tm_feat a float) */
  int Cq, Ckv, H, K; float scale, tm_feat, ln_eps; int pad_;
  const float* q_conv_Bp; const float* q_conv_b; CdsegLinW q_lin; CdsegLnW q_cpe_ln;
  const float* kv_conv_Bp; const float* kv_conv_b; CdsegLinW kv_lin; CdsegLnW kv_cpe_ln;
  CdsegLnW q_norm1, kv_norm1, q_norm2; CdsegLinW q, kv, proj, fc1, fc2;
} CdsegCrossW;
#define CDSEG_MAX_STAGES 8
typedef struct CdsegNetW {
  int condition, T_dim, n_enc, n_dec, c_enc, c_dec, pad0_, pad1_;
  CdsegStemW n_stem, c_stem;
  CdsegStageW n_enc_st[CDSEG_MAX_STAGES], n_dec_st[CDSEG_MAX_STAGES], c_enc_st[CDSEG_MAX_STAGES], c_dec_st[CDSEG_MAX_STAGES]; /* dec: execution order */
  CdsegLinW n_head, c_head;
  const float* fc_t1_W; const float* fc_t1_b; const float* fc_t2_W; const float* fc_t2_b;
  CdsegCrossW tm;
} CdsegNetW;
typedef struct CdsegForwardArgs {
  const CdsegNetW* w; const CdsegPlanLevel* levels; int n_lv_n, n_lv_c;   /* levels[0 .. n_lv_n): CN, then n_lv_c NN levels */
  int64_t N; int B, attn_mode;
  const float* n_feat; const float* c_feat;       /* fp32 [N, cin], caller numbering */
  const float* t_emb;                             /* fp32 [B, T_dim]: one timestep-embedding row per scene (NULL: no timestep branch) */
  float* n_out; float* c_out;                     /* fp32 [N, n_head.N] / [N, c_head.N], caller numbering */
  void* arena_main; size_t arena_main_bytes; void* arena_side; size_t arena_side_bytes;   /* >= cdseg_net_arena_bytes */
  void* stream_main; void* stream_side;           /* stream_side NULL or == stream_main: everything on one stream */
  void** block_events;                            /* optional: 6 cudaEvent_t per Block in execution order (CN blocks, then NN blocks), see CdsegBlockArgs.ev */
} CdsegForwardArgs;
int cdseg_net_arena_bytes(const CdsegForwardArgs* args, size_t* main_bytes, size_t* side_bytes);
int cdseg_net_forward(const CdsegForwardArgs* args);
/* debug switches of the executor: bit 0 / 1 serialise the encoders / decoders of the two networks, bit 2 logs arena allocations */
void cdseg_net_set_debug(int flags);
/* sizeof(CdsegBlockArgs, CdsegPatchMap, CdsegPlanLevel, CdsegLinW, CdsegLnW, CdsegBlockW, CdsegPoolW, CdsegUnpoolW, CdsegStageW, CdsegStemW,
 * CdsegCrossW, CdsegNetW, CdsegForwardArgs) in that order; returns how many there are.  For bindings to check their struct mirrors. */
int cdseg_struct_sizes(size_t* out, int n);
/* test hooks: the split-K / tap-split heuristic of net_exec.cu and block_exec.cu (must equal cdsegnet_b200/ops.py::pick_split) */
int cdseg_debug_pick_split(int64_t tiles, int T);
int cdseg_debug_pick_split_block(int64_t tiles, int T);

/* CUDA events for live per-kernel timing inside the timed region (bench.py) */
void* cdseg_event_create(void);
void cdseg_event_destroy(void* e);
int cdseg_event_elapsed_ms(void* e0, void* e1, float* ms);

/* ---- DefaultSegmentorV2 wrapper: criteria (forward values) and diffusion samplers ------------------------------------
 * Criteria of pointcept/models/losses/builder.py:14-51 over MSELoss (misc.py:24-94, batch_sample_point <= 0),
 * CrossEntropyLoss (misc.py:97-129) and LovaszLoss mode="multiclass" (lovasz.py:118-165, 244-272), all with one shared
 * ignore_index.  n_pred fp32 [n, C] logits, n_target int64 [n]; c_pred / c_target fp32 [n, Cc] (Noise-Network output and
 * its diffusion target) or NULL with has_mse = 0.  out5 (device, fp32): [0] MSE, [1] CE, [2] Lovasz (mean over the classes
 * present among the valid labels), [3] their sum (loss_type "EW" and every eval pass), [4] sqrt(MSE * (CE + Lovasz))
 * (loss_type "GLS", task_num = 2, training pass).  Caller allocates the workspace; nothing is copied to the host. */
size_t cdseg_criteria_workspace_bytes(int64_t n, int C);
int cdseg_criteria(const float* n_pred, const int64_t* n_target, int64_t n, int C, int64_t ignore_index, const float* c_pred,
                   const float* c_target, int Cc, int mse_use_ignore, float w_mse, float w_ce, float w_lov, int has_mse,
                   int has_ce, int has_lov, float* out5, void* workspace, size_t workspace_bytes, void* stream);
/* continuous_q_sample, default.py:216-222: out = sqrt_ab[b] * x0 + sqrt_1mab[b] * noise, b = batch[row] (batch NULL: b = 0) */
int cdseg_q_sample(const float* x0, const float* noise, const int32_t* batch, const float* sqrt_ab, const float* sqrt_1mab,
                   int64_t n, int C, float* out, void* stream);
/* continuous_p_ddim_sample, default.py:192-214, one timestep shared by all rows: the four scalars are sqrt(Alpha_bar[t]),
 * sqrt(1 - Alpha_bar[t]) and the same at t - 1; last != 0 returns the x0 estimate (t == 0) */
int cdseg_ddim_step(const float* x_t, const float* pred, int64_t total, float sqrt_ab, float sqrt_1mab, float sqrt_ab_prev,
                    float sqrt_1mab_prev, int target_is_x0, int last, float* out, void* stream);
/* y = (y + a * x) * scale  (the "avg" accumulation of inference_ddim, default.py:342, 359) */
int cdseg_axpy_scale(float* y, const float* x, float a, float scale, int64_t total, void* stream);

/* ---- test-time fragment pipeline --------------------------------------------------------------
 * GridSample(mode="test"), pointcept/datasets/transform.py:796-933: grid = floor(coord / grid_size) - min; key = FNV64-1A
 * (hash_fnv != 0) or the ravel hash of grid; order = stable argsort(key) (the reference's np.argsort leaves the order of
 * points inside one voxel unspecified; this library fixes it to ascending point index); voxel_of_point = GridSample's
 * `inverse`; start int32 [n + 1] (first V + 1 entries used, start[V] = n) and count int32 [n] (first V used) describe the
 * runs; stats int32 [8] on the device: [0] V = voxels, [1] F = max count = number of fragments, [2..4] / [5..7] min / max of
 * floor(coord / grid_size).  legacy_f32 != 0 divides in float32 (NumPy 1.x value-based casting of np.array(grid_size));
 * 0 divides in float64 (NumPy >= 2).  coord fp32 [n,3], or fp64 [n,3] with coord_f64 != 0 (the TTA rotations of transform.py:259-294
 * leave float64 coordinates); grid int32 [n,3]; key int64 [n] (uint64 bit pattern). */
size_t cdseg_grid_sample_workspace_bytes(int64_t n);
int cdseg_grid_sample_plan(const void* coord, int coord_f64, int64_t n, double grid_size, int hash_fnv, int legacy_f32, int32_t* grid,
                           int64_t* key, int32_t* order, int32_t* voxel_of_point, int32_t* start, int32_t* count,
                           int32_t* stats, void* workspace, size_t workspace_bytes, void* stream);
/* index[f][v] = order[start[v] + f % count[v]] for f < F, v < V: row f is fragment f's `index` (transform.py:868-870) */
int cdseg_fragment_index(const int32_t* order, const int32_t* start, int V, int F, int32_t* index, void* stream);
/* vote accumulation, pointcept/engines/test.py:252-257: pred[index[r], :] += softmax(logits[r, :]) */
int cdseg_vote_softmax_add(const float* logits, const int32_t* index, int64_t n, int C, float* pred, void* stream);
/* out[r] = argmax_c x[r, c] (lowest index on ties), test.py:268 */
int cdseg_argmax_rows(const float* x, int64_t n, int C, int64_t* out, void* stream);

/* ---- pointops.knn_query ---------------------------------------------------------------------------
 * Replaces knn_query_cuda_launcher(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2)
 * (libs/pointops/src/knn_query/knn_query_cuda_kernel.h:13, kernel .cu:60-104): for each of the m query points new_xyz fp32
 * [m,3] the nsample (<= 128) nearest points of xyz fp32 [n,3] inside the same batch (offset / new_offset int32 [B],
 * cumulative), idx int32 [m,nsample] ascending by squared distance (-1 = batch smaller than nsample), dist2 fp32
 * [m,nsample] SQUARED distances (1e10 for missing) -- the Python wrapper takes the sqrt like functions/query.py:26.
 * Exact (uniform cell grid + shell search instead of the reference's O(m n) scan); ties go to the lower point index.
 * Added to the reference signature: B, n, a caller-allocated workspace, the stream and the int status. */
size_t cdseg_knn_workspace_bytes(int64_t n);
int cdseg_knn_query(int m, int nsample, const float* xyz, const float* new_xyz, const int32_t* offset, const int32_t* new_offset,
                    int B, int64_t n, int32_t* idx, float* dist2, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training side of the criteria row (the backward of the NETWORK is not built yet) -----------
 * cdseg_criteria plus the gradients of the selected combination (gls = 0: EW sum / eval; gls = 1: sqrt(MSE * (CE + Lovasz)),
 * losses/builder.py:37-49) w.r.t. the two network outputs: grad_n_pred fp32 [n, C] (CE backward + the Lovasz gradient pulled
 * through the softmax; the Jaccard gradient is a constant of the sort order exactly as autograd sees lovasz.py:141-143) and
 * grad_c_pred fp32 [n, Cc] (NULL when has_mse = 0).  Same workspace size as cdseg_criteria. */
int cdseg_criteria_grad(const float* n_pred, const int64_t* n_target, int64_t n, int C, int64_t ignore_index, const float* c_pred,
                        const float* c_target, int Cc, int mse_use_ignore, float w_mse, float w_ce, float w_lov, int has_mse,
                        int has_ce, int has_lov, int gls, float* out5, float* grad_n_pred, float* grad_c_pred, void* workspace,
                        size_t workspace_bytes, void* stream);
/* one AdamW step (torch.optim.AdamW semantics, the optimizer of configs/scannet/CDSegNet.py:143 built by utils/optimizer.py:20-55)
 * over n_chunks table entries {float* param; const float* grad; float* exp_avg; float* exp_avg_sq; int64 numel} (device memory,
 * one entry per <= 64K-element chunk of a tensor): one launch per parameter group; step counts from 1 */
int cdseg_adamw_step(const void* table, int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif
