"""torch-tensor wrappers over the C-ABI kernels (include/cdseg_b200.h).

torch is used here for device memory and streams only: every wrapper allocates the
outputs (caller-allocates convention of the C ABI), passes raw device pointers and
the current CUDA stream, and raises on a non-zero status.  No wrapper synchronises.
"""
import ctypes
import numpy as np
import torch

from . import _lib
from ._lib import check

ORDER_IDS = {"z": 0, "z-trans": 1, "hilbert": 2, "hilbert-trans": 3}
PROFILE = None     # bench.py sets this to a list: per Block a dict(ev=[6 raw cudaEvent_t], n, C, H, pairs, has_t) (see Block._native)


def _p(t, dtype=None):
    """raw device pointer of a tensor (None -> NULL).  The checks are cheap attribute reads; a CPU tensor, a
    non-contiguous view or a wrong dtype must raise (there is no CPU fallback and no silent conversion)."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous() or (dtype is not None and t.dtype is not dtype):
        _bad(t, dtype)
    return t.data_ptr()


def _bad(t, dtype):
    if not t.is_cuda:
        raise _lib.CdsegError("cdsegnet_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise _lib.CdsegError("tensor must be contiguous")
    raise _lib.CdsegError(f"expected {dtype}, got {t.dtype}")


_STREAM = [None]        # raw cudaStream_t pinned by stream_scope(); torch.cuda.current_stream() costs ~4 us per call


def _stream():
    s = _STREAM[0]
    return s if s is not None else torch.cuda.current_stream().cuda_stream


class stream_scope:
    """`with stream_scope(stream):` makes `stream` torch's current stream AND pins its raw handle for the kernel
    wrappers, so the ~1000 launches of a forward do not each query torch for the current stream."""

    def __init__(self, stream):
        self.stream = stream
        self.ctx = torch.cuda.stream(stream)

    def __enter__(self):
        self.ctx.__enter__()
        self.prev = _STREAM[0]
        _STREAM[0] = self.stream.cuda_stream
        return self

    def __exit__(self, *a):
        _STREAM[0] = self.prev
        return self.ctx.__exit__(*a)


def _ws(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------- serialization
def grid_max(grid):
    out = torch.empty(1, dtype=torch.int32, device=grid.device)
    check(_lib.load().cdseg_grid_max(_p(grid, torch.int32), grid.numel(), _p(out), _stream()), "grid_max")
    return out


def offset2batch(offset, n):
    batch = torch.empty(n, dtype=torch.int32, device=offset.device)
    check(_lib.load().cdseg_offset2batch(_p(offset, torch.int64), offset.numel(), n, _p(batch), _stream()), "offset2batch")
    return batch


def encode_codes(grid, batch, depth, orders):
    n = grid.shape[0]
    k = len(orders)
    ids = (ctypes.c_int * k)(*[ORDER_IDS[o] for o in orders])
    codes = torch.empty((k, n), dtype=torch.int64, device=grid.device)
    check(_lib.load().cdseg_encode_codes(_p(grid, torch.int32), _p(batch, torch.int32), n, depth, ids, k, _p(codes), _stream()),
          "encode_codes")
    return codes


def argsort_rows(codes, nbits):
    k, n = codes.shape
    lib = _lib.load()
    order = torch.empty((k, n), dtype=torch.int32, device=codes.device)
    inverse = torch.empty((k, n), dtype=torch.int32, device=codes.device)
    nb = lib.cdseg_argsort_workspace_bytes(k, n)
    ws = _ws(nb, codes.device)
    check(lib.cdseg_argsort_rows(_p(codes, torch.int64), k, n, nbits, _p(order), _p(inverse), _p(ws), nb, _stream()),
          "argsort_rows")
    return order, inverse


def gather_rows(src, idx):
    """src[idx] for a contiguous 2-D (or 1-D) tensor with 4-byte-multiple rows and an int32 index vector"""
    n = idx.shape[0]
    out = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    row_bytes = src.element_size() * (src[0].numel() if src.dim() > 1 else 1)
    check(_lib.load().cdseg_gather_rows(_p(src), _p(idx, torch.int32), n, row_bytes, _p(out), _stream()), "gather_rows")
    return out


def renumber(perm, inv_perm, grid, batch, code, order, inverse):
    """level-0 tables in curve-order numbering (see include/cdseg_b200.h) -> (i_grid, i_batch, i_code, i_order, i_inverse)"""
    k, N = code.shape
    outs = (torch.empty_like(grid), torch.empty_like(batch), torch.empty_like(code), torch.empty_like(order), torch.empty_like(inverse))
    check(_lib.load().cdseg_renumber(_p(perm, torch.int32), _p(inv_perm, torch.int32), _p(grid, torch.int32), _p(batch, torch.int32),
                                     _p(code, torch.int64), _p(order, torch.int32), _p(inverse, torch.int32), k, N, *[_p(o) for o in outs],
                                     _stream()), "renumber")
    return outs


def patch_maps(order_row, scene_count, K):
    """order_row int32 [n] (one curve).  Returns dict(slot_src, slot_dst, point_slot, patch_len, T, Kp)."""
    lib = _lib.load()
    B = len(scene_count)
    cnt = (ctypes.c_int64 * B)(*[int(c) for c in scene_count])
    T = ctypes.c_int(0)
    check(lib.cdseg_patch_count(cnt, B, K, ctypes.byref(T)), "patch_count")
    T = T.value
    Kp = (K + 127) // 128 * 128
    dev = order_row.device
    n = int(sum(int(c) for c in scene_count))
    slot_src = torch.empty(T * Kp, dtype=torch.int32, device=dev)
    slot_dst = torch.empty(T * Kp, dtype=torch.int32, device=dev)
    point_slot = torch.empty(n, dtype=torch.int32, device=dev)
    patch_len = torch.empty(T, dtype=torch.int32, device=dev)
    check(lib.cdseg_patch_maps(_p(order_row, torch.int32), cnt, B, K, Kp, _p(slot_src), _p(slot_dst), _p(point_slot),
                               _p(patch_len), _stream()), "patch_maps")
    # algorithmic (query, key) pairs of this partition, padding rows excluded
    pairs = int(sum(int(c) * (K if int(c) > K else int(c)) for c in scene_count))
    return dict(slot_src=slot_src, slot_dst=slot_dst, point_slot=point_slot, patch_len=patch_len, T=T, Kp=Kp, K=K,
                pairs=pairs)


# ---------------------------------------------------------------- pooling
def pool_plan(code, order, n_dev, n_host, c0, pooling_depth, grid, batch, B, cap_child, m_dev=None, offset=None):
    """One sync-free pooling level.  code/order: [k, ld].  Returns the child arrays (capacity cap_child).
    m_dev int32 [1] / offset int64 [B]: optional pre-zeroed outputs (the Plan hands out slices of one buffer per forward)."""
    lib = _lib.load()
    k, ld = code.shape
    dev = code.device
    cap_parent = ld
    out = dict(
        cluster=torch.empty(cap_parent, dtype=torch.int32, device=dev),
        idx_ptr=torch.empty(cap_child + 1, dtype=torch.int32, device=dev),
        head=torch.empty(cap_child, dtype=torch.int32, device=dev),
        code=torch.empty((k, cap_child), dtype=torch.int64, device=dev),
        order=torch.empty((k, cap_child), dtype=torch.int32, device=dev),
        inverse=torch.empty((k, cap_child), dtype=torch.int32, device=dev),
        grid=torch.empty((cap_child, 3), dtype=torch.int32, device=dev),
        batch=torch.empty(cap_child, dtype=torch.int32, device=dev),
        m_dev=m_dev if m_dev is not None else torch.zeros(1, dtype=torch.int32, device=dev),
        offset=offset if offset is not None else torch.zeros(B, dtype=torch.int64, device=dev),
    )
    nb = lib.cdseg_pool_plan_workspace_bytes(k, ld)
    ws = _ws(nb, dev)
    check(lib.cdseg_pool_plan(_p(code, torch.int64), _p(order, torch.int32), k, ld, _p(n_dev), int(n_host), c0,
                              pooling_depth, _p(grid, torch.int32), _p(batch, torch.int32), _p(out["cluster"]),
                              _p(out["idx_ptr"]), _p(out["head"]), _p(out["code"]), _p(out["order"]),
                              _p(out["inverse"]), cap_child, _p(out["grid"]), _p(out["batch"]), _p(out["m_dev"]),
                              _p(out["offset"]), _p(ws), nb, _stream()), "pool_plan")
    return out


def pool_reduce(x, coord, members, idx_ptr, m, bn_scale=None, bn_shift=None, gelu=False):
    C = x.shape[1]
    out = torch.empty((m, C), dtype=torch.float32, device=x.device)
    out_coord = torch.empty((m, 3), dtype=torch.float32, device=x.device) if coord is not None else None
    check(_lib.load().cdseg_pool_reduce(_p(x, torch.float32), _p(coord), _p(members, torch.int32), _p(idx_ptr, torch.int32),
                                        m, C, _p(bn_scale), _p(bn_shift), int(gelu), _p(out), _p(out_coord), _stream()),
          "pool_reduce")
    return out, out_coord


def unpool_add(a, b, cluster, alpha=1.0):
    n, C = a.shape
    out = torch.empty_like(a)
    check(_lib.load().cdseg_unpool_add(_p(a, torch.float32), _p(b, torch.float32), _p(cluster, torch.int32), n, C,
                                       float(alpha), _p(out), _stream()), "unpool_add")
    return out


# ---------------------------------------------------------------- submanifold conv
def nbr_build(grid, batch, ksize):
    lib = _lib.load()
    n = grid.shape[0]
    nbr = torch.empty((n, ksize ** 3), dtype=torch.int32, device=grid.device)
    nb = lib.cdseg_nbr_workspace_bytes(n)
    ws = _ws(nb, grid.device)
    check(lib.cdseg_nbr_build(_p(grid, torch.int32), _p(batch, torch.int32), n, ksize, _p(nbr), _p(ws), nb, _stream()),
          "nbr_build")
    return nbr


def subm_conv(x, nbr, wt, bias, ksize, ep_scale=None, ep_shift=None, ep_gelu=False):
    """wt: fp32 [k^3, Ci, Co] (tap-major transposed weight)."""
    n, Ci = x.shape
    Co = wt.shape[2]
    out = torch.empty((n, Co), dtype=torch.float32, device=x.device)
    check(_lib.load().cdseg_subm_conv(_p(x, torch.float32), _p(nbr, torch.int32), _p(wt, torch.float32), _p(bias),
                                      _p(ep_scale), _p(ep_shift), int(ep_gelu), n, Ci, Co, ksize, _p(out), _stream()),
          "subm_conv")
    return out


def conv_im2col_tc(x8, nbr, Bp, N, bias, act):
    """tiny-C_in sparse conv as an im2col GEMM (x8: fp32 [n, 8] zero-padded input, nbr int32 [n, taps])."""
    n = nbr.shape[0]
    out = torch.empty((n, N), dtype=torch.float32, device=x8.device)
    check(_lib.load().cdseg_conv_im2col_tc(_p(x8, torch.float32), _p(nbr, torch.int32), nbr.shape[1], _p(Bp, torch.float32),
                                           n, N, _p(bias), act, _p(out), N, _stream()), "conv_im2col_tc")
    return out


# ---------------------------------------------------------------- attention
ATTN_KERNEL = 3       # tcgen05 attention generation behind mode "f16": 3 = attn_tc3.cu (deferred fold, double-buffered P / O),
                      # 2 = attn_tc2.cu (round 1), 1 = attn_tc.cu (whole-patch K/V image); 2 and 1 are kept as A/B comparators
ATTN_MODES = ("f16", "exact", "tc32")      # index = CDSEG_ATTN_F16 / _EXACT / _TC32 (include/cdseg_b200.h)


def attn_pack(src, col0, C, nwhich, pm, H, mode="f16", has_v=True):
    """Gather rows of src (fp32 [n, ld]) by pm['slot_src'] into per-(head, patch) operand tiles of the attention kernel of `mode`:
      "f16"   fp16 core-matrix tiles, V 32 wide with the ones column (flash-branch numerics, ptv3.py:282-289)
      "tc32"  fp16 hi | lo halves of q / k, V 48 wide [v_hi | 1 | v_lo] (fp32-class results on the tensor cores)
      "exact" plain fp32 rows (SIMT dense-branch kernel, ptv3.py:264-280)"""
    lib = _lib.load()
    T, Kp = pm["T"], pm["Kp"]
    dev = src.device
    if mode == "exact":
        bufs = [torch.empty((H, T, Kp, 16), dtype=torch.float32, device=dev) for _ in range(nwhich)]
        ptrs = [_p(b) for b in bufs] + [None] * (3 - nwhich)
        check(lib.cdseg_attn_pack_f32(_p(src, torch.float32), src.shape[1], col0, C, nwhich, _p(pm["slot_src"]), H, T, Kp, *ptrs,
                                      _stream()), "attn_pack")
        return bufs
    if mode == "tc32":
        bufs = [torch.empty((H, T, Kp, 48) if (has_v and i == nwhich - 1) else (2, H, T, Kp, 16), dtype=torch.float16, device=dev)
                for i in range(nwhich)]
        ptrs = [_p(b) for b in bufs] + [None] * (3 - nwhich)
        check(lib.cdseg_attn_pack_split(_p(src, torch.float32), src.shape[1], col0, C, nwhich, _p(pm["slot_src"]), H, T, Kp, *ptrs,
                                        int(has_v), _stream()), "attn_pack_split")
        return bufs
    if mode != "f16":
        raise ValueError(f"attention mode {mode!r} not in {ATTN_MODES}")
    v32 = ATTN_KERNEL >= 2 and has_v
    bufs = [torch.empty((H, T, Kp, 32 if (v32 and i == nwhich - 1) else 16), dtype=torch.float16, device=dev) for i in range(nwhich)]
    ptrs = [_p(b) for b in bufs] + [None] * (3 - nwhich)
    check(lib.cdseg_attn_pack_f16v(_p(src, torch.float32), src.shape[1], col0, C, nwhich, _p(pm["slot_src"]), H, T, Kp, *ptrs,
                                   int(v32), _stream()), "attn_pack")
    return bufs


def attn(q, k, v, pm, H, scale, n_out, mode="f16"):
    """softmax(q k^T * scale) v per (patch, head); rows scattered to original point order."""
    lib = _lib.load()
    C = H * 16
    out = torch.empty((n_out, C), dtype=torch.float32, device=q.device)
    args = (_p(q), _p(k), _p(v), _p(pm["patch_len"]), _p(pm["slot_dst"]), H, pm["T"], pm["Kp"], float(scale))
    if mode == "exact":
        st = lib.cdseg_attn_exact(*args, _p(out), C, _stream())
    elif mode == "tc32":
        st = lib.cdseg_attn_tc3(*args, 1, _p(out), C, _stream())
    elif ATTN_KERNEL == 3:
        st = lib.cdseg_attn_tc3(*args, 0, _p(out), C, _stream())
    else:
        st = (lib.cdseg_attn_tc2 if v.shape[-1] == 32 else lib.cdseg_attn_tc)(*args, _p(out), C, _stream())
    check(st, "attn")
    return out


# ---------------------------------------------------------------- row-wise
def add_layernorm(a, b=None, t=None, batch=None, gamma=None, beta=None, eps=1e-5, want_sum=True, want_ln=True):
    n, C = a.shape
    y = torch.empty_like(a) if want_sum else None
    ln = torch.empty_like(a) if want_ln else None
    check(_lib.load().cdseg_add_layernorm(_p(a, torch.float32), _p(b), _p(t), _p(batch), _p(gamma), _p(beta), float(eps),
                                          n, C, _p(y), _p(ln), _stream()), "add_layernorm")
    return y, ln


def scale_shift_act(x, scale=None, shift=None, act=0):
    n, C = x.shape
    out = torch.empty_like(x)
    check(_lib.load().cdseg_scale_shift_act(_p(x, torch.float32), _p(scale), _p(shift), act, n, C, _p(out), _stream()),
          "scale_shift_act")
    return out


def small_linear(x, W, bias, act=0):
    R, K = x.shape
    O = W.shape[0]
    out = torch.empty((R, O), dtype=torch.float32, device=x.device)
    check(_lib.load().cdseg_small_linear(_p(x, torch.float32), _p(W, torch.float32), _p(bias), act, R, K, O, _p(out),
                                         _stream()), "small_linear")
    return out


def rows_uniform_flag(x, batch, offset, flag):
    n, C = x.shape
    check(_lib.load().cdseg_rows_uniform(_p(x, torch.float32), _p(batch, torch.int32), _p(offset, torch.int64), n, C,
                                         _p(flag), _stream()), "rows_uniform")


def launch_count():
    return int(_lib.load().cdseg_launch_count())


def launch_count_reset():
    _lib.load().cdseg_launch_count_reset()


# ---------------------------------------------------------------- tensor-core GEMM (3xTF32)
def gemm_pack_b(W):
    """W fp32 [T, K, N] -> packed operand blocks for gemm_tc (cache the result per weight)."""
    lib = _lib.load()
    T, K, N = W.shape
    Bp = torch.empty(lib.cdseg_gemm_packed_b_floats(T, K, N), dtype=torch.float32, device=W.device)
    check(lib.cdseg_gemm_pack_b(_p(W, torch.float32), T, K, N, _p(Bp), _stream()), "gemm_pack_b")
    return Bp


def tile_tap_mask(nbr):
    M, T = nbr.shape
    mask = torch.empty((M + 127) // 128, dtype=torch.int32, device=nbr.device)
    check(_lib.load().cdseg_tile_tap_mask(_p(nbr, torch.int32), M, T, _p(mask), _stream()), "tile_tap_mask")
    return mask


def post_attn(o, x1, proj, ln, fc1, fc2, eps=1e-5):
    """x2 = x1 + proj(o); out = x2 + fc2(GELU(fc1(LN(x2)))) in one kernel.  proj/fc1/fc2 = (packed blocks, bias); ln = (gamma, beta)"""
    n, C = o.shape
    out = torch.empty_like(o)
    check(_lib.load().cdseg_post_attn(_p(o, torch.float32), _p(x1, torch.float32), n, C, _p(proj[0]), _p(proj[1]), _p(ln[0]), _p(ln[1]),
                                      float(eps), _p(fc1[0]), _p(fc1[1]), _p(fc2[0]), _p(fc2[1]), _p(out), _stream()), "post_attn")
    return out


CONV_PLAN_UCAP = 384      # CDSEG_CONV_PLAN_UCAP (include/cdseg_b200.h)


def conv_tile_plan(nbr):
    """per 128-row tile: distinct neighbour rows + local index of every (row, tap) (uint8 records, see the header)"""
    lib = _lib.load()
    n = nbr.shape[0]
    plan = torch.empty(lib.cdseg_conv_plan_bytes(n), dtype=torch.uint8, device=nbr.device)
    check(lib.cdseg_conv_tile_plan(_p(nbr, torch.int32), n, _p(plan), _stream()), "conv_tile_plan")
    return plan


def pre_attn(conv_in, x, nbr, tile_mask, plan, conv, lin, cpe_ln, n1_ln, qkv, tproj=None, batch=None, eps=1e-5):
    """x1 = x + LN_cpe(lin(conv3(conv_in))) (+ tproj[batch]); qkv = qkv_lin(LN_1(x1)) in one kernel.
    conv/lin/qkv = (packed blocks, bias); *_ln = (gamma, beta).  Returns (x1, qkv)."""
    n, C = x.shape
    x1 = torch.empty_like(x)
    out = torch.empty((n, 3 * C), dtype=torch.float32, device=x.device)
    check(_lib.load().cdseg_pre_attn(_p(conv_in, torch.float32), _p(x, torch.float32), n, C, _p(nbr, torch.int32), _p(tile_mask),
                                     _p(plan, torch.uint8), _p(conv[0]), _p(conv[1]), _p(lin[0]), _p(lin[1]), _p(cpe_ln[0]), _p(cpe_ln[1]), _p(tproj),
                                     _p(batch), _p(n1_ln[0]), _p(n1_ln[1]), float(eps), _p(qkv[0]), _p(qkv[1]), _p(x1), _p(out),
                                     _stream()), "pre_attn")
    return x1, out


def set_gemm_precision(kind):
    """"fp32": every dense layer / conv on the 3-term fp16 hi/lo split (fp32-class results, default); "fp16": fp16 operands with fp32
    accumulation, one MMA per product term -- the numerics of the reference's autocast training (engines/train.py:226)"""
    if kind not in ("fp32", "fp16"):
        raise ValueError(kind)
    _lib.load().cdseg_set_gemm_precision(int(kind == "fp16"))


def set_fused_mask(mask):
    _lib.load().cdseg_set_fused_mask(int(mask))


GEMM_MODE = "tc"      # "tc": tcgen05 split-fp16 kernels for conv + linears ; "simt": round-1a SIMT conv + cuBLAS SGEMM
NATIVE_BLOCKS = True  # run each PTv3 Block through the C++ executor (cdseg_block_forward) instead of launch-by-launch
NATIVE_NET = True     # run the whole feature phase through ONE C-ABI call (cdseg_net_forward, csrc/net_exec.cu); False: per-module calls

_ARENAS = {}


def arena(nbytes, device):
    """grow-only scratch arena per (device, CUDA stream): blocks on one stream run back to back, so they can share it"""
    key = (device.index, _stream())
    a = _ARENAS.get(key)
    if a is None or a.numel() < nbytes:
        a = torch.empty(int(nbytes * 1.25), dtype=torch.uint8, device=device)
        _ARENAS[key] = a
    return a


def pick_split(tiles, T):
    """taps/K-slices are split over grid.z until ~200 CTAs exist (levels with few rows would otherwise
    leave most of the 148 SMs idle while one CTA streams the whole weight)"""
    if tiles >= 120 or T == 1:
        return 1
    if T == 27:      # deep-level convs: the (possibly uneven) split that fills one wave of 2 x 148 CTAs; same rule as csrc/net_exec.cu
        best, best_cost = 1, None
        for s in range(1, T + 1):
            cost = -(-tiles * s // 296) * (-(-T // s) + 1)
            if best_cost is None or cost < best_cost:
                best, best_cost = s, cost
        return best
    want = -(-200 // tiles)
    for s in range(1, T + 1):
        if T % s == 0 and s >= want:
            return s
    return T


def gemm_tc(A, Bp, N, K, idx=None, tile_mask=None, bias=None, res=None, act=0, nsplit=1, M=None, T=None):
    """out[M,N] = act(bias + sum_t A[idx[:,t]] @ W_t) + res, fp32 in/out, tcgen05 3xTF32 inside.
    Without idx, T > 1 means: A row m is cut into T K-slices of width K (split-K of a Linear)."""
    lib = _lib.load()
    if T is None:
        T = idx.shape[1] if idx is not None else 1
    M = (idx.shape[0] if idx is not None else A.shape[0]) if M is None else M
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    nb = lib.cdseg_gemm_tc_workspace_bytes(M, N, nsplit)
    ws = _ws(nb, A.device) if nb else None
    check(lib.cdseg_gemm_tc(_p(A, torch.float32), A.shape[1], _p(idx), T, _p(tile_mask), _p(Bp, torch.float32), M, N, K,
                            _p(bias), _p(res), res.shape[1] if res is not None else 0, act, _p(out), N, nsplit, _p(ws), nb,
                            _stream()), "gemm_tc")
    return out


# ---------------------------------------------------------------- wrapper: criteria + diffusion samplers (losses.cu)
def criteria(n_pred, n_target, ignore_index=-1, c_pred=None, c_target=None, mse_use_ignore=True, weights=(1.0, 1.0, 1.0),
             has=(True, True, True)):
    """forward values of MSE / CE / Lovasz and their EW / GLS combinations -> fp32 [5] on the device
    ([0] MSE, [1] CE, [2] Lovasz, [3] sum, [4] sqrt(MSE * (CE + Lovasz))); see include/cdseg_b200.h"""
    lib = _lib.load()
    n, C = n_pred.shape
    has_mse = bool(has[0]) and c_pred is not None and c_target is not None
    out = torch.empty(5, dtype=torch.float32, device=n_pred.device)
    nb = lib.cdseg_criteria_workspace_bytes(n, C)
    ws = _ws(nb, n_pred.device)
    check(lib.cdseg_criteria(_p(n_pred, torch.float32), _p(n_target, torch.int64), n, C, int(ignore_index),
                             _p(c_pred, torch.float32) if has_mse else None, _p(c_target, torch.float32) if has_mse else None,
                             c_pred.shape[1] if has_mse else 0, int(bool(mse_use_ignore)), float(weights[0]), float(weights[1]),
                             float(weights[2]), int(has_mse), int(bool(has[1])), int(bool(has[2])), _p(out), _p(ws), nb, _stream()),
          "criteria")
    return out


def q_sample(x0, noise, batch, sqrt_ab, sqrt_1mab):
    """x_t = sqrt_ab[batch] * x0 + sqrt_1mab[batch] * noise (default.py:216-222); sqrt_* fp32 [B] on the device"""
    out = torch.empty_like(x0)
    n, C = x0.shape
    check(_lib.load().cdseg_q_sample(_p(x0, torch.float32), _p(noise, torch.float32), _p(batch, torch.int32), _p(sqrt_ab, torch.float32),
                                     _p(sqrt_1mab, torch.float32), n, C, _p(out), _stream()), "q_sample")
    return out


def ddim_step(x_t, pred, sqrt_ab, sqrt_1mab, sqrt_ab_prev, sqrt_1mab_prev, target_is_x0, last):
    out = torch.empty_like(x_t)
    check(_lib.load().cdseg_ddim_step(_p(x_t, torch.float32), _p(pred, torch.float32), x_t.numel(), float(sqrt_ab), float(sqrt_1mab),
                                      float(sqrt_ab_prev), float(sqrt_1mab_prev), int(target_is_x0), int(last), _p(out), _stream()),
          "ddim_step")
    return out


def axpy_scale_(y, x, a=1.0, scale=1.0):
    check(_lib.load().cdseg_axpy_scale(_p(y, torch.float32), _p(x, torch.float32), float(a), float(scale), y.numel(), _stream()),
          "axpy_scale")
    return y


# ---------------------------------------------------------------- test-time fragment pipeline (fragments.cu)
def grid_sample_plan(coord, grid_size, hash_type="fnv", legacy_f32=False):
    """GridSample(mode="test") plan of one raw scene (transform.py:825-870).  coord fp32 or fp64 [n,3] on the device.
    -> dict(grid_coord int32 [n,3], key int64 [n], order int32 [n], inverse int32 [n], start int32 [V+1], count int32 [V],
            n_voxels, n_fragments, min_grid [3])  -- ONE host sync (8 ints) to learn V and F."""
    lib = _lib.load()
    n = coord.shape[0]
    dev = coord.device
    i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
    grid, key, order, vox, start, count, stats = i32(n, 3), torch.empty(n, dtype=torch.int64, device=dev), i32(n), i32(n), i32(n + 1), i32(n), i32(8)
    nb = lib.cdseg_grid_sample_workspace_bytes(n)
    ws = _ws(nb, dev)
    if coord.dtype not in (torch.float32, torch.float64):
        raise _lib.CdsegError(f"coord must be float32 or float64, got {coord.dtype}")
    check(lib.cdseg_grid_sample_plan(_p(coord), int(coord.dtype is torch.float64), n, float(grid_size), int(hash_type == "fnv"), int(bool(legacy_f32)),
                                     _p(grid), _p(key), _p(order), _p(vox), _p(start), _p(count), _p(stats), _p(ws), nb, _stream()),
          "grid_sample_plan")
    st = stats.cpu().tolist()
    V, F = st[0], st[1]
    return dict(grid_coord=grid, key=key, order=order, inverse=vox, start=start[:V + 1], count=count[:V], n_voxels=V, n_fragments=F,
                min_grid=st[2:5])


def fragment_index(order, start, n_voxels, n_fragments):
    """int32 [F, V]: row f = the `index` of fragment f"""
    index = torch.empty((n_fragments, n_voxels), dtype=torch.int32, device=order.device)
    check(_lib.load().cdseg_fragment_index(_p(order, torch.int32), _p(start, torch.int32), n_voxels, n_fragments, _p(index), _stream()),
          "fragment_index")
    return index


def vote_softmax_add_(pred, logits, index):
    check(_lib.load().cdseg_vote_softmax_add(_p(logits, torch.float32), _p(index, torch.int32), logits.shape[0], logits.shape[1],
                                             _p(pred, torch.float32), _stream()), "vote_softmax_add")
    return pred


def argmax_rows(x):
    out = torch.empty(x.shape[0], dtype=torch.int64, device=x.device)
    check(_lib.load().cdseg_argmax_rows(_p(x, torch.float32), x.shape[0], x.shape[1], _p(out), _stream()), "argmax_rows")
    return out


def criteria_grad(n_pred, n_target, ignore_index=-1, c_pred=None, c_target=None, mse_use_ignore=True, weights=(1.0, 1.0, 1.0),
                  has=(True, True, True), gls=False):
    """criteria() plus d loss / d n_pred and d loss / d c_pred of the EW (gls=False) or GLS combination -> (out5, grad_n, grad_c)"""
    lib = _lib.load()
    n, C = n_pred.shape
    has_mse = bool(has[0]) and c_pred is not None and c_target is not None
    out = torch.empty(5, dtype=torch.float32, device=n_pred.device)
    gn = torch.empty_like(n_pred)
    gc = torch.empty_like(c_pred) if has_mse else None
    nb = lib.cdseg_criteria_workspace_bytes(n, C)
    ws = _ws(nb, n_pred.device)
    check(lib.cdseg_criteria_grad(_p(n_pred, torch.float32), _p(n_target, torch.int64), n, C, int(ignore_index),
                                  _p(c_pred, torch.float32) if has_mse else None, _p(c_target, torch.float32) if has_mse else None,
                                  c_pred.shape[1] if has_mse else 0, int(bool(mse_use_ignore)), float(weights[0]), float(weights[1]),
                                  float(weights[2]), int(has_mse), int(bool(has[1])), int(bool(has[2])), int(bool(gls)), _p(out), _p(gn),
                                  _p(gc), _p(ws), nb, _stream()), "criteria_grad")
    return out, gn, gc


def reduce_ln(part, nsplit, bias=None, ln1=None, res=None, t=None, batch=None, ln2=None, eps=1e-5, want_y=True):
    """v = bias + sum_z part[z] -> [LN(ln1)] -> + res (+ t[batch]) -> y -> LN(ln2): (y or None, ln or None).  part: [nsplit, n, C] fp32"""
    _, n, C = part.shape
    y = torch.empty((n, C), dtype=torch.float32, device=part.device) if want_y else None
    o = torch.empty((n, C), dtype=torch.float32, device=part.device) if ln2 is not None else None
    check(_lib.load().cdseg_reduce_ln(_p(part, torch.float32), nsplit, _p(bias), _p(ln1[0]) if ln1 else None, _p(ln1[1]) if ln1 else None,
                                      _p(res), _p(t), _p(batch), _p(ln2[0]) if ln2 else None, _p(ln2[1]) if ln2 else None, float(eps), n, C,
                                      _p(y), _p(o), _stream()), "reduce_ln")
    return y, o
