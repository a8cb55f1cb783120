"""Host side of the native feature-phase executor (csrc/net_exec.cu, include/cdseg_b200.h "CdsegNetW" / "CdsegForwardArgs").

`NetWeights(model)` describes the module tree of a `PointTransformerV3` (same parameters, same names -- ptv3.py:1340-1846 of the
reference) to the C side ONCE: packed tensor-core operands, folded eval-mode BatchNorms, LayerNorm vectors.  It is rebuilt only
when a parameter / buffer changed (version counters, device moves).  `forward_native` then runs one whole forward with TWO
C-ABI calls: `cdseg_plan_build` (structure.Plan) and `cdseg_net_forward`.
"""
import ctypes

import torch

from . import _lib, ops
from ._lib import _F, _I, _L, _P, _Z, PlanLevel, check

MAX_STAGES = 8


class LinW(ctypes.Structure):
    _fields_ = [("Bp", _P), ("bias", _P), ("K", _I), ("N", _I)]


class LnW(ctypes.Structure):
    _fields_ = [("g", _P), ("b", _P)]


class BlockW(ctypes.Structure):
    _fields_ = [("C", _I), ("H", _I), ("T_dim", _I), ("order_index", _I), ("scale", _F), ("ln_eps", _F),
                ("conv_Bp", _P), ("conv_b", _P), ("lin", LinW), ("cpe_ln", LnW), ("t_W", _P), ("t_b", _P),
                ("n1", LnW), ("qkv", LinW), ("proj", LinW), ("n2", LnW), ("fc1", LinW), ("fc2", LinW)]


class PoolW(ctypes.Structure):
    _fields_ = [("proj", LinW), ("bn_scale", _P), ("bn_shift", _P)]


class UnpoolW(ctypes.Structure):
    _fields_ = [("proj", LinW), ("proj_skip", LinW), ("cat_a", LinW), ("cat_b", LinW), ("cat", _I), ("alpha", _F)]


class StageW(ctypes.Structure):
    _fields_ = [("n_blocks", _I), ("has_pool", _I), ("has_up", _I), ("level", _I), ("blocks", ctypes.POINTER(BlockW)),
                ("pool", PoolW), ("up", UnpoolW)]


class StemW(ctypes.Structure):
    _fields_ = [("Bp", _P), ("shift", _P), ("cin", _I), ("cout", _I), ("ksize", _I), ("pad_", _I)]


class CrossW(ctypes.Structure):
    _fields_ = [("Cq", _I), ("Ckv", _I), ("H", _I), ("K", _I), ("scale", _F), ("tm_feat", _F), ("ln_eps", _F), ("pad_", _I),
                ("q_conv_Bp", _P), ("q_conv_b", _P), ("q_lin", LinW), ("q_cpe_ln", LnW),
                ("kv_conv_Bp", _P), ("kv_conv_b", _P), ("kv_lin", LinW), ("kv_cpe_ln", LnW),
                ("q_norm1", LnW), ("kv_norm1", LnW), ("q_norm2", LnW), ("q", LinW), ("kv", LinW), ("proj", LinW), ("fc1", LinW),
                ("fc2", LinW)]


class NetW(ctypes.Structure):
    _fields_ = [("condition", _I), ("T_dim", _I), ("n_enc", _I), ("n_dec", _I), ("c_enc", _I), ("c_dec", _I), ("pad0_", _I), ("pad1_", _I),
                ("n_stem", StemW), ("c_stem", StemW),
                ("n_enc_st", StageW * MAX_STAGES), ("n_dec_st", StageW * MAX_STAGES), ("c_enc_st", StageW * MAX_STAGES),
                ("c_dec_st", StageW * MAX_STAGES), ("n_head", LinW), ("c_head", LinW),
                ("fc_t1_W", _P), ("fc_t1_b", _P), ("fc_t2_W", _P), ("fc_t2_b", _P), ("tm", CrossW)]


class ForwardArgs(ctypes.Structure):
    _fields_ = [("w", ctypes.POINTER(NetW)), ("levels", ctypes.POINTER(PlanLevel)), ("n_lv_n", _I), ("n_lv_c", _I), ("N", _L), ("B", _I),
                ("attn_mode", _I), ("n_feat", _P), ("c_feat", _P), ("t_emb", _P), ("n_out", _P), ("c_out", _P),
                ("arena_main", _P), ("arena_main_bytes", _Z), ("arena_side", _P), ("arena_side_bytes", _Z),
                ("stream_main", _P), ("stream_side", _P), ("block_events", ctypes.POINTER(_P))]


def _f32(t):
    """raw pointer of a contiguous fp32 CUDA tensor (None -> NULL)"""
    return ops._p(t.detach() if t is not None else None, torch.float32)


class NetWeights:
    """C-side description of one PointTransformerV3's weights; keeps every packed tensor alive."""

    def __init__(self, model):
        from . import ptv3
        self.keep = []
        self.blocks = []                     # execution order: CN (enc stages, then dec stages), then NN
        w = NetW()
        self.w = w

        def lin(weight, bias=None, bn=None, cols=None):
            if bn is None and cols is None:
                Bp, b = ptv3.packed_linear(weight, bias)                  # a missing bias becomes zeros (the fused kernels always add one)
            else:
                Bp, b = ptv3.packed_folded(weight, bias, bn, cols)       # eval-mode BatchNorm folded; bias may stay None
            self.keep += [Bp, b]
            K = weight.shape[1] if cols is None else cols[1] - cols[0]
            return LinW(_f32(Bp), _f32(b), int(K), int(weight.shape[0]))

        def ln(m):
            return LnW(_f32(m.weight), _f32(m.bias))

        def conv_pack(conv):
            Bp = ptv3._PACK.get((id(conv.weight), "conv"), [conv.weight], lambda: ops.gemm_pack_b(conv.wt()))
            self.keep.append(Bp)
            return _f32(Bp), _f32(conv.bias)

        def block(b):
            bw = BlockW()
            a = b.attn
            bw.C, bw.H, bw.T_dim, bw.order_index = a.channels, a.num_heads, b.T_dim, a.order_index
            bw.scale, bw.ln_eps = float(a.scale), float(b.norm1[0].eps)
            bw.conv_Bp, bw.conv_b = conv_pack(b.cpe[0])
            bw.lin, bw.cpe_ln = lin(b.cpe[1].weight, b.cpe[1].bias), ln(b.cpe[2])
            if b.T_dim != -1:
                bw.t_W, bw.t_b = _f32(b.t_mlp.weight), _f32(b.t_mlp.bias)
            bw.n1, bw.n2 = ln(b.norm1[0]), ln(b.norm2[0])
            bw.qkv, bw.proj = lin(a.qkv.weight, a.qkv.bias), lin(a.proj.weight, a.proj.bias)
            bw.fc1, bw.fc2 = lin(b.mlp[0].fc1.weight, b.mlp[0].fc1.bias), lin(b.mlp[0].fc2.weight, b.mlp[0].fc2.bias)
            return bw

        def stage(dst, mod, level, net):
            blocks = [m for name, m in mod._modules.items() if name.startswith("block")]
            arr = (BlockW * max(len(blocks), 1))(*[block(b) for b in blocks])
            self.keep.append(arr)
            dst.n_blocks, dst.level, dst.blocks = len(blocks), level, ctypes.cast(arr, ctypes.POINTER(BlockW))
            for b in blocks:
                self.blocks.append((net, level, b))
            if "down" in mod._modules:
                d = mod._modules["down"]
                sc, sh = ptv3.bn_fold(d.norm[0])
                self.keep += [sc, sh]
                dst.has_pool = 1
                dst.pool = PoolW(lin(d.proj.weight, d.proj.bias), _f32(sc), _f32(sh))
            if "up" in mod._modules:
                u = mod._modules["up"]
                dst.has_up = 1
                uw = UnpoolW()
                uw.proj = lin(u.proj[0].weight, u.proj[0].bias, bn=u.proj[1])
                uw.proj_skip = lin(u.proj_skip[0].weight, u.proj_skip[0].bias, bn=u.proj_skip[1])
                uw.cat, uw.alpha = int(u.mode != "add"), float(u.alpha)
                if u.mode != "add":
                    wc = u.proj_cat[0].weight
                    uw.cat_a = lin(wc, None, cols=(0, u.cout))
                    uw.cat_b = lin(wc, u.proj_cat[0].bias, cols=(u.cout, 2 * u.cout))
                dst.up = uw

        def stem(e):
            Bp, shift = e.packed()
            self.keep += [Bp, shift]
            c = e.stem.conv
            return StemW(_f32(Bp), _f32(shift), c.cin, c.cout, c.k, 0)

        def net(prefix, enc, dec, enc_st, dec_st):
            ne = len(enc)
            for s in range(ne):
                stage(enc_st[s], enc[s], s, prefix)
            for j in range(len(dec)):                    # registered (= executed) in reversed stage order
                stage(dec_st[j], dec[j], ne - 2 - j, prefix)
            return ne, len(dec)

        w.condition, w.T_dim = int(model.condition), int(model.T_dim if model.condition else -1)
        w.n_stem = stem(model._n_embedding)
        w.n_enc, w.n_dec = net("n", model._n_enc, model._n_dec, w.n_enc_st, w.n_dec_st)
        w.n_head = lin(model._n_head.weight, model._n_head.bias)
        if model.condition:
            w.c_stem = stem(model._c_embedding)
            w.c_enc, w.c_dec = net("c", model._c_enc, model._c_dec, w.c_enc_st, w.c_dec_st)
            w.c_head = lin(model._c_head.weight, model._c_head.bias)
            if model.T_dim != -1:
                w.fc_t1_W, w.fc_t1_b = _f32(model.fc_t1.weight), _f32(model.fc_t1.bias)
                w.fc_t2_W, w.fc_t2_b = _f32(model.fc_t2.weight), _f32(model.fc_t2.bias)
            cb = model._tm_dec0.cross_block2
            t = w.tm
            a = cb.attn
            t.Cq, t.Ckv, t.H, t.K = a.C, cb.kv_cpe[1].weight.shape[0], a.H, a.K
            t.scale, t.tm_feat, t.ln_eps = float(a.scale), float(cb.tm_feat), float(cb.q_norm1[0].eps)
            t.q_conv_Bp, t.q_conv_b = conv_pack(cb.q_cpe[0])
            t.q_lin, t.q_cpe_ln = lin(cb.q_cpe[1].weight, cb.q_cpe[1].bias), ln(cb.q_cpe[2])
            t.kv_conv_Bp, t.kv_conv_b = conv_pack(cb.kv_cpe[0])
            t.kv_lin, t.kv_cpe_ln = lin(cb.kv_cpe[1].weight, cb.kv_cpe[1].bias), ln(cb.kv_cpe[2])
            t.q_norm1, t.kv_norm1, t.q_norm2 = ln(cb.q_norm1[0]), ln(cb.kv_norm1[0]), ln(cb.q_norm2[0])
            t.q, t.kv, t.proj = lin(a.q.weight, a.q.bias), lin(a.kv.weight, a.kv.bias), lin(a.proj.weight, a.proj.bias)
            t.fc1, t.fc2 = lin(cb.mlp[0].fc1.weight, cb.mlp[0].fc1.bias), lin(cb.mlp[0].fc2.weight, cb.mlp[0].fc2.bias)


def supported(model):
    """the native executor covers the shipped topologies: 16-channel heads everywhere, tensor-core stems, no exotic options"""
    ok = ops.GEMM_MODE == "tc" and ops.NATIVE_BLOCKS and ops.NATIVE_NET
    ok = ok and model._n_embedding.stem.conv.cin <= 8 and len(model._n_enc) <= MAX_STAGES
    if model.condition:
        ok = ok and model._c_embedding.stem.conv.cin <= 8
    return ok


def _signature(model):
    ts = getattr(model, "_net_tensors", None)
    if ts is None:
        ts = list(model.parameters()) + list(model.buffers())
        model._net_tensors = ts
    v = 0
    for t in ts:
        v += t._version
    return (v, ts[0].data_ptr(), ts[-1].data_ptr(), len(ts))


def weights(model):
    sig = _signature(model)
    hit = getattr(model, "_net_weights", None)
    if hit is None or hit[0] != sig:
        model._net_tensors = None
        sig = _signature(model)
        hit = (sig, NetWeights(model))
        model._net_weights = hit
    return hit[1]


_ARENAS = {}


def _arena(key, nbytes, device):
    a = _ARENAS.get(key)
    if a is None or a.numel() < nbytes:
        a = torch.empty(int(nbytes * 1.2) + 4096, dtype=torch.uint8, device=device)
        _ARENAS[key] = a
    return a


def forward_native(model, plan, n_feat, c_feat, t_emb, mode, main, side, nw=None, outs=None):
    """feature phase of one forward through cdseg_net_forward.  n_feat / c_feat: fp32 [N, cin] in the caller's numbering; t_emb: fp32
    [B, T_dim] (one row per scene) or None.  Returns (n_out, c_out) in the caller's numbering."""
    lib = _lib.load()
    nw = nw if nw is not None else weights(model)
    dev = n_feat.device
    N, B = n_feat.shape[0], plan.n_levels[0].B
    a = ForwardArgs()
    a.w = ctypes.pointer(nw.w)
    a.levels = ctypes.cast(plan.desc, ctypes.POINTER(PlanLevel))
    a.n_lv_n, a.n_lv_c = len(plan.n_levels), len(plan.c_levels) if plan.c_levels else 0
    a.N, a.B, a.attn_mode = N, B, ops.ATTN_MODES.index(mode)
    n_out, c_out = outs if outs is not None else (None, None)
    if n_out is None:
        n_out = torch.empty((N, nw.w.n_head.N), dtype=torch.float32, device=dev)
    a.n_feat, a.n_out = _f32(n_feat), n_out.data_ptr()
    if model.condition:
        if c_out is None:
            c_out = torch.empty((N, nw.w.c_head.N), dtype=torch.float32, device=dev)
        a.c_feat, a.c_out = _f32(c_feat), c_out.data_ptr()
        a.t_emb = _f32(t_emb)
    a.stream_main = main.cuda_stream
    two = model.condition and side is not None and side is not main
    a.stream_side = side.cuda_stream if two else None
    mb, sb = ctypes.c_size_t(0), ctypes.c_size_t(0)
    check(lib.cdseg_net_arena_bytes(ctypes.byref(a), ctypes.byref(mb), ctypes.byref(sb)), "net_arena_bytes")
    am = _arena((dev.index, "main"), mb.value, dev)
    a.arena_main, a.arena_main_bytes = am.data_ptr(), am.numel()
    if two:
        as_ = _arena((dev.index, "side"), sb.value, dev)
        a.arena_side, a.arena_side_bytes = as_.data_ptr(), as_.numel()
        plan.arena.record_stream(side)                   # tables read by the side stream
        for t in (c_feat, t_emb, c_out):
            if t is not None:
                t.record_stream(side)
    if main.cuda_stream != torch.cuda.current_stream(dev).cuda_stream:       # the caller forked `main` off its own stream (ptv3.py priority streams)
        plan.arena.record_stream(main)
        for t in (n_feat, c_feat, t_emb, n_out, c_out):
            if t is not None:
                t.record_stream(main)
    evs = None
    if ops.PROFILE is not None:                          # bench.py: CUDA events around the pre / attention / post kernels of every block
        nb = len(nw.blocks)
        evs = (_P * (6 * nb))(*[lib.cdseg_event_create() for _ in range(6 * nb)])
        a.block_events = ctypes.cast(evs, ctypes.POINTER(_P))
        for i, (net, level, b) in enumerate(nw.blocks):
            L = (plan.n_levels if net == "n" else plan.c_levels)[level]
            pm = L.d.pm[b.attn.order_index]
            ops.PROFILE.append(dict(ev=[evs[6 * i + j] for j in range(6)], n=L.n, C=b.attn.channels, H=b.attn.num_heads,
                                    pairs=int(pm.pairs), has_t=b.T_dim != -1))
    check(lib.cdseg_net_forward(ctypes.byref(a)), "net_forward")
    return n_out, c_out
