"""Floating-point side of the oracle: a *functional* CPU/fp32 restatement of the
CDSegNet PTv3 dual-network forward that consumes a reference-compatible
``state_dict`` (same parameter names as the upstream model).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Pure torch + numpy, no spconv /
torch_scatter / flash_attn / addict / timm.

Reference lines restated (ptv3.py = pointcept/models/point_transformer_v3/
point_transformer_v3m1_base.py):
  * Point / serialization / sparsify   pointcept/models/utils/structure.py:39-140
  * PointSequential dispatch           pointcept/models/modules.py:58-83
  * SerializedAttention                ptv3.py:125-296  (dense branch 264-280 = fp32 ground truth;
                                        flash branch 282-289 emulated by ``attn_mode="flash16"``)
  * MLP / Block                        ptv3.py:299-428
  * SerializedPooling / Unpooling      ptv3.py:431-630
  * Embedding                          ptv3.py:633-663
  * SerializedCrossAttention/CrossBlock/TransferModule   ptv3.py:859-1337
  * PointTransformerV3.forward         ptv3.py:1757-1845
  * calc_t_emb                         pointcept/utils/comm.py:21-39
  * DefaultSegmentorV2.inference       pointcept/models/default.py:371-422

Third-party semantics restated ("parity unpinned"):
  * spconv SubMConv3d: out[i] = b + sum_{d in {-r..r}^3} W[:, a, b, c, :] . in[j]
    for active j with grid[j] = grid[i] + (a-r, b-r, c-r) and the same batch id;
    weight layout [C_out, k, k, k, C_in], axes in grid_coord (x, y, z) order.
  * torch_scatter.segment_csr: per CSR segment max / mean.
  * flash_attn varlen: exact softmax attention inside each cu_seqlens segment,
    fp16 inputs/outputs, fp32 accumulation.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from . import serialization_np as S


class OPoint(dict):
    """attribute dict standing in for addict.Dict / Point (structure.py:14-45)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


# ----------------------------------------------------------------------------
# structure
# ----------------------------------------------------------------------------

def make_point(d):
    p = OPoint(d)
    if "batch" not in p and "offset" in p:
        p["batch"] = torch.from_numpy(S.offset2batch(p["offset"].numpy()))
    elif "offset" not in p and "batch" in p:
        p["offset"] = torch.cumsum(torch.bincount(p["batch"]), 0).long()
    return p


def serialize(point, orders, perm_fn=None):
    """structure.py:47-102."""
    g = point["grid_coord"].numpy()
    code, order, inverse, depth = S.serialization(g, point["batch"].numpy(), orders)
    if perm_fn is not None:
        perm = perm_fn(code.shape[0])
        code, order, inverse = code[perm], order[perm], inverse[perm]
    point["serialized_depth"] = depth
    point["serialized_code"] = torch.from_numpy(np.ascontiguousarray(code))
    point["serialized_order"] = torch.from_numpy(np.ascontiguousarray(order))
    point["serialized_inverse"] = torch.from_numpy(np.ascontiguousarray(inverse))


# ----------------------------------------------------------------------------
# third-party op restatements
# ----------------------------------------------------------------------------

def subm_conv3d(feat, batch, grid, weight, bias):
    """spconv.SubMConv3d restated (see module docstring).  weight [Co,k,k,k,Ci]."""
    k = weight.shape[1]
    r = k // 2
    g = grid.numpy().astype(np.int64)
    b = batch.numpy().astype(np.int64)
    lim = int(g.max()) + 2 * r + 2 if len(g) else 1
    def key(bb, c):
        return ((bb * lim + c[:, 0] + r) * lim + c[:, 1] + r) * lim + c[:, 2] + r
    keys = key(b, g)
    srt = np.argsort(keys, kind="stable")
    skeys = keys[srt]
    out = torch.zeros(feat.shape[0], weight.shape[0], dtype=feat.dtype)
    for a in range(k):
        for bb in range(k):
            for c in range(k):
                q = g + np.array([a - r, bb - r, c - r], dtype=np.int64)
                ok = (q >= -r).all(1) & (q < lim - r).all(1)
                qk = key(b, q)
                pos = np.searchsorted(skeys, qk)
                pos = np.minimum(pos, len(skeys) - 1)
                hit = ok & (skeys[pos] == qk)
                if not hit.any():
                    continue
                dst = torch.from_numpy(np.nonzero(hit)[0])
                src = torch.from_numpy(srt[pos[hit]])
                out.index_add_(0, dst, feat[src] @ weight[:, a, bb, c, :].t())
    if bias is not None:
        out = out + bias
    return out


def segment_max(src, cluster, m):
    out = torch.full((m, src.shape[1]), -float("inf"), dtype=src.dtype)
    out.scatter_reduce_(0, cluster[:, None].expand(-1, src.shape[1]), src, reduce="amax")
    return out


def segment_mean(src, cluster, counts):
    out = torch.zeros((len(counts), src.shape[1]), dtype=src.dtype)
    out.index_add_(0, cluster, src)
    return out / counts[:, None].to(src.dtype)


BN_TRAIN = False      # True: BatchNorm normalises with the statistics of the batch (nn.BatchNorm1d in train mode) -- the training oracle


def bn_eval(x, sd, p, eps=1e-3):
    """nn.BatchNorm1d(eps=1e-3, momentum=0.01) (ptv3.py:1435): running statistics in eval mode, batch statistics (biased variance) when
    BN_TRAIN is set (the running-statistics update is not restated: it does not enter the loss or the gradients)."""
    if BN_TRAIN:
        mean, var = x.mean(0), x.var(0, unbiased=False)
        return (x - mean) / torch.sqrt(var + eps) * sd[p + "weight"] + sd[p + "bias"]
    return (x - sd[p + "running_mean"]) / torch.sqrt(sd[p + "running_var"] + eps) * sd[p + "weight"] + sd[p + "bias"]


def ln(x, sd, p):
    return F.layer_norm(x, (x.shape[1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


def lin(x, sd, p):
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


# ----------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------

def _attend(q, k, v, scale, mode):
    """q [B,H,Lq,d], k/v [B,H,Lk,d] -> [B,H,Lq,d]."""
    if mode == "dense":                       # ptv3.py:264-280 in fp32
        attn = (q * scale) @ k.transpose(-2, -1)
        attn = torch.softmax(attn, dim=-1)
        return attn @ v
    if mode == "flash16":                     # ptv3.py:282-289: qkv.half(), fp32 accumulate, fp16 out
        q, k, v = q.half().float(), k.half().float(), v.half().float()
        s = (q @ k.transpose(-2, -1)) * scale
        p = torch.exp(s - s.amax(-1, keepdim=True))
        l = p.sum(-1, keepdim=True)
        o = (p.half().float() @ v) / l
        return o.half().float()
    raise ValueError(mode)


def varlen_attention(q, k, v, cu, H, scale, mode):
    """q,k,v [Npad, C] in padded slot order, cu = patch boundaries -> [Npad, C]."""
    C = q.shape[1]
    d = C // H
    out = torch.empty_like(q)
    cu = [int(c) for c in cu]
    lens = {}
    for i in range(len(cu) - 1):
        lens.setdefault(cu[i + 1] - cu[i], []).append(cu[i])
    for L, starts in lens.items():
        if L == 0:
            continue
        idx = (torch.tensor(starts)[:, None] + torch.arange(L)[None, :]).reshape(-1)
        def shp(t):
            return t[idx].reshape(len(starts), L, H, d).permute(0, 2, 1, 3)
        o = _attend(shp(q), shp(k), shp(v), scale, mode)
        out[idx] = o.permute(0, 2, 1, 3).reshape(-1, C)
    return out


def padding_maps(point, K):
    """ptv3.py:188-244; cached on the point exactly like the reference (keys
    "pad"/"unpad"/"cu_seqlens_key"), so a decoder block re-uses the maps its
    stage's encoder built even if its own patch size differed."""
    if "pad" not in point:
        pad, unpad, cu = S.patch_maps(point["offset"].numpy(), K)
        point["pad"], point["unpad"], point["cu_seqlens_key"] = (
            torch.from_numpy(pad), torch.from_numpy(unpad), torch.from_numpy(cu))
    return point["pad"], point["unpad"], point["cu_seqlens_key"]


def serialized_attention(sd, p, point, H, K, order_index, mode):
    """ptv3.py:246-296."""
    feat = point["feat"]
    C = feat.shape[1]
    scale = (C // H) ** -0.5
    pad, unpad, cu = padding_maps(point, K)
    order = point["serialized_order"][order_index][pad]
    inverse = unpad[point["serialized_inverse"][order_index]]
    qkv = lin(feat, sd, p + "qkv.")[order]
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    o = varlen_attention(q, k, v, cu, H, scale, mode)
    o = o[inverse]
    return lin(o, sd, p + "proj.")


def serialized_cross_attention(sd, p, q_point, kv_point, H, Kq, order_index, mode):
    """ptv3.py:989-1055.  NOTE the reference quirk: kv rows are gathered with
    *q's* pad map (ptv3.py:1008-1010)."""
    C = q_point["feat"].shape[1]
    scale = (C // H) ** -0.5
    pad, unpad, cu = padding_maps(q_point, Kq)
    q_order = q_point["serialized_order"][order_index][pad]
    q_inverse = unpad[q_point["serialized_inverse"][order_index]]
    kv_order = kv_point["serialized_order"][order_index][pad]
    q = lin(q_point["feat"], sd, p + "q.")[q_order]
    kv = lin(kv_point["feat"], sd, p + "kv.")[kv_order]
    k, v = kv[:, :C], kv[:, C:]
    o = varlen_attention(q, k, v, cu, H, scale, mode)
    o = o[q_inverse].float()
    return lin(o, sd, p + "proj.")


# ----------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------

def cpe(sd, p, point):
    """PointSequential(SubMConv3d k3, Linear, LayerNorm) -- ptv3.py:355-365."""
    # The conv reads point.sparse_conv_feat.features, which the reference only re-syncs with
    # point.feat through PointSequential / Block end (modules.py:66-77, ptv3.py:427).  After
    # SerializedUnpooling the two differ (see unpooling()); "conv_in" carries that stale tensor.
    x = subm_conv3d(point.pop("conv_in", point["feat"]), point["batch"], point["grid_coord"],
                    sd[p + "0.weight"], sd.get(p + "0.bias"))
    x = lin(x, sd, p + "1.")
    return ln(x, sd, p + "2.")


def mlp(sd, p, x):
    return lin(F.gelu(lin(x, sd, p + "fc1.")), sd, p + "fc2.")


def block(sd, p, point, H, K, order_index, has_t, mode):
    """ptv3.py:399-428 (pre_norm=True, eval: DropPath = identity)."""
    shortcut = point["feat"]
    point["feat"] = shortcut + cpe(sd, p + "cpe.", point)
    shortcut = point["feat"]
    if has_t and "t_emb" in point:
        point["feat"] = shortcut + lin(point["t_emb"], sd, p + "t_mlp.")
        shortcut = point["feat"]
    point["feat"] = ln(point["feat"], sd, p + "norm1.0.")
    point["feat"] = shortcut + serialized_attention(sd, p + "attn.", point, H, K, order_index, mode)
    shortcut = point["feat"]
    point["feat"] = shortcut + mlp(sd, p + "mlp.0.", ln(point["feat"], sd, p + "norm2.0."))
    return point


def pooling(sd, p, point, stride, has_t, perm_fn):
    """ptv3.py:464-555."""
    plan = S.pool_plan(point["serialized_code"].numpy(), stride, point["serialized_depth"])
    code, order, inverse = plan["code"], plan["order"], plan["inverse"]
    if perm_fn is not None:
        perm = perm_fn(code.shape[0])
        code, order, inverse = code[perm], order[perm], inverse[perm]
    cluster = torch.from_numpy(plan["cluster"])
    counts = torch.from_numpy(plan["counts"])
    head = torch.from_numpy(plan["head_indices"])
    m = len(counts)
    d = OPoint(
        feat=segment_max(lin(point["feat"], sd, p + "proj."), cluster, m),
        coord=segment_mean(point["coord"], cluster, counts),
        grid_coord=point["grid_coord"][head] >> plan["pooling_depth"],
        serialized_code=torch.from_numpy(np.ascontiguousarray(code)),
        serialized_order=torch.from_numpy(np.ascontiguousarray(order)),
        serialized_inverse=torch.from_numpy(np.ascontiguousarray(inverse)),
        serialized_depth=plan["depth"],
        batch=point["batch"][head],
    )
    if has_t:
        d["t_emb"] = point["t_emb"][head]
    d["pooling_inverse"] = cluster
    d["pooling_parent"] = OPoint(point)      # addict copies nested dicts
    d["idx_ptr"] = torch.from_numpy(plan["idx_ptr"])
    new = make_point(d)
    new["feat"] = F.gelu(bn_eval(new["feat"], sd, p + "norm.0."))
    return new


def unpooling(sd, p, point, mode, scale, scale_i):
    """ptv3.py:601-630."""
    parent = point.pop("pooling_parent")
    inverse = point.pop("pooling_inverse")
    up = F.gelu(bn_eval(lin(point["feat"], sd, p + "proj.0."), sd, p + "proj.1."))
    skip = F.gelu(bn_eval(lin(parent["feat"], sd, p + "proj_skip.0."), sd, p + "proj_skip.1."))
    # reference quirk: everything below assigns parent.feat directly (ptv3.py:608-625), so
    # parent.sparse_conv_feat keeps the *unscaled, unfused* proj_skip output and the first
    # decoder block's CPE convolves that tensor, not the fused features.
    parent["conv_in"] = skip
    if scale:                                   # universal_scalling, ptv3.py:34-35
        skip = skip * 2 ** (-0.5)
    if scale_i is not None:                     # exponentially_scalling, ptv3.py:37-38, 610-611
        skip = skip * 0.8 ** (scale_i - 1)      # scale_i=False -> 0.8**-1 = 1.25 (reference quirk)
    if mode == "add":
        parent["feat"] = skip + up[inverse]
    elif mode == "cat":
        parent["feat"] = lin(torch.cat([skip, up[inverse]], dim=-1), sd, p + "proj_cat.0.")
    else:
        raise ValueError(mode)
    return parent


def cross_block(sd, p, q_point, kv_point, H, Kq, mode, tm_feat=1.0):
    """ptv3.py:1179-1223 with pre_norm=True and tm_feat a float."""
    q_short = q_point["feat"]
    q_point["feat"] = q_short + cpe(sd, p + "q_cpe.", q_point)
    q_short = q_point["feat"]
    kv_point["feat"] = kv_point["feat"] + cpe(sd, p + "kv_cpe.", kv_point)
    q_point["feat"] = ln(q_point["feat"], sd, p + "q_norm1.0.")
    kv_point["feat"] = ln(kv_point["feat"], sd, p + "kv_norm1.0.")
    a = serialized_cross_attention(sd, p + "attn.", q_point, kv_point, H, Kq, 0, mode)
    q_point["feat"] = q_short + tm_feat * a
    q_short = q_point["feat"]
    q_point["feat"] = q_short + mlp(sd, p + "mlp.0.", ln(q_point["feat"], sd, p + "q_norm2.0."))
    return q_point


def calc_t_emb(ts, dim):
    """pointcept/utils/comm.py:21-39."""
    half = dim // 2
    f = torch.exp(torch.arange(half) * -(np.log(10000) / (half - 1)))
    e = ts * f
    return torch.cat((torch.sin(e), torch.cos(e)), 1)


def swish(x):
    return x * torch.sigmoid(x)


# ----------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------

DEFAULT_CFG = dict(
    c_in_channels=6, n_in_channels=6, order=("z", "z-trans", "hilbert", "hilbert-trans"),
    c_stride=(4, 4), c_enc_depths=(2, 2, 2), c_enc_channels=(32, 64, 128), c_enc_num_head=(2, 4, 8),
    c_enc_patch_size=(1024, 1024, 1024), c_dec_depths=(2, 2), c_dec_channels=(64, 64),
    c_dec_num_head=(4, 4), c_dec_patch_size=(1024, 1024),
    n_stride=(2, 2, 2, 2), n_enc_depths=(2, 2, 2, 6, 6), n_enc_channels=(32, 64, 128, 256, 512),
    n_enc_num_head=(2, 4, 8, 16, 32), n_enc_patch_size=(1024,) * 5, n_dec_depths=(2, 2, 2, 2),
    n_dec_channels=(64, 64, 128, 256), n_dec_num_head=(4, 4, 8, 16), n_dec_patch_size=(1024,) * 4,
    num_classes=20, T_dim=128, condition=True, skip_connection_mode="cat",
    skip_connection_scale=True, skip_connection_scale_i=False, tm_feat=1.0, shuffle_orders=True,
)


def torch_randperm(k):
    """the reference's own draw: CPU global generator (structure.py:95, ptv3.py:502)."""
    return torch.randperm(k).numpy()


def identity_perm(k):
    return np.arange(k)


def _stage(sd, prefix, s, point, depth, heads, patch, n_orders, has_t, mode, stride, perm_fn):
    if s > 0:
        point = pooling(sd, f"{prefix}enc{s}.down.", point, stride, has_t, perm_fn)
    for i in range(depth):
        point = block(sd, f"{prefix}enc{s}.block{i}.", point, heads, patch, i % n_orders, has_t, mode)
    return point


def _dstage(sd, prefix, s, point, depth, heads, patch, n_orders, has_t, mode, skip_mode, scale, scale_i):
    point = unpooling(sd, f"{prefix}dec{s}.up.", point, skip_mode, scale, scale_i)
    for i in range(depth):
        point = block(sd, f"{prefix}dec{s}.block{i}.", point, heads, patch, i % n_orders, has_t, mode)
    return point


def forward(sd, cfg, c_in=None, n_in=None, attn_mode="dense", perm_fn=None, trace=None):
    with torch.set_grad_enabled(any(torch.is_tensor(v) and v.requires_grad for v in sd.values())):     # autograd only for the training oracle
        return _forward(sd, cfg, c_in, n_in, attn_mode, perm_fn, trace)


def _forward(sd, cfg, c_in=None, n_in=None, attn_mode="dense", perm_fn=None, trace=None):
    """PointTransformerV3.forward (ptv3.py:1757-1845), eval mode.

    sd: state_dict of the *backbone* (keys like ``_n_enc.enc0.block0...``).
    perm_fn(k) -> permutation of range(k) stands in for the CPU ``torch.randperm``
    draws (call order = reference module execution order); None = draw from torch's
    global CPU generator like the reference.  NOTE (reference quirk): every
    SerializedPooling is built with its default ``shuffle_orders=True`` (ptv3.py:1473-1481,
    1624-1633 pass no such kwarg), so pooling shuffles even when the model-level flag
    is False; the model-level flag only gates Point.serialization (ptv3.py:1765,1768).
    trace: optional dict that receives intermediate Points."""
    c = dict(DEFAULT_CFG); c.update(cfg)
    if perm_fn is None:
        perm_fn = torch_randperm
    ser_perm = perm_fn if c["shuffle_orders"] else None
    orders = list(c["order"]); no = len(orders)
    tdim = c["T_dim"]

    def embed(prefix, point):
        x = subm_conv3d(point["feat"], point["batch"], point["grid_coord"], sd[prefix + "stem.conv.weight"], None)
        point["feat"] = F.gelu(bn_eval(x, sd, prefix + "stem.norm."))
        return point

    # CN = code prefix n_ (dominant); NN = code prefix c_ (auxiliary, has t_emb)
    n_skip_mode = "cat" if c["skip_connection_mode"] == "cat_all" else "add"
    c_skip_mode = "add" if c["skip_connection_mode"] == "add" else "cat"
    nd_ch = list(c["n_dec_channels"]) + [c["n_enc_channels"][-1]]

    def n_enc(s, pt):
        return _stage(sd, "_n_enc.", s, pt, c["n_enc_depths"][s], c["n_enc_num_head"][s], c["n_enc_patch_size"][s],
                      no, False, attn_mode, c["n_stride"][s - 1] if s else None, perm_fn)

    def n_dec(s, pt):
        return _dstage(sd, "_n_dec.", s, pt, c["n_dec_depths"][s], c["n_dec_num_head"][s], c["n_dec_patch_size"][s],
                       no, False, attn_mode, n_skip_mode, False,
                       (s + 1) if c["skip_connection_scale_i"] else None)

    def c_enc(s, pt):
        return _stage(sd, "_c_enc.", s, pt, c["c_enc_depths"][s], c["c_enc_num_head"][s], c["c_enc_patch_size"][s],
                      no, tdim != -1, attn_mode, c["c_stride"][s - 1] if s else None, perm_fn)

    def c_dec(s, pt):
        # NN unpool: skip_connection_scale from cfg, skip_connection_scale_i left at its
        # default False (ptv3.py:1666-1674) -> the 1.25 quirk.
        return _dstage(sd, "_c_dec.", s, pt, c["c_dec_depths"][s], c["c_dec_num_head"][s], c["c_dec_patch_size"][s],
                       no, tdim != -1, attn_mode, c_skip_mode, c["skip_connection_scale"], False)

    if not c["condition"]:
        n = make_point(n_in)
        serialize(n, orders, ser_perm)
        n = embed("_n_embedding.", n)
        ns = len(c["n_enc_depths"])
        for s in range(ns):
            n = n_enc(s, n)
            if trace is not None:
                trace[f"n_enc{s}"] = OPoint(n)
        for s in reversed(range(ns - 1)):
            n = n_dec(s, n)
        n["feat"] = lin(n["feat"], sd, "_n_head.")
        return n

    assert len(c["c_enc_depths"]) == 3 and len(c["n_enc_depths"]) == 5, "interleave below is the reference's fixed 3/5-stage schedule"
    cp = make_point(c_in)
    npt = make_point(n_in)
    serialize(cp, orders, ser_perm)
    serialize(npt, orders, ser_perm)
    if tdim != -1 and "t_emb" in cp:
        cp["t_emb"] = swish(lin(swish(lin(cp["t_emb"], sd, "fc_t1.")), sd, "fc_t2."))
    cp = embed("_c_embedding.", cp)
    npt = embed("_n_embedding.", npt)
    cp = c_enc(0, cp); npt = n_enc(0, npt)
    cp = c_enc(1, cp); npt = n_enc(1, npt); npt = n_enc(2, npt)
    cp = c_enc(2, cp); npt = n_enc(3, npt); npt = n_enc(4, npt)
    if trace is not None:
        trace["c_enc2"] = OPoint(cp); trace["n_enc4"] = OPoint(npt)
    npt = cross_block(sd, "_tm_dec0.cross_block2.", npt, cp, c["n_enc_num_head"][-1], c["n_enc_patch_size"][-1],
                      attn_mode, c["tm_feat"])
    if trace is not None:
        trace["n_tm"] = OPoint(npt)
    # decoders are registered in reversed stage order: _c_dec[0] == dec1, _n_dec[0] == dec3
    cp = c_dec(1, cp)
    npt = n_dec(3, npt); npt = n_dec(2, npt)
    cp = c_dec(0, cp)
    npt = n_dec(1, npt); npt = n_dec(0, npt)
    cp["feat"] = lin(cp["feat"], sd, "_c_head.")
    npt["feat"] = lin(npt["feat"], sd, "_n_head.")
    return cp, npt


@torch.no_grad()
def segmentor_inference(sd, cfg, input_dict, noise, T=1000, attn_mode="dense", perm_fn=None):
    """DefaultSegmentorV2.inference(eval=False) (default.py:371-422) with the
    N(0,1) draw of default.py:393 injected as `noise`.  sd keys carry the
    ``backbone.`` prefix."""
    bsd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    base = dict(coord=input_dict["coord"], grid_coord=input_dict["grid_coord"], offset=input_dict["offset"])
    c = dict(DEFAULT_CFG); c.update(cfg)
    if not c["condition"]:
        n = forward(bsd, cfg, n_in=dict(base, feat=input_dict["feat"]), attn_mode=attn_mode, perm_fn=perm_fn)
        return n["feat"]
    N = input_dict["feat"].shape[0]
    ts = (T - 1) * torch.ones((N, 1), dtype=torch.int64)
    c_in = dict(base, feat=noise, t_emb=calc_t_emb(ts, c["T_dim"]))
    n_in = dict(base, feat=input_dict["feat"])
    _, n = forward(bsd, cfg, c_in, n_in, attn_mode=attn_mode, perm_fn=perm_fn)
    return n["feat"]
