"""Serialized point structure on the GPU: the B200 counterpart of the reference's
``Point.serialization`` / ``Point.sparsify`` / ``SerializedPooling`` bookkeeping
(pointcept/models/utils/structure.py:39-140; ptv3.py:464-505; ptv3.py:188-244).

Design (see DESIGN.md "Plan phase"): everything that depends only on ``grid_coord`` /
``offset`` -- curve codes, the four sorted orders, the whole pooling hierarchy of BOTH
networks, cluster ids, pooled codes/orders -- is computed up front by sync-free kernels
(pooled counts stay in device memory), followed by ONE device->host copy of the level
sizes.  Neighbour tables and patch slot maps are then built lazily per level and cached
(the analogue of spconv's ``indice_key`` and of the reference's cached "pad"/"unpad").

Row bookkeeping: arrays are stored in *physical* curve order (the order of the ``order``
argument); the reference's ``shuffle_orders`` row permutations (CPU ``torch.randperm``,
structure.py:95, ptv3.py:502) only permute ``rowmap`` (logical row -> physical row).
"""
import numpy as np
import torch

from . import ops


class Level:
    """One resolution level of one network."""

    def __init__(self):
        self.n = None            # number of points (host int, known after finalize())
        self.cap = None          # allocated rows
        self.B = None
        self.grid = None         # int32 [cap,3]
        self.batch = None        # int32 [cap]
        self.offset_host = None  # np.int64 [B] cumulative
        self.offset_dev = None   # int64 [B]
        self.code = None         # int64 [k, cap]  physical rows
        self.order = None        # int32 [k, cap]
        self.inverse = None      # int32 [k, cap]
        self.depth = None
        self.rowmap = None       # list: logical row -> physical row
        self.parent = None
        self.cluster = None      # int32 [parent.cap]   parent point -> this level's point (pooling_inverse)
        self.idx_ptr = None      # int32 [cap+1]
        self.head = None         # int32 [cap]
        self.c0 = None           # physical row of the parent that defined the clusters
        self.pooling_depth = None
        self.m_dev = None
        self.perm = None         # level 0 only: internal id r <-> original point perm[r] (see Plan)
        self.inv_perm = None
        self.orig = None         # level 0 only: (code, order, inverse) in the caller's numbering
        self._nbr = {}
        self._pm = {}
        self._pad_K = None       # patch size the reference would have cached its pad maps with

    # ---- lazily built, cached structures --------------------------------------------
    def scene_count(self):
        return np.diff(self.offset_host, prepend=0)

    def nbr(self, ksize):
        if ksize not in self._nbr:
            self._nbr[ksize] = ops.nbr_build(self.grid[: self.n], self.batch[: self.n], ksize)
        return self._nbr[ksize]

    def tile_mask(self, ksize):
        """per 128-row tile bitmask of the taps that occur (lets the conv GEMM skip absent taps)"""
        key = ("mask", ksize)
        if key not in self._nbr:
            self._nbr[key] = ops.tile_tap_mask(self.nbr(ksize))
        return self._nbr[key]

    def conv_plan(self, ksize):
        """per 128-row tile: distinct neighbour rows + local indices (operand cache plan of the fused pre-attention kernel)"""
        key = ("plan", ksize)
        if key not in self._nbr:
            self._nbr[key] = ops.conv_tile_plan(self.nbr(ksize))
        return self._nbr[key]

    def patch_maps(self, order_index, K):
        """slot maps of logical curve `order_index`.  Like the reference (ptv3.py:191-244 caches
        "pad"/"unpad" on the Point), the FIRST patch size used at a level sticks."""
        if self._pad_K is None:
            self._pad_K = K
        K = self._pad_K
        prow = self.rowmap[order_index]
        if prow not in self._pm:
            self._pm[prow] = ops.patch_maps(self.order[prow][: self.n], self.scene_count(), K)
        return self._pm[prow]

    def members(self):
        """parent points sorted by cluster (reference `indices`, ptv3.py:487) -- the parent's c0 order."""
        return self.parent.order[self.c0][: self.parent.n]

    # ---- reference-shaped views (int64, logical row order) for API parity / tests ------
    def serialized(self, what):
        """reference-shaped [k, n] int64 view in the CALLER's point numbering and logical row order"""
        if self.orig is not None:
            t = self.orig[{"code": 0, "order": 1, "inverse": 2}[what]]
        else:
            t = {"code": self.code, "order": self.order, "inverse": self.inverse}[what]
        return torch.stack([t[r][: self.n] for r in self.rowmap]).long()

    def pooling_inverse(self):
        """parent point -> pooled point (reference `cluster` / "pooling_inverse"), caller's numbering"""
        c = self.cluster[: self.parent.n]
        return c[self.parent.inv_perm.long()] if self.parent.inv_perm is not None else c

    def reference_pad_maps(self, K):
        """(pad, unpad, cu_seqlens) exactly as ptv3.py:188-244 would build them (host side, tiny)."""
        cnt = self.scene_count()
        padded = np.where(cnt > K, (cnt + K - 1) // K * K, cnt)
        start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        pstart = np.concatenate([[0], np.cumsum(padded)[:-1]])
        pad, unpad, cu = [], [], []
        for b in range(len(cnt)):
            j = np.arange(padded[b])
            pad.append(start[b] + np.where(j < cnt[b], j, j - K))
            unpad.append(pstart[b] + np.arange(cnt[b]))
            cu.append(np.arange(pstart[b], pstart[b] + padded[b], K))
        cu.append(np.array([padded.sum()]))
        return (np.concatenate(pad).astype(np.int64), np.concatenate(unpad).astype(np.int64),
                np.concatenate(cu).astype(np.int32))


def _draw(k, perm_fn):
    p = perm_fn(k)
    return [int(v) for v in (p.tolist() if hasattr(p, "tolist") else p)]


def torch_randperm(k):
    """the reference's own draw: CPU global generator (structure.py:95, ptv3.py:502)."""
    return torch.randperm(k)


class Plan:
    """Serialization + pooling hierarchy of the CN (code prefix n_) and, if `c_strides`
    is given, the NN (code prefix c_).  Mirrors the reference's RNG call order so that a
    seeded run reproduces the reference's shuffles."""

    def __init__(self, grid_coord, offset, orders, n_strides, c_strides=None, shuffle_orders=True, perm_fn=None,
                 extra_flags=None):
        perm_fn = perm_fn or torch_randperm
        dev = grid_coord.device
        k = len(orders)
        grid = grid_coord.to(torch.int32).contiguous()
        N = grid.shape[0]
        B = offset.numel()
        offset = offset.to(torch.int64).contiguous()
        # --- sync #1: depth + offsets (the reference syncs here too: structure.py:66) ---------
        gm = ops.grid_max(grid)
        head = torch.cat([gm.to(torch.int64), offset]).cpu().numpy()
        depth = int(head[0]).bit_length()
        if depth > 16:
            raise ValueError("serialized depth > 16 (structure.py:74)")
        offset_host = head[1:].astype(np.int64)
        if int(offset_host[-1]) != N:
            raise ValueError("offset[-1] != number of points")
        nbits = 3 * depth + max(0, int(B - 1).bit_length())
        batch = ops.offset2batch(offset, N)
        code = ops.encode_codes(grid, batch, depth, orders)
        order, inverse = ops.argsort_rows(code, nbits)
        # Internal numbering = rank along the first curve: point r of the working set is the caller's
        # point perm[r].  Every level-0 kernel (conv tiles, patch gathers, LayerNorm rows) then walks
        # memory in space-filling-curve order; inputs are gathered once and the logits scattered back.
        perm, inv_perm = order[0], inverse[0]
        i_grid, i_batch, i_code, i_order, i_inverse = ops.renumber(perm, inv_perm, grid, batch, code, order, inverse)

        def level0():
            L = Level()
            L.n = L.cap = N; L.B = B; L.grid = i_grid; L.batch = i_batch
            L.offset_host = offset_host; L.offset_dev = offset
            L.code, L.order, L.inverse, L.depth = i_code, i_order, i_inverse, depth
            L.perm, L.inv_perm, L.orig = perm, inv_perm, (code, order, inverse)
            return L

        # RNG order of the reference forward (ptv3.py:1761-1794): c.serialization, n.serialization,
        # then the poolings in module execution order.
        self.c_levels = None
        if c_strides is not None:
            c0 = level0()
            c0.rowmap = _draw(k, perm_fn) if shuffle_orders else list(range(k))
            self.c_levels = [c0]
        n0 = level0()
        if c_strides is not None:
            n0._nbr = c0._nbr          # same points, same numbering: share the neighbour tables / tap masks

        n0.rowmap = _draw(k, perm_fn) if shuffle_orders else list(range(k))
        self.n_levels = [n0]

        n_pool = len(n_strides) + (len(c_strides) if c_strides is not None else 0)
        n_flag = extra_flags.numel() if extra_flags is not None else 0
        # every pooled level's point count (int32) and per-scene offsets (int64) land in two small buffers: one zero fill and
        # one device-to-host copy each per forward instead of a fill / cast / cat per level
        cnt_buf = torch.zeros(max(n_pool, 1), dtype=torch.int32, device=dev)
        off_buf = torch.zeros((max(n_pool, 1), B), dtype=torch.int64, device=dev)
        slot = [0]

        def pool(levels, stride):
            par = levels[-1]
            pd = (int(np.ceil(stride)) - 1).bit_length()
            if pd > par.depth:
                pd = 0
            ch = Level()
            ch.parent, ch.c0, ch.pooling_depth = par, par.rowmap[0], pd
            ch.cap, ch.B, ch.depth = par.cap, B, par.depth - pd
            i = slot[0]
            slot[0] += 1
            out = ops.pool_plan(par.code, par.order, par.m_dev, par.n if par.m_dev is None else 0, ch.c0, pd, par.grid,
                                par.batch, B, ch.cap, cnt_buf[i:i + 1], off_buf[i])
            ch.slot = i
            ch.cluster, ch.idx_ptr, ch.head = out["cluster"], out["idx_ptr"], out["head"]
            ch.code, ch.order, ch.inverse = out["code"], out["order"], out["inverse"]
            ch.grid, ch.batch, ch.m_dev, ch.offset_dev = out["grid"], out["batch"], out["m_dev"], out["offset"]
            perm = _draw(k, perm_fn)                    # SerializedPooling always shuffles (ptv3.py:501-505)
            ch.rowmap = [par.rowmap[p] for p in perm]
            levels.append(ch)

        if c_strides is not None:
            assert len(c_strides) == 2 and len(n_strides) == 4, "reference schedule is 3 NN / 5 CN stages"
            pool(self.c_levels, c_strides[0]); pool(self.n_levels, n_strides[0]); pool(self.n_levels, n_strides[1])
            pool(self.c_levels, c_strides[1]); pool(self.n_levels, n_strides[2]); pool(self.n_levels, n_strides[3])
        else:
            for s in n_strides:
                pool(self.n_levels, s)

        # --- sync #2: pooled sizes + offsets (+ caller flags) ------------------------------
        pooled = [L for L in (self.c_levels or [])[1:] + self.n_levels[1:]]
        self.flags = None
        if pooled or extra_flags is not None:
            cnt_host = cnt_buf.cpu().numpy()
            off_host = off_buf.cpu().numpy()
            for L in pooled:
                L.n = int(cnt_host[L.slot])
                L.offset_host = off_host[L.slot].astype(np.int64)
            if extra_flags is not None:
                self.flags = extra_flags.cpu().numpy().astype(np.int64).reshape(-1)
