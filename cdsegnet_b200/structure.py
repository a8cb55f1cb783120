"""Serialized point structure on the GPU: the B200 counterpart of the reference's
``Point.serialization`` / ``Point.sparsify`` / ``SerializedPooling`` bookkeeping
(pointcept/models/utils/structure.py:39-140; ptv3.py:464-505; ptv3.py:188-244).

Design (see DESIGN.md "Plan phase"): everything that depends only on ``grid_coord`` /
``offset`` -- curve codes, the four sorted orders, the whole pooling hierarchy of BOTH
networks, cluster ids, pooled codes/orders, neighbour tables, tap masks, conv tile plans
and patch slot maps -- is computed up front by ONE C-ABI call (``cdseg_plan_build``,
csrc/plan_exec.cu) that enqueues the sync-free kernels from C++ and synchronises twice
(depth + offsets; pooled level sizes).  The Python objects below are thin views over the
descriptors that call fills in; tensors are materialised lazily, only for callers that
read them (tests, exported ``Point`` fields, the launch-by-launch debug path).

Row bookkeeping: arrays are stored in *physical* curve order (the order of the ``order``
argument); the reference's ``shuffle_orders`` row permutations (CPU ``torch.randperm``,
structure.py:95, ptv3.py:502) only permute ``rowmap`` (logical row -> physical row).
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from ._lib import check

_DT = {torch.int32: 4, torch.int64: 8, torch.uint8: 1, torch.float32: 4}


class Level:
    """One resolution level of one network: a view over a CdsegPlanLevel descriptor."""

    def __init__(self, plan, index, k):
        # no reference to the Plan itself: Plan -> Level -> Plan would be a cycle, and a cycle keeps the (hundreds of MB) arena alive
        # until Python's cyclic collector happens to run -- every forward would then cudaMalloc a fresh arena
        self._mem = plan.mem
        self._i, self._k = index, k
        self.d = plan.desc[index]
        self.parent = None
        self._cache = {}
        self._nbr = {}
        self._pm = {}
        self._pad_K = None       # patch size the reference would have cached its pad maps with

    # ---- host scalars --------------------------------------------------------------------
    n = property(lambda s: int(s.d.n))
    cap = property(lambda s: int(s.d.cap))
    B = property(lambda s: int(s.d.B))
    depth = property(lambda s: int(s.d.depth))
    c0 = property(lambda s: int(s.d.c0))
    pooling_depth = property(lambda s: int(s.d.pooling_depth))
    rowmap = property(lambda s: [int(s.d.rowmap[i]) for i in range(s._k)])

    @property
    def offset_host(self):
        return np.array([self.d.offset_host[b] for b in range(self.d.B)], dtype=np.int64)

    # ---- device arrays (lazy views into the plan arena) ------------------------------------
    def _view(self, key, ptr, shape, dtype):
        t = self._cache.get(key)
        if t is None:
            t = self._mem.view(ptr, shape, dtype) if ptr else None
            self._cache[key] = t
        return t

    grid = property(lambda s: s._view("grid", s.d.grid, (s.cap, 3), torch.int32))
    batch = property(lambda s: s._view("batch", s.d.batch, (s.cap,), torch.int32))
    code = property(lambda s: s._view("code", s.d.code, (s._k, s.cap), torch.int64))
    order = property(lambda s: s._view("order", s.d.order, (s._k, s.cap), torch.int32))
    inverse = property(lambda s: s._view("inverse", s.d.inverse, (s._k, s.cap), torch.int32))
    cluster = property(lambda s: s._view("cluster", s.d.cluster, (s.cap,), torch.int32))
    idx_ptr = property(lambda s: s._view("idx_ptr", s.d.idx_ptr, (s.cap + 1,), torch.int32))
    head = property(lambda s: s._view("head", s.d.head, (s.cap,), torch.int32))
    perm = property(lambda s: s._view("perm", s.d.perm, (s.cap,), torch.int32))
    inv_perm = property(lambda s: s._view("inv_perm", s.d.inv_perm, (s.cap,), torch.int32))

    @property
    def orig(self):
        """level 0 only: (code, order, inverse) in the caller's numbering"""
        if not self.d.o_code:
            return None
        return (self._view("o_code", self.d.o_code, (self._k, self.cap), torch.int64),
                self._view("o_order", self.d.o_order, (self._k, self.cap), torch.int32),
                self._view("o_inverse", self.d.o_inverse, (self._k, self.cap), torch.int32))

    # ---- indice tables / slot maps: prebuilt by the plan call when the model asked for them, else built on demand ----------
    def scene_count(self):
        return np.diff(self.offset_host, prepend=0)

    def nbr(self, ksize):
        if ksize not in self._nbr:
            if ksize == 3 and self.d.nbr3:
                self._nbr[3] = self._mem.view(self.d.nbr3, (self.n, 27), torch.int32)
            elif self.d.nbr_stem and ksize == self.d.stem_ksize:
                self._nbr[ksize] = self._mem.view(self.d.nbr_stem, (self.n, ksize ** 3), torch.int32)
            else:
                self._nbr[ksize] = ops.nbr_build(self.grid[: self.n], self.batch[: self.n], ksize)
        return self._nbr[ksize]

    def tile_mask(self, ksize):
        """per 128-row tile bitmask of the taps that occur (lets the conv GEMM skip absent taps)"""
        key = ("mask", ksize)
        if key not in self._nbr:
            if ksize == 3 and self.d.tile_mask3:
                self._nbr[key] = self._mem.view(self.d.tile_mask3, ((self.n + 127) // 128,), torch.int32)
            else:
                self._nbr[key] = ops.tile_tap_mask(self.nbr(ksize))
        return self._nbr[key]

    def conv_plan(self, ksize):
        """per 128-row tile: distinct neighbour rows + local indices (operand cache plan of the fused pre-attention kernel)"""
        key = ("plan", ksize)
        if key not in self._nbr:
            if ksize == 3 and self.d.conv_plan3:
                nb = int(_lib.load().cdseg_conv_plan_bytes(self.n))
                self._nbr[key] = self._mem.view(self.d.conv_plan3, (nb,), torch.uint8)
            else:
                self._nbr[key] = ops.conv_tile_plan(self.nbr(ksize))
        return self._nbr[key]

    def patch_maps(self, order_index, K):
        """slot maps of logical curve `order_index`.  Like the reference (ptv3.py:191-244 caches
        "pad"/"unpad" on the Point), the FIRST patch size used at a level sticks."""
        if self._pad_K is None:
            self._pad_K = int(self.d.K) if self.d.K > 0 else K
        K = self._pad_K
        prow = self.rowmap[order_index]
        if prow not in self._pm:
            m = self.d.pm[order_index]
            if m.T > 0 and m.K == K and (self.d.pm_mask >> order_index) & 1:
                T, Kp = int(m.T), int(m.Kp)
                v = self._mem.view
                self._pm[prow] = dict(slot_src=v(m.slot_src, (T * Kp,), torch.int32), slot_dst=v(m.slot_dst, (T * Kp,), torch.int32),
                                      point_slot=v(m.point_slot, (self.n,), torch.int32), patch_len=v(m.patch_len, (T,), torch.int32),
                                      T=T, Kp=Kp, K=K, pairs=int(m.pairs))
            else:
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


class _Mem:
    """the plan's device arena + raw-pointer -> tensor-view conversion (shared by the Plan and its Levels)"""

    def __init__(self, arena, keep):
        self.arena, self.keep, self.base = arena, keep, arena.data_ptr()

    def view(self, ptr, shape, dtype):
        nbytes = int(np.prod(shape)) * _DT[dtype]
        off = ptr - self.base
        if off < 0 or off + nbytes > self.arena.numel():
            if ptr == self.keep[1].data_ptr():
                return self.keep[1]
            raise _lib.CdsegError("plan pointer outside the arena")
        return self.arena[off: off + nbytes].view(dtype).view(shape)


class Plan:
    """Serialization + pooling hierarchy of the CN (code prefix n_) and, if `c_strides`
    is given, the NN (code prefix c_).  Mirrors the reference's RNG call order so that a
    seeded run reproduces the reference's shuffles.

    `spec` (optional, from the model): dict(n=[...], c=[...]) with one dict(K=patch size, mask=logical curves whose slot maps are
    needed, conv_plan=bool, stem=kernel size or 0) per level; the tables are then built inside the same native call.  Without it
    (direct users, tests) they are built on first use."""

    def __init__(self, grid_coord, offset, orders, n_strides, c_strides=None, shuffle_orders=True, perm_fn=None,
                 extra_flags=None, spec=None, aux=None):
        perm_fn = perm_fn or torch_randperm
        dev = grid_coord.device
        if dev.type != "cuda":
            raise _lib.CdsegError("cdsegnet_b200 kernels need CUDA tensors (there is no CPU fallback)")
        k = len(orders)
        grid = grid_coord.to(torch.int32).contiguous()
        N = grid.shape[0]
        B = offset.numel()
        offset = offset.to(torch.int64).contiguous()
        self._keep = (grid, offset)
        n_n, n_c = len(n_strides) + 1, (len(c_strides) + 1 if c_strides is not None else 0)
        n_lv = n_n + n_c
        self.desc = (_lib.PlanLevel * n_lv)()
        d = self.desc
        # RNG order of the reference forward (ptv3.py:1761-1794): c.serialization, n.serialization, then the poolings in module
        # execution order (c1, n1, n2, c2, n3, n4 for the shipped 3 / 5 stage schedule).
        ident = list(range(k))
        if c_strides is not None:
            rm_c0 = _draw(k, perm_fn) if shuffle_orders else ident
        rm_n0 = _draw(k, perm_fn) if shuffle_orders else ident
        rm = {("n", 0): rm_n0}
        if c_strides is not None:
            rm[("c", 0)] = rm_c0
            if len(c_strides) == 2 and len(n_strides) == 4:
                sched = [("c", 1), ("n", 1), ("n", 2), ("c", 2), ("n", 3), ("n", 4)]
            else:
                raise NotImplementedError("the reference interleaves a 3-stage Noise Network with a 5-stage Conditional Network "
                                          "(ptv3.py:1785-1794); other depth pairs have no defined shuffle order")
        else:
            sched = [("n", s) for s in range(1, n_n)]
        for net, s in sched:
            perm = _draw(k, perm_fn)                    # SerializedPooling always shuffles (ptv3.py:501-505)
            rm[(net, s)] = [rm[(net, s - 1)][p] for p in perm]
        index = {("n", s): s for s in range(n_n)}
        index.update({("c", s): n_n + s for s in range(n_c)})
        K_all = []
        for (net, s), i in index.items():
            L = d[i]
            L.parent = -1 if s == 0 else index[(net, s - 1)]
            strides = n_strides if net == "n" else c_strides
            L.stride = int(np.ceil(strides[s - 1])) if s > 0 else 0
            for r in range(k):
                L.rowmap[r] = rm[(net, s)][r]
            sp = spec[net][s] if spec is not None else None
            if sp is not None:
                L.K, L.pm_mask, L.want_conv_plan, L.stem_ksize = int(sp["K"]), int(sp["mask"]), int(bool(sp["conv_plan"])), int(sp.get("stem", 0))
                K_all.append(int(sp["K"]))
        lib = _lib.load()
        stem = max([int(d[i].stem_ksize) for i in range(n_lv)] + [0])
        need = lib.cdseg_plan_arena_bytes(N, B, k, n_lv - (2 if n_c else 1), 2 if n_c else 1, stem, min(K_all) if K_all else 1024,
                                          max(K_all) if K_all else 1024)
        # one block from torch's caching allocator (an upper bound: the pooled sizes are only known inside the call); exported
        # Points keep lazy views into it, so it is NOT shared between plans
        self.arena = torch.empty(int(need), dtype=torch.uint8, device=dev)
        self.mem = _Mem(self.arena, self._keep)
        self._base = self.mem.base
        ids = (ctypes.c_int * k)(*[ops.ORDER_IDS[o] for o in orders])
        n_flags = extra_flags.numel() if extra_flags is not None else 0
        fh = (ctypes.c_int32 * max(n_flags, 1))()
        check(lib.cdseg_plan_build(grid.data_ptr(), offset.data_ptr(), N, B, ids, k, d, n_lv,
                                   extra_flags.data_ptr() if n_flags else None, n_flags, fh, self._base, self.arena.numel(),
                                   ops._stream(), aux.cuda_stream if aux is not None else None), "plan_build")
        # aux (a torch.cuda.Stream): the indice tables are built there and only the per-level `ready` events order them (cdseg_net_forward
        # waits on those); any other reader calls join_aux() first
        self.aux = aux
        if aux is not None:
            self.arena.record_stream(aux)
        self.flags = np.array([fh[i] for i in range(n_flags)], dtype=np.int64) if n_flags else None
        self.n_levels = [Level(self, index[("n", s)], k) for s in range(n_n)]
        self.c_levels = [Level(self, index[("c", s)], k) for s in range(n_c)] if n_c else None
        for levels in (self.n_levels, self.c_levels or []):
            for s in range(1, len(levels)):
                levels[s].parent = levels[s - 1]
        if self.c_levels:
            self.n_levels[0]._nbr = self.c_levels[0]._nbr          # same points, same numbering: share lazily built tables too

    def join_aux(self):
        """make the current stream wait for the tables built on the aux stream (no-op without one)"""
        if self.aux is not None:
            check(_lib.load().cdseg_plan_finish(), "plan_finish")      # launches an aux-stream build leaves for later (no-op after a native forward)
            torch.cuda.current_stream().wait_stream(self.aux)
            self.aux = None

    def view(self, ptr, shape, dtype):
        """tensor view of arena memory (or of the caller's offset tensor) at raw device pointer `ptr`"""
        return self.mem.view(ptr, shape, dtype)
