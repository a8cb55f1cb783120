"""GPU parity suite (-m gpu): every C-ABI kernel against the CPU oracle on the same seeded inputs.
Integer / index work is compared bit-exactly; floating point with the tolerance stated in the test."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import serialization_np as S
from oracle import ptv3_oracle as O

pytestmark = pytest.mark.gpu
ORDERS = ("z", "z-trans", "hilbert", "hilbert-trans")
DEV = "cuda"


def cu(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return x.to(dtype) if dtype is not None else x


@pytest.fixture(scope="module")
def ops(lib):
    from cdsegnet_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("depth", (1, 3, 9, 11, 16))
def test_encode_bit_exact(ops, depth):
    z = np.load(os.path.join(GOLDEN, "codes.npz"))
    g, b = cu(z[f"grid_{depth}"], torch.int32), cu(z[f"batch_{depth}"], torch.int32)
    codes = ops.encode_codes(g, b, depth, ORDERS).cpu().numpy()
    for r, o in enumerate(ORDERS):
        assert np.array_equal(codes[r], z[f"code_{depth}_{o}"]), (depth, o)
    assert int(ops.grid_max(g).item()) == int(z[f"grid_{depth}"].max())


@pytest.mark.parametrize("n,bits", [(1, 5), (31, 9), (4096, 27), (4097, 30), (120000, 27), (300000, 40), (70000, 51)])
def test_argsort_bit_exact(ops, n, bits):
    rng = np.random.default_rng(n)
    codes = rng.integers(0, 1 << bits, size=(4, n), dtype=np.int64)
    codes[1] = rng.integers(0, 7, size=n)            # heavy duplicates -> exercises stability
    order, inverse = ops.argsort_rows(cu(codes), bits)
    order, inverse = order.cpu().numpy(), inverse.cpu().numpy()
    ref = np.argsort(codes, axis=1, kind="stable")
    assert np.array_equal(order, ref)
    for r in range(4):
        assert np.array_equal(inverse[r][order[r]], np.arange(n))


def test_argsort_empty_and_zero_bits(ops):
    o, i = ops.argsort_rows(torch.zeros((4, 0), dtype=torch.int64, device=DEV), 10)
    assert o.shape == (4, 0)
    o, i = ops.argsort_rows(torch.zeros((2, 100), dtype=torch.int64, device=DEV), 0)
    assert np.array_equal(o.cpu().numpy()[0], np.arange(100))


def _scene(counts, seed=0):
    from cdsegnet_b200 import synth
    return synth.collate([synth.small_room(c, seed + i) for i, c in enumerate(counts)])


@pytest.mark.timeout(300)
@pytest.mark.parametrize("counts", [(2000,), (1500, 700, 1100), "scannet120k", "scannet3x"])
def test_plan_vs_oracle(ops, counts):
    """serialization + the whole pooling hierarchy (CN strides 2,2,2,2 / NN strides 4,4) bit-exact,
    including the reference's shuffle bookkeeping -- small rooms, the benchmarked 120k-point scene (BASELINE config 2) and a ragged
    batch of three ScanNet-shaped scenes"""
    from cdsegnet_b200.structure import Plan
    from cdsegnet_b200 import synth
    if counts == "scannet120k":
        sc = synth.collate([synth.scannet_scene(120000, 0)])
    elif counts == "scannet3x":
        sc = synth.collate([synth.scannet_scene(n, 3 + i, room_m=(5.0, 4.0, 3.0), n_boxes=5) for i, n in enumerate((40000, 33000, 900))])
    else:
        sc = _scene(counts)
    counts = np.diff(sc["offset"], prepend=0)
    rng = np.random.default_rng(7)
    perms = [rng.permutation(4) for _ in range(8)]
    it = iter(perms)
    plan = Plan(cu(sc["grid_coord"]), cu(sc["offset"]), ORDERS, (2, 2, 2, 2), (4, 4), True, lambda k: next(it))
    batch = S.offset2batch(sc["offset"])
    code, order, inverse, depth = S.serialization(sc["grid_coord"], batch)
    pi = iter(perms)

    def shuffle(c, o, i):
        p = next(pi)
        return c[p], o[p], i[p]
    c0 = shuffle(code, order, inverse)
    n0 = shuffle(code, order, inverse)

    def check(L, c, o, i, d):
        assert L.depth == d and L.n == c.shape[1]
        assert np.array_equal(L.serialized("code").cpu().numpy(), c)
        assert np.array_equal(L.serialized("order").cpu().numpy(), o)
        assert np.array_equal(L.serialized("inverse").cpu().numpy(), i)

    check(plan.c_levels[0], *c0, depth)
    check(plan.n_levels[0], *n0, depth)
    state = {"c": (c0, depth, sc["grid_coord"], batch), "n": (n0, depth, sc["grid_coord"], batch)}
    sched = [("c", 1, 4), ("n", 1, 2), ("n", 2, 2), ("c", 2, 4), ("n", 3, 2), ("n", 4, 2)]
    for net, li, stride in sched:
        (c, o, i), d, g, b = state[net]
        pl = S.pool_plan(c, stride, d)
        cc, oo, ii = shuffle(pl["code"], pl["order"], pl["inverse"])
        L = (plan.c_levels if net == "c" else plan.n_levels)[li]
        check(L, cc, oo, ii, pl["depth"])
        par_n = c.shape[1]
        assert np.array_equal(L.pooling_inverse().cpu().numpy(), pl["cluster"])           # pooling_inverse
        assert np.array_equal(L.idx_ptr[: L.n + 1].cpu().numpy(), pl["idx_ptr"])
        head = L.head[: L.n].cpu().numpy()                                                # any member is a valid head
        members = L.members().cpu().numpy()                                               # internal numbering
        if L.parent.perm is not None:
            members = L.parent.perm.cpu().numpy()[members]
            head = L.parent.perm.cpu().numpy()[head]
        assert np.array_equal(pl["cluster"][members], np.repeat(np.arange(L.n), pl["counts"]))
        assert np.array_equal(pl["cluster"][head], np.arange(L.n))
        g2 = g[pl["head_indices"]] >> pl["pooling_depth"]
        b2 = b[pl["head_indices"]]
        assert np.array_equal(L.grid[: L.n].cpu().numpy(), g2) and np.array_equal(L.batch[: L.n].cpu().numpy(), b2)
        assert np.array_equal(L.offset_host, np.cumsum(np.bincount(b2, minlength=len(counts))))
        state[net] = ((cc, oo, ii), pl["depth"], g2, b2)


@pytest.mark.parametrize("counts,K", [((5, 12), 4), ((10,), 4), ((3, 11), 4), ((1024, 3000, 3001), 1024), ((130,), 128),
                                      ((127, 300), 128)])
def test_patch_maps_vs_reference_padding(ops, counts, K):
    n = sum(counts)
    order = np.random.default_rng(0).permutation(n).astype(np.int32)
    pm = ops.patch_maps(cu(order), np.array(counts), K)
    pad, unpad, cu_seq = S.patch_maps(np.cumsum(counts), K)
    src, dst, ps, plen = (pm[k].cpu().numpy() for k in ("slot_src", "slot_dst", "point_slot", "patch_len"))
    Kp, T = pm["Kp"], pm["T"]
    assert T == len(cu_seq) - 1 and np.array_equal(plen, np.diff(cu_seq))
    # packed slot t*Kp + j  <->  reference padded slot cu[t] + j
    ref_slot = np.concatenate([cu_seq[t] + np.arange(plen[t]) for t in range(T)])
    packed = np.concatenate([t * Kp + np.arange(plen[t]) for t in range(T)])
    assert np.array_equal(src[packed], order[pad[ref_slot]])                # == serialized_order[pad]
    mask = np.ones(T * Kp, bool); mask[packed] = False
    assert (src[mask] == -1).all() and (dst[mask] == -1).all()
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    to_packed = np.empty(int(cu_seq[-1]), np.int64); to_packed[ref_slot] = packed
    assert np.array_equal(ps, to_packed[unpad[inv]])                        # == unpad[serialized_inverse]
    real = dst >= 0
    assert real.sum() == n and np.array_equal(np.sort(dst[real]), np.arange(n))
    assert np.array_equal(ps[dst[real]], np.nonzero(real)[0])


@pytest.mark.parametrize("ks", (3, 5))
def test_nbr_table(ops, ks):
    sc = _scene((1800, 900))
    g, b = sc["grid_coord"], S.offset2batch(sc["offset"])
    nbr = ops.nbr_build(cu(g), cu(b, torch.int32), ks).cpu().numpy()
    key = {(int(bb), *map(int, gg)): i for i, (bb, gg) in enumerate(zip(b, g))}
    r = ks // 2
    rng = np.random.default_rng(0)
    for i in rng.integers(0, len(g), 300):
        for t in range(ks ** 3):
            a, bb, c = t // (ks * ks), (t // ks) % ks, t % ks
            q = (int(b[i]), int(g[i, 0]) + a - r, int(g[i, 1]) + bb - r, int(g[i, 2]) + c - r)
            assert nbr[i, t] == key.get(q, -1)


@pytest.mark.parametrize("ci,co,ks", [(6, 32, 5), (4, 32, 5), (16, 16, 3), (32, 32, 3), (64, 64, 3), (96, 96, 3), (128, 128, 3)])
def test_subm_conv_vs_oracle(ops, ci, co, ks):
    sc = _scene((1500, 600))
    g, b = sc["grid_coord"], S.offset2batch(sc["offset"])
    gen = torch.Generator().manual_seed(ci * 100 + co)
    x = torch.randn(len(g), ci, generator=gen)
    w = torch.randn(co, ks, ks, ks, ci, generator=gen) / (ks ** 3 * ci * 0.4) ** 0.5
    bias = torch.randn(co, generator=gen) if ks == 3 else None
    ref = O.subm_conv3d(x, torch.from_numpy(b), torch.from_numpy(g), w, bias)
    nbr = ops.nbr_build(cu(g), cu(b, torch.int32), ks)
    wt = w.reshape(co, ks ** 3, ci).permute(1, 2, 0).contiguous().to(DEV)
    out = ops.subm_conv(x.to(DEV), nbr, wt, bias.to(DEV) if bias is not None else None, ks).cpu()
    assert (out - ref).abs().max() < 2e-5          # fp32 accumulate, different summation order only


def test_pool_reduce_unpool_rowwise(ops):
    sc = _scene((2000, 1200))
    batch = S.offset2batch(sc["offset"])
    code, order, inverse, depth = S.serialization(sc["grid_coord"], batch)
    pl = S.pool_plan(code, 2, depth)
    m, n, C = len(pl["counts"]), len(batch), 48
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(n, C, generator=gen); coord = torch.from_numpy(sc["coord"])
    scale, shift = torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen)
    cl = torch.from_numpy(pl["cluster"])
    ref = torch.nn.functional.gelu(O.segment_max(x, cl, m) * scale + shift)
    refc = O.segment_mean(coord, cl, torch.from_numpy(pl["counts"]))
    out, outc = ops.pool_reduce(x.to(DEV), coord.to(DEV), cu(pl["indices"], torch.int32), cu(pl["idx_ptr"], torch.int32),
                                m, scale.to(DEV), shift.to(DEV), True)
    assert (out.cpu() - ref).abs().max() < 1e-5 and (outc.cpu() - refc).abs().max() < 1e-5
    up = torch.randn(m, C, generator=gen)
    got = ops.unpool_add(x.to(DEV), up.to(DEV), cu(pl["cluster"], torch.int32), 1.25 * 2 ** -0.5).cpu()
    assert (got - (x * (1.25 * 2 ** -0.5) + up[cl])).abs().max() < 1e-6


@pytest.mark.parametrize("C", (16, 32, 64, 96, 128, 256, 512))
def test_add_layernorm(ops, C):
    gen = torch.Generator().manual_seed(C)
    n, B = 777, 3
    a, b = torch.randn(n, C, generator=gen), torch.randn(n, C, generator=gen)
    tt = torch.randn(B, C, generator=gen); batch = torch.randint(0, B, (n,), generator=gen).sort().values
    g, be = torch.randn(C, generator=gen), torch.randn(C, generator=gen)
    y, ln = ops.add_layernorm(a.to(DEV), b.to(DEV), tt.to(DEV), batch.int().to(DEV), g.to(DEV), be.to(DEV), 1e-5)
    ry = a + b + tt[batch]
    assert (y.cpu() - ry).abs().max() < 1e-6
    assert (ln.cpu() - torch.nn.functional.layer_norm(ry, (C,), g, be, 1e-5)).abs().max() < 2e-5
    _, ln2 = ops.add_layernorm(a.to(DEV), gamma=g.to(DEV), beta=be.to(DEV), want_sum=False)
    assert (ln2.cpu() - torch.nn.functional.layer_norm(a, (C,), g, be, 1e-5)).abs().max() < 2e-5


def test_small_linear_and_scale_shift(ops):
    gen = torch.Generator().manual_seed(5)
    x, W, b = torch.randn(3, 128, generator=gen), torch.randn(512, 128, generator=gen) / 11, torch.randn(512, generator=gen)
    ref = torch.nn.functional.linear(x, W, b); ref = ref * torch.sigmoid(ref)
    assert (ops.small_linear(x.to(DEV), W.to(DEV), b.to(DEV), act=2).cpu() - ref).abs().max() < 1e-5
    y = torch.randn(1001, 64, generator=gen); s, h = torch.randn(64, generator=gen), torch.randn(64, generator=gen)
    got = ops.scale_shift_act(y.to(DEV), s.to(DEV), h.to(DEV), 1).cpu()
    assert (got - torch.nn.functional.gelu(y * s + h)).abs().max() < 1e-5
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    off = torch.tensor([400, 1001], device=DEV)
    rows = torch.cat([y[:1].expand(400, -1), y[1:2].expand(601, -1)]).contiguous().to(DEV)
    ops.rows_uniform_flag(rows, ops.offset2batch(off, 1001), off, flag)
    assert int(flag.item()) == 0
    rows[700, 3] += 1
    ops.rows_uniform_flag(rows, ops.offset2batch(off, 1001), off, flag)
    assert int(flag.item()) == 1


def _attn_case(ops, counts, K, H, mode, seed=0, oracle_mode=None):
    """mode: ops.ATTN_MODES.  The reference is the oracle's dense fp32 formula ("tc32", "exact") or its fp16 emulation of the
    flash branch ("f16") on the same gathered rows (ptv3.py:258-290)."""
    n, C = sum(counts), H * 16
    gen = torch.Generator().manual_seed(seed)
    qkv = torch.randn(n, 3 * C, generator=gen) * 1.5
    order = torch.randperm(n, generator=gen).numpy()
    pad, unpad, cu_seq = S.patch_maps(np.cumsum(counts), K)
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    g = qkv[torch.from_numpy(order[pad])]
    omode = oracle_mode or ("flash16" if mode == "f16" else "dense")
    ref = O.varlen_attention(g[:, :C], g[:, C:2 * C], g[:, 2 * C:], cu_seq, H, 0.25, omode)[torch.from_numpy(unpad[inv])]
    pm = ops.patch_maps(cu(order.astype(np.int32)), np.array(counts), K)
    q, k, v = ops.attn_pack(qkv.to(DEV), 0, C, 3, pm, H, mode)
    out = ops.attn(q, k, v, pm, H, 0.25, n, mode)
    torch.cuda.synchronize()
    return out.cpu(), ref


# shapes: short / ragged / sub-patch scenes, the 2500- and 5000-point multi-patch cases, and the deep levels of the shipped config
# (H = 16 at 3 804 points, H = 32 at one unpadded 991-point sequence: SURVEY.md App. B "stage/patch geometry")
ATTN_SHAPES = [((128,), 128, 1), ((300,), 128, 2), ((1000, 77, 129), 128, 4), ((64,), 64, 1), ((2500,), 1024, 2), ((991,), 1024, 8),
               ((40, 900), 256, 3), ((5000,), 1024, 4), ((3804,), 1024, 16), ((991,), 1024, 32), ((991, 1030), 1024, 32)]


@pytest.mark.parametrize("counts,K,H", ATTN_SHAPES)
def test_attention_exact_vs_dense_oracle(ops, counts, K, H):
    out, ref = _attn_case(ops, counts, K, H, "exact")
    assert (out - ref).abs().max() < 2e-5          # fp32 both sides


@pytest.mark.timeout(120)
@pytest.mark.parametrize("counts,K,H", ATTN_SHAPES)
def test_attention_tc32_vs_dense_oracle(ops, counts, K, H):
    """tcgen05 kernel with hi/lo-split operands and probabilities (cdseg_attn_tc3 mode 1) against the dense fp32 formula
    (ptv3.py:264-280): 22-bit operands + ex2.approx leave ~1e-6 relative; 2e-5 abs on outputs of magnitude ~1"""
    out, ref = _attn_case(ops, counts, K, H, "tc32")
    assert (out - ref).abs().max() < 2e-5


@pytest.mark.timeout(120)
@pytest.mark.parametrize("gen", [1, 2])
@pytest.mark.parametrize("counts,K,H", [((300,), 128, 2), ((1000, 77, 129), 128, 4), ((2500,), 1024, 2), ((5000,), 1024, 4)])
def test_attention_tcgen05_older_kernels_still_correct(ops, counts, K, H, gen):
    """the first- (whole-patch K/V image) and second-generation (K/V ring) kernels stay in the library as A/B comparators"""
    ops.ATTN_KERNEL = gen
    try:
        out, ref = _attn_case(ops, counts, K, H, "f16")
    finally:
        ops.ATTN_KERNEL = 3
    assert (out - ref).abs().max() < 2e-3


@pytest.mark.timeout(120)
@pytest.mark.parametrize("counts,K,H", ATTN_SHAPES)
def test_attention_tcgen05_vs_flash_oracle(ops, counts, K, H):
    """fp16 operands / fp32 accumulate / fp16 probabilities, like flash_attn: tolerance 2e-3 abs on
    outputs of magnitude ~1 (fp16 rounding of P and of the reference's fp16 output)"""
    out, ref = _attn_case(ops, counts, K, H, "f16")
    assert (out - ref).abs().max() < 2e-3


@pytest.mark.timeout(120)
@pytest.mark.parametrize("poly", [1, 3])
def test_attention_fma_pipe_exponentials(ops, lib, poly):
    """1 / 3 of every 8 exponentials evaluated by the FMA-pipe polynomial (cdseg_attn_set_poly): same bar as the MUFU path"""
    lib.cdseg_attn_set_poly(poly)
    try:
        out, ref = _attn_case(ops, (2500,), 1024, 2, "f16")
    finally:
        lib.cdseg_attn_set_poly(0)
    assert (out - ref).abs().max() < 2e-3


@pytest.mark.timeout(120)
@pytest.mark.parametrize("counts,K,H", [((1000, 77, 129), 128, 4), ((2500,), 1024, 2), ((5000,), 1024, 4)])
def test_attention_tcgen05_vs_real_flash_attn(ops, counts, K, H):
    """against the upstream kernel itself (flash_attn is installed on the GPU box): same inputs, same
    varlen partition; both are fp16-in / fp32-accumulate / fp16-out, so they may differ by one fp16 ulp of
    the output (2^-10 relative: 1.95e-3 at |o| in [2,4)) on a few elements and agree elsewhere"""
    fa = pytest.importorskip("flash_attn")
    n, C = sum(counts), H * 16
    gen = torch.Generator().manual_seed(1)
    qkv = torch.randn(n, 3 * C, generator=gen) * 1.5
    order = torch.randperm(n, generator=gen).numpy()
    pad, unpad, cu_seq = S.patch_maps(np.cumsum(counts), K)
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    g = qkv[torch.from_numpy(order[pad])].to(DEV)
    ref = fa.flash_attn_varlen_qkvpacked_func(g.half().reshape(-1, 3, H, 16), cu(cu_seq), max_seqlen=K, dropout_p=0,
                                              softmax_scale=0.25).reshape(-1, C).float()[cu(unpad[inv])]
    pm = ops.patch_maps(cu(order.astype(np.int32)), np.array(counts), K)
    q, k, v = ops.attn_pack(qkv.to(DEV), 0, C, 3, pm, H, "f16")
    out = ops.attn(q, k, v, pm, H, 0.25, n, "f16")
    d = (out - ref).abs()
    assert (d <= 2.0 ** -10 * ref.abs().clamp(min=0.5)).all().item()      # <= 1 fp16 ulp
    assert d.mean().item() < 5e-5


# ------------------------------------------------------------------ tcgen05 3xTF32 GEMM
@pytest.mark.timeout(120)
@pytest.mark.parametrize("M,K,N,act,use_res,T", [(1000, 32, 96, 0, False, 1), (5000, 32, 128, 1, False, 1), (129, 16, 16, 0, True, 1),
                                                  (300, 64, 20, 0, False, 1), (777, 96, 288, 1, True, 1), (816, 2048, 512, 0, True, 32),
                                                  (816, 512, 1536, 0, False, 8), (3023, 256, 1024, 1, False, 1), (1, 64, 6, 0, False, 1),
                                                  # few row tiles + K split requested: served by narrow (32 / 64 column) output tiles, no partial sums
                                                  (990, 512, 512, 1, True, 8), (3800, 256, 256, 0, True, 4), (3800, 1024, 256, 0, True, 16),
                                                  (990, 512, 96, 0, False, 8), (300, 256, 160, 1, True, 4)])
@pytest.mark.parametrize("narrow", (0, 1))
def test_gemm_tc_linear_vs_fp64(ops, lib, M, K, N, act, use_res, T, narrow):
    """fp32-faithful: 3xTF32 split keeps the result within ~1e-5 relative of an fp64 reference (cuBLAS SGEMM class);
    narrow = the narrow-output-tile alternative to split-K (cdseg_gemm_tc_set_narrow)"""
    if narrow and T == 1:
        pytest.skip("the switch only changes launches that ask for a K split")
    lib.cdseg_gemm_tc_set_narrow(narrow)
    gen = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=gen); w = torch.randn(N, K, generator=gen) / K ** 0.5
    b = torch.randn(N, generator=gen); r = torch.randn(M, N, generator=gen) if use_res else None
    ref = x.double() @ w.double().t() + b.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    if use_res:
        ref = ref + r.double()
    Bp = ops.gemm_pack_b(w.t().contiguous()[None].to(DEV))
    tiles = -(-M // 128) * -(-N // 128)
    ns = ops.pick_split(tiles, T) if T > 1 else 1
    try:
        out = ops.gemm_tc(x.to(DEV), Bp, N, K // T, bias=b.to(DEV), res=r.to(DEV) if use_res else None, act=act, nsplit=ns, T=T)
        torch.cuda.synchronize()
    finally:
        lib.cdseg_gemm_tc_set_narrow(0)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 2e-5 * max(1.0, ref.abs().max().item()), err


@pytest.mark.timeout(120)
@pytest.mark.parametrize("M,K,N,act,use_res", [(120000, 32, 64, 0, False), (120000, 64, 20, 0, False), (120001, 64, 6, 0, False),
                                               (52190, 64, 64, 1, True), (16384, 32, 200, 0, False), (20000, 64, 130, 1, True)])
def test_skinny_linear_vs_fp64(ops, M, K, N, act, use_res):
    """the level-0 linears (K = 32 / 64, >= 16 384 rows) take the FP32-pipe streaming kernel inside cdseg_gemm_tc: same contract
    (packed operand blocks in, bias / GELU / residual fused), fp32 products -> within 2e-6 relative of fp64"""
    gen = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=gen); w = torch.randn(N, K, generator=gen) / K ** 0.5
    b = torch.randn(N, generator=gen); r = torch.randn(M, N, generator=gen) if use_res else None
    ref = x.double() @ w.double().t() + b.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    if use_res:
        ref = ref + r.double()
    Bp = ops.gemm_pack_b(w.t().contiguous()[None].to(DEV))
    out = ops.gemm_tc(x.to(DEV), Bp, N, K, bias=b.to(DEV), res=r.to(DEV) if use_res else None, act=act)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 5e-6 * max(1.0, ref.abs().max().item()), err


@pytest.mark.timeout(120)
@pytest.mark.parametrize("c,ns", [(16, 1), (32, 1), (64, 3), (96, 1), (128, 9), (256, 27)])
def test_gemm_tc_conv_vs_oracle(ops, c, ns):
    sc = _scene((1500, 600))
    g, b = sc["grid_coord"], S.offset2batch(sc["offset"])
    gen = torch.Generator().manual_seed(c)
    x = torch.randn(len(g), c, generator=gen)
    w = torch.randn(c, 3, 3, 3, c, generator=gen) / (27 * c * 0.4) ** 0.5
    bias = torch.randn(c, generator=gen)
    ref = O.subm_conv3d(x, torch.from_numpy(b), torch.from_numpy(g), w, bias)
    nbr = ops.nbr_build(cu(g), cu(b, torch.int32), 3)
    mask = ops.tile_tap_mask(nbr)
    nb = nbr.cpu().numpy()
    exp_mask = np.array([np.bitwise_or.reduce((nb[i:i + 128] >= 0).astype(np.int64) << np.arange(27), axis=None)
                         for i in range(0, len(nb), 128)])
    assert np.array_equal(mask.cpu().numpy().astype(np.int64) & 0x7ffffff, exp_mask)
    Bp = ops.gemm_pack_b(w.reshape(c, 27, c).permute(1, 2, 0).contiguous().to(DEV))
    out = ops.gemm_tc(x.to(DEV), Bp, c, c, idx=nbr, tile_mask=mask, bias=bias.to(DEV), nsplit=ns)
    torch.cuda.synchronize()
    assert (out.cpu() - ref).abs().max() < 1e-4      # 3xTF32: ~2^-21 relative per product, |out| up to ~5


@pytest.mark.timeout(120)
@pytest.mark.parametrize("ci,co", [(6, 32), (4, 32), (6, 48)])
def test_stem_conv_im2col_tc_vs_oracle(ops, ci, co):
    """Embedding stem (k=5, tiny C_in) through the tensor-core im2col GEMM == oracle conv + folded BN + exact GELU"""
    sc = _scene((1500, 700))
    g, b = sc["grid_coord"], S.offset2batch(sc["offset"])
    gen = torch.Generator().manual_seed(ci * 7 + co)
    x = torch.randn(len(g), ci, generator=gen)
    w = torch.randn(co, 5, 5, 5, ci, generator=gen) / (125 * ci * 0.4) ** 0.5
    scale, shift = torch.rand(co, generator=gen) + 0.5, torch.randn(co, generator=gen)
    ref = torch.nn.functional.gelu(O.subm_conv3d(x, torch.from_numpy(b), torch.from_numpy(g), w, None) * scale + shift)
    nbr = ops.nbr_build(cu(g), cu(b, torch.int32), 5)
    wt = w.reshape(co, 125, ci).permute(1, 2, 0) * scale                      # [125, ci, co]
    wp = torch.zeros(128, 8, co)
    wp[:125, :ci] = wt
    Bp = ops.gemm_pack_b(wp.reshape(32, 32, co).contiguous().to(DEV))
    x8 = torch.nn.functional.pad(x, (0, 8 - ci)).to(DEV)
    out = ops.conv_im2col_tc(x8, nbr, Bp, co, shift.to(DEV), 1)
    torch.cuda.synchronize()
    assert (out.cpu() - ref).abs().max() < 1e-4


# ------------------------------------------------------------------ fused row-tile kernels (chained GEMMs in tensor memory)
def _lin(gen, cin, cout):
    return torch.randn(cout, cin, generator=gen) / cin ** 0.5, 0.3 * torch.randn(cout, generator=gen)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("n,C", [(1000, 32), (128, 32), (1, 32), (5000, 64), (777, 128), (40000, 32), (33000, 64), (20000, 128)])
def test_post_attn_chain_vs_fp64(ops, n, C):
    """proj + residual + LayerNorm + fc1 + GELU + fc2 + residual in ONE kernel == the fp64 composition (ptv3.py:290-296, 416-424)"""
    gen = torch.Generator().manual_seed(n + C)
    o, x1 = torch.randn(n, C, generator=gen), 2.0 * torch.randn(n, C, generator=gen) + 0.5
    wp, bp = _lin(gen, C, C); w1, b1 = _lin(gen, C, 4 * C); w2, b2 = _lin(gen, 4 * C, C)
    g, be = torch.rand(C, generator=gen) + 0.5, 0.2 * torch.randn(C, generator=gen)
    d = lambda a: a.double()
    x2 = d(x1) + d(o) @ d(wp).t() + d(bp)
    h = torch.nn.functional.layer_norm(x2, (C,), d(g), d(be), 1e-5)
    ref = x2 + torch.nn.functional.gelu(h @ d(w1).t() + d(b1)) @ d(w2).t() + d(b2)
    pk = lambda w, b: (ops.gemm_pack_b(w.t().contiguous()[None].to(DEV)), b.to(DEV))
    out = ops.post_attn(o.to(DEV), x1.to(DEV), pk(wp, bp), (g.to(DEV), be.to(DEV)), pk(w1, b1), pk(w2, b2), 1e-5)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 3e-5 * max(1.0, ref.abs().max().item()), err


def _curve_sorted(sc):
    """the scene's points renumbered along the z curve (what structure.Plan does to every level): neighbours of a 128-row tile
    are then few distinct rows"""
    g, b = sc["grid_coord"], S.offset2batch(sc["offset"])
    code, order, _, _ = S.serialization(g, b)
    return np.ascontiguousarray(g[order[0]]), np.ascontiguousarray(b[order[0]])


@pytest.mark.parametrize("counts,curve", [((1500, 600), True), ((1500, 600), False), ((130,), True), ((9000,), True)])
def test_conv_tile_plan_bit_exact(ops, counts, curve):
    """per 128-row tile: ascending distinct neighbour rows + local index of every (row, tap) (include/cdseg_b200.h)"""
    sc = _scene(counts)
    g, b = _curve_sorted(sc) if curve else (sc["grid_coord"], S.offset2batch(sc["offset"]))
    nbr = ops.nbr_build(cu(g), cu(b, torch.int32), 3)
    plan = ops.conv_tile_plan(nbr).cpu().numpy()
    nb = nbr.cpu().numpy()
    UCAP = ops.CONV_PLAN_UCAP
    rec = 16 + 4 * UCAP + 128 * 27 * 2
    assert plan.size == -(-len(nb) // 128) * rec
    cached = 0
    for t in range(-(-len(nb) // 128)):
        r = plan[t * rec:(t + 1) * rec]
        rows = nb[t * 128:(t + 1) * 128]
        uniq = np.unique(rows[rows >= 0])
        assert int(r[:4].view(np.int32)[0]) == len(uniq)
        if len(uniq) > UCAP:
            continue
        cached += 1
        u = r[16:16 + 4 * UCAP].view(np.int32)
        assert np.array_equal(u[:len(uniq)], uniq) and np.all(u[len(uniq):] == -1)
        li = r[16 + 4 * UCAP:].view(np.int16).reshape(128, 27)[:len(rows)]
        exp = np.where(rows >= 0, np.searchsorted(uniq, np.maximum(rows, 0)), -1)
        assert np.array_equal(li, exp)
    assert cached > 0 or not curve


@pytest.mark.timeout(180)
@pytest.mark.parametrize("curve", (True, False))
@pytest.mark.parametrize("counts,C,with_t,stale", [((1500, 600), 32, False, False), ((1500, 600), 32, True, True), ((3000,), 64, True, False),
                                                    ((900, 500), 128, False, True), ((130,), 64, False, False)])
def test_pre_attn_chain_vs_fp64(ops, counts, C, with_t, stale, curve):
    """cpe conv + Linear + LayerNorm + residual (+ per-scene t) + norm1 + qkv in ONE kernel == the fp64 composition
    (ptv3.py:355-362, 400-413, 258); `stale` = the conv reads another tensor than the residual (unpooling quirk);
    `curve` = rows numbered along the z curve (tiles use the shared-memory neighbour cache) or in generation order
    (neighbourhoods too scattered for the cache: the direct-gather fallback)"""
    sc = _scene(counts)
    g, b = _curve_sorted(sc) if curve else (sc["grid_coord"], S.offset2batch(sc["offset"]))
    n = len(g)
    gen = torch.Generator().manual_seed(C + n)
    x = torch.randn(n, C, generator=gen)
    cin = torch.randn(n, C, generator=gen) if stale else x
    wc = torch.randn(C, 3, 3, 3, C, generator=gen) / (27 * C * 0.4) ** 0.5
    bc = 0.3 * torch.randn(C, generator=gen)
    wl, bl = _lin(gen, C, C); wq, bq = _lin(gen, C, 3 * C)
    cg, cb = torch.rand(C, generator=gen) + 0.5, 0.2 * torch.randn(C, generator=gen)
    g1, b1 = torch.rand(C, generator=gen) + 0.5, 0.2 * torch.randn(C, generator=gen)
    tproj = torch.randn(len(counts), C, generator=gen) if with_t else None
    d = lambda a: a.double()
    y = O.subm_conv3d(d(cin), torch.from_numpy(b), torch.from_numpy(g), d(wc), d(bc))
    y = torch.nn.functional.layer_norm(y @ d(wl).t() + d(bl), (C,), d(cg), d(cb), 1e-5)
    x1 = d(x) + y + (d(tproj)[torch.from_numpy(b).long()] if with_t else 0)
    qkv = torch.nn.functional.layer_norm(x1, (C,), d(g1), d(b1), 1e-5) @ d(wq).t() + d(bq)
    nbr = ops.nbr_build(cu(g), cu(b, torch.int32), 3)
    mask = ops.tile_tap_mask(nbr)
    pk = lambda w, bb: (ops.gemm_pack_b(w.t().contiguous()[None].to(DEV)), bb.to(DEV))
    conv = (ops.gemm_pack_b(wc.reshape(C, 27, C).permute(1, 2, 0).contiguous().to(DEV)), bc.to(DEV))
    o1, oq = ops.pre_attn(cin.to(DEV), x.to(DEV), nbr, mask, ops.conv_tile_plan(nbr), conv, pk(wl, bl), (cg.to(DEV), cb.to(DEV)), (g1.to(DEV), b1.to(DEV)),
                          pk(wq, bq), tproj.to(DEV) if with_t else None, cu(b, torch.int32) if with_t else None, 1e-5)
    torch.cuda.synchronize()
    e1 = (o1.cpu().double() - x1).abs().max().item()
    eq = (oq.cpu().double() - qkv).abs().max().item()
    assert e1 < 3e-5 * max(1.0, x1.abs().max().item()), e1
    assert eq < 5e-5 * max(1.0, qkv.abs().max().item()), eq


@pytest.mark.parametrize("n,C,ns,with_ln1,with_t", [(990, 512, 8, True, True), (3800, 256, 4, False, False), (130, 96, 3, True, False), (7, 16, 1, True, True)])
def test_reduce_ln_vs_fp64(ops, n, C, ns, with_ln1, with_t):
    """split-K partial sums + bias -> [LayerNorm] -> + residual (+ per-scene t) -> LayerNorm in one launch == the fp64 composition"""
    gen = torch.Generator().manual_seed(n + C)
    part = torch.randn(ns, n, C, generator=gen)
    bias, res = torch.randn(C, generator=gen), torch.randn(n, C, generator=gen)
    g1, b1, g2, b2 = (torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen), torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen))
    B = 3
    t = torch.randn(B, C, generator=gen) if with_t else None
    batch = torch.sort(torch.randint(0, B, (n,), generator=gen)).values.int()
    d = lambda a: a.double()
    v = d(part).sum(0) + d(bias)
    if with_ln1:
        v = torch.nn.functional.layer_norm(v, (C,), d(g1), d(b1), 1e-5)
    y = d(res) + v + (d(t)[batch.long()] if with_t else 0)
    ln = torch.nn.functional.layer_norm(y, (C,), d(g2), d(b2), 1e-5)
    dev = lambda a: a.to(DEV) if a is not None else None
    gy, gl = ops.reduce_ln(dev(part), ns, dev(bias), (dev(g1), dev(b1)) if with_ln1 else None, dev(res), dev(t), dev(batch) if with_t else None,
                           (dev(g2), dev(b2)))
    assert (gy.cpu().double() - y).abs().max() < 2e-5 * max(1.0, y.abs().max().item())
    assert (gl.cpu().double() - ln).abs().max() < 5e-5 * max(1.0, ln.abs().max().item())
