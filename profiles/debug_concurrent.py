"""Which kernel pair misbehaves when co-resident?  Each candidate kernel runs repeatedly on stream A while another runs on stream B;
outputs are compared bitwise with the solo result."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops, synth
from oracle import serialization_np as S
dev = "cuda"
torch.manual_seed(0)
sc = synth.collate([synth.scannet_scene(120000, 0)])
g = torch.from_numpy(sc["grid_coord"]).to(dev)
off = torch.from_numpy(sc["offset"]).to(dev)
from cdsegnet_b200.structure import Plan
plan = Plan(g, off, ("z", "z-trans", "hilbert", "hilbert-trans"), (2, 2, 2, 2), None, False)
L = plan.n_levels[0]
n = L.n
nbr, mask, cplan = L.nbr(3), L.tile_mask(3), L.conv_plan(3)

def make_pre(C):
    gen = torch.Generator().manual_seed(C)
    x = torch.randn(n, C, generator=gen).to(dev)
    wc = (torch.randn(27, C, C, generator=gen) / (27 * C * 0.4) ** 0.5).to(dev)
    lin = lambda a, b: (ops.gemm_pack_b((torch.randn(b, a, generator=gen) / a ** 0.5).t().contiguous()[None].to(dev)), torch.randn(b, generator=gen).to(dev))
    conv = (ops.gemm_pack_b(wc), torch.randn(C, generator=gen).to(dev))
    l1, lq = lin(C, C), lin(C, 3 * C)
    ln = lambda: (torch.rand(C, generator=gen).to(dev) + 0.5, torch.randn(C, generator=gen).to(dev))
    a, b = ln(), ln()
    return lambda: ops.pre_attn(x, x, nbr, mask, cplan, conv, l1, a, b, lq)

def make_post(C):
    gen = torch.Generator().manual_seed(C + 1)
    o, x1 = torch.randn(n, C, generator=gen).to(dev), torch.randn(n, C, generator=gen).to(dev)
    lin = lambda a, b: (ops.gemm_pack_b((torch.randn(b, a, generator=gen) / a ** 0.5).t().contiguous()[None].to(dev)), torch.randn(b, generator=gen).to(dev))
    pr, f1, f2 = lin(C, C), lin(C, 4 * C), lin(4 * C, C)
    ln = (torch.rand(C, generator=gen).to(dev) + 0.5, torch.randn(C, generator=gen).to(dev))
    return lambda: (ops.post_attn(o, x1, pr, ln, f1, f2),)

def make_attn(H, mode):
    C = 16 * H
    gen = torch.Generator().manual_seed(H)
    qkv = torch.randn(n, 3 * C, generator=gen).to(dev)
    pm = L.patch_maps(0, 1024)
    q, k, v = ops.attn_pack(qkv, 0, C, 3, pm, H, mode)
    return lambda: (ops.attn(q, k, v, pm, H, 0.25, n, mode),)

def make_gemm(K, N):
    gen = torch.Generator().manual_seed(K + N)
    x = torch.randn(n, K, generator=gen).to(dev)
    Bp = ops.gemm_pack_b((torch.randn(N, K, generator=gen) / K ** 0.5).t().contiguous()[None].to(dev))
    return lambda: (ops.gemm_tc(x, Bp, N, K),)

cands = {"pre32": make_pre(32), "pre64": make_pre(64), "post32": make_post(32), "attn_tc32_H2": make_attn(2, "tc32"), "attn_f16_H2": make_attn(2, "f16"),
         "attn_tc32_H4": make_attn(4, "tc32"), "gemm32x64": make_gemm(32, 64)}
solo = {}
for k, f in cands.items():
    outs = f(); torch.cuda.synchronize()
    solo[k] = [o.clone() for o in outs]
    outs2 = f(); torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(solo[k], outs2)), ("not deterministic alone", k)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
names = list(cands)
for va in names:
    for vb in names:
        bad_a = bad_b = 0
        for rep in range(6):
            torch.cuda.synchronize()
            with ops.stream_scope(sa):
                oa = [cands[va]() for _ in range(3)]
            with ops.stream_scope(sb):
                ob = [cands[vb]() for _ in range(3)]
            torch.cuda.synchronize()
            bad_a += sum(not all(torch.equal(x, y) for x, y in zip(o, solo[va])) for o in oa)
            bad_b += sum(not all(torch.equal(x, y) for x, y in zip(o, solo[vb])) for o in ob)
        if bad_a or bad_b:
            print(f"A={va:14s} B={vb:14s}: A wrong {bad_a}/18, B wrong {bad_b}/18", flush=True)
print("done")
