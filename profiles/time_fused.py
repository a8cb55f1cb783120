"""Fused row-tile kernels against the launch sequences they replace, alone on the device (CUDA events, L2-warm and
L2-flushed):  python profiles/time_fused.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, cold, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


def lin(cin, cout):
    w = torch.randn(cout, cin, device=dev) / cin ** 0.5
    return ops.gemm_pack_b(w.t().contiguous()[None]), torch.randn(cout, device=dev)


for n, C in ((120000, 32), (120000, 64), (52190, 64), (14640, 128), (14640, 64)):
    o, x1 = torch.randn(n, C, device=dev), torch.randn(n, C, device=dev)
    proj, fc1, fc2 = lin(C, C), lin(C, 4 * C), lin(4 * C, C)
    g, b = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)

    def old():
        att = ops.gemm_tc(o, proj[0], C, C, bias=proj[1])
        x2, h = ops.add_layernorm(x1, att, gamma=g, beta=b)
        hid = ops.gemm_tc(h, fc1[0], 4 * C, C, bias=fc1[1], act=1)
        return ops.gemm_tc(hid, fc2[0], C, 4 * C, bias=fc2[1], res=x2)

    def new():
        return ops.post_attn(o, x1, proj, (g, b), fc1, fc2)
    err = (old() - new()).abs().max().item()
    alg = 3 * n * C * 4
    for cold in (False, True):
        t_old, t_new = timeit(old, cold), timeit(new, cold)
        print(f"post_attn n={n} C={C} {'cold' if cold else 'warm'}: 4 launches {t_old:7.1f} us -> fused {t_new:7.1f} us "
              f"({alg / t_new * 1e-3:6.0f} GB/s algorithmic, max|old-new| {err:.1e})")


# ---- pre-attention chain on a real neighbour table
import numpy as np
import bench
sc = bench.make_scene(0)
grid = torch.from_numpy(np.ascontiguousarray(sc["grid_coord"])).int().to(dev)
n = grid.shape[0]
batch = torch.zeros(n, dtype=torch.int32, device=dev)
# curve order first, like the model does internally (rows of a tile are spatial neighbours)
codes = ops.encode_codes(grid, batch, 9, ["z"])
order, _ = ops.argsort_rows(codes, 27)
grid = grid[order[0].long()].contiguous()
nbr = ops.nbr_build(grid, batch, 3)
mask = ops.tile_tap_mask(nbr)
plan = ops.conv_tile_plan(nbr)
U = plan.view(-1, plan.numel() // ((n + 127) // 128))[:, :4].contiguous().view(torch.int32).flatten()
print(f"conv tile plan: {U.numel()} tiles, distinct neighbour rows per tile mean {U.float().mean().item():.1f} max {U.max().item()} (cache {ops.CONV_PLAN_UCAP})")
for C in (32, 64):
    x = torch.randn(n, C, device=dev)
    wc = torch.randn(27, C, C, device=dev) / (27 * C * 0.4) ** 0.5
    conv = (ops.gemm_pack_b(wc), torch.randn(C, device=dev))
    l1, qk = lin(C, C), lin(C, 3 * C)
    cg, cb, g1, b1 = (torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev), torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev))

    def old():
        y = ops.gemm_tc(x, conv[0], C, C, idx=nbr, tile_mask=mask, bias=conv[1])
        y = ops.gemm_tc(y, l1[0], C, C, bias=l1[1])
        _, y = ops.add_layernorm(y, gamma=cg, beta=cb, want_sum=False)
        x1, h = ops.add_layernorm(x, y, gamma=g1, beta=b1)
        return x1, ops.gemm_tc(h, qk[0], 3 * C, C, bias=qk[1])

    def new():
        return ops.pre_attn(x, x, nbr, mask, plan, conv, l1, (cg, cb), (g1, b1), qk)
    (a1, aq), (b1_, bq) = old(), new()
    err = max((a1 - b1_).abs().max().item(), (aq - bq).abs().max().item())
    for cold in (False, True):
        t_old, t_new = timeit(old, cold), timeit(new, cold)
        print(f"pre_attn n={n} C={C} {'cold' if cold else 'warm'}: 5 launches {t_old:7.1f} us -> fused {t_new:7.1f} us (max|old-new| {err:.1e})")
