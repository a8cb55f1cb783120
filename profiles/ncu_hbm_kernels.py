"""One plan phase + one pooling / gather pass of the 120k scene for ncu: the HBM-side kernels north_star lists (key encode, radix
sort, patch gather, grid-pool plan / reduce, row gathers).  python profiles/ncu_hbm_kernels.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops, synth
from cdsegnet_b200.structure import Plan
dev = "cuda"
sc = synth.collate([synth.scannet_scene(120000, 0)])
g = torch.from_numpy(sc["grid_coord"]).to(dev); off = torch.from_numpy(sc["offset"]).to(dev)
spec = dict(n=[dict(K=1024, mask=15, conv_plan=True, stem=5 if s == 0 else 0) for s in range(5)],
            c=[dict(K=1024, mask=15, conv_plan=True, stem=5 if s == 0 else 0) for s in range(3)])
for _ in range(2):
    plan = Plan(g, off, ("z", "z-trans", "hilbert", "hilbert-trans"), (2, 2, 2, 2), (4, 4), True, spec=spec)
L0, L1 = plan.n_levels[0], plan.n_levels[1]
x = torch.randn(L0.n, 64, device=dev)
for _ in range(2):
    f, _ = ops.pool_reduce(x, None, L1.members(), L1.idx_ptr, L1.n, None, None, False)
    y = ops.gather_rows(x, L0.inv_perm)
    u = ops.unpool_add(x, f, L1.cluster[: L0.n], 1.0)
    pm = L0.patch_maps(0, 1024)
    qkv = torch.randn(L0.n, 96, device=dev)
    ops.attn_pack(qkv, 0, 32, 3, pm, 2, "f16"); ops.attn_pack(qkv, 0, 32, 3, pm, 2, "tc32")
torch.cuda.synchronize()
print("ok", L0.n, L1.n)
