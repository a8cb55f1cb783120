"""one launch of each attention mode at the stage-0 shape for ncu (--set full): python profiles/ncu_attn_r2.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops
dev = "cuda"
n, H = 120000, 2
C = 16 * H
g = torch.Generator(device=dev).manual_seed(0)
order = torch.randperm(n, device=dev, generator=g).int()
pm = ops.patch_maps(order, np.array([n]), 1024)
qkv = torch.randn(n, 3 * C, device=dev, generator=g)
for mode in ("f16", "tc32"):
    q, k, v = ops.attn_pack(qkv, 0, C, 3, pm, H, mode)
    for _ in range(3):
        ops.attn(q, k, v, pm, H, 0.25, n, mode)
torch.cuda.synchronize()
print("ok")
