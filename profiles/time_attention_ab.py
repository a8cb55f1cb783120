"""A/B timing of attention builds / poll back-off (CDSEG_LIB, CDSEG_ATTN_SLEEP): stage-0 shapes only"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (n, H) in ((120000, 2), (120000, 4), (3804, 16)):
    C = 16 * H
    g = torch.Generator(device=dev).manual_seed(0)
    order = torch.randperm(n, device=dev, generator=g).int()
    pm = ops.patch_maps(order, np.array([n]), 1024)
    qkv = torch.randn(n, 3 * C, device=dev, generator=g)
    out = []
    for mode in ("f16", "tc32"):
        q, k, v = ops.attn_pack(qkv, 0, C, 3, pm, H, mode)
        ts = []
        for i in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.attn(q, k, v, pm, H, 0.25, n, mode); e1.record(); torch.cuda.synchronize()
            if i >= 2: ts.append(e0.elapsed_time(e1))
        out.append(f"{mode} {1e3*float(np.median(ts)):7.1f} us")
    print(f"lib={os.path.basename(os.environ.get('CDSEG_LIB', 'default')):28s} sleep={os.environ.get('CDSEG_ATTN_SLEEP', '0'):>4s} n={n:6d} H={H:2d}: " + "   ".join(out), flush=True)
