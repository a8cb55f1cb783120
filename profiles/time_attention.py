"""Isolated timing of the stage-0 patch-attention kernel (118 patches x 1024 keys, H heads), CUDA events, L2 flushed."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops
dev = "cuda"
n, K = 120000, 1024
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for H in (2, 4):
    C = 16 * H
    order = torch.randperm(n, device=dev).int()
    pm = ops.patch_maps(order, np.array([n]), K)
    qkv = torch.randn(n, 3 * C, device=dev)
    for v2 in (True, False):
        ops.ATTN_V2 = v2
        q, k, v = ops.attn_pack(qkv, 0, C, 3, pm, H)
        ts = []
        for i in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.attn(q, k, v, pm, H, 0.25, n); e1.record(); torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        ex = pm["pairs"] * H
        print(f"H={H} kernel={'attn_tc2' if v2 else 'attn_tc '} OCC={os.environ.get('CDSEG_ATTN_OCC','4')}: {ms*1e3:7.1f} us  "
              f"{4.0*pm['pairs']*C/ms/1e9:7.1f} TFLOP/s  {ex/ms/1e9:6.2f} Texp/s ({100*ex/ms/1e9/4.653:.0f}% of MUFU peak)")
