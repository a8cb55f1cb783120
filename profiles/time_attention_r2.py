"""Isolated timing of the patch-attention kernels (round 2): generations 2 (attn_tc2.cu) and 3 (attn_tc3.cu) in the fp16 mode, the
fp32-faithful "tc32" mode and the FMA-pipe exponential variants, at the stage-0 shapes of the 120k scene (118 patches x 1024 keys,
H = 2 / 4) and at the deep levels (H = 16 at 3 804 points, H = 32 at 991).  CUDA events, L2 flushed, median of 10."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(n, K, H, mode, kernel=3, poly=0, reps=12):
    C = 16 * H
    g = torch.Generator(device=dev).manual_seed(0)
    order = torch.randperm(n, device=dev, generator=g).int()
    pm = ops.patch_maps(order, np.array([n]), K)
    qkv = torch.randn(n, 3 * C, device=dev, generator=g)
    ops.ATTN_KERNEL = kernel
    lib.cdseg_attn_set_poly(poly)
    q, k, v = ops.attn_pack(qkv, 0, C, 3, pm, H, mode)
    ts = []
    for i in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = ops.attn(q, k, v, pm, H, 0.25, n, mode); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    ops.ATTN_KERNEL = 3
    lib.cdseg_attn_set_poly(0)
    ms = float(np.median(ts))
    ex = pm["pairs"] * H
    return ms, 4.0 * pm["pairs"] * C / ms / 1e9, ex / ms / 1e9, o


for (n, H) in ((120000, 2), (120000, 4), (52190, 4), (14640, 8), (3804, 16), (991, 32)):
    ref = None
    for label, mode, kernel, poly in (("tc2 f16 (round 1)", "f16", 2, 0), ("tc3 f16", "f16", 3, 0), ("tc3 f16 poly1", "f16", 3, 1),
                                      ("tc3 f16 poly2", "f16", 3, 2), ("tc3 f16 poly3", "f16", 3, 3), ("tc3 tc32", "tc32", 3, 0),
                                      ("exact SIMT", "exact", 3, 0)):
        if mode == "exact" and n > 60000:
            continue
        ms, tf, te, o = run(n, 1024, H, mode, kernel, poly)
        if ref is None:
            ref = o
        print(f"n={n:6d} H={H:2d} {label:18s}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s  {te:6.2f} Texp/s ({100*te/4.653:3.0f}% of MUFU peak)  "
              f"max|o - o_tc2| = {(o - ref).abs().max().item():.2e}", flush=True)
