"""error pattern of post_kernel / pre_kernel outputs under two-stream concurrency"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops, synth
dev = "cuda"
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "debug_concurrent.py")).read().split("cands = ")[0])
def pattern(name, got, ref):
    d = (got != ref)
    rows = d.any(1).nonzero().flatten().cpu().numpy()
    cols = d.any(0).nonzero().flatten().cpu().numpy()
    tiles = np.unique(rows // 128)
    mx = (got - ref).abs().max().item()
    print(f"   {name}: {len(rows)} rows in {len(tiles)} tiles {tiles[:8]}, cols {cols.min()}..{cols.max()} ({len(cols)}), rows in tile {np.unique(rows % 128)[:6]}..{(rows % 128).max()}, "
          f"max|d| {mx:.3e}, nan {torch.isnan(got).sum().item()}", flush=True)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
for nameA, fa, nameB, fb in (("post32", make_post(32), "post32", make_post(32)), ("post64", make_post(64), "post64", make_post(64)),
                             ("pre32", make_pre(32), "pre32", make_pre(32)), ("pre32", make_pre(32), "post32", make_post(32)),
                             ("pre32", make_pre(32), "attn_tc32", make_attn(2, "tc32")), ("post32", make_post(32), "attn_tc32", make_attn(2, "tc32")),
                             ("post128", make_post(128), "post128", make_post(128))):
    ra = [o.clone() for o in fa()]; rb = [o.clone() for o in fb()]; torch.cuda.synchronize()
    bad = 0
    for rep in range(40):
        torch.cuda.synchronize()
        with ops.stream_scope(sa):
            oa = [fa() for _ in range(3)]
        with ops.stream_scope(sb):
            ob = [fb() for _ in range(3)]
        torch.cuda.synchronize()
        for o in oa:
            for x, y in zip(o, ra):
                if not torch.equal(x, y):
                    bad += 1
                    if bad <= 3: pattern(nameA + "(A)", x, y)
        for o in ob:
            for x, y in zip(o, rb):
                if not torch.equal(x, y):
                    bad += 1
                    if bad <= 3: pattern(nameB + "(B)", x, y)
    print(f"{nameA} || {nameB}: {bad} wrong outputs of 240+", flush=True)
