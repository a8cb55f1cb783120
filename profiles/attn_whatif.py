"""What-if timing of the mode-0 attention kernel (attn_tc3.cu): each variant drops ONE piece of the per-chunk work (results are
wrong on purpose) so that the time it frees shows what the kernel is bound by.  Stage-0 shape of the 120k scene, H = 2 and 4."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NAMES = {0: "full kernel", 1: "no exponentials (FFMA result stored)", 2: "S read from TMEM only in chunk 0", 3: "no P stores to shared memory",
         4: "no row max", 5: "no P.V MMAs (chunk 0 only)", 6: "no O fold loads from TMEM"}
for n, H in ((120000, 2), (120000, 4)):
    C = 16 * H
    g = torch.Generator(device=dev).manual_seed(0)
    order = torch.randperm(n, device=dev, generator=g).int()
    pm = ops.patch_maps(order, np.array([n]), 1024)
    qkv = torch.randn(n, 3 * C, device=dev, generator=g)
    q, k, v = ops.attn_pack(qkv, 0, C, 3, pm, H, "f16")
    for dbg in range(7):
        lib.cdseg_attn_set_debug(dbg)
        ts = []
        for i in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.attn(q, k, v, pm, H, 0.25, n, "f16"); e1.record(); torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        print(f"H={H} variant {dbg} ({NAMES[dbg]:38s}): {1e3*float(np.median(ts)):7.1f} us", flush=True)
    lib.cdseg_attn_set_debug(0)
