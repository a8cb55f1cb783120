"""Deep-level Linears (n = 3 000 / 900 rows, C = 256 / 512) as the executor launches them (ptv3.linear: split-K / narrow-tile heuristic):
launch time by CUDA events (20 back-to-back launches, so the ~2-3 us launch-to-launch cost is inside) and the clock64 timeline of CTA 0
(profiling hook cdseg_gemm_tc_set_trace).  Where do 13-27 us go when the tensor work is ~1 us?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops, _lib, ptv3
dev = "cuda"
lib = _lib.load()
names = ["start", "loads_issued", "chunk0_converted", "chunk0_in_tmem", "producer_done", "acc_complete", "epi_pass0", "epi_pass1",
         "epi_pass2", "epi_pass3", "mma_A_ready", "mma_B_ready", "mma_issued", "end"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, K, N, act, what) in ((3000, 256, 256, 0, "C=256 proj / cpe linear"), (3000, 256, 768, 0, "C=256 qkv"), (3000, 256, 1024, 1, "C=256 fc1"),
                             (3000, 1024, 256, 0, "C=256 fc2"), (900, 512, 512, 0, "C=512 proj"), (900, 512, 1536, 0, "C=512 qkv"),
                             (900, 512, 2048, 1, "C=512 fc1"), (900, 2048, 512, 0, "C=512 fc2")):
    x = torch.randn(M, K, device=dev); w = torch.nn.Parameter(torch.randn(N, K, device=dev) / K ** 0.5); b = torch.nn.Parameter(torch.randn(N, device=dev))
    res = torch.randn(M, N, device=dev)
    for _ in range(3):
        ptv3.linear(x, w, b, act=act, res=res)
    torch.cuda.synchronize()
    lib.cdseg_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ptv3.linear(x, w, b, act=act, res=res)
    e1.record(); torch.cuda.synchronize()
    nl = lib.cdseg_launch_count() / 20
    # cold (L2 flushed) single launch
    flush.fill_(1); torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(); ptv3.linear(x, w, b, act=act, res=res); c1.record(); torch.cuda.synchronize()
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.cdseg_gemm_tc_set_trace(buf.data_ptr(), 0)
    ptv3.linear(x, w, b, act=act, res=res); torch.cuda.synchronize()
    lib.cdseg_gemm_tc_set_trace(None, 0)
    t = buf.cpu().numpy()
    print(f"{what:24s} M={M} K={K} N={N}: {e0.elapsed_time(e1) * 50:.1f} us per Linear warm ({nl:.0f} launches), {c0.elapsed_time(c1) * 1e3:.1f} us cold; "
          f"CTA 0 cycles: " + ", ".join(f"{nm}={int(v - t[0])}" for nm, v in zip(names, t) if v))
