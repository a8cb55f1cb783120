"""clock64 timeline of single CTAs of gemm_tc launches (profiling hook cdseg_gemm_tc_set_trace): where does a CTA of a
1-iteration Linear (120k x 32 -> N) spend its time?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdsegnet_b200 import ops, _lib
dev = "cuda"
lib = _lib.load()
n = 120000
names = ["start", "loads_issued", "chunk0_converted", "chunk0_in_tmem", "producer_done", "acc_complete", "epi_pass0", "epi_pass1",
         "epi_pass2", "epi_pass3", "mma_A_ready", "mma_B_ready", "mma_issued", "end"]
for (K, N, act) in ((32, 32, 0), (32, 96, 0), (32, 128, 1), (128, 32, 0)):
    x = torch.randn(n, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    Bp = ops.gemm_pack_b(w.t().contiguous()[None])
    for _ in range(3):
        ops.gemm_tc(x, Bp, N, K, bias=b, act=act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm_tc(x, Bp, N, K, bias=b, act=act)
    e1.record(); torch.cuda.synchronize()
    alg = n * (K + N) * 4
    print(f"K={K} N={N} gelu={act}: untraced {e0.elapsed_time(e1) * 50:.1f} us/launch, {alg / (e0.elapsed_time(e1) * 50e-6) / 1e9:.0f} GB/s algorithmic")
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    for cta in (3, 500):
        buf.zero_()
        lib.cdseg_gemm_tc_set_trace(buf.data_ptr(), cta)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm_tc(x, Bp, N, K, bias=b, act=act); e1.record(); torch.cuda.synchronize()
        t = buf.cpu().numpy()
        print(f"K={K} N={N} gelu={act} CTA {cta}: kernel {e0.elapsed_time(e1)*1e3:.1f} us; cycles since CTA start: " +
              ", ".join(f"{nm}={int(v - t[0])}" for nm, v in zip(names, t) if v))
lib.cdseg_gemm_tc_set_trace(None, 0)
