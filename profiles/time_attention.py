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

# ---- FMA-pipe exponentials: POLY of every 8 exponentials by polynomial instead of MUFU.EX2 (cdseg_attn_set_poly)
from cdsegnet_b200 import _lib
lib = _lib.load()
for H in (2, 4):
    C = 16 * H
    order = torch.randperm(n, device=dev).int()
    pm = ops.patch_maps(order, np.array([n]), K)
    qkv = torch.randn(n, 3 * C, device=dev)
    ops.ATTN_V2 = True
    q, k, v = ops.attn_pack(qkv, 0, C, 3, pm, H)
    lib.cdseg_attn_set_poly(0)
    base = ops.attn(q, k, v, pm, H, 0.25, n).clone()
    for poly in (0, 1, 2, 3):
        lib.cdseg_attn_set_poly(poly)
        ts = []
        for i in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = ops.attn(q, k, v, pm, H, 0.25, n); e1.record(); torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        ex = pm["pairs"] * H
        print(f"H={H} poly={poly}/8: {ms*1e3:7.1f} us  {4.0*pm['pairs']*C/ms/1e9:7.1f} TFLOP/s  {ex/ms/1e9:6.2f} Texp/s ({100*ex/ms/1e9/4.653:.0f}% of the MUFU-only peak)"
              f"  max|o - o_poly0| = {(o - base).abs().max().item():.2e}")
lib.cdseg_attn_set_poly(0)

# ---- upstream-kernel comparator: flash_attn 2.8.3's varlen kernel on the same stage-0 problem, alone and inside the
# reference's own sequence qkv[order] -> .half() -> flash_attn_varlen_qkvpacked_func -> feat[inverse] (ptv3.py:258-290)
try:
    import flash_attn
except Exception as e:                                     # not installed: nothing to compare against
    flash_attn = None
    print("flash_attn not importable:", e)
if flash_attn is not None:
    for H in (2, 4):
        C = 16 * H
        order = torch.randperm(n, device=dev)
        inverse = torch.empty_like(order); inverse[order] = torch.arange(n, device=dev)
        pm = ops.patch_maps(order.int(), np.array([n]), K)
        npad = -(-n // K) * K
        pad = torch.cat([torch.arange(n, device=dev), torch.arange(n - (npad - n), n, device=dev)])[:npad]      # same sizes as the reference's pad map
        unpad = torch.arange(n, device=dev)
        cu_seqlens = torch.arange(0, npad + 1, K, device=dev, dtype=torch.int32)
        qkv = torch.randn(n, 3 * C, device=dev)

        def ref_seq():
            g = qkv.half()[order][pad]
            o = flash_attn.flash_attn_varlen_qkvpacked_func(g.reshape(-1, 3, H, 16), cu_seqlens, max_seqlen=K, dropout_p=0.0, softmax_scale=0.25)
            return o.reshape(-1, C)[unpad][inverse].float()

        g = qkv.half()[order][pad].reshape(-1, 3, H, 16).contiguous()
        kern = lambda: flash_attn.flash_attn_varlen_qkvpacked_func(g, cu_seqlens, max_seqlen=K, dropout_p=0.0, softmax_scale=0.25)

        def ours_seq():
            ops.ATTN_V2 = True
            q, k, v = ops.attn_pack(qkv, 0, C, 3, pm, H)
            return ops.attn(q, k, v, pm, H, 0.25, n)
        for name, fn in (("flash_attn varlen kernel alone", kern), ("reference sequence (gather, half, flash_attn, gather)", ref_seq),
                         ("this repo: pack_heads + attn_tc2", ours_seq)):
            ts = []
            for i in range(12):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            print(f"H={H} {name}: {ms * 1e3:7.1f} us  ({4.0 * n * K * C / ms / 1e9:6.1f} TFLOP/s on {n} x {K} pairs)")
