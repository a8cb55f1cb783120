"""clock64 timeline of one CTA of the fused pre-attention kernel (profiling hook cdseg_pre_attn_set_trace): where do the ~27 us
per 128-row tile go?  Stage-0 shaped input (120k-point scene in curve order)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cdsegnet_b200 import ops, _lib
dev = "cuda"
lib = _lib.load()
names = ["tile_start", "fill_arrived", "split_done", "taps_done", "conv_acc", "lin_operand", "lin_acc", "x1_done", "qkv_operand", "qkv_acc0",
         "qkv_stored", "tile_end"]
sc = bench.make_scene(0)
grid = torch.from_numpy(np.ascontiguousarray(sc["grid_coord"])).int().to(dev)
n = grid.shape[0]
batch = torch.zeros(n, dtype=torch.int32, device=dev)
codes = ops.encode_codes(grid, batch, 9, ["z"])
order, _ = ops.argsort_rows(codes, 27)
grid = grid[order[0].long()].contiguous()
nbr = ops.nbr_build(grid, batch, 3)
mask = ops.tile_tap_mask(nbr)
plan = ops.conv_tile_plan(nbr)
lin = lambda ci, co: (ops.gemm_pack_b((torch.randn(co, ci, device=dev) / ci ** 0.5).t().contiguous()[None]), torch.randn(co, device=dev))
for C in (32, 64):
    x = torch.randn(n, C, device=dev)
    conv = (ops.gemm_pack_b(torch.randn(27, C, C, device=dev) / (27 * C * 0.4) ** 0.5), torch.randn(C, device=dev))
    l1, qk = lin(C, C), lin(C, 3 * C)
    ln = lambda: (torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev))
    a, b = ln(), ln()
    run = lambda: ops.pre_attn(x, x, nbr, mask, plan, conv, l1, a, b, qk)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    for cta in (0, 150):
        buf = torch.zeros(6 * 12 + 18, dtype=torch.int64, device=dev)
        lib.cdseg_pre_attn_set_trace(buf.data_ptr(), cta)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        lib.cdseg_pre_attn_set_trace(None, 0)
        full = buf.cpu().numpy()
        t, w = full[:72].reshape(6, 12), full[72:].reshape(6, 3)
        print(f"C={C} CTA {cta}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us, popc(mask) of its first tile = {bin(int(mask[cta]) & 0x7ffffff).count('1')}")
        for i, row in enumerate(t):
            if row[0] == 0:
                break
            print(f"  tile {i}: " + ", ".join(f"{nm}=+{int(v - row[0])}" for nm, v in zip(names[1:], row[1:]) if v) +
                  (f"   (gap to next tile start {int(t[i + 1][0] - row[11])})" if i + 1 < 6 and t[i + 1][0] else ""))
            print(f"          tap loop of thread 0: a_empty waits {int(w[i][0])}, gather + store issue {int(w[i][1])}, wait::st + arrive {int(w[i][2])}")
