"""which kernel of a level-0 Block produces the first wrong values under two-stream overlap?  Both branches stop after
CDSEG_NET_BLOCK_LIMIT blocks; the block scratch (x1, qkv, packed operands, o) of the last executed block is compared between a
serialised and an overlapped run."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
import cdsegnet_b200 as cb
from cdsegnet_b200 import configs, ops, synth, netexec, _lib
from cdsegnet_b200.segmentor import calc_t_emb
from helpers import replay, t
DEV = "cuda"
lib = _lib.load()
sc = synth.collate([synth.scannet_scene(120000, 0)])
torch.manual_seed(0)
seg = cb.build_model(configs.segmentor_cfg()); bench.random_weights(seg); seg = seg.to(DEV).eval()
seg.backbone.attention_mode = os.environ.get("DBG_ATTN", "tc32")
seg.backbone.overlap_streams = True
n = len(sc["coord"])
rng = np.random.default_rng(5)
noise = t(rng.standard_normal((n, 6)).astype(np.float32)).to(DEV)
perms = [rng.permutation(4) for _ in range(8)]
inp = {k: t(sc[k]).to(DEV) for k in ("coord", "grid_coord", "offset", "feat")}
base = {k: inp[k] for k in ("coord", "grid_coord", "offset")}
def fwd():
    ts = 999 * torch.ones((1, 1), dtype=torch.int64, device=DEV)
    seg.backbone(dict(base, feat=noise, t_emb=calc_t_emb(ts, 128)), dict(base, feat=inp["feat"]), perm_fn=replay(perms))
    torch.cuda.synchronize()
C, row = 32, 120000 * 32 * 4
regions = [("y1", 0, row), ("y2", row, row), ("x1 (pre)", 2 * row, row), ("h", 3 * row, row), ("o (attn)", 4 * row, row), ("att", 5 * row, row),
           ("qkv (pre)", 6 * row, 3 * row), ("hid", 9 * row, 4 * row)]
pk = 2 * 118 * 1024 * 16 * 2
off = 13 * row
mode = seg.backbone.attention_mode
units = {"f16": (1, 1, 2), "exact": (2, 2, 2), "tc32": (2, 2, 3)}[mode]
for name, u in zip(("q pack", "k pack", "v pack"), units):
    regions.append((name, off, u * pk)); off += u * pk
# activations behind scratch + workspace (allocation order of net_exec.cu): main: x8, stem out, block0 out, block1 out ; side: t1, t_scene, x8, stem, block0, block1
ACT_MAIN = 541184256 + 24764416
ACT_SIDE = 541184256 + 0 + 2048 + 512
acts = {"main": [("stem out (post of nothing)", ACT_MAIN + 3840000, row), ("block0 out (post)", ACT_MAIN + 3840000 + row, row), ("block1 out (post)", ACT_MAIN + 3840000 + 2 * row, row)],
        "side": [("stem out", ACT_SIDE + 3840000, row), ("block0 out (post)", ACT_SIDE + 3840000 + row, row), ("block1 out (post)", ACT_SIDE + 3840000 + 2 * row, row)]}
lib.cdseg_net_set_debug(3)
fwd()
am = netexec._ARENAS[(inp["feat"].device.index, "main")]; as_ = netexec._ARENAS[(inp["feat"].device.index, "side")]
ref_m, ref_s = am.clone(), as_.clone()
lib.cdseg_net_set_debug(0)
for rep in range(40):
    fwd()
    for name, arena, ref in (("main", am, ref_m), ("side", as_, ref_s)):
        out = []
        for rn, o, nb in regions:
            a, b = arena[o:o + nb], ref[o:o + nb]
            bad = (a != b)
            if bool(bad.any()):
                idx = bad.nonzero().flatten()
                out.append(f"{rn}: {int(bad.sum())} bytes differ, first at byte {int(idx[0])} (row {int(idx[0]) // (nb // 120000) if 'pack' not in rn else -1})")
        for rn, o, nb in acts[name]:
            a = arena[o:o + nb].view(torch.float32); b = ref[o:o + nb].view(torch.float32)
            bad = (a != b).nonzero().flatten()
            if len(bad):
                rows = torch.unique(bad // 32)
                r0 = int(rows[0])
                out.append(f"ACTIVATION {rn}: {len(rows)} rows differ {rows[:12].tolist()}, max|d| {float((a - b).abs().max()):.3e}, row {r0}: got {[round(float(v), 5) for v in a[r0 * 32:r0 * 32 + 3]]} ref {[round(float(v), 5) for v in b[r0 * 32:r0 * 32 + 3]]}")
        print(rep, name, "scratch:", out if out else "identical", flush=True)
        if out:                                  # nature of the wrong values in x1 / qkv (fused pre-attention kernel outputs)
            for rn, o, nb in regions:
                if "(pre)" not in rn:
                    continue
                a = arena[o:o + nb].view(torch.float32); b = ref[o:o + nb].view(torch.float32)
                w = nb // 4 // 120000
                bad = (a != b).nonzero().flatten()
                rows = torch.unique(bad // w)
                print("   ", rn, "width", w, "wrong rows:", rows[:24].tolist(), "... total", len(rows), "| tiles", torch.unique(rows // 128)[:12].tolist())
                for r in rows[:3].tolist():
                    cols = (a[r * w:(r + 1) * w] != b[r * w:(r + 1) * w]).nonzero().flatten()
                    print("       row", r, "cols", cols[:8].tolist(), f"({len(cols)} of {w})", "got", [round(float(v), 5) for v in a[r * w + cols[:4]]],
                          "ref", [round(float(v), 5) for v in b[r * w + cols[:4]]])
                # are the wrong rows equal to OTHER rows of the reference (shifted / swapped tiles)?
                r0 = int(rows[0])
                match = ((b.view(-1, w) - a[r0 * w:(r0 + 1) * w]).abs().max(1).values < 1e-6).nonzero().flatten()
                print("       row", r0, "of the racy run equals reference rows:", match[:8].tolist())
