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
lib.cdseg_net_set_debug(3)
fwd()
am = netexec._ARENAS[(inp["feat"].device.index, "main")]; as_ = netexec._ARENAS[(inp["feat"].device.index, "side")]
ref_m, ref_s = am.clone(), as_.clone()
lib.cdseg_net_set_debug(0)
for rep in range(8):
    fwd()
    for name, arena, ref in (("main", am, ref_m), ("side", as_, ref_s)):
        out = []
        for rn, o, nb in regions:
            a, b = arena[o:o + nb], ref[o:o + nb]
            bad = (a != b)
            if bool(bad.any()):
                idx = bad.nonzero().flatten()
                out.append(f"{rn}: {int(bad.sum())} bytes differ, first at byte {int(idx[0])} (row {int(idx[0]) // (nb // 120000) if 'pack' not in rn else -1})")
        print(rep, name, "scratch:", out if out else "identical", flush=True)
