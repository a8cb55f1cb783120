import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
import cdsegnet_b200 as cb
from cdsegnet_b200 import configs, ops, synth
from cdsegnet_b200.segmentor import calc_t_emb
from helpers import replay, t
DEV = "cuda"
sc = synth.collate([synth.scannet_scene(120000, 0)])
torch.manual_seed(0)
seg = cb.build_model(configs.segmentor_cfg()); bench.random_weights(seg); seg = seg.to(DEV).eval()
seg.backbone.attention_mode = os.environ.get("DBG_ATTN", "tc32")
ops.set_fused_mask(int(os.environ.get("DBG_FUSED", "7")))
n = len(sc["coord"])
rng = np.random.default_rng(5)
noise = rng.standard_normal((n, 6)).astype(np.float32)
perms = [rng.permutation(4) for _ in range(8)]
inp = {k: t(sc[k]).to(DEV) for k in ("coord", "grid_coord", "offset", "feat")}
base = {k: inp[k] for k in ("coord", "grid_coord", "offset")}
def via_seg():
    seg.backbone.perm_fn = replay(perms)
    return seg.inference(inp, eval=False, noise=t(noise))["seg_logits"].cpu().numpy()
def via_backbone():
    ts = 999 * torch.ones((1, 1), dtype=torch.int64, device=DEV)
    c, nn_ = seg.backbone(dict(base, feat=t(noise).to(DEV), t_emb=calc_t_emb(ts, 128)), dict(base, feat=inp["feat"]), perm_fn=replay(perms))
    return nn_["feat"].cpu().numpy()
ops.NATIVE_NET = False
ref = via_backbone()
for name, fn in (("backbone per-module", via_backbone), ("seg per-module", via_seg)):
    o = fn(); d = np.abs(o - ref); print(name, d.max(), int((d.max(1) > 1e-4).sum()))
ops.NATIVE_NET = True
for rep in range(4):
    o = via_backbone(); d = np.abs(o - ref); print(os.environ.get("DBG_ATTN"), os.environ.get("DBG_FUSED"), rep, "backbone native", d.max(), int((d.max(1) > 1e-4).sum()), flush=True)
