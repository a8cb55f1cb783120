"""find the first activation that differs between a serialised and an overlapped two-stream native forward"""
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
seg.backbone.attention_mode = "tc32"
seg.backbone.overlap_streams = True
n = len(sc["coord"])
rng = np.random.default_rng(5)
noise = rng.standard_normal((n, 6)).astype(np.float32)
perms = [rng.permutation(4) for _ in range(8)]
inp = {k: t(sc[k]).to(DEV) for k in ("coord", "grid_coord", "offset", "feat")}
base = {k: inp[k] for k in ("coord", "grid_coord", "offset")}
noise_d = t(noise).to(DEV)
def fwd():
    ts = 999 * torch.ones((1, 1), dtype=torch.int64, device=DEV)
    c, nn_ = seg.backbone(dict(base, feat=noise_d, t_emb=calc_t_emb(ts, 128)), dict(base, feat=inp["feat"]), perm_fn=replay(perms))
    torch.cuda.synchronize()
    return nn_["feat"].clone(), c["feat"].clone()
lib.cdseg_net_set_debug(3 | 4)
sys.stderr.flush()
fd = os.open("/tmp/alloc.log", os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
old = os.dup(2); os.dup2(fd, 2)
fwd()                                   # logs the allocation table once (serialised)
os.dup2(old, 2); os.close(fd)
import re
allocs = {}
for line in open("/tmp/alloc.log"):
    m = re.match(r"\[net\] arena (\S+) alloc off=(\d+) bytes=(\d+)", line)
    if m:
        allocs.setdefault(m.group(1), []).append((int(m.group(2)), int(m.group(3))))
lib.cdseg_net_set_debug(3)
rn, rc = fwd()
am = netexec._ARENAS[(torch.device(DEV).index if torch.device(DEV).index is not None else inp["feat"].device.index, "main")]
as_ = netexec._ARENAS[(inp["feat"].device.index, "side")]
ref_m, ref_s = am.clone(), as_.clone()
for flags in (0, 8, 16, 24):
    lib.cdseg_net_set_debug(3 | flags)
    rn, rc = fwd()
    lib.cdseg_net_set_debug(flags)
    bad = 0
    for rep in range(6):
        on, oc = fwd()
        bad += int((on - rn).abs().max().item() > 1e-3 or (oc - rc).abs().max().item() > 1e-3)
    print("flags", flags, "(8: side attention exact, 16: main attention exact):", bad, "of 6 forwards wrong", flush=True)
