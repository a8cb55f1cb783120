import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cdsegnet_b200 as cb
from cdsegnet_b200 import configs, synth
DEV = "cuda"
sc = synth.collate([synth.scannet_scene(int(sys.argv[1]) if len(sys.argv) > 1 else 120000, 0)])
torch.manual_seed(0)
seg = cb.build_model(configs.segmentor_cfg()); bench.random_weights(seg); seg = seg.to(DEV).eval()
seg.backbone.attention_mode = "tc32"
inp = {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in sc.items()}
for _ in range(2):
    out = seg.inference(inp, eval=False)["seg_logits"]
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
