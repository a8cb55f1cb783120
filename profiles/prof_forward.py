"""Run N full CDSegNet forwards on a 120k-point synthetic scene (same workload as bench.py) so that ncu
can attach to individual kernels:   ncu ... -k regex:attn_tc_kernel -s 37 -c 1 python profiles/prof_forward.py 2
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cdsegnet_b200 as cb  # noqa: E402
from cdsegnet_b200 import configs  # noqa: E402

n_fwd = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
seg = cb.build_model(configs.segmentor_cfg())
bench.random_weights(seg)
seg = seg.to(dev).eval()
seg.backbone.attention_mode = sys.argv[2] if len(sys.argv) > 2 else "tc32"
sc = bench.make_scene(0)
inp = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in sc.items()}
noise = torch.randn(len(sc["coord"]), 6, device=dev)
for _ in range(n_fwd):
    out = seg.inference(inp, eval=False, noise=noise)["seg_logits"]
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().mean()))
