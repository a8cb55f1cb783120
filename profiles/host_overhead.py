"""How long does the HOST need to enqueue one forward, vs. the GPU time of the step?  (is the step launch-bound?)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cdsegnet_b200 as cb
from cdsegnet_b200 import configs, ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
seg = cb.build_model(configs.segmentor_cfg()); bench.random_weights(seg); seg = seg.to(dev).eval()
sc = bench.make_scene(0)
inp = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in sc.items()}
noise = torch.randn(len(sc["coord"]), 6, device=dev)
for overlap in (True, False):
    seg.backbone.overlap_streams = overlap
    for _ in range(3):
        seg.inference(inp, eval=False, noise=noise)
    torch.cuda.synchronize()
    enq, tot = [], []
    for _ in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seg.inference(inp, eval=False, noise=noise)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        enq.append(t1 - t0); tot.append(t2 - t0)
    print(f"overlap_streams={overlap}: host enqueue {1e3*np.median(enq):.2f} ms, enqueue+drain {1e3*np.median(tot):.2f} ms per forward")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    seg.inference(inp, eval=False, noise=noise)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
