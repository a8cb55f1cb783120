"""Where the host time of one forward goes (round 2): enqueue time of the plan phase and of the feature phase, per-call cost of
cdseg_block_forward, and the GPU-idle gaps seen by CUDA events around each stage."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cdsegnet_b200 as cb
from cdsegnet_b200 import configs, ops, ptv3

dev = torch.device("cuda", 0)
torch.manual_seed(0)
seg = cb.build_model(configs.segmentor_cfg()); bench.random_weights(seg); seg = seg.to(dev).eval()
seg.backbone.attention_mode = sys.argv[1] if len(sys.argv) > 1 else "tc32"
sc = bench.make_scene(0)
res = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in sc.items()}
res["grid_coord"] = res["grid_coord"].int()
noise = torch.randn(len(sc["coord"]), 6, device=dev)
for _ in range(3):
    seg.inference(res, eval=False, noise=noise)
torch.cuda.synchronize()

# 1) enqueue vs drain
for overlap in (True, False):
    seg.backbone.overlap_streams = overlap
    seg.inference(res, eval=False, noise=noise); torch.cuda.synchronize()
    t_enq, t_all = [], []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seg.inference(res, eval=False, noise=noise)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_enq.append(t1 - t0); t_all.append(t2 - t0)
    print(f"overlap_streams={overlap}: host enqueue {1e3*np.median(t_enq):.2f} ms, enqueue+drain {1e3*np.median(t_all):.2f} ms per forward")
seg.backbone.overlap_streams = True

# 2) time of Block._native per call (host side only) and of the Plan
orig = ptv3.Block._native
acc = {"t": 0.0, "n": 0}
def timed_native(self, *a, **k):
    t0 = time.perf_counter(); r = orig(self, *a, **k); acc["t"] += time.perf_counter() - t0; acc["n"] += 1; return r
ptv3.Block._native = timed_native
from cdsegnet_b200 import structure
origP = structure.Plan.__init__
accP = {"t": 0.0}
def timed_plan(self, *a, **k):
    t0 = time.perf_counter(); origP(self, *a, **k); accP["t"] += time.perf_counter() - t0
structure.Plan.__init__ = timed_plan
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    seg.inference(res, eval=False, noise=noise)
t_host = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
print(f"host per forward {1e3*t_host:.2f} ms: Plan.__init__ (incl. its 2 syncs) {1e3*accP['t']/5:.2f} ms, Block._native {1e3*acc['t']/5:.2f} ms over {acc['n']//5} calls "
      f"({1e6*acc['t']/max(acc['n'], 1):.1f} us per block call), everything else {1e3*(t_host - accP['t']/5 - acc['t']/5):.2f} ms")
ptv3.Block._native = orig
structure.Plan.__init__ = origP

# 3) cProfile of one forward
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable(); seg.inference(res, eval=False, noise=noise); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])

# 4) host timeline of one forward: when each C-ABI call starts / returns, relative to the start of seg.inference
from cdsegnet_b200 import _lib
lib = _lib.load()
marks = []
def wrap(name):
    fn = getattr(lib, name)
    def w(*a):
        t0 = time.perf_counter(); r = fn(*a); marks.append((name, t0, time.perf_counter())); return r
    setattr(lib, name, w)
for nm in ("cdseg_plan_build", "cdseg_net_arena_bytes", "cdseg_net_forward"):
    wrap(nm)
rows = []
for _ in range(6):
    torch.cuda.synchronize()
    marks.clear()
    t0 = time.perf_counter()
    seg.inference(res, eval=False, noise=noise)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rows.append([(n, 1e6 * (a - t0), 1e6 * (b - t0)) for n, a, b in marks] + [("inference returns", 1e6 * (t1 - t0), 1e6 * (t2 - t0))])
med = rows[-1]
print("host timeline of the last forward (us from the start of seg.inference): call, enter, return")
for n, a, b in med:
    print(f"  {n:24s} {a:9.1f} {b:9.1f}")
