"""In-situ kernel timeline of one forward (round 2, VERDICT item 5: launch-gap analysis).

torch.profiler's CUDA activity (CUPTI concurrent-kernel tracing) records every kernel of the process, the ones launched from
libcdseg_b200.so included, with start time / duration / stream and WITHOUT serialising them (unlike ncu).  From the last profiled
forward this prints, per stream: busy time, idle time between consecutive kernels (histogram), and for the whole step: wall time
from first kernel start to last kernel end, time with 0 / 1 / 2 streams busy, the plan phase's extent and its two host syncs.

  python profiles/timeline_r2.py [tc32|f16] [out.csv]
"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cdsegnet_b200 as cb
from cdsegnet_b200 import configs

dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "tc32"
out = sys.argv[2] if len(sys.argv) > 2 else None
torch.manual_seed(0)
seg = cb.build_model(configs.segmentor_cfg()); bench.random_weights(seg); seg = seg.to(dev).eval()
seg.backbone.attention_mode = mode
sc = bench.make_scene(0)
res = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in sc.items()}
res["grid_coord"] = res["grid_coord"].int()
noise = torch.randn(len(sc["coord"]), 6, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    seg.inference(res, eval=False, noise=noise)
torch.cuda.synchronize()

from torch.profiler import profile, ProfilerActivity
STEPS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        flush.fill_(1)
        torch.cuda.synchronize()
        seg.inference(res, eval=False, noise=noise)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = []
for e in ev:
    tr = e.time_range
    rows.append((tr.start, tr.end, e.name, getattr(e, "stream", None) if hasattr(e, "stream") else None))
# stream ids are not on FunctionEvent in every torch build: fall back to the chrome trace
if not rows or rows[0][3] is None:
    import tempfile
    p = os.path.join(tempfile.mkdtemp(), "t.json")
    prof.export_chrome_trace(p)
    tr = json.load(open(p))["traceEvents"]
    rows = [(t["ts"], t["ts"] + t["dur"], t["name"], t["args"].get("stream", t.get("tid"))) for t in tr
            if t.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in t]
rows.sort()
# split into steps at the L2-flush fill kernels (vectorized_elementwise / FillFunctor)
cuts = [i for i, r in enumerate(rows) if "FillFunctor" in r[2] or "fill" in r[2].lower()]
last = rows[cuts[-1] + 1:] if cuts else rows
t0 = last[0][0]
if out:
    with open(out, "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for s, e, n, st in last:
            f.write(f"{s - t0:.3f},{e - s:.3f},{st},\"{n[:90]}\"\n")
streams = sorted({r[3] for r in last}, key=str)
wall = max(r[1] for r in last) - t0
print(f"mode {mode}: last forward {len(last)} device activities over {wall / 1e3:.3f} ms (first start -> last end), streams {streams}")
bounds = [0, 1, 2, 3, 5, 10, 20, 50, 1e9]
for st in streams:
    rs = [r for r in last if r[3] == st]
    busy = sum(r[1] - r[0] for r in rs)
    gaps = np.array([max(0.0, rs[i + 1][0] - rs[i][1]) for i in range(len(rs) - 1)])
    span = rs[-1][1] - rs[0][0]
    print(f" stream {st}: {len(rs)} activities, span {span / 1e3:.3f} ms, busy {busy / 1e3:.3f} ms, idle between activities {gaps.sum() / 1e3:.3f} ms "
          f"(median gap {np.median(gaps):.2f} us)")
    h = np.histogram(gaps, bins=bounds)[0]
    print("   gap histogram (us): " + ", ".join(f"[{bounds[i]},{bounds[i + 1] if bounds[i + 1] < 1e9 else 'inf'}): {h[i]} = {gaps[(gaps >= bounds[i]) & (gaps < bounds[i + 1])].sum():.0f} us"
                                               for i in range(len(h))))
    big = sorted(((gaps[i], i) for i in range(len(gaps))), reverse=True)[:8]
    for g, i in big:
        print(f"     gap {g:8.1f} us after {rs[i][2][:60]} (t = {(rs[i][1] - t0) / 1e3:.3f} ms) before {rs[i + 1][2][:60]}")
# concurrency profile: time with k streams busy
pts = []
for s, e, n, st in last:
    pts.append((s, 1)); pts.append((e, -1))
pts.sort()
lvl, prev, acc = 0, pts[0][0], {}
for t, d in pts:
    acc[lvl] = acc.get(lvl, 0.0) + (t - prev); prev = t; lvl += d
print(" time with k kernels in flight: " + ", ".join(f"k={k}: {v / 1e3:.3f} ms" for k, v in sorted(acc.items())))
# per-kernel in-situ durations (compare with the ncu-serialised list)
agg = {}
for s, e, n, st in last:
    k = n.split("(")[0][:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e - s
print(" in-situ kernel time by name (top 14):")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"   {t:9.1f} us  {c:4d} x  {k}")
