"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel shares of ONE forward (the last
one in the file, delimited by encode_kernel launches) and the largest (kernel, grid) groups.
  python profiles/launch_summary.py gpurun_out/launches3.csv [top_groups]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
idx = [i for i, x in enumerate(rows) if "encode_kernel" in x["Kernel Name"]]
fw = rows[idx[-1]:]
if len(idx) >= 2 and len(fw) < idx[-1] - idx[-2]:      # capture cut off mid-forward (-c limit): take the last complete one
    fw = rows[idx[-2]:idx[-1]]
tot = sum(float(x["Metric Value"]) for x in fw) / 1e3
print(f"last forward: {len(fw)} launches, {tot:.1f} us serialised")
name = lambda k: re.sub(r"\(.*", "", k).replace("void ", "")[:44]
byk = collections.defaultdict(lambda: [0, 0.0])
byg = collections.defaultdict(lambda: [0, 0.0])
for x in fw:
    t = float(x["Metric Value"]) / 1e3
    k = name(x["Kernel Name"])
    byk[k][0] += 1; byk[k][1] += t
    byg[(k, x["Grid Size"])][0] += 1; byg[(k, x["Grid Size"])][1] += t
print("\n| kernel | launches | us | share |\n|---|---|---|---|")
for k, v in sorted(byk.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
print("\n| kernel | grid | launches | us | avg us |\n|---|---|---|---|---|")
for k, v in sorted(byg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"| {k[0]} | {k[1]} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} |")
