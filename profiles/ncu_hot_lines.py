"""Hot source lines of an .ncu-rep captured with --import-source on (kernels built with -lineinfo):
warp-stall samples per CUDA source line.   python profiles/ncu_hot_lines.py rep.ncu-rep [top] [kernel-id]"""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3:
    cmd += ["--launch-skip", sys.argv[3], "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows, fpath, hdr, fn = [], None, None, None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fpath, hdr = os.path.basename(r[1]), None
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
        si = hdr.index("Warp Stall Sampling (All Samples)")
        ii = hdr.index("Instructions Executed")
    elif hdr and r[0].isdigit():
        try:
            rows.append((int(r[si]), int(r[ii]), fpath, int(r[0]), r[1].strip()[:140]))
        except ValueError:
            pass
tot = sum(x[0] for x in rows)
print(f"# {fn[:100] if fn else '?'}\ntotal warp-stall samples {tot}\n\n| samples | warp insts | line | source |\n|---|---|---|---|")
for s, n, f, l, src in sorted(rows, key=lambda x: -x[0])[:top]:
    print(f"| {100 * s / max(tot, 1):.1f}% | {n} | {f}:{l} | `{src}` |")
