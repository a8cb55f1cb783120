"""dram__bytes_read.sum + dram__bytes_write.sum (and duration) of selected kernels from .ncu-rep files -> profiles/r02_ncu_traffic.json,
which bench.py reads for the `traffic` fields.   python profiles/ncu_traffic.py"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr = r[0]
    return [dict(zip(hdr, x)) for x in r[2:]]
def traffic(row):
    f = lambda k: float(row[k].replace(",", ""))
    units = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
res = {"source": "ncu --set full --clock-control none, round 2 (profiles/r02_ncu_*.md)"}
def pick(rep, name_part, key):
    p = os.path.join(ROOT, "gpurun_out", rep)
    if not os.path.exists(p):
        return
    out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    for x in r[2:]:
        d = dict(zip(hdr, x))
        if name_part in d.get("Kernel Name", ""):
            u = dict(zip(hdr, units))
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
            res[key] = rd + wr
            return
pick("r02_attn_f16.ncu-rep", "attn_tc3", "attn_f16")
pick("r02_attn_tc32.ncu-rep", "attn_tc3", "attn_tc32")
pick("r02_pre.ncu-rep", "pre_kernel", "pre")
pick("r02_post.ncu-rep", "post_kernel", "post")
json.dump(res, open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
print(res)
