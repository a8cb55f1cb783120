"""CPU emulation of candidate attention numerics through the oracle (which operands must be split into fp16 hi + lo halves for the
forward to stay within 1e-3 of the dense fp32 branch?): tc32 = q, k, v split, P single fp16; tc32p = P split as well (what
attn_tc3.cu mode 1 implements); tc32v16 / tc32qk16 = v / q,k left in single fp16.   python profiles/attention_numerics.py"""
import sys, time, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import ptv3_oracle as O
from helpers import load_case, oracle_forward
def split22(x):
    hi = x.half().float(); lo = (x - hi).half().float(); return hi + lo
orig = O._attend
def attend(q, k, v, scale, mode):
    if mode in ("tc32", "tc32p", "tc32v16", "tc32qk16"):
        qq, kk = (split22(q), split22(k)) if mode != "tc32qk16" else (q.half().float(), k.half().float())
        s = (qq @ kk.transpose(-2, -1)) * scale
        p = torch.exp(s - s.amax(-1, keepdim=True))
        p16 = p.half().float()
        if mode == "tc32p":
            p16 = p16 + (p - p16).half().float()
        l = p16.sum(-1, keepdim=True)
        vv = split22(v) if mode != "tc32v16" else v.half().float()
        return (p16 @ vv) / l
    return orig(q, k, v, scale, mode)
O._attend = attend
for name in ["case1_single", "case2_batch2"]:
    z, cfg, shapes = load_case(name)
    _, ref = oracle_forward(z, cfg, shapes, "dense")
    for mode in ["flash16", "tc32", "tc32p", "tc32v16", "tc32qk16"]:
        _, got = oracle_forward(z, cfg, shapes, mode)
        print(name, mode, "max|d| vs dense = %.3e" % np.abs(got - ref).max(), "vs reference golden %.3e" % np.abs(got - z["n_feat"]).max())
