import json
import os

import numpy as np
import torch

from oracle import ptv3_oracle as O
from oracle.weights import synth_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"ptv3_{name}.npz"))
    j = json.load(open(os.path.join(GOLDEN, f"ptv3_{name}.json")))
    return z, j["cfg"], j["shapes"]


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def replay(perms):
    perms = [np.asarray(p) for p in perms]

    def fn(k):
        return perms.pop(0)
    fn.remaining = perms
    return fn


def oracle_forward(z, cfg, shapes, mode, perms=None, trace=None):
    sd = synth_state_dict(shapes)
    base = dict(coord=t(z["coord"]), grid_coord=t(z["grid_coord"]).long(), offset=t(z["offset"]))
    pf = replay(z["perms"] if perms is None else perms)
    if cfg["condition"]:
        n = len(z["coord"])
        ts = 999 * torch.ones((n, 1), dtype=torch.int64)
        c, nn_ = O.forward(sd, cfg, dict(base, feat=t(z["noise"]), t_emb=O.calc_t_emb(ts, cfg["T_dim"])),
                           dict(base, feat=t(z["feat"])), attn_mode=mode, perm_fn=pf, trace=trace)
        return c["feat"].numpy(), nn_["feat"].numpy()
    nn_ = O.forward(sd, cfg, n_in=dict(base, feat=t(z["feat"])), attn_mode=mode, perm_fn=pf, trace=trace)
    return None, nn_["feat"].numpy()
