"""Golden fixture for the TRAINING pass made by EXECUTING THE REFERENCE'S OWN SOURCES on CPU:
`DefaultSegmentorV2.forward` (pointcept/models/default.py:424-493) in train mode -- batch-statistics BatchNorm, GLS criteria -- followed
by `loss.backward()`; the loss, the random draws and the parameter gradients the reference's autograd produces are recorded.
DropPath is configured off (drop_path=0.0) so that the pass is a deterministic function of the recorded draws.
Third-party ops are the shims of make_golden.py (their BACKWARD is torch autograd through the shim: unpinned like their forward).

Run once in the authoring container:   python tests/golden/make_golden_train.py      -> train.npz / train.json
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as MG                       # noqa: E402
from make_golden_wrapper import CRITERIA       # noqa: E402
from oracle.weights import synth_state_dict   # noqa: E402
from cdsegnet_b200 import synth                # noqa: E402

FULL = ("_n_head.weight", "_n_head.bias", "_c_head.weight", "fc_t1.weight", "_n_embedding.stem.conv.weight", "_n_embedding.stem.norm.weight",
        "_n_enc.enc0.block0.attn.qkv.weight", "_n_enc.enc4.block1.mlp.0.fc2.weight", "_n_enc.enc2.down.norm.0.bias",
        "_c_enc.enc1.block0.t_mlp.weight", "_c_dec.dec0.up.proj_cat.0.weight", "_n_dec.dec1.block0.cpe.0.weight", "_tm_dec0.cross_block2.attn.kv.weight")


def main():
    ptv3, comm, ser = MG.load_reference()
    # the timm shim of make_golden.py refuses train mode; with p = 0 DropPath is the identity in both modes
    sys.modules["timm.models.layers"].DropPath.forward = lambda self, x: x
    torch.Tensor.cuda = lambda self, *a, **k: self
    default = importlib.import_module("pointcept.models.default")
    Seg = default.DefaultSegmentorV2
    scene = synth.collate([synth.scannet_scene(1500, 16, room_m=(2.4, 2.0, 1.6), n_boxes=2),
                           synth.scannet_scene(1100, 17, room_m=(2.0, 2.0, 1.6), n_boxes=2)])
    cfg = dict(MG.SMALL_CFG, drop_path=0.0)
    ps = MG.patch_sizes_for(scene, 64)
    cfg.update(n_enc_patch_size=tuple(ps), n_dec_patch_size=tuple(ps[:4]), c_enc_patch_size=(ps[0], ps[2], ps[4]),
               c_dec_patch_size=(ps[0], ps[2]))
    N = len(scene["coord"])
    rng = np.random.default_rng(9)
    segment = rng.integers(0, 15, size=N)
    segment[rng.random(N) < 0.06] = -1
    wrapper_kw = dict(criteria=CRITERIA, loss_type="GLS", task_num=2, num_classes=20, T=1000, beta_start=0, beta_end=1000,
                      noise_schedule="cosine", T_dim=128, dm=True, dm_input="xt", dm_target="noise", dm_min_snr=None,
                      condition=True, c_in_channels=6)
    torch.manual_seed(0)
    model = Seg(backbone=dict(type="PT-v3m1", **cfg), **wrapper_kw)
    shapes = {k: list(v.shape) for k, v in model.backbone.state_dict().items()}
    model.backbone.load_state_dict(synth_state_dict(shapes), strict=True)
    model.train()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    inp = dict(coord=t(scene["coord"]), grid_coord=t(scene["grid_coord"]).long(), offset=t(scene["offset"]), feat=t(scene["feat"]),
               segment=t(segment))
    # record the draws of the pass: torch.randint (timesteps), torch.normal (noise), then the poolings' torch.randperm
    rec, perms = {}, []
    real_randint, real_normal, real_randperm = torch.randint, torch.normal, torch.randperm

    def randint(*a, **k):
        v = real_randint(*a, **k); rec["ts"] = v.numpy().copy(); return v

    def normal(*a, **k):
        v = real_normal(*a, **k); rec["noise"] = v.numpy().copy(); return v

    def randperm(n, *a, **k):
        p = real_randperm(n, *a, **k); perms.append(p.numpy().copy()); return p
    torch.manual_seed(555)
    torch.randint, torch.normal, torch.randperm = randint, normal, randperm
    try:
        loss = model(inp)["loss"]
    finally:
        torch.randint, torch.normal, torch.randperm = real_randint, real_normal, real_randperm
    loss.backward()
    rec["loss"] = np.float64(float(loss))
    rec["perms"] = np.stack(perms)
    names, norms = [], []
    for n, p in model.backbone.named_parameters():
        names.append(n)
        norms.append(float(p.grad.norm()) if p.grad is not None else -1.0)
        if n in FULL:
            rec["grad__" + n] = p.grad.numpy().copy()
    rec["grad_norms"] = np.array(norms, dtype=np.float64)
    rec.update({k: v for k, v in scene.items()})
    rec["segment"] = segment
    np.savez_compressed(os.path.join(HERE, "train.npz"), **rec)
    jcfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
    with open(os.path.join(HERE, "train.json"), "w") as f:
        json.dump(dict(cfg=jcfg, shapes=shapes, wrapper=wrapper_kw, grad_names=names), f, indent=0)
    print("loss", rec["loss"], "ts", rec["ts"].ravel(), "params", len(names), "without grad", sum(1 for v in norms if v < 0),
          "max grad norm", max(norms))


if __name__ == "__main__":
    main()
