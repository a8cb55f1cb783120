"""Golden fixtures for the DefaultSegmentorV2 wrapper (§8 row a17) made by EXECUTING THE REFERENCE'S OWN SOURCES on CPU:
  pointcept/models/default.py          (DefaultSegmentorV2: diffusion schedule, q/p samplers, inference, inference_ddim, forward)
  pointcept/models/losses/{builder,misc,lovasz}.py   (Criteria EW / GLS, MSELoss, CrossEntropyLoss, LovaszLoss)
on top of the shimmed backbone of make_golden.py.  The reference calls `.cuda()` on every tensor it creates; here
`torch.Tensor.cuda` is patched to the identity so the very same statements run on the host (the random draws come from
torch's CPU generator in the reference as well: `torch.normal(...).cuda()`, default.py:393, 455, 462).

Run once in the authoring container:   python tests/golden/make_golden_wrapper.py      -> wrapper.npz
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
from oracle.weights import synth_state_dict   # noqa: E402
from cdsegnet_b200 import synth                # noqa: E402

CRITERIA = [dict(type="MSELoss", loss_weight=1.0, ignore_index=-1, batch_sample_point=-1),
            dict(type="CrossEntropyLoss", loss_weight=1.0, ignore_index=-1),
            dict(type="LovaszLoss", mode="multiclass", loss_weight=1.0, ignore_index=-1)]


def main():
    ptv3, comm, ser = MG.load_reference()
    torch.Tensor.cuda = lambda self, *a, **k: self
    default = importlib.import_module("pointcept.models.default")
    losses = importlib.import_module("pointcept.models.losses")
    Seg = default.DefaultSegmentorV2
    rec = {}

    # ---- diffusion schedules (default.py:75-189) ---------------------------------------------------------------
    for sched, T, b0, b1 in (("linear", 1000, 1e-4, 0.02), ("cosine", 1000, 1e-4, 0.02), ("sigmoid", 1000, 1e-4, 0.02),
                             ("linear", 50, 1e-4, 0.02), ("cosine", 20, 1e-4, 0.02),
                             ("cosine", 1000, 0, 1000)):                   # the shipped config (configs/scannet/CDSegNet.py:23-27)
        # ("laplace" raises inside the reference itself: torch.cat of 0-d tensors, default.py:184)
        m = Seg.__new__(Seg)
        Beta, Alpha, Alpha_bar, Sigma, SNR = Seg.get_diffusion_hyperparams(m, noise_schedule=sched, T=T, beta_start=b0, beta_end=b1)
        for nme, v in (("Beta", Beta), ("Alpha", Alpha), ("Alpha_bar", Alpha_bar), ("Sigma", Sigma), ("SNR", SNR)):
            rec[f"sched_{sched}_{T}_{b1:g}_{nme}"] = v.numpy()
    for T, step in ((1000, 1), (1000, 5), (1000, 20), (50, 7)):
        rec[f"times_{T}_{step}"] = np.ascontiguousarray(Seg.get_time_schedule(None, T, step))

    # ---- q_sample / p_ddim_sample on fixed inputs (default.py:192-222) -------------------------------------------
    g = torch.Generator().manual_seed(3)
    x0, eps_, pred = (torch.randn(500, 6, generator=g) for _ in range(3))
    for target in ("noise", "x0"):
        m = Seg.__new__(Seg)
        torch.nn.Module.__init__(m)
        m.dm_target = target
        m.Beta, m.Alpha, m.Alpha_bar, m.Sigma, m.SNR = (v.float() for v in Seg.get_diffusion_hyperparams(
            m, noise_schedule="cosine", T=1000, beta_start=0, beta_end=1000))
        for tv in (0, 1, 499, 999):
            ts = tv * torch.ones((500, 1), dtype=torch.int64)
            rec[f"qsample_{tv}"] = m.continuous_q_sample(x0, ts, eps_).numpy()
            rec[f"pddim_{target}_{tv}"] = m.continuous_p_ddim_sample(x0, ts, pred).numpy()
    rec["samp_x0"], rec["samp_eps"], rec["samp_pred"] = x0.numpy(), eps_.numpy(), pred.numpy()

    # ---- criteria on random predictions (losses/builder.py:14-51, misc.py:24-129, lovasz.py) ------------------------
    for i, (n, C, p_ignore, n_absent) in enumerate(((4000, 20, 0.1, 3), (1, 20, 0.0, 0), (777, 13, 0.5, 6), (300, 20, 0.97, 0))):  # (all-ignored batches are degenerate in the reference: CE = nan, Lovasz returns an empty tensor)
        g = torch.Generator().manual_seed(10 + i)
        logits = 3.0 * torch.randn(n, C, generator=g)
        labels = torch.randint(0, C - n_absent, (n,), generator=g)
        labels[torch.rand(n, generator=g) < p_ignore] = -1
        c_pred, c_target = torch.randn(n, 6, generator=g), torch.randn(n, 6, generator=g)
        rec[f"loss{i}_logits"], rec[f"loss{i}_labels"] = logits.numpy(), labels.numpy()
        rec[f"loss{i}_c_pred"], rec[f"loss{i}_c_target"] = c_pred.numpy(), c_target.numpy()
        for lt, mode in (("EW", "train"), ("GLS", "train"), ("GLS", "eval")):
            crit = losses.build_criteria(CRITERIA, loss_type=lt, task_num=2)
            point = dict(n_pred=logits, n_target=labels, c_pred=c_pred, c_target=c_target, loss_mode=mode)
            parts = [float(c(point)) for c in crit.criteria]
            rec[f"loss{i}_parts"] = np.array(parts, dtype=np.float64)
            val = crit(dict(point))
            rec[f"loss{i}_{lt}_{mode}"] = np.float64(float(val))
        crit = losses.build_criteria(CRITERIA[1:], loss_type="EW", task_num=2)      # eval pass without the diffusion branch
        rec[f"loss{i}_noc"] = np.float64(float(crit(dict(n_pred=logits, n_target=labels, loss_mode="eval"))))
    rec["n_loss_cases"] = 4

    # ---- the whole wrapper on the small dual network ------------------------------------------------------------
    scene = synth.collate([synth.scannet_scene(2600, 6, room_m=(3.0, 2.4, 1.6), n_boxes=3),
                           synth.scannet_scene(1900, 7, room_m=(2.4, 2.4, 1.6), n_boxes=2)])
    cfg = dict(MG.SMALL_CFG)
    ps = MG.patch_sizes_for(scene, 64)
    cfg.update(n_enc_patch_size=tuple(ps), n_dec_patch_size=tuple(ps[:4]), c_enc_patch_size=(ps[0], ps[2], ps[4]),
               c_dec_patch_size=(ps[0], ps[2]))
    N = len(scene["coord"])
    seg_rng = np.random.default_rng(5)
    segment = seg_rng.integers(0, 17, size=N)
    segment[seg_rng.random(N) < 0.08] = -1
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    wrapper_kw = dict(criteria=CRITERIA, loss_type="GLS", task_num=2, num_classes=20, T=1000, beta_start=0, beta_end=1000,
                      noise_schedule="cosine", T_dim=128, dm=True, dm_input="xt", dm_target="noise", dm_min_snr=None,
                      condition=True, c_in_channels=6)
    torch.manual_seed(0)
    model = Seg(backbone=dict(type="PT-v3m1", **cfg), **wrapper_kw).eval()
    shapes = {k: list(v.shape) for k, v in model.backbone.state_dict().items()}
    model.backbone.load_state_dict(synth_state_dict(shapes), strict=True)

    def inputs():
        return dict(coord=t(scene["coord"]), grid_coord=t(scene["grid_coord"]).long(), offset=t(scene["offset"]),
                    feat=t(scene["feat"]), segment=t(segment))
    with torch.no_grad():
        torch.manual_seed(111)
        out = model.inference(inputs(), eval=True)
        rec["w_inference_logits"], rec["w_inference_loss"] = out["seg_logits"].numpy(), np.float64(float(out["loss"]))
        torch.manual_seed(222)
        out = model.inference(inputs(), eval=True, noise_level=0.05)
        rec["w_inference_nl_logits"], rec["w_inference_nl_loss"] = out["seg_logits"].numpy(), np.float64(float(out["loss"]))
        for mode in ("avg", "final"):
            torch.manual_seed(333)
            out = model.inference_ddim(inputs(), T=1000, step=3, report=100, eval=True, mode=mode)
            rec[f"w_ddim_{mode}_logits"], rec[f"w_ddim_{mode}_loss"] = out["seg_logits"].numpy(), np.float64(float(out["loss"]))
        torch.manual_seed(444)
        out = model(inputs())                      # training-mode criteria (GLS) on the eval-mode network
        rec["w_forward_loss"] = np.float64(float(out["loss"]))
    rec.update({f"w_{k}": v for k, v in scene.items()})
    rec["w_segment"] = segment
    np.savez_compressed(os.path.join(HERE, "wrapper.npz"), **rec)
    jcfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
    with open(os.path.join(HERE, "wrapper.json"), "w") as f:
        json.dump(dict(cfg=jcfg, shapes=shapes, wrapper={k: v for k, v in wrapper_kw.items()}), f, indent=0)
    for k in sorted(rec):
        if k.startswith("w_") and k.endswith("loss"):
            print(k, rec[k])
    print("loss cases:", [float(rec[f"loss{i}_GLS_train"]) for i in range(4)])


if __name__ == "__main__":
    main()
