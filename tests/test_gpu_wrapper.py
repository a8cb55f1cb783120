"""GPU parity (-m gpu) of the DefaultSegmentorV2 wrapper row (§8 a17): criteria kernels, diffusion samplers and the three
entry points (inference eval=True, inference_ddim, forward) against the outputs of the reference's own default.py and
losses/*.py (tests/golden/wrapper.npz) and the CPU oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import load_case, replay, t
from oracle import wrapper_oracle as W
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"
Z = np.load(os.path.join(GOLDEN, "wrapper.npz"))
J = json.load(open(os.path.join(GOLDEN, "wrapper.json")))


@pytest.fixture(scope="module")
def ops(lib):
    from cdsegnet_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("i", range(4))
def test_criteria_kernels_vs_reference(ops, i):
    """CE + Lovasz (radix sort per class) + masked MSE + EW / GLS in one call == the reference's Criteria (1e-5 relative)"""
    g = lambda k: torch.from_numpy(Z[f"loss{i}_{k}"]).to(DEV)
    out = ops.criteria(g("logits"), g("labels"), -1, g("c_pred"), g("c_target")).cpu().numpy().astype(np.float64)
    ref = Z[f"loss{i}_parts"]
    assert np.allclose(out[:3], ref, rtol=2e-5, atol=1e-6), (out, ref)
    assert abs(out[3] - float(Z[f"loss{i}_EW_train"])) < 2e-5 * abs(out[3])
    assert abs(out[3] - float(Z[f"loss{i}_GLS_eval"])) < 2e-5 * abs(out[3])
    assert abs(out[4] - float(Z[f"loss{i}_GLS_train"])) < 2e-5 * abs(out[4])
    noc = ops.criteria(g("logits"), g("labels"), -1).cpu().numpy()
    assert noc[0] == 0.0 and abs(noc[3] - float(Z[f"loss{i}_noc"])) < 2e-5 * abs(noc[3])


@pytest.mark.parametrize("n,C", [(120000, 20), (50000, 200), (33, 16)])
def test_criteria_kernels_vs_oracle_full_size(ops, n, C):
    """BASELINE sizes (120k points x 20 classes; ScanNet200's 200 classes) against the CPU oracle"""
    g = torch.Generator().manual_seed(n + C)
    logits = 2.0 * torch.randn(n, C, generator=g)
    labels = torch.randint(0, max(1, C - 2), (n,), generator=g)
    labels[torch.rand(n, generator=g) < 0.07] = -1
    cp, ct = torch.randn(n, 6, generator=g), torch.randn(n, 6, generator=g)
    loss, parts = W.criteria(dict(n_pred=logits, n_target=labels, c_pred=cp, c_target=ct, loss_mode="train"), "GLS", 2)
    out = ops.criteria(logits.to(DEV), labels.to(DEV), -1, cp.to(DEV), ct.to(DEV)).cpu().numpy()
    assert np.allclose(out[:3], [float(p) for p in parts], rtol=5e-5), (out, parts)
    assert abs(out[4] - float(loss)) < 5e-5 * abs(float(loss))


def test_criteria_weights_and_partial_sets(ops):
    g = lambda k: torch.from_numpy(Z[f"loss0_{k}"]).to(DEV)
    base = Z["loss0_parts"]
    out = ops.criteria(g("logits"), g("labels"), -1, g("c_pred"), g("c_target"), weights=(0.5, 2.0, 3.0)).cpu().numpy()
    assert np.allclose(out[:3], base * [0.5, 2.0, 3.0], rtol=2e-5)
    out = ops.criteria(g("logits"), g("labels"), -1, g("c_pred"), g("c_target"), has=(True, True, False)).cpu().numpy()
    assert out[2] == 0.0 and np.allclose(out[:2], base[:2], rtol=2e-5)
    out = ops.criteria(g("logits"), g("labels"), -1, g("c_pred"), g("c_target"), mse_use_ignore=False).cpu().numpy()
    ref = float(((torch.from_numpy(Z["loss0_c_pred"]) - torch.from_numpy(Z["loss0_c_target"])) ** 2).mean())
    assert abs(out[0] - ref) < 2e-5 * ref


def test_samplers_vs_reference(ops):
    """q_sample / DDIM update kernels are bit-exact against default.py:192-222 (fp32, round-to-nearest at every step)"""
    import cdsegnet_b200 as cb
    x0, eps, pred = (torch.from_numpy(Z[k]).to(DEV) for k in ("samp_x0", "samp_eps", "samp_pred"))
    for target in ("noise", "x0"):
        seg = cb.DefaultSegmentorV2(backbone=None, dm=True, dm_target=target, noise_schedule="cosine", beta_start=0, beta_end=1000)
        for tv in (0, 1, 499, 999):
            ts = tv * torch.ones((500, 1), dtype=torch.int64)
            q = seg.continuous_q_sample(x0, ts, eps).cpu().numpy()
            assert np.array_equal(q, Z[f"qsample_{tv}"], equal_nan=True), tv
            p = seg.continuous_p_ddim_sample(x0, ts, pred).cpu().numpy()
            assert np.array_equal(p, Z[f"pddim_{target}_{tv}"], equal_nan=True), (target, tv)


def _wrapper():
    import cdsegnet_b200 as cb
    seg = cb.build_model(dict(type="DefaultSegmentorV2", backbone=dict(type="PT-v3m1", **dict(J["cfg"], enable_flash=False)),
                              **J["wrapper"]))
    seg.backbone.load_state_dict(synth_state_dict(J["shapes"]), strict=True)
    return seg.to(DEV).eval()


def _inputs():
    return dict(coord=t(Z["w_coord"]).to(DEV), grid_coord=t(Z["w_grid_coord"]).to(DEV), offset=t(Z["w_offset"]).to(DEV),
                feat=t(Z["w_feat"]).to(DEV), segment=t(Z["w_segment"]).to(DEV))


def _check(out, key, tol=1e-3):
    logits = out["seg_logits"].cpu().numpy()
    assert np.abs(logits - Z[f"w_{key}_logits"]).max() < tol, key
    assert abs(float(out["loss"]) - float(Z[f"w_{key}_loss"])) < tol, key


def test_wrapper_inference_eval_vs_reference():
    """inference(eval=True): logits within 1e-3 and the evaluator's loss (CE + Lovasz) within 1e-3 of the reference run;
    the Noise-Network input is DRAWN (CPU generator, default.py:393), not injected"""
    seg = _wrapper()
    torch.manual_seed(111)
    _check(seg.inference(_inputs(), eval=True), "inference")


def test_wrapper_inference_noise_level_vs_reference(monkeypatch):
    seg = _wrapper()
    # the reference perturbs the input with torch.randn_like on the tensor's own device; the golden run drew it on the host
    monkeypatch.setattr(torch, "randn_like", lambda x: torch.randn(x.shape).to(x.device))
    torch.manual_seed(222)
    _check(seg.inference(_inputs(), eval=True, noise_level=0.05), "inference_nl")


@pytest.mark.parametrize("mode", ("avg", "final"))
def test_wrapper_inference_ddim_vs_reference(mode):
    """3-step DDIM (4 backbone passes, the last one at t = -1 like the reference) -- logits and loss within 2e-3"""
    seg = _wrapper()
    torch.manual_seed(333)
    _check(seg.inference_ddim(_inputs(), T=1000, step=3, report=100, eval=True, mode=mode), f"ddim_{mode}", tol=2e-3)


def test_wrapper_forward_criteria_vs_reference():
    """forward(): per-scene random timestep, q-sampled NN input, GLS criteria -- loss within 1e-3 of the reference run"""
    seg = _wrapper()
    torch.manual_seed(444)
    loss = float(seg(_inputs())["loss"])
    assert abs(loss - float(Z["w_forward_loss"])) < 1e-3, loss
    seg.train()
    with pytest.raises(NotImplementedError):
        seg(_inputs())


# ------------------------------------------------------------------ training side: criteria gradients + optimizer (§8(f) rank 2)
@pytest.mark.parametrize("i", range(4))
@pytest.mark.parametrize("gls", (False, True))
def test_criteria_gradients_vs_autograd(ops, i, gls):
    """d loss / d logits and d loss / d c_pred of the EW sum and of the GLS loss == torch autograd through the oracle's restatement of the
    reference criteria (fp32; Lovasz gradient = Jaccard gradient at the sorted position, pulled through the softmax)"""
    g = lambda k: torch.from_numpy(Z[f"loss{i}_{k}"])
    logits, cp = g("logits").clone().requires_grad_(True), g("c_pred").clone().requires_grad_(True)
    loss, _ = W.criteria(dict(n_pred=logits, n_target=g("labels"), c_pred=cp, c_target=g("c_target"), loss_mode="train"), "GLS" if gls else "EW", 2)
    loss.backward()
    out, gn, gc = ops.criteria_grad(g("logits").to(DEV), g("labels").to(DEV), -1, g("c_pred").to(DEV), g("c_target").to(DEV), gls=gls)
    assert abs(float(out[4 if gls else 3]) - float(loss.detach())) < 2e-5 * abs(float(loss.detach()))
    for got, ref in ((gn, logits.grad), (gc, cp.grad)):
        scale = float(ref.abs().max())
        assert float((got.cpu() - ref).abs().max()) < 1e-4 * scale + 1e-9, (float((got.cpu() - ref).abs().max()), scale)


def test_criteria_value_and_grad_api(ops):
    from cdsegnet_b200.losses import build_criteria
    crit = build_criteria(J["wrapper"]["criteria"], "GLS", 2)
    g = lambda k: torch.from_numpy(Z[f"loss0_{k}"]).to(DEV)
    point = dict(n_pred=g("logits"), n_target=g("labels"), c_pred=g("c_pred"), c_target=g("c_target"), loss_mode="train")
    loss, gn, gc = crit.value_and_grad(point)
    assert abs(float(loss) - float(Z["loss0_GLS_train"])) < 2e-5 * float(loss) and gn.shape == point["n_pred"].shape and gc.shape == point["c_pred"].shape
    loss_e, gn_e, _ = crit.value_and_grad(dict(point, loss_mode="eval"))
    assert abs(float(loss_e) - float(Z["loss0_GLS_eval"])) < 2e-5 * float(loss_e)
    assert float(gn[g("labels") == -1].abs().max()) == 0.0                    # ignored points carry no gradient


def test_fused_adamw_vs_torch_and_param_groups():
    """build_optimizer: keyword groups like utils/optimizer.py:20-55; FusedAdamW (one launch per group) == torch.optim.AdamW over several
    steps with changing lr / betas (what OneCycleLR does to the groups)"""
    import cdsegnet_b200 as cb
    from cdsegnet_b200.optim import build_optimizer, FusedAdamW
    torch.manual_seed(0)
    m = cb.PointTransformerV3(**dict(J["cfg"])).to(DEV)
    ref = {n: p.detach().cpu().clone().requires_grad_(True) for n, p in m.named_parameters()}
    opt = build_optimizer(dict(type="AdamW", lr=0.002, weight_decay=0.05), m, [dict(keyword="block", lr=0.0002)])
    assert isinstance(opt, FusedAdamW) and len(opt.param_groups) == 2
    names1 = [n for n, _ in m.named_parameters() if "block" in n]
    assert len(opt.param_groups[1]["params"]) == len(names1) > 0 and opt.param_groups[1]["lr"] == 0.0002 and opt.param_groups[0]["lr"] == 0.002
    topt = torch.optim.AdamW([dict(params=[ref[n] for n, _ in m.named_parameters() if "block" not in n], lr=0.002),
                              dict(params=[ref[n] for n in names1], lr=0.0002)], lr=0.002, weight_decay=0.05)
    gen = torch.Generator().manual_seed(1)
    for it in range(4):
        for n, p in m.named_parameters():
            gr = torch.randn(p.shape, generator=gen) * 0.1
            p.grad = gr.to(DEV)
            ref[n].grad = gr.clone()
        for o in (opt, topt):
            for gi, grp in enumerate(o.param_groups):                                  # a scheduler moving lr and beta1 between steps
                grp["lr"] = (0.002 if gi == 0 else 0.0002) * (1 + 0.3 * it)
                grp["betas"] = (0.9 - 0.01 * it, 0.999)
        opt.step(); topt.step()
    torch.cuda.synchronize()
    worst = max(float((p.detach().cpu() - ref[n]).abs().max()) for n, p in m.named_parameters())
    assert worst < 2e-6, worst


def test_optimizer_step_invalidates_packed_weight_caches():
    """FusedAdamW writes parameters through raw pointers; the forward's packed tensor-core operand caches are keyed on
    Parameter._version, so a step must bump it: step -> forward must equal the forward of a freshly built model holding the
    updated weights (stale packed weights after optimizer.step())."""
    import copy
    import cdsegnet_b200 as cb
    from cdsegnet_b200.optim import build_optimizer
    z, cfg, shapes = load_case("case3_cn_only")
    m = cb.PointTransformerV3(**dict(cfg, enable_flash=False))
    m.load_state_dict(synth_state_dict(shapes), strict=True)
    m = m.to(DEV).eval()
    base = dict(coord=t(z["coord"]).to(DEV), grid_coord=t(z["grid_coord"]).to(DEV), offset=t(z["offset"]).to(DEV))

    def fwd(model):
        out = model(n_point=dict(base, feat=t(z["feat"]).to(DEV)), perm_fn=replay(z["perms"]))["feat"]
        torch.cuda.synchronize()
        return out.cpu().numpy()
    before = fwd(m)                                       # fills the pack caches
    opt = build_optimizer(dict(type="AdamW", lr=0.05, weight_decay=0.0), m, [dict(keyword="block", lr=0.02)])
    g = torch.Generator().manual_seed(3)
    for p in m.parameters():
        p.grad = torch.randn(p.shape, generator=g).to(DEV)
    opt.step()
    after = fwd(m)
    fresh = cb.PointTransformerV3(**dict(cfg, enable_flash=False))
    fresh.load_state_dict({k: v.detach().cpu() for k, v in m.state_dict().items()}, strict=True)
    ref = fwd(fresh.to(DEV).eval())
    assert np.abs(after - before).max() > 1e-2            # the step changed the network ...
    assert np.abs(after - ref).max() < 1e-5               # ... and the forward saw the new weights


def test_fused_adamw_lagging_parameters_and_reloaded_state():
    """parameters whose gradient is None on some iterations keep their own step count (per-parameter bias correction, like
    torch.optim.AdamW), and load_state_dict -- which replaces the moment tensors -- must not leave a stale chunk table behind"""
    from cdsegnet_b200.optim import FusedAdamW
    torch.manual_seed(3)
    ps = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in ((70000,), (33, 17), (5,))]
    ref = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ps]
    opt = FusedAdamW(ps, lr=0.01, weight_decay=0.02)
    topt = torch.optim.AdamW(ref, lr=0.01, weight_decay=0.02)
    gen = torch.Generator().manual_seed(4)
    for it in range(6):
        if it == 3:                                            # checkpoint round trip: new exp_avg / exp_avg_sq tensors
            opt.load_state_dict(opt.state_dict())
        for i, (p, r) in enumerate(zip(ps, ref)):
            if i == 1 and it in (1, 2, 4):                     # this parameter sits out three iterations
                p.grad = None; r.grad = None
                continue
            g = torch.randn(p.shape, generator=gen)
            p.grad = g.to(DEV); r.grad = g.clone()
        opt.step(); topt.step()
    torch.cuda.synchronize()
    for p, r in zip(ps, ref):
        assert float((p.detach().cpu() - r.detach()).abs().max()) < 2e-6
