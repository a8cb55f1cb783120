"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the DefaultSegmentorV2 wrapper math --
diffusion schedule, samplers and criteria -- of the reference.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it.

Pinned: tests/golden/wrapper.npz holds the outputs of the reference's own pointcept/models/default.py and
pointcept/models/losses/*.py executed on CPU (tests/golden/make_golden_wrapper.py); tests/test_cpu_oracle.py checks
every function below against them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------- schedule
def diffusion_betas(kind="linear", start=1e-4, stop=0.02, T=1000):
    """default.py:127-189 (fp64).  "laplace" raises in the reference itself (torch.cat of 0-d tensors, line 184)."""
    if kind == "linear":
        scale = 1000 / T
        return torch.linspace(scale * start, scale * stop, T, dtype=torch.float64)
    if kind == "cosine":                                   # note: t runs over [start, stop] / T, not [0, 1] (line 145)
        t = torch.linspace(start, stop, T + 1, dtype=torch.float64) / T
        ac = torch.cos((t + 0.008) / 1.008 * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        return torch.clip(1 - ac[1:] / ac[:-1], 0, 0.999)
    if kind == "sigmoid":
        lo, hi = -3, 3
        t = torch.linspace(0, T, T + 1, dtype=torch.float64) / T
        v0, v1 = torch.tensor(float(lo)).sigmoid(), torch.tensor(float(hi)).sigmoid()
        ac = (-((t * (hi - lo) + lo)).sigmoid() + v1) / (v1 - v0)
        ac = ac / ac[0]
        return torch.clip(1 - ac[1:] / ac[:-1], 0, 0.999)
    raise NotImplementedError(kind)


def diffusion_hyperparams(kind, beta_start, beta_end, T):
    """default.py:75-125 -> Beta, Alpha, Alpha_bar, Sigma, SNR (fp64 [T])"""
    beta = diffusion_betas(kind, beta_start, beta_end, T)
    alpha = 1 - beta
    abar = alpha + 0
    bt = beta + 0
    for t in range(1, T):
        abar[t] *= abar[t - 1]
        bt[t] *= (1 - abar[t - 1]) / (1 - abar[t])
    sigma = torch.sqrt(bt)
    sigma[0] = 0.0
    return beta, alpha, abar, sigma, abar / (1 - abar)


def time_schedule(T=1000, step=5):
    """default.py:224-226"""
    return np.linspace(-1, T - 1, num=step + 1, dtype=int)[::-1]


def q_sample(abar32, x0, ts, noise):
    """default.py:216-222; abar32 = Alpha_bar.float(); ts int64 [N,1]"""
    return torch.sqrt(abar32[ts]) * x0 + torch.sqrt(1 - abar32[ts]) * noise


def p_ddim_sample(abar32, x_t, ts, pred, target="noise"):
    """default.py:192-214 (negative timesteps index from the end, like the reference's tensor indexing)"""
    if target == "noise":
        x0 = (x_t - torch.sqrt(1 - abar32[ts]) * pred) / torch.sqrt(abar32[ts])
        eps = pred
    else:
        x0 = pred
        eps = (x_t - torch.sqrt(abar32[ts]) * x0) / torch.sqrt(1 - abar32[ts])
    if ts[0] == 0:
        return x0
    return torch.sqrt(abar32[ts - 1]) * x0 + torch.sqrt(1 - abar32[ts - 1]) * eps


# ---------------------------------------------------------------------------------------------- criteria
def mse_loss(c_pred, c_target, n_target, ignore_index=-1, weight=1.0):
    """misc.py:24-94 with batch_sample_point <= 0 (configs/scannet/CDSegNet.py:118).  `if self.ignore_index:` is a truth
    test: None and 0 both disable the mask; the snr weight is dead code (hasattr on a dict is always False)."""
    if ignore_index:
        valid = n_target != ignore_index
        c_pred, c_target = c_pred[valid], c_target[valid]
    return ((c_pred - c_target) ** 2).mean() * weight


def ce_loss(n_pred, n_target, ignore_index=-1, weight=1.0):
    """misc.py:97-129"""
    if ignore_index:
        valid = n_target != ignore_index
        n_pred, n_target = n_pred[valid], n_target[valid]
    return F.cross_entropy(n_pred, n_target) * weight


def lovasz_grad(gt_sorted):
    """lovasz.py:22-33"""
    p = len(gt_sorted)
    gts = gt_sorted.sum()
    inter = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jac = 1.0 - inter / union
    if p > 1:
        jac[1:p] = jac[1:p] - jac[0:-1]
    return jac


def lovasz_loss(n_pred, n_target, ignore_index=-1, weight=1.0):
    """lovasz.py:244-272 mode="multiclass" -> 89-165: softmax, drop ignored points, per class PRESENT among the labels the
    Lovasz extension of the Jaccard loss, mean over those classes"""
    prob = n_pred.softmax(dim=1)
    if ignore_index is not None:
        valid = n_target != ignore_index
        prob, n_target = prob[valid], n_target[valid]
    if prob.numel() == 0:
        return prob * 0.0
    losses = []
    for c in n_target.unique():
        fg = (n_target == c).type_as(prob)
        err = (fg - prob[:, c]).abs()
        err_sorted, perm = torch.sort(err, 0, descending=True)
        losses.append(torch.dot(err_sorted, lovasz_grad(fg[perm])))
    return sum(losses) / len(losses) * weight


def criteria(point, loss_type="EW", task_num=2, ignore_index=-1, weights=(1.0, 1.0, 1.0)):
    """builder.py:14-51 over [MSELoss, CrossEntropyLoss, LovaszLoss] (configs/scannet/CDSegNet.py:117-121).
    Returns (loss, [mse, ce, lovasz])."""
    parts = []
    if "c_pred" in point and "c_target" in point:
        parts.append(mse_loss(point["c_pred"], point["c_target"], point["n_target"], ignore_index, weights[0]))
    else:
        parts.append(0.0)
    parts.append(ce_loss(point["n_pred"], point["n_target"], ignore_index, weights[1]))
    parts.append(lovasz_loss(point["n_pred"], point["n_target"], ignore_index, weights[2]))
    if point["loss_mode"] == "eval" or loss_type == "EW":
        return parts[0] + parts[1] + parts[2], parts
    if task_num == 1:
        loss = parts[0] + parts[1]
    else:
        loss = parts[0] * (parts[1] + parts[2])
    return torch.pow(torch.as_tensor(loss), 1.0 / task_num), parts


# ---------------------------------------------------------------------------------------------- training pass
def training_loss(bsd, cfg, input_dict, ts, noise, abar32, loss_type="GLS", task_num=2, ignore_index=-1, dm_target="noise",
                  perm_fn=None, attn_mode="dense"):
    """DefaultSegmentorV2.forward (default.py:424-493) for condition=True, dm=True with the two random draws injected: ts int64 [B,1]
    (timestep per scene) and noise fp32 [N, C_in].  bsd = backbone state_dict (tensors may require grad: autograd then yields the
    reference's parameter gradients); BatchNorm follows ptv3_oracle.BN_TRAIN.  Returns (loss, [mse, ce, lovasz])."""
    from . import ptv3_oracle as O
    base = dict(coord=input_dict["coord"], grid_coord=input_dict["grid_coord"], offset=input_dict["offset"])
    counts = torch.diff(input_dict["offset"], prepend=input_dict["offset"].new_zeros(1))
    batch = torch.repeat_interleave(torch.arange(len(counts)), counts)
    x0 = input_dict["feat"]
    t_emb = O.calc_t_emb(ts, cfg["T_dim"])[batch]
    x_t = q_sample(abar32, x0, ts[batch], noise)
    c_out, n_out = O.forward(bsd, cfg, dict(base, feat=x_t, t_emb=t_emb), dict(base, feat=x0), attn_mode=attn_mode, perm_fn=perm_fn)
    point = dict(c_pred=c_out["feat"], c_target=noise if dm_target == "noise" else x0, n_pred=n_out["feat"], n_target=input_dict["segment"],
                 loss_mode="train")
    return criteria(point, loss_type, task_num, ignore_index)
