"""``DefaultSegmentorV2`` (CNF, variant (1)) boundary: pointcept/models/default.py:13-494.

Implements the single-step inference (SSI) path ``inference(input_dict, eval, noise_level)``
(default.py:371-422) on top of the B200 backbone, with the reference's RNG coupling kept
(the N(0,1) draw for the Noise-Network input comes from torch's CPU generator exactly like
default.py:393) but with the timestep embedding built once per SCENE instead of per point.
``forward`` (training loss) and ``inference_ddim`` are the next §8(f) rows.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .registry import build_model


def calc_t_emb(ts, t_emb_dim):
    """pointcept/utils/comm.py:21-39.  ts: int64 [R,1] -> fp32 [R, t_emb_dim]."""
    assert t_emb_dim % 2 == 0
    half = t_emb_dim // 2
    f = torch.exp(torch.arange(half) * -(np.log(10000) / (half - 1))).to(ts.device)
    e = ts * f
    return torch.cat((torch.sin(e), torch.cos(e)), 1)


class DefaultSegmentorV2(nn.Module):
    def __init__(self, backbone=None, criteria=None, loss_type="EW", task_num=2, num_classes=20, T=1000,
                 beta_start=0.0001, beta_end=0.02, noise_schedule="linear", T_dim=128, dm=False, dm_input="xt",
                 dm_target="noise", dm_min_snr=None, condition=False, c_in_channels=6):
        super().__init__()
        self.backbone = build_model(backbone) if isinstance(backbone, dict) else backbone
        self.criteria_cfg = criteria            # losses are stock torch; not on the inference hot path
        self.num_classes, self.T, self.T_dim = num_classes, T, T_dim
        self.condition, self.dm, self.dm_input, self.dm_target = condition, dm, dm_input, dm_target
        self.c_in_channels = c_in_channels

    @torch.no_grad()
    def inference(self, input_dict, eval=True, noise_level=None, noise=None):
        """-> dict(seg_logits=[N, num_classes]).  `noise` optionally injects the NN input
        (otherwise drawn like default.py:393: CPU generator, then moved to the GPU)."""
        if eval:
            raise NotImplementedError("eval=True (loss on the validation pass) needs the criteria; "
                                      "use eval=False as tools/test_CDSegNet_*.py do (engines/test.py:214-218)")
        feat = input_dict["feat"]
        if noise_level is not None:              # add_gaussian_noise, default.py:225-233
            feat = feat + noise_level * torch.randn_like(feat)
        base = dict(coord=input_dict["coord"], grid_coord=input_dict["grid_coord"], offset=input_dict["offset"])
        if not self.condition:
            n_point = self.backbone(n_point=dict(base, feat=feat))
            return dict(seg_logits=n_point["feat"])
        c_target = feat if self.c_in_channels == feat.shape[-1] else input_dict["coord"]
        c_feat, t = c_target, 0
        if self.dm and self.dm_input == "xt":
            if noise is None:
                noise = torch.normal(0, 1, size=c_target.shape, dtype=torch.float32)
            c_feat = noise.to(feat.device, non_blocking=True)
            t = self.T - 1
        c_point = dict(base, feat=c_feat)
        if self.T_dim != -1:
            B = input_dict["offset"].numel()
            ts = t * torch.ones((B, 1), dtype=torch.int64, device=feat.device)
            c_point["t_emb"] = calc_t_emb(ts, self.T_dim)          # one row per scene
        c_point, n_point = self.backbone(c_point, dict(base, feat=feat))
        return dict(seg_logits=n_point["feat"])
