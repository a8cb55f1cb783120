"""Test-time fragment pipeline on the device.

Mirrors, for the test path only:
  GridSample(mode="test")            pointcept/datasets/transform.py:796-905 (hashes :907-933)
  CenterShift(apply_z=False) + Collect(keys=("coord","grid_coord","index"), feat_keys=...) + collate_fn
                                     transform.py:142-155, 26-50; datasets/utils.py:15-41
  the tester's voting loop           pointcept/engines/test.py:198-268

The reference voxelises on the host (numpy argsort / unique, a python loop per fragment), collates every fragment and
copies it to the GPU; here the raw scene is uploaded once, the plan is a handful of kernels (ops.grid_sample_plan), all
fragment index rows come from one gather and the votes are accumulated by a fused softmax + scatter-add kernel.
Tie order inside a voxel: the reference's np.argsort (unstable) leaves it unspecified; this path uses ascending point
index (a stable sort), see oracle/fragments_np.py.
"""
import numpy as np
import torch

from . import ops


def _dev(x, device, dtype=None):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    x = x.to(device, non_blocking=True)
    return x.to(dtype) if dtype is not None and x.dtype is not dtype else x


class GridSample:
    """`GridSample(..., mode="test")` with the reference's constructor; `__call__` returns the list of fragment dicts
    (`index`, optionally `grid_coord`, the gathered `keys`, every other entry passed through) as CUDA tensors."""

    def __init__(self, grid_size=0.05, hash_type="fnv", mode="test", keys=("coord", "color", "normal", "segment"), return_inverse=False,
                 return_grid_coord=False, return_min_coord=False, return_displacement=False, project_displacement=False,
                 legacy_f32=False, device="cuda"):
        if mode != "test":
            raise NotImplementedError("cdsegnet_b200.fragments.GridSample implements mode='test' (the tester's voxelize step)")
        if return_displacement or project_displacement:
            raise NotImplementedError("return_displacement is not used by the CDSegNet test configs")
        self.grid_size, self.hash_type, self.mode, self.keys = grid_size, hash_type, mode, tuple(keys)
        self.return_inverse, self.return_grid_coord, self.return_min_coord = return_inverse, return_grid_coord, return_min_coord
        self.legacy_f32, self.device = legacy_f32, device

    def plan(self, coord):
        coord = _dev(coord, self.device).contiguous()
        p = ops.grid_sample_plan(coord, self.grid_size, self.hash_type, self.legacy_f32)
        p["index"] = ops.fragment_index(p["order"], p["start"], p["n_voxels"], p["n_fragments"])
        return p

    def __call__(self, data_dict):
        assert "coord" in data_dict
        data = {k: (_dev(v, self.device) if isinstance(v, (np.ndarray, torch.Tensor)) else v) for k, v in data_dict.items()}
        data["coord"] = data["coord"].contiguous()
        p = self.plan(data["coord"])
        if self.return_inverse:                                # transform.py:873-875: set on the scene dict, so every part carries it
            data_dict["inverse"] = data["inverse"] = p["inverse"].long()
        parts = []
        for f in range(p["n_fragments"]):
            idx = p["index"][f]
            il = idx.long()
            part = dict(index=il)
            if self.return_grid_coord:
                part["grid_coord"] = p["grid_coord"][il]
            if self.return_min_coord:
                part["min_coord"] = (torch.tensor(p["min_grid"], dtype=torch.float64) * self.grid_size).reshape(1, 3)
            for k, v in data.items():
                part[k] = v[il] if k in self.keys else v
            parts.append(part)
        return parts


def center_shift(coord, apply_z=True):
    """transform.py:142-155 on a device tensor, in the tensor's own dtype like numpy there"""
    mn, mx = coord.min(0).values, coord.max(0).values
    shift = torch.stack([(mn[0] + mx[0]) / 2, (mn[1] + mx[1]) / 2, mn[2] if apply_z else torch.zeros_like(mn[2])])
    return coord - shift


def collect_fragment(part, feat_keys=("color", "normal"), apply_z=False):
    """post_transform of the shipped test configs on one fragment + collate_fn of a single-fragment batch ->
    the model's input_dict (coord, grid_coord int, index, feat, offset)"""
    coord = center_shift(part["coord"], apply_z)
    feat = torch.cat([part[k].float() for k in feat_keys], dim=1)
    n = coord.shape[0]
    return dict(coord=coord.float().contiguous(), grid_coord=part["grid_coord"], index=part["index"], feat=feat.contiguous(),
                offset=torch.tensor([n], dtype=torch.int64, device=coord.device))


class FragmentVoter:
    """the tester's per-scene loop (engines/test.py:198-268): for every augmentation and fragment run the model, add the
    softmax of its logits into pred[index], return the argmax labels.

      voter = FragmentVoter(model, num_classes=20, voxelize=GridSample(grid_size=0.02, keys=("coord","color","normal"), return_grid_coord=True))
      labels = voter(dict(coord=..., color=..., normal=...), augs=[None, rotate_z(0.5), ...])
    """

    def __init__(self, model, num_classes, voxelize, feat_keys=("color", "normal"), inference_mode="SSI", noise_level=None, step=1):
        self.model, self.num_classes, self.voxelize, self.feat_keys = model, num_classes, voxelize, tuple(feat_keys)
        self.inference_mode, self.noise_level, self.step = inference_mode, noise_level, step

    def _logits(self, input_dict):
        if self.inference_mode == "SSI":
            return self.model.inference(input_dict, eval=False, noise_level=self.noise_level)["seg_logits"]
        mode = {"MSAI": "avg", "MSFI": "final"}[self.inference_mode]
        return self.model.inference_ddim(input_dict, eval=False, noise_level=self.noise_level, mode=mode, step=self.step)["seg_logits"]

    @torch.no_grad()
    def votes(self, data_dict, augs=(None,)):
        dev = self.voxelize.device
        base = {k: (_dev(v, dev) if isinstance(v, (np.ndarray, torch.Tensor)) else v) for k, v in data_dict.items()}
        n = base["coord"].shape[0]
        pred = torch.zeros((n, self.num_classes), dtype=torch.float32, device=dev)
        for aug in augs:
            d = dict(base)
            if aug is not None:
                d = aug(d)
            for part in self.voxelize(d):
                inp = collect_fragment(part, self.feat_keys)
                logits = self._logits(inp)
                ops.vote_softmax_add_(pred, logits.contiguous(), inp["index"].int().contiguous())
        return pred

    def __call__(self, data_dict, augs=(None,)):
        return ops.argmax_rows(self.votes(data_dict, augs))


def rotate_z(angle_half_turns, scale=1.0):
    """RandomRotateTargetAngle(angle=[a], axis="z", center=[0,0,0], p=1) (+ RandomScale([s, s])) of the shipped TTA list
    (configs/scannet/CDSegNet.py:282-370; transform.py:259-309): coord and normal rotate, coord scales; float64 results like
    np.dot(float32, float64) there"""
    def aug(d):
        a = float(angle_half_turns) * np.pi
        c, s = np.cos(a), np.sin(a)
        R = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=torch.float64, device=d["coord"].device)
        d = dict(d)
        d["coord"] = (d["coord"].double() @ R.t()) * scale
        if "normal" in d:
            d["normal"] = d["normal"].double() @ R.t()
        return d
    return aug
