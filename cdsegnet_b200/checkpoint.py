"""Loading reference checkpoints.

The reference's tester / evaluator load `checkpoint["state_dict"]` with `strict=True`, stripping the DDP `module.` prefix on a
single process and adding it under DDP (pointcept/engines/test.py:67-84; hooks/misc.py:215-238 also rewrites a keyword).
`load_checkpoint` does the same for a `cdsegnet_b200.DefaultSegmentorV2` (or backbone): the parameter names are the reference's
(tests/test_cpu_oracle.py::test_state_dict_contract), so released `model_best.pth` files load unchanged.

spconv weight layout: SubMConv3d weights are stored `[C_out, k, k, k, C_in]` (spconv 2.x) and consumed tap-major here; the
reference tree does not pin spconv's tap order (SURVEY.md §8c "parity unpinned"), so `verify_conv_layout` is provided to
check a checkpoint against a forward of the real reference before trusting its logits.
"""
from collections import OrderedDict

import torch


def normalize_state_dict(state_dict, wrapped=False, keywords="", replacement=None):
    """test.py:70-78 / hooks/misc.py:226-238: `module.` prefix handling (+ optional keyword replacement)"""
    out = OrderedDict()
    for key, value in state_dict.items():
        if not key.startswith("module."):
            key = "module." + key
        if keywords and keywords in key:
            key = key.replace(keywords, replacement if replacement is not None else keywords)
        if not wrapped:
            key = key[7:]
        out[key] = value
    return out


def load_checkpoint(model, checkpoint, strict=True, keywords="", replacement=None, map_location="cpu"):
    """checkpoint: path, the dict torch.load returns (with "state_dict"), or a bare state_dict.
    Returns dict(epoch=..., best_metric_value=..., missing_keys=[...], unexpected_keys=[...])."""
    if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__"):
        checkpoint = torch.load(checkpoint, map_location=map_location, weights_only=False)
    sd = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    wrapped = isinstance(model, torch.nn.parallel.DistributedDataParallel)
    info = model.load_state_dict(normalize_state_dict(sd, wrapped, keywords, replacement), strict=strict)
    meta = {k: checkpoint[k] for k in ("epoch", "best_metric_value") if isinstance(checkpoint, dict) and k in checkpoint}
    return dict(meta, missing_keys=list(info.missing_keys), unexpected_keys=list(info.unexpected_keys))


def verify_conv_layout(model, input_dict, reference_logits, tol=1e-3):
    """max |logits - reference_logits| of `model.inference(input_dict, eval=False)` must be below tol when the spconv tap
    order assumed here (x-major taps, weight [C_out, kx, ky, kz, C_in]) matches the checkpoint's"""
    out = model.inference(input_dict, eval=False)["seg_logits"]
    err = float((out - reference_logits.to(out.device)).abs().max())
    if err > tol:
        raise RuntimeError(f"checkpoint logits differ from the reference by {err:.3e} > {tol}: spconv tap order / weight layout mismatch?")
    return err
