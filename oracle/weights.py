"""Deterministic synthetic weights keyed by parameter name (TEST INFRASTRUCTURE).

The golden fixtures under tests/golden/ were produced by running the reference
model with exactly these weights (tests/golden/make_golden.py), so fixtures need
to carry only inputs/outputs and the name->shape table, not a state_dict.
"""
import zlib
import numpy as np
import torch


def synth_tensor(name, shape, dtype=torch.float32):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    shape = tuple(int(s) for s in shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        a = 0.1 * rng.standard_normal(shape)
    elif leaf == "running_var":
        a = rng.uniform(0.5, 1.5, shape)
    elif leaf == "weight" and len(shape) == 1:          # LayerNorm / BatchNorm scale
        a = 1.0 + 0.1 * rng.standard_normal(shape)
    elif leaf == "weight" and len(shape) == 5:          # SubMConv3d [Co,k,k,k,Ci]; ~40 % of taps are occupied
        fan_in = 0.4 * np.prod(shape[1:])
        a = rng.standard_normal(shape) / np.sqrt(fan_in)
    elif leaf == "weight":                              # Linear [out,in]
        a = rng.standard_normal(shape) / np.sqrt(shape[-1])
    elif leaf == "bias":
        a = 0.05 * rng.standard_normal(shape)
    else:
        a = 0.1 * rng.standard_normal(shape)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def synth_state_dict(shapes):
    """shapes: mapping name -> shape (e.g. from model.state_dict() or a JSON table)."""
    return {k: synth_tensor(k, v) for k, v in shapes.items()}
