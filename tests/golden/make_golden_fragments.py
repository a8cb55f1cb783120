"""Golden fixtures for the test-time fragment pipeline made by EXECUTING THE REFERENCE'S OWN
pointcept/datasets/transform.py (GridSample, CenterShift) and pointcept/datasets/utils.py (collate_fn) on CPU.
Run once in the authoring container:   python tests/golden/make_golden_fragments.py   -> fragments.npz
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def raw_scene(n, seed, extent=(3.0, 2.4, 1.6)):
    """raw scan-like cloud: several points per 2 cm voxel on a few planes (so that count.max() > 1), negative coordinates included"""
    rng = np.random.default_rng(seed)
    face = rng.integers(0, 3, n)
    p = rng.random((n, 3)) * np.array(extent)
    p[face == 0, 2] = 0.0
    p[face == 1, 0] = 0.0
    p[face == 2, 1] = extent[1]
    p += rng.normal(0, 0.004, p.shape)
    p -= np.array([1.0, 0.7, 0.1])
    return p.astype(np.float32)


def main():
    for name, path in (("pointcept", f"{REF}/pointcept"), ("pointcept.utils", f"{REF}/pointcept/utils"),
                       ("pointcept.datasets", f"{REF}/pointcept/datasets")):
        m = types.ModuleType(name); m.__path__ = [path]; sys.modules[name] = m
    T = importlib.import_module("pointcept.datasets.transform")
    U = importlib.import_module("pointcept.datasets.utils")
    rec = {"numpy_version": np.__version__}
    cases = [(6000, 0.02, "fnv", (3.0, 2.4, 1.6)), (6000, 0.05, "ravel", (3.0, 2.4, 1.6)), (50, 0.02, "fnv", (0.2, 0.2, 0.1)),
             (20000, 0.02, "fnv", (1.0, 0.8, 0.6))]
    for i, (n, gs, ht, ext) in enumerate(cases):
        coord = raw_scene(n, i, ext)
        rng = np.random.default_rng(100 + i)
        color = rng.random((n, 3)).astype(np.float32)
        normal = rng.standard_normal((n, 3)).astype(np.float32)
        gsamp = T.GridSample(grid_size=gs, hash_type=ht, mode="test", keys=("coord", "color", "normal"), return_grid_coord=True,
                             return_inverse=True)
        d = dict(coord=coord.copy(), color=color, normal=normal, name="scene")
        parts = gsamp(d)
        # the deterministic intermediates, recomputed with the reference's own static methods
        scaled = coord / np.array(gs)
        grid = np.floor(scaled).astype(int)
        grid -= grid.min(0)
        key = gsamp.hash(grid)
        rec[f"c{i}_coord"], rec[f"c{i}_grid_size"], rec[f"c{i}_hash"] = coord, gs, ht
        rec[f"c{i}_color"], rec[f"c{i}_normal"] = color, normal
        rec[f"c{i}_grid"], rec[f"c{i}_key"] = grid, key
        rec[f"c{i}_inverse"] = d["inverse"]
        rec[f"c{i}_n_fragments"] = len(parts)
        rec[f"c{i}_index"] = np.stack([p["index"] for p in parts])
        rec[f"c{i}_part_grid"] = np.stack([p["grid_coord"] for p in parts])
        # post_transform of the shipped test config on fragment 0: CenterShift(apply_z=False) -> ToTensor -> Collect, then collate_fn
        post = T.Compose([dict(type="CenterShift", apply_z=False), dict(type="ToTensor"),
                          dict(type="Collect", keys=("coord", "grid_coord", "index"), feat_keys=("color", "normal"))])
        frag = post({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in parts[0].items()})
        batch = U.collate_fn([frag])
        for k in ("coord", "grid_coord", "index", "feat", "offset"):
            rec[f"c{i}_in_{k}"] = batch[k].numpy()
    rec["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(HERE, "fragments.npz"), **rec)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in rec.items() if k.startswith("c0_")})


if __name__ == "__main__":
    main()
