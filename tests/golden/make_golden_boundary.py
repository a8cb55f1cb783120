"""Golden fixture for the drop-in boundary (§8b), made by EXECUTING THE REFERENCE'S OWN LOADER on its own config files:
  pointcept/utils/config.py    Config.fromfile (mmcv-style, `_base_` inheritance)      -- addict / yapf shimmed (not installed)
  configs/{scannet,scannet200,nuscenes}/CDSegNet.py                                       -- read where they lie
  pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py                    -- PointTransformerV3(**cfg.model.backbone)
For every config it records `cfg.model` exactly as the reference's Config produced it and the reference model's own
state_dict name -> shape table, i.e. the contract `build_model(cfg.model)` + `load_state_dict(strict=True)` (engines/test.py:58-87)
holds a drop-in to.  /root/reference does not exist on the GPU box; tests/test_cpu_boundary.py reads only the fixture.

Run once in the authoring container:   python tests/golden/make_golden_boundary.py     -> boundary.json
"""
import importlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as MG          # noqa: E402

REF = MG.REF
CONFIGS = {"scannet": "configs/scannet/CDSegNet.py", "scannet200": "configs/scannet200/CDSegNet.py", "nuscenes": "configs/nuscenes/CDSegNet.py"}


def install_yapf_shim():
    y = types.ModuleType("yapf"); yl = types.ModuleType("yapf.yapflib"); ya = types.ModuleType("yapf.yapflib.yapf_api")
    ya.FormatCode = lambda text, **kw: (text, True)        # only used by Config.pretty_text
    sys.modules.update({"yapf": y, "yapf.yapflib": yl, "yapf.yapflib.yapf_api": ya})


def plain(x):
    if isinstance(x, dict):
        return {k: plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [plain(v) for v in x]
    return x


def main():
    ptv3, comm, ser = MG.load_reference()
    install_yapf_shim()
    # configs/scannet200/CDSegNet.py imports its label names from pointcept.datasets.preprocessing...: namespace packages so that
    # pointcept/datasets/__init__.py (needs termcolor, SharedArray, ...) never runs
    for name in ("pointcept.datasets", "pointcept.datasets.preprocessing", "pointcept.datasets.preprocessing.scannet",
                 "pointcept.datasets.preprocessing.scannet.meta_data"):
        m = types.ModuleType(name); m.__path__ = [os.path.join(REF, *name.split("."))]; sys.modules[name] = m
    config = importlib.import_module("pointcept.utils.config")
    out = {}
    for name, rel in CONFIGS.items():
        cfg = config.Config.fromfile(os.path.join(REF, rel))
        model = plain(cfg.model.to_dict() if hasattr(cfg.model, "to_dict") else dict(cfg.model))
        bb = dict(model["backbone"]); bb.pop("type")
        ref = ptv3.PointTransformerV3(**bb)
        shapes = {k: list(v.shape) for k, v in ref.state_dict().items()}
        n_par = sum(p.numel() for p in ref.parameters())
        out[name] = dict(config_file=rel, model=model, shapes=shapes, n_parameters=n_par,
                         optimizer=plain(cfg.optimizer), param_dicts=plain(cfg.param_dicts) if "param_dicts" in cfg else None)
        print(name, len(shapes), "tensors,", n_par, "parameters")
    json.dump(out, open(os.path.join(HERE, "boundary.json"), "w"))


if __name__ == "__main__":
    main()
