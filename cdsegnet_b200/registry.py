"""Model registry boundary (pointcept/models/builder.py:8-16; pointcept/utils/registry.py:9-56, 238-316).

If the reference's ``pointcept`` package is importable, the B200 classes are registered
into ITS ``MODELS`` registry with ``force=True`` under the reference's own keys
("PT-v3m1", "DefaultSegmentorV2"), so ``build_model(cfg.model)`` in the reference's
tools/*.py returns them unchanged.  Otherwise a minimal registry
with the same ``build(cfg)`` semantics is used.
"""


class Registry:
    def __init__(self, name):
        self.name = name
        self._module_dict = {}

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._module_dict and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._module_dict[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def build(self, cfg):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise KeyError('cfg must be a dict containing the key "type"')
        args = dict(cfg)
        t = args.pop("type")
        cls = self.get(t) if isinstance(t, str) else t
        if cls is None:
            raise KeyError(f"{t} is not in the {self.name} registry")
        return cls(**args)


try:                                    # reference present: plug into its registry
    from pointcept.models.builder import MODELS  # type: ignore
    USING_POINTCEPT_REGISTRY = True
except Exception:                       # standalone
    MODELS = Registry("models")
    USING_POINTCEPT_REGISTRY = False


def build_model(cfg):
    return MODELS.build(cfg)


def register(name, cls):
    MODELS.register_module(name=name, force=True, module=cls)
    return cls
