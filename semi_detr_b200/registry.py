"""Minimal stand-in for the mmcv ``Registry`` / ``build_from_cfg`` pair the reference's configs rely on
(DETECTORS, HEADS, TRANSFORMER, BBOX_ASSIGNERS, MATCH_COST, LOSSES, HOOKS ...).
Only what the hot path needs: ``type=`` strings resolve to the classes of this package unchanged."""


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def get(self, key):
        if key not in self._modules:
            raise KeyError(f"{key} is not in the {self.name} registry")
        return self._modules[key]

    def __contains__(self, key):
        return key in self._modules

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict) or "type" not in cfg:
        raise TypeError(f"cfg must be a dict with a 'type' key, got {cfg!r}")
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop("type")
    cls = registry.get(obj_type) if isinstance(obj_type, str) else obj_type
    return cls(**args)


MATCH_COST = Registry("Match Cost")
BBOX_ASSIGNERS = Registry("bbox_assigner")
HOOKS = Registry("hook")
LOSSES = Registry("loss")
HEADS = Registry("head")
DETECTORS = Registry("detector")
TRANSFORMER = Registry("Transformer")
POSITIONAL_ENCODING = Registry("position encoding")
BACKBONES = Registry("backbone")
