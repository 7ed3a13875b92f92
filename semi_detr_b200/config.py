"""Python-file configs with ``_base_`` inheritance -- just enough of mmcv's ``Config.fromfile`` for the reference's
``configs/dino_detr/*.py`` and ``configs/detr_ssod/*.py`` to load unchanged, plus the reference's own post-processing
(``detr_ssod/utils/patch.py:69-81`` ``patch_config``: ``cfg_name``, ``${...}`` substitution through
``detr_ssod/utils/vars.py:15-35``, and ``semi_wrapper`` taking the place of ``model``).

Only the ``model=`` / ``semi_wrapper=`` dicts are consumed here (``build_detector``); runner, data pipeline, logging
and evaluation keys are loaded and kept but nothing in this package acts on them (out of scope).

Merge rules restated from mmcv 1.3.16 ``Config._merge_a_into_b`` / ``_file2dict``:
 * a config file is executed as Python; its public, non-module, non-callable globals are the config;
 * ``_base_`` (a path or a list of paths, relative to the file) is loaded first, bases may not define the same
   top-level key twice, the child is merged INTO the base;
 * dict-into-dict merges recursively; a child dict carrying ``_delete_=True`` replaces the base value instead;
 * anything else overwrites.
"""
import copy
import os
import re
import types

_VAR = re.compile(r"\$\{[a-zA-Z\d_.]*\}")
DELETE_KEY = "_delete_"
BASE_KEY = "_base_"


class ConfigDict(dict):
    """dict with attribute access (``cfg.model.bbox_head.type``)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"'ConfigDict' object has no attribute '{name}'") from None

    def __setattr__(self, name, value):
        self[name] = value

    def to_dict(self):
        return _plain(self)


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_plain(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_plain(x) for x in v)
    return v


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict((k, _wrap(x)) for k, x in v.items())
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


def _merge(child, base):
    """child INTO a copy of base"""
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get(DELETE_KEY, False):
            out[k] = _merge(v, out[k])
        elif isinstance(v, dict):
            v = {kk: vv for kk, vv in v.items() if kk != DELETE_KEY}
            out[k] = _merge(v, {})
        else:
            out[k] = v
    return out


def _file2dict(filename):
    filename = os.path.abspath(os.path.expanduser(filename))
    if not os.path.isfile(filename):
        raise FileNotFoundError(f"config file {filename} does not exist")
    if not filename.endswith(".py"):
        raise IOError("only .py configs are supported")
    with open(filename, "r") as f:
        source = f.read()
    scope = {"__file__": filename, "__name__": "_sdb_config_"}
    exec(compile(source, filename, "exec"), scope)
    cfg = {k: v for k, v in scope.items()
           if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType, type))}
    bases = cfg.pop(BASE_KEY, None)
    if bases is None:
        return cfg
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        bd = _file2dict(os.path.join(os.path.dirname(filename), b))
        dup = merged.keys() & bd.keys()
        if dup:
            raise KeyError(f"duplicate key(s) {sorted(dup)} in the bases of {filename}")
        merged.update(bd)
    return _merge(cfg, merged)


def _get_value(cfg, chained_key):
    for k in chained_key.split("."):
        cfg = cfg[k]
    return cfg


def resolve(cfg, base=None):
    """``${a.b}`` substitution: a string that IS one variable takes the variable's value (any type, e.g.
    ``model="${model}"``); variables inside a longer string are formatted in."""
    if base is None:
        base = cfg
    if isinstance(cfg, dict):
        return {k: resolve(v, base) for k, v in cfg.items()}
    if isinstance(cfg, list):
        return [resolve(v, base) for v in cfg]
    if isinstance(cfg, tuple):
        return tuple(resolve(v, base) for v in cfg)
    if isinstance(cfg, str):
        names = _VAR.findall(cfg)
        if len(names) == 1 and len(names[0]) == len(cfg):
            return copy.deepcopy(_get_value(base, names[0][2:-1]))
        for n in names:
            cfg = cfg.replace(n, str(_get_value(base, n[2:-1])))
        return cfg
    return cfg


class Config:
    """``Config.fromfile(path)`` -> attribute/dict access to the merged config; ``cfg.filename`` is kept."""

    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, "_cfg_dict", _wrap(cfg_dict or {}))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def fromfile(filename):
        return Config(_file2dict(filename), filename=os.path.abspath(filename))

    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_cfg_dict"), name)

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, name, default=None):
        return self._cfg_dict.get(name, default)

    def pop(self, name, *default):
        return self._cfg_dict.pop(name, *default)

    def to_dict(self):
        return self._cfg_dict.to_dict()


def patch_config(cfg):
    """The reference's ``patch_config`` without its side effects on the environment: adds ``cfg_name``, resolves
    ``${...}``, and lets ``semi_wrapper`` replace ``model`` (detr_ssod/utils/patch.py:69-81)."""
    d = cfg.to_dict()
    d["cfg_name"] = os.path.splitext(os.path.basename(cfg.filename))[0]
    d = resolve(d)
    out = Config(d, filename=cfg.filename)
    if out.get("semi_wrapper") is not None:
        out.model = out.semi_wrapper
        out.pop("semi_wrapper")
    return out


def build_detector(model_cfg, train_cfg=None, test_cfg=None):
    """``mmdet.models.build_detector`` for the registry of this package: ``DINODETR`` or the ``DinoDetrSSOD`` wrapper."""
    from . import dino, ssod  # noqa: F401  (registers the classes)
    from .registry import DETECTORS
    cfg = _plain(model_cfg)
    return DETECTORS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg)
                           if (train_cfg is not None or test_cfg is not None) else None)
