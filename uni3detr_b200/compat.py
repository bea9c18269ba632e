"""Registry / Config shim.

The reference builds everything through the OpenMMLab registries
(``build_model(cfg.model)`` -> ``type=`` lookups, extra_tools/test.py:193) and the config
files are plain python dicts (projects/configs/uni3detr/*.py). mmcv / mmdet / mmdet3d are
not installable in this environment, so this module provides the minimum of that API
surface: ``Registry.register_module``, ``build_from_cfg``, ``ConfigDict`` and
``Config.fromfile`` with ``_base_`` inheritance. When the real packages are importable,
``register_with_openmmlab()`` additionally registers the drop-in classes under the real
registries so extra_tools/{train,test}.py find them by the same names.
"""
import copy
import importlib
import os
import types


class ConfigDict(dict):
    """dict with attribute access (like mmcv.ConfigDict)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_config(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [to_config(v) for v in obj]
    if isinstance(obj, tuple):
        return tuple(to_config(v) for v in obj)
    return obj


def _merge(base, new):
    out = copy.deepcopy(base)
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            v = copy.deepcopy(v)
            if isinstance(v, dict):
                v.pop("_delete_", None)
            out[k] = v
    return out


class Config:
    """Subset of mmcv.Config: python config files, `_base_` inheritance, attribute access."""

    def __init__(self, cfg_dict, filename=None):
        object.__setattr__(self, "_cfg", to_config(cfg_dict))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def _load(filename):
        filename = os.path.abspath(filename)
        with open(filename) as f:
            src = f.read()
        scope = {"__file__": filename}
        exec(compile(src, filename, "exec"), scope)
        cfg = {k: v for k, v in scope.items()
               if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType))}
        bases = cfg.pop("_base_", [])
        if isinstance(bases, str):
            bases = [bases]
        merged = {}
        for b in bases:
            path = os.path.join(os.path.dirname(filename), b)
            if not os.path.exists(path):
                # the reference configs inherit mmdet3d's default_runtime.py, which is not part of
                # the reference tree (uni3detr_sunrgbd.py:1-3); it carries no model keys.
                continue
            merged = _merge(merged, Config._load(path))
        return _merge(merged, cfg)

    @staticmethod
    def fromfile(filename):
        return Config(Config._load(filename), filename)

    def __getattr__(self, name):
        return getattr(self._cfg, name)

    def __getitem__(self, name):
        return self._cfg[name]

    def get(self, name, default=None):
        return self._cfg.get(name, default)

    def __contains__(self, name):
        return name in self._cfg


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._modules and not force and self._modules[key] is not cls:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self._modules.get(key)

    def __contains__(self, key):
        return key in self._modules

    def build(self, cfg, **default_args):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    if cfg is None:
        return None
    if not isinstance(cfg, dict) or "type" not in cfg:
        raise TypeError(f"cfg must be a dict with a `type` key, got {cfg!r}")
    args = dict(cfg)
    typ = args.pop("type")
    if isinstance(typ, str):
        cls = registry.get(typ)
        if cls is None:
            raise KeyError(f"{typ} is not in the {registry.name} registry")
    else:
        cls = typ
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    return cls(**args)


# the registries the reference's modules register into
DETECTORS = Registry("detector")
MIDDLE_ENCODERS = Registry("middle_encoder")
VOXEL_ENCODERS = Registry("voxel_encoder")
BACKBONES = Registry("backbone")
NECKS = Registry("neck")
HEADS = Registry("head")
TRANSFORMER = Registry("transformer")
TRANSFORMER_LAYER_SEQUENCE = Registry("transformer_layer_sequence")
TRANSFORMER_LAYER = Registry("transformer_layer")
ATTENTION = Registry("attention")
FEEDFORWARD_NETWORK = Registry("feedforward_network")
BBOX_CODERS = Registry("bbox_coder")
POSITIONAL_ENCODING = Registry("positional_encoding")
LOSSES = Registry("loss")

_MODEL_REGISTRIES = [DETECTORS, MIDDLE_ENCODERS, VOXEL_ENCODERS, BACKBONES, NECKS, HEADS]


def build_model(cfg, train_cfg=None, test_cfg=None):
    """mmdet3d.models.build_model equivalent for the names this package provides."""
    cfg = to_config(copy.deepcopy(dict(cfg)))
    if train_cfg is not None:
        cfg.setdefault("train_cfg", train_cfg)
    if test_cfg is not None:
        cfg.setdefault("test_cfg", test_cfg)
    for reg in _MODEL_REGISTRIES:
        if cfg["type"] in reg:
            return build_from_cfg(cfg, reg)
    raise KeyError(f"{cfg['type']} is not registered")


def register_with_openmmlab():
    """If mmcv/mmdet/mmdet3d exist, mirror our classes into their registries (force=True)."""
    try:
        mm_det = importlib.import_module("mmdet.models.builder")
        mm_3d = importlib.import_module("mmdet3d.models.builder")
        mm_tr = importlib.import_module("mmcv.cnn.bricks.registry")
        mm_ut = importlib.import_module("mmdet.models.utils.builder")
        mm_bb = importlib.import_module("mmdet.core.bbox.builder")
    except Exception:
        return False
    pairs = [(DETECTORS, mm_det.DETECTORS), (MIDDLE_ENCODERS, mm_3d.MIDDLE_ENCODERS),
             (VOXEL_ENCODERS, mm_3d.VOXEL_ENCODERS), (BACKBONES, mm_det.BACKBONES),
             (NECKS, mm_det.NECKS), (HEADS, mm_det.HEADS), (TRANSFORMER, mm_ut.TRANSFORMER),
             (TRANSFORMER_LAYER_SEQUENCE, mm_tr.TRANSFORMER_LAYER_SEQUENCE),
             (ATTENTION, mm_tr.ATTENTION), (BBOX_CODERS, mm_bb.BBOX_CODERS)]
    for ours, theirs in pairs:
        for key, cls in ours._modules.items():
            theirs.register_module(name=key, force=True, module=cls)
    return True
