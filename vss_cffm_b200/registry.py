"""Plugin surface of the reference: mmcv-style registries under the reference's names.

Reference: mmseg/models/builder.py:3-66 (five ``Registry`` objects + ``build_*``), which relies on
``mmcv.utils.Registry`` / ``build_from_cfg`` (mmcv 1.3.0, absent from this image).  A config names
a class by the string ``type=...`` and the remaining keys become constructor kwargs.
"""
import warnings


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return key in self._module_dict

    def __repr__(self):
        return f"Registry(name={self._name}, items={sorted(self._module_dict)})"

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        if not (isinstance(cls, type) or callable(cls)):
            raise TypeError(f"module must be a class or callable, got {type(cls)}")
        name = name or cls.__name__
        if not force and name in self._module_dict:
            raise KeyError(f"{name} is already registered in {self._name}")
        self._module_dict[name] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module
        if isinstance(name, type):                     # bare @REG.register_module form
            self._register(name)
            return name

        def deco(cls):
            self._register(cls, name, force)
            return cls
        return deco


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f"cfg must be a dict, got {type(cfg)}")
    if "type" not in cfg and not (default_args and "type" in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", got {cfg}')
    args = dict(cfg)
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    typ = args.pop("type")
    if isinstance(typ, str):
        cls = registry.get(typ)
        if cls is None:
            raise KeyError(f"{typ} is not in the {registry.name} registry")
    elif isinstance(typ, type) or callable(typ):
        cls = typ
    else:
        raise TypeError(f"type must be a str or class, got {type(typ)}")
    return cls(**args)


BACKBONES = Registry("backbone")
NECKS = Registry("neck")
HEADS = Registry("head")
LOSSES = Registry("loss")
SEGMENTORS = Registry("segmentor")


def build(cfg, registry, default_args=None):
    if isinstance(cfg, (list, tuple)):
        import torch.nn as nn
        return nn.Sequential(*[build_from_cfg(c, registry, default_args) for c in cfg])
    return build_from_cfg(cfg, registry, default_args)


def build_backbone(cfg):
    return build(cfg, BACKBONES)


def build_neck(cfg):
    return build(cfg, NECKS)


def build_head(cfg):
    return build(cfg, HEADS)


def build_loss(cfg):
    return build(cfg, LOSSES)


def build_segmentor(cfg, train_cfg=None, test_cfg=None):
    """mmseg/models/builder.py:56-66."""
    if train_cfg is not None or test_cfg is not None:
        warnings.warn("train_cfg and test_cfg is deprecated, please specify them in model", UserWarning)
    assert cfg.get("train_cfg") is None or train_cfg is None, "train_cfg specified in both outer field and model field"
    assert cfg.get("test_cfg") is None or test_cfg is None, "test_cfg specified in both outer field and model field"
    return build(cfg, SEGMENTORS, dict(train_cfg=train_cfg, test_cfg=test_cfg))
