"""Plumbing-only stand-ins that let the UNMODIFIED reference (/root/reference) import in
the build container, where mmcv / timm / IPython / fast_pytorch_kmeans are absent.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_goldens.py`` (run in the build
container, where /root/reference exists) to execute the real reference on CPU and
write ``tests/golden/*.npz``.  Nothing here is arithmetic of the hot path: every
tensor op is executed by the reference's own files.  The only "math" restated is
``mmcv.cnn.ConvModule`` = Conv2d -> (Sync)BatchNorm -> ReLU (mmcv 1.3.0,
call site reference mmseg/models/decode_heads/cffm_head.py:61-66).

Nothing in the product package, the ``-m gpu`` tests, ``smoke()`` or ``bench.py``
imports this module.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("CFFM_REFERENCE_ROOT", "/root/reference")
_SHIM_ROOTS = ("mmcv", "timm", "IPython", "fast_pytorch_kmeans", "terminaltables",
               "matplotlib", "addict", "yapf", "prettytable")


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _Dummy(metaclass=_DummyMeta):
    """Accepts construction, subclassing, attribute access and use as a decorator."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and (isinstance(a[0], type) or callable(a[0])):
            return a[0]
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __iter__(self):
        return iter(())


# ----------------------------------------------------------------------------- mmcv.utils
class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def __contains__(self, key):
        return key in self._module_dict

    def _register_module(self, module_class, module_name=None, force=False):
        name = module_name or module_class.__name__
        self._module_dict[name] = module_class

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register_module(module, name, force)
            return module
        if isinstance(name, type):  # deprecated bare-decorator form
            self._register_module(name)
            return name

        def _reg(cls):
            self._register_module(cls, name, force)
            return cls
        return _reg


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f"{obj_type} is not in the {registry.name} registry")
    else:
        obj_cls = obj_type
    return obj_cls(**args)


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value


def _to_cfgdict(x):
    if isinstance(x, dict):
        return ConfigDict({k: _to_cfgdict(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_to_cfgdict(v) for v in x]
    if isinstance(x, tuple):
        return tuple(_to_cfgdict(v) for v in x)
    return x


def _merge(a, b):
    """b overrides a; dicts merge recursively unless b carries _delete_=True."""
    b = dict(b)
    if b.pop("_delete_", False):
        return b
    out = dict(a)
    for k, v in b.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def _load_cfg_file(path):
    path = os.path.abspath(path)
    ns = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), ns)
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not isinstance(v, types.ModuleType)}
    base = cfg.pop("_base_", None)
    if base is None:
        return cfg
    if isinstance(base, str):
        base = [base]
    merged = {}
    for b in base:
        merged = _merge(merged, _load_cfg_file(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


class Config:
    def __init__(self, d, filename=None):
        object.__setattr__(self, "_cfg_dict", _to_cfgdict(d))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def fromfile(filename):
        return Config(_load_cfg_file(filename), filename)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def get(self, *a):
        return self._cfg_dict.get(*a)


def is_tuple_of(seq, expected_type):
    return isinstance(seq, tuple) and all(isinstance(i, expected_type) for i in seq)


def _identity_decorator(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]

    def wrap(fn):
        return fn
    return wrap


# ------------------------------------------------------------------------------ mmcv.cnn
def build_norm_layer(cfg, num_features, postfix=""):
    t = cfg["type"]
    if t in ("BN", "BN2d"):
        return "bn" + str(postfix), nn.BatchNorm2d(num_features, eps=cfg.get("eps", 1e-5))
    if t == "SyncBN":
        return "bn" + str(postfix), nn.SyncBatchNorm(num_features, eps=cfg.get("eps", 1e-5))
    if t == "GN":
        return "gn" + str(postfix), nn.GroupNorm(cfg["num_groups"], num_features)
    if t == "LN":
        return "ln" + str(postfix), nn.LayerNorm(num_features)
    raise KeyError(t)


class ConvModule(nn.Module):
    """conv -> norm -> act with mmcv 1.3.0 naming (conv / bn / activate), bias='auto'."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"),
                 inplace=True, with_spectral_norm=False, padding_mode="zeros",
                 order=("conv", "norm", "act")):
        super().__init__()
        assert order == ("conv", "norm", "act") and conv_cfg is None
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode="fan_out", nonlinearity="relu")
        if bias:
            nn.init.constant_(self.conv.bias, 0)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            assert act_cfg["type"] == "ReLU"
            self.activate = nn.ReLU(inplace=inplace)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = getattr(self, self.norm_name)(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def normal_init(module, mean=0, std=1, bias=0):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.normal_(module.weight, mean, std)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
    nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def xavier_init(module, gain=1, bias=0, distribution="normal"):
    nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


# ------------------------------------------------------------------------------- timm
class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.training, "shim DropPath is eval-only"
        return x


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


_REAL = {
    "mmcv": dict(__version__="1.3.0", Config=Config, ConfigDict=ConfigDict),
    "mmcv.utils": dict(Registry=Registry, build_from_cfg=build_from_cfg, ConfigDict=ConfigDict,
                       Config=Config, is_tuple_of=is_tuple_of,
                       deprecated_api_warning=_identity_decorator,
                       get_logger=lambda *a, **k: __import__("logging").getLogger("mmseg"),
                       print_log=lambda *a, **k: None),
    "mmcv.utils.parrots_wrapper": dict(_BatchNorm=nn.modules.batchnorm._BatchNorm,
                                       SyncBatchNorm=nn.SyncBatchNorm,
                                       DataLoader=torch.utils.data.DataLoader),
    "mmcv.cnn": dict(ConvModule=ConvModule, build_norm_layer=build_norm_layer, normal_init=normal_init,
                     constant_init=constant_init, kaiming_init=kaiming_init, xavier_init=xavier_init,
                     UPSAMPLE_LAYERS=Registry("upsample layer")),
    "mmcv.runner": dict(auto_fp16=_identity_decorator, force_fp32=_identity_decorator,
                        load_checkpoint=lambda *a, **k: None, get_dist_info=lambda: (0, 1)),
    "timm.models.layers": dict(DropPath=DropPath, to_2tuple=to_2tuple,
                               trunc_normal_=nn.init.trunc_normal_),
    "timm.models.registry": dict(register_model=lambda fn: fn),
    "timm.models.vision_transformer": dict(_cfg=lambda **k: dict(k)),
    "IPython": dict(embed=lambda *a, **k: None),
}


class _ShimModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _SHIM_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _ShimModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        for k, v in _REAL.get(module.__name__, {}).items():
            setattr(module, k, v)


_installed = False


def install():
    """Register the stand-ins and put the reference tree on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}; goldens can only be "
                           "regenerated in the build container")
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
