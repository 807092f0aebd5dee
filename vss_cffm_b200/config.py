"""Loader for the reference's python config files (local_configs/cffm/**), unchanged.

Implements the subset of ``mmcv.Config.fromfile`` those files use:
python-file execution, recursive ``_base_`` list inheritance with dict merge, and
``_delete_=True`` overrides.  No ``{{_base_.x}}`` substitution, imports or lambdas occur in them.
"""
import os
import types


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(x):
    if isinstance(x, dict):
        return ConfigDict({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    return x


def _merge(new, base):
    """Values of ``new`` override ``base``; dicts merge recursively unless ``_delete_`` is set."""
    out = dict(base)
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(v, out[k])
        elif isinstance(v, dict):
            out[k] = {kk: vv for kk, vv in v.items() if kk != "_delete_"}
        else:
            out[k] = v
    return out


def _load_file(path):
    path = os.path.abspath(os.path.expanduser(path))
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    if not path.endswith(".py"):
        raise IOError("Only py type are supported by this loader")
    ns = {"__file__": path}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), ns)
    cfg = {k: v for k, v in ns.items()
           if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType, type))}
    bases = cfg.pop("_base_", None)
    if bases is None:
        return cfg
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        sub = _load_file(os.path.join(os.path.dirname(path), b))
        dup = set(merged) & set(sub)
        if dup:
            raise KeyError(f"Duplicate key is not allowed among bases: {sorted(dup)}")
        merged.update(sub)
    return _merge(cfg, merged)


class Config:
    def __init__(self, cfg_dict=None, filename=None):
        super().__setattr__("_cfg_dict", _wrap(cfg_dict or {}))
        super().__setattr__("_filename", filename)

    @staticmethod
    def fromfile(filename):
        return Config(_load_file(filename), filename)

    @property
    def filename(self):
        return self._filename

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, name, default=None):
        return self._cfg_dict.get(name, default)

    def keys(self):
        return self._cfg_dict.keys()

    def to_dict(self):
        return self._cfg_dict
