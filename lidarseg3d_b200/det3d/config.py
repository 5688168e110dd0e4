"""Config loader mirroring det3d.torchie.Config (det3d/torchie/utils/config.py:12-162): a config is a python
module; its public names become an attribute dict.  ``addict`` is not required: a minimal shim is installed so
that the reference configs' ``from addict.addict import Dict`` resolves."""
import os.path as osp
import sys
import types
from importlib import import_module


class ConfigDict(dict):
    """Attribute-style dict (addict.Dict subset): nested dicts are wrapped, missing attr -> AttributeError."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError("'{}' object has no attribute '{}'".format(self.__class__.__name__, name))

    def __setattr__(self, name, value):
        self[name] = value

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def copy(self):
        return ConfigDict(self)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}


def install_addict_shim():
    if "addict" in sys.modules:
        return
    try:
        import addict  # noqa: F401
        return
    except ImportError:
        pass
    pkg = types.ModuleType("addict")
    sub = types.ModuleType("addict.addict")
    pkg.Dict = sub.Dict = ConfigDict
    pkg.addict = sub
    sys.modules["addict"] = pkg
    sys.modules["addict.addict"] = sub


class Config(object):
    @staticmethod
    def fromfile(filename):
        filename = osp.abspath(osp.expanduser(filename))
        if not osp.isfile(filename):
            raise FileNotFoundError('file "{}" does not exist'.format(filename))
        if not filename.endswith(".py"):
            raise IOError("Only py type is supported now!")
        module_name = osp.basename(filename)[:-3]
        if "." in module_name:
            raise ValueError("Dots are not allowed in config file path.")
        install_addict_shim()
        config_dir = osp.dirname(filename)
        sys.path.insert(0, config_dir)
        try:
            sys.modules.pop(module_name, None)
            mod = import_module(module_name)
        finally:
            sys.path.pop(0)
        cfg_dict = {k: v for k, v in mod.__dict__.items() if not k.startswith("__") and not isinstance(v, types.ModuleType)}
        return Config(cfg_dict, filename=filename)

    def __init__(self, cfg_dict=None, filename=None):
        if cfg_dict is None:
            cfg_dict = dict()
        elif not isinstance(cfg_dict, dict):
            raise TypeError("cfg_dict must be a dict, but got {}".format(type(cfg_dict)))
        object.__setattr__(self, "_cfg_dict", ConfigDict({k: v for k, v in cfg_dict.items()
                                                          if not callable(v) or isinstance(v, dict)}))
        object.__setattr__(self, "_filename", filename)
        text = ""
        if filename:
            with open(filename, "r") as f:
                text = f.read()
        object.__setattr__(self, "_text", text)

    @property
    def filename(self):
        return self._filename

    @property
    def text(self):
        return self._text

    def __repr__(self):
        return "Config (path: {}): {}".format(self.filename, self._cfg_dict.__repr__())

    def __len__(self):
        return len(self._cfg_dict)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    def __setitem__(self, name, value):
        self._cfg_dict[name] = value

    def __iter__(self):
        return iter(self._cfg_dict)

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)
