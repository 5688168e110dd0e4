"""Name -> class plugin registry with the behaviour of the reference's det3d.utils.registry (det3d/utils/registry.py:6-78)
and the model registries of det3d/models/registry.py:3-16.

Contract kept:
  * ``@REG.register_module`` on a class stores it under ``cls.__name__`` and returns the class; a non-class raises
    ``TypeError``; a duplicate name raises ``KeyError("<name> is already registered in <registry>")``.
  * ``REG.get(name)`` returns the class or ``None``; ``REG.name`` / ``REG.module_dict`` are read-only views.
  * ``build_from_cfg(cfg, registry, default_args)``: ``cfg["type"]`` is a registered name or a class, the remaining keys are
    constructor kwargs, ``default_args`` only fill keys that are missing; an unknown name raises
    ``KeyError("<type> is not in the <registry> registry")``, any other ``type`` value raises ``TypeError``.
"""
import inspect


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def __repr__(self):
        return "{}(name={}, items={})".format(type(self).__name__, self._name, list(self._module_dict))

    def get(self, key):
        return self._module_dict.get(key)

    def _register_module(self, module_class):
        if not inspect.isclass(module_class):
            raise TypeError("module must be a class, but got {}".format(type(module_class)))
        key = module_class.__name__
        if key in self._module_dict:
            raise KeyError("{} is already registered in {}".format(key, self._name))
        self._module_dict[key] = module_class

    def register_module(self, cls):
        self._register_module(cls)
        return cls


def _resolve(obj_type, registry):
    if inspect.isclass(obj_type):
        return obj_type
    if not isinstance(obj_type, str):
        raise TypeError("type must be a str or valid type, but got {}".format(type(obj_type)))
    found = registry.get(obj_type)
    if found is None:
        raise KeyError("{} is not in the {} registry".format(obj_type, registry.name))
    return found


def build_from_cfg(cfg, registry, default_args=None):
    assert isinstance(cfg, dict) and "type" in cfg
    assert default_args is None or isinstance(default_args, dict)
    kwargs = {k: v for k, v in cfg.items() if k != "type"}
    cls = _resolve(cfg["type"], registry)
    for k, v in (default_args or {}).items():
        kwargs.setdefault(k, v)
    return cls(**kwargs)


# det3d/models/registry.py:3-16
(READERS, BACKBONES, IMG_BACKBONES, IMG_HEADS, NECKS, HEADS, LOSSES, DETECTORS, SECOND_STAGE, ROI_HEAD,
 POINT_HEADS) = (Registry(n) for n in ("reader", "backbone", "img_backbone", "img_head", "neck", "head", "loss", "detector",
                                       "second_stage", "roi_head", "point_head"))
