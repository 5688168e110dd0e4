"""Host-side mirror of the reference's det3d plugin API for the segmentation forward path.

``install_alias()`` publishes this package under the reference's import names (``det3d.utils``, ``det3d.models``,
``det3d.torchie``) so code written against the reference (``from det3d.models import build_detector``;
``from det3d.torchie import Config``) resolves to the B200-native implementation.
"""
import sys
import types

from . import builder, registry
from .builder import (build_backbone, build_detector, build_img_backbone, build_img_head, build_point_head,  # noqa: F401
                      build_reader)
from .config import Config, ConfigDict, install_addict_shim  # noqa: F401
from .registry import (BACKBONES, DETECTORS, IMG_BACKBONES, IMG_HEADS, POINT_HEADS, READERS, Registry,  # noqa: F401
                       build_from_cfg)
from . import readers, backbones, img_backbones, img_heads, point_heads, detectors  # noqa: F401,E402  (registration)


def install_alias(name="det3d"):
    """Register ``det3d``-named module aliases (no-op when a real det3d is already imported)."""
    if name in sys.modules and not getattr(sys.modules[name], "_ls3d_alias", False):
        return sys.modules[name]
    me = sys.modules[__name__]
    root = types.ModuleType(name)
    root._ls3d_alias = True
    utils = types.ModuleType(name + ".utils")
    utils.Registry, utils.build_from_cfg = Registry, build_from_cfg
    utils.registry = registry
    models = types.ModuleType(name + ".models")
    for k in dir(builder):
        if k.startswith("build"):
            setattr(models, k, getattr(builder, k))
    for k in dir(registry):
        if k.isupper():
            setattr(models, k, getattr(registry, k))
    models.builder, models.registry = builder, registry
    torchie = types.ModuleType(name + ".torchie")
    torchie.Config, torchie.ConfigDict = Config, ConfigDict
    root.utils, root.models, root.torchie = utils, models, torchie
    root.impl = me
    sys.modules[name] = root
    sys.modules[name + ".utils"] = utils
    sys.modules[name + ".utils.registry"] = registry
    sys.modules[name + ".models"] = models
    sys.modules[name + ".models.builder"] = builder
    sys.modules[name + ".models.registry"] = registry
    sys.modules[name + ".torchie"] = torchie
    return root
