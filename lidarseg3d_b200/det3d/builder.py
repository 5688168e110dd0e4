"""Builder entry points with the names and semantics of det3d/models/builder.py:19-63.

``build`` turns a cfg dict into an object through its registry; a list of cfg dicts becomes an ``nn.Sequential`` of the
built items (builder.py:19-24).  The public ``build_<kind>(cfg)`` functions are generated from one table so that the mapping
kind -> registry lives in a single place; ``build_detector`` additionally forwards ``train_cfg`` / ``test_cfg`` as default
constructor arguments (builder.py:62-63, called from tools/train.py:139 and tools/dist_test.py:128).
"""
from torch import nn

from . import registry as _reg
from .registry import build_from_cfg


def build(cfg, registry, default_args=None):
    if not isinstance(cfg, list):
        return build_from_cfg(cfg, registry, default_args)
    return nn.Sequential(*(build_from_cfg(item, registry, default_args) for item in cfg))


_KINDS = {
    "reader": _reg.READERS,
    "backbone": _reg.BACKBONES,
    "img_backbone": _reg.IMG_BACKBONES,
    "img_head": _reg.IMG_HEADS,
    "neck": _reg.NECKS,
    "head": _reg.HEADS,
    "loss": _reg.LOSSES,
    "point_head": _reg.POINT_HEADS,
    "roi_head": _reg.ROI_HEAD,
    "second_stage_module": _reg.SECOND_STAGE,
}


def _make_builder(kind, registry):
    def builder(cfg):
        return build(cfg, registry)

    builder.__name__ = builder.__qualname__ = "build_" + kind
    builder.__doc__ = "Build a {} from its cfg dict (or a list of them) through the {!r} registry.".format(kind, registry.name)
    return builder


for _kind, _registry in _KINDS.items():
    globals()["build_" + _kind] = _make_builder(_kind, _registry)
del _kind, _registry


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return build(cfg, _reg.DETECTORS, {"train_cfg": train_cfg, "test_cfg": test_cfg})


__all__ = ["build", "build_detector"] + ["build_" + k for k in _KINDS]
