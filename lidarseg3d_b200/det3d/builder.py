"""Builders mirroring det3d/models/builder.py:19-63 (a list cfg becomes nn.Sequential)."""
from torch import nn

from .registry import (BACKBONES, DETECTORS, HEADS, IMG_BACKBONES, IMG_HEADS, LOSSES, NECKS, POINT_HEADS, READERS,
                       ROI_HEAD, SECOND_STAGE, build_from_cfg)


def build(cfg, registry, default_args=None):
    if isinstance(cfg, list):
        return nn.Sequential(*[build_from_cfg(c, registry, default_args) for c in cfg])
    return build_from_cfg(cfg, registry, default_args)


def build_second_stage_module(cfg):
    return build(cfg, SECOND_STAGE)


def build_roi_head(cfg):
    return build(cfg, ROI_HEAD)


def build_reader(cfg):
    return build(cfg, READERS)


def build_backbone(cfg):
    return build(cfg, BACKBONES)


def build_img_backbone(cfg):
    return build(cfg, IMG_BACKBONES)


def build_img_head(cfg):
    return build(cfg, IMG_HEADS)


def build_neck(cfg):
    return build(cfg, NECKS)


def build_head(cfg):
    return build(cfg, HEADS)


def build_loss(cfg):
    return build(cfg, LOSSES)


def build_point_head(cfg):
    return build(cfg, POINT_HEADS)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return build(cfg, DETECTORS, dict(train_cfg=train_cfg, test_cfg=test_cfg))
