"""SegNet / SegMSeg3DNet behind the DETECTORS registry (reference det3d/models/detectors/seg_net.py:12-107,
seg_mseg3d_net.py:8-147): same constructor, ``forward(example, return_loss)`` contract and ``example`` wire format."""
import torch
from torch import nn

from . import builder
from .registry import DETECTORS


def load_checkpoint(model, filename, strict=False):
    """det3d/torchie/trainer/checkpoint.py:122-173: {"state_dict": ...} or a raw state dict, optional module. prefix."""
    ck = torch.load(filename, map_location="cpu")
    sd = ck.get("state_dict", ck) if isinstance(ck, dict) else ck
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    model.load_state_dict(sd, strict=strict)
    return ck


class _SegBase(nn.Module):
    def __init__(self, reader, backbone, train_cfg=None, test_cfg=None):
        super().__init__()
        self.reader = builder.build_reader(reader)
        self.backbone = builder.build_backbone(backbone)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg

    def init_weights(self, pretrained=None):
        if pretrained is None:
            return
        try:
            load_checkpoint(self, pretrained, strict=False)
            print("init weight from {}".format(pretrained))
        except Exception:
            print("no pretrained model at {}".format(pretrained))

    def _lidar_branch(self, example):
        num_voxels = example["num_voxels"]
        batch_size = len(num_voxels)
        points = example["points"][:, 0:4]
        data = dict(features=example["voxels"], num_voxels=example["num_points"], voxel_coords=example["coordinates"],
                    batch_size=batch_size, input_shape=example["shape"][0], points=points)
        data["voxel_features"] = self.reader(data["features"], data["num_voxels"], data["voxel_coords"])
        return self.backbone(data)


@DETECTORS.register_module
class SegNet(_SegBase):
    def __init__(self, reader, backbone, point_head, neck=None, bbox_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None, **kwargs):
        super().__init__(reader, backbone, train_cfg, test_cfg)
        self.point_head = builder.build_point_head(point_head)
        self.init_weights(pretrained=pretrained)

    def forward(self, example, return_loss=True, **kwargs):
        if return_loss:
            raise NotImplementedError("lidarseg3d_b200: forward (inference) path only; call with return_loss=False")
        with torch.no_grad():
            data = self._lidar_branch(example)
            data = self.point_head(batch_dict=data, return_loss=False)
            self.last_batch_dict = data
            return self.point_head.predict(example=example, test_cfg=self.test_cfg)


@DETECTORS.register_module
class SegMSeg3DNet(_SegBase):
    def __init__(self, reader, backbone, img_backbone, img_head, point_head, neck=None, bbox_head=None, train_cfg=None,
                 test_cfg=None, pretrained=None, **kwargs):
        super().__init__(reader, backbone, train_cfg, test_cfg)
        self.img_backbone = builder.build_img_backbone(img_backbone)
        self.img_head = builder.build_img_head(img_head)
        self.point_head = builder.build_point_head(point_head)

    def forward(self, example, return_loss=True, **kwargs):
        if return_loss:
            raise NotImplementedError("lidarseg3d_b200: forward (inference) path only; call with return_loss=False")
        with torch.no_grad():
            batch_size = len(example["num_voxels"])
            images = example["images"]
            num_cams, hi, wi = images.shape[1], images.shape[3], images.shape[4]
            images = images.view(-1, 3, hi, wi).contiguous(memory_format=torch.channels_last)
            img_data = dict(inputs=self.img_backbone(images), batch_size=batch_size)
            img_data = self.img_head(batch_dict=img_data, return_loss=False)
            feats = img_data["image_features"]                                   # [B*ncam, C, ho, wo]
            _, c, ho, wo = feats.shape
            data = self._lidar_branch(example)
            data["points_cuv"] = example["points_cuv"]
            data["image_features"] = feats.view(batch_size, num_cams, c, ho, wo) if feats.is_contiguous() else \
                feats.reshape(batch_size, num_cams, c, ho, wo)
            data["_ls3d_image_features_nhwc"] = feats.permute(0, 2, 3, 1).contiguous().view(batch_size, num_cams, ho, wo, c)
            data["image_logits"] = img_data["image_logits"]
            data["camera_semantic_embeddings"] = img_data.get("camera_semantic_embeddings", None)
            data["metadata"] = example.get("metadata", None)
            data = self.point_head(batch_dict=data, return_loss=False)
            self.last_batch_dict = data
            return self.point_head.predict(example=example, test_cfg=self.test_cfg)
