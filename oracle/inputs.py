"""The reference's input pipeline for one batch, on the CPU: what its DataLoader workers hand to ``model(example)``.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's parity block and CPU arm).  Chains the pinned oracle pieces:
voxelizer (oracle/voxelize.py <- point_cloud_ops.py:112-184 + collate.py:141-150), projection / resize / coordinate
normalisation (oracle/camera.py <- loading.py:361-416, img_transforms.py:78-99, segpreprocess.py:654-671) and the image
normalisation (oracle/nets.py::image_input_transform <- img_transforms.py:18-29)."""
import numpy as np
import torch

from . import camera as oc
from . import nets as on
from . import voxelize as ov


def grid_shape(voxel_size, pc_range):
    vs = np.asarray(voxel_size, np.float32)
    rg = np.asarray(pc_range, np.float32)
    return np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)


def cpu_example(frames, voxel_size, pc_range, calib=None, images_u8=None, net_hw=None, img_mean=None, img_std=None,
                max_points=5, max_voxels=300000):
    """frames: list of [N_i, F] fp32 numpy; calib: list of dict(ref_to_global, cams_from_global, intrinsics, img_hw);
    images_u8 [B, ncam, H, W, 3] uint8 raw camera images.  Returns the ``example`` dict of CPU tensors."""
    frames = [np.asarray(f, np.float32) for f in frames]
    vox = [ov.points_to_voxel(f, voxel_size, pc_range, max_points, max_voxels) for f in frames]
    v, c, n, nv, pts = ov.collate_frames([(a, b, cc, f) for (a, b, cc), f in zip(vox, frames)])
    B = len(frames)
    ex = dict(voxels=torch.from_numpy(v), coordinates=torch.from_numpy(c), num_points=torch.from_numpy(n),
              num_voxels=torch.from_numpy(nv), shape=np.stack([grid_shape(voxel_size, pc_range)] * B), points=torch.from_numpy(pts),
              metadata=[dict(token=i) for i in range(B)])
    if calib is not None:
        ex["points_cuv"] = torch.from_numpy(np.concatenate([
            oc.project_points(f[:, :3], cb["ref_to_global"], cb["cams_from_global"], cb["intrinsics"], cb["img_hw"], net_hw)
            for f, cb in zip(frames, calib)]))
    if images_u8 is not None:
        raw = np.asarray(images_u8)
        nh, nw = int(net_hw[0]), int(net_hw[1])
        if raw.shape[-3:-1] != (nh, nw):
            raw = np.stack([np.stack([oc.resize_bilinear_u8(im, (nw, nh)) for im in fr]) for fr in raw])
        ex["images_u8"] = torch.from_numpy(raw)
        ex["images"] = torch.from_numpy(on.image_input_transform(raw, img_mean, img_std))
    return ex
