"""On-device construction of the reference's ``example`` wire format.

The reference voxelizes on the CPU inside DataLoader workers (SegVoxelization -> collate_kitti,
det3d/datasets/pipelines/segpreprocess.py:148-177, det3d/torchie/parallel/collate.py:91-170); here raw frames go
to the GPU once and ``ls3d_voxelize`` produces the same tensors (bit-exact) for the whole batch.
"""
import numpy as np
import torch

from . import ops


def build_example(frames, voxel_size, pc_range, max_points=5, max_voxels=300000, images=None, points_cuv=None,
                  metadata=None, device="cuda", non_blocking=True, images_u8=None, img_mean=None, img_std=None,
                  image_dtype=torch.float32):
    """frames: list of [N_i, F] fp32 (numpy or pinned/cuda tensors).  Returns the ``example`` dict on ``device``:
    voxels, coordinates (b,z,y,x), num_points, num_voxels, shape, points (b | features), [images, points_cuv, metadata]."""
    offs = [0]
    dev_frames = []
    for f in frames:
        t = torch.as_tensor(f)
        dev_frames.append(t.to(device, non_blocking=non_blocking))
        offs.append(offs[-1] + t.shape[0])
    pts = torch.cat(dev_frames, 0).float().contiguous() if len(dev_frames) > 1 else dev_frames[0].float().contiguous()
    B = len(frames)
    vox = ops.voxelize(pts, offs, voxel_size, pc_range, max_points, max_voxels)
    bidx = torch.repeat_interleave(torch.arange(B, device=pts.device, dtype=torch.float32),
                                   torch.tensor(np.diff(offs), device=pts.device))
    grid = np.round((np.asarray(pc_range[3:], np.float32) - np.asarray(pc_range[:3], np.float32))
                    / np.asarray(voxel_size, np.float32)).astype(np.int64)
    ex = dict(voxels=vox["voxels"], coordinates=vox["coordinates"], num_points=vox["num_points"],
              num_voxels=vox["num_voxels"], shape=np.stack([grid] * B), points=torch.cat([bidx[:, None], pts], 1),
              metadata=metadata if metadata is not None else [dict(token=i) for i in range(B)])
    if images_u8 is not None:
        # raw resized camera images [B, ncam, H, W, 3] uint8: 1 byte per value over PCIe, normalised on the device
        # (segpreprocess.py:621-637 on the loader's CPU in the reference) straight into the stem's channels-last layout
        u8 = torch.as_tensor(images_u8).to(device, non_blocking=non_blocking)
        ex["images"] = ops.normalize_images_u8(u8, img_mean, img_std, image_dtype)
    elif images is not None:
        ex["images"] = torch.as_tensor(images).to(device, non_blocking=non_blocking)
    if points_cuv is not None:
        ex["points_cuv"] = torch.as_tensor(points_cuv).to(device, non_blocking=non_blocking)
    return ex
