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
                  image_dtype=torch.float32, net_hw=None, calib=None):
    """frames: list of [N_i, F] fp32 (numpy or pinned/cuda tensors).  Returns the ``example`` dict on ``device``:
    voxels, coordinates (b,z,y,x), num_points, num_voxels, shape, points (b | features), [images, points_cuv, metadata].

    Camera inputs, all produced on the device from what the loader reads from disk:
      images_u8 [B, ncam, H, W, 3] uint8 - decoded camera images; resized to ``net_hw`` when that differs from (H, W)
        (cv2.resize semantics, img_transforms.py:78-99) and normalised (img_transforms.py:18-29);
      calib = list (one per frame) of dict(cams_from_global [ncam,4,4], intrinsics [ncam,3,3], ref_to_global [4,4] (optional:
        without it the first entry is cam_from_lidar), img_hw (H, W)) -> ``points_cuv`` by the loader's projection rules
        (loading.py:373-416, segpreprocess.py:654-671) unless ``points_cuv`` is given."""
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
        if net_hw is not None and tuple(int(v) for v in net_hw) != tuple(u8.shape[-3:-1]):
            ex["images"] = ops.resize_images_u8(u8, net_hw, img_mean, img_std, image_dtype)
        else:
            ex["images"] = ops.normalize_images_u8(u8, img_mean, img_std, image_dtype)
    elif images is not None:
        ex["images"] = torch.as_tensor(images).to(device, non_blocking=non_blocking)
    if points_cuv is not None:
        ex["points_cuv"] = torch.as_tensor(points_cuv).to(device, non_blocking=non_blocking)
    elif calib is not None:
        assert len(calib) == B and net_hw is not None
        cuv = []
        for i, c in enumerate(calib):
            cuv.append(ops.project_points(dev_frames[i] if B > 1 else pts, c["cams_from_global"], c["intrinsics"], c["img_hw"],
                                          net_hw, ref_to_global=c.get("ref_to_global")))
        ex["points_cuv"] = torch.cat(cuv, 0) if B > 1 else cuv[0]
    return ex


class HostStager:
    """Two-deep upload pipeline for loader output held in pinned host memory: ``stage(batch)`` enqueues the host->device
    copies of the NEXT batch on a side stream while the current one computes; ``take()`` makes the compute stream wait for
    them.  (The reference overlaps the same copies with DataLoader workers + non_blocking ``example_to_device``,
    det3d/torchie/apis/train.py:28-64.)"""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.pending = None

    def stage(self, batch):
        """batch: dict of pinned host tensors / lists of them (other values pass through)."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            out = {}
            for k, v in batch.items():
                if torch.is_tensor(v):
                    out[k] = v.to(self.device, non_blocking=True)
                elif isinstance(v, (list, tuple)) and v and torch.is_tensor(v[0]):
                    out[k] = [t.to(self.device, non_blocking=True) for t in v]
                else:
                    out[k] = v
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.pending = (out, ev)

    def take(self):
        out, ev = self.pending
        self.pending = None
        torch.cuda.current_stream(self.device).wait_event(ev)
        for v in out.values():                        # the consumer stream now owns the buffers (caching-allocator safety)
            for t in (v if isinstance(v, (list, tuple)) else [v]):
                if torch.is_tensor(t):
                    t.record_stream(torch.cuda.current_stream(self.device))
        return out
