"""Hard voxelization oracle (numpy).

TEST INFRASTRUCTURE ONLY: the checker for tests/, __graft_entry__.smoke() and bench.py's CPU baseline / reference arm;
nothing under lidarseg3d_b200/ imports it and the product path has no CPU fallback.

Restates ``points_to_voxel`` / ``_points_to_voxel_reverse_kernel``
(reference det3d/ops/point_cloud/point_cloud_ops.py:7-55,112-184) as called by ``VoxelGenerator.generate``
(det3d/core/input/voxel_generator.py:5-30) from ``SegVoxelization`` (det3d/datasets/pipelines/
segpreprocess.py:148-177), plus ``collate_kitti``'s batch-index padding (det3d/torchie/parallel/collate.py:141-150).
"""
import numpy as np


def grid_size_of(voxel_size, pc_range):
    """round((hi-lo)/vs) in fp32 (point_cloud_ops.py:26-30, voxel_generator.py:9-11)."""
    vs = np.asarray(voxel_size, dtype=np.float32)
    rg = np.asarray(pc_range, dtype=np.float32)
    return np.round((rg[3:] - rg[:3]) / vs).astype(np.int32)


def voxel_coords_of(points, voxel_size, pc_range):
    """c_j = floor((p_j - lo_j) / vs_j) in fp32 with true division; valid iff 0 <= c_j < grid_j
    (point_cloud_ops.py:33-41).  Returns (coords_xyz int32 [N,3], valid bool [N])."""
    vs = np.asarray(voxel_size, dtype=np.float32)
    rg = np.asarray(pc_range, dtype=np.float32)
    grid = grid_size_of(voxel_size, pc_range)
    p = np.ascontiguousarray(points[:, :3], dtype=np.float32)
    c = np.floor((p - rg[None, :3]) / vs[None, :])  # fp32 sub, fp32 div, floor
    valid = np.all((c >= 0) & (c < grid[None, :].astype(np.float32)), axis=1)
    return c.astype(np.int64), valid


def points_to_voxel_loop(points, voxel_size, pc_range, max_points, max_voxels):
    """Literal sequential restatement (point_cloud_ops.py:31-55); slow, for small cases."""
    grid = grid_size_of(voxel_size, pc_range)
    coords, valid = voxel_coords_of(points, voxel_size, pc_range)
    F = points.shape[1]
    lut = {}
    voxels, coors, nump = [], [], []
    for i in range(points.shape[0]):
        if not valid[i]:
            continue
        key = (int(coords[i, 2]), int(coords[i, 1]), int(coords[i, 0]))  # reversed: z, y, x
        vid = lut.get(key, -1)
        if vid == -1:
            vid = len(voxels)
            if vid >= max_voxels:
                continue
            lut[key] = vid
            voxels.append(np.zeros((max_points, F), dtype=points.dtype))
            coors.append(key)
            nump.append(0)
        if nump[vid] < max_points:
            voxels[vid][nump[vid]] = points[i]
            nump[vid] += 1
    M = len(voxels)
    return (np.stack(voxels) if M else np.zeros((0, max_points, F), points.dtype),
            np.asarray(coors, dtype=np.int32).reshape(M, 3),
            np.asarray(nump, dtype=np.int32))


def points_to_voxel(points, voxel_size, pc_range, max_points=5, max_voxels=300000):
    """Vectorised equivalent of the sequential rule: voxel id = order of first appearance; first
    ``max_points`` points per voxel in arrival order; voxels beyond ``max_voxels`` (and their points) dropped."""
    points = np.ascontiguousarray(points)
    grid = grid_size_of(voxel_size, pc_range).astype(np.int64)
    coords, valid = voxel_coords_of(points, voxel_size, pc_range)
    idx = np.nonzero(valid)[0]
    c = coords[idx]
    lin = (c[:, 2] * grid[1] + c[:, 1]) * grid[0] + c[:, 0]
    uniq, first, inv = np.unique(lin, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique voxels by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    vid = rank[inv]                                     # voxel id per valid point
    keep = vid < max_voxels
    idx, vid, c = idx[keep], vid[keep], c[keep]
    M = int(min(order.size, max_voxels))
    # slot of each point inside its voxel = number of earlier points of the same voxel
    o2 = np.argsort(vid, kind="stable")
    sv = vid[o2]
    start = np.r_[0, np.nonzero(np.diff(sv))[0] + 1] if sv.size else np.zeros(0, np.int64)
    seg_start = np.zeros(sv.size, dtype=np.int64)
    seg_start[start] = start
    seg_start = np.maximum.accumulate(seg_start)
    slot = np.empty(sv.size, dtype=np.int64)
    slot[o2] = np.arange(sv.size) - seg_start
    F = points.shape[1]
    voxels = np.zeros((M, max_points, F), dtype=points.dtype)
    ok = slot < max_points
    voxels[vid[ok], slot[ok]] = points[idx[ok]]
    num = np.minimum(np.bincount(vid, minlength=M), max_points).astype(np.int32)
    coors = np.zeros((M, 3), dtype=np.int32)
    fo = first[order][:M]
    cc = coords[np.nonzero(valid)[0][fo]]
    coors[:, 0], coors[:, 1], coors[:, 2] = cc[:, 2], cc[:, 1], cc[:, 0]
    return voxels, coors, num


def collate_frames(frames):
    """collate_kitti's padding of the batch index (collate.py:141-150): list of
    (voxels, coordinates, num_points, points) -> batched arrays."""
    voxels = np.concatenate([f[0] for f in frames], 0)
    num_points = np.concatenate([f[2] for f in frames], 0)
    coords = np.concatenate([np.pad(f[1], ((0, 0), (1, 0)), constant_values=i) for i, f in enumerate(frames)], 0)
    points = np.concatenate([np.pad(f[3], ((0, 0), (1, 0)), constant_values=i) for i, f in enumerate(frames)], 0)
    num_voxels = np.asarray([f[0].shape[0] for f in frames], dtype=np.int64)
    return voxels, coords.astype(np.int32), num_points, num_voxels, points.astype(np.float32)
