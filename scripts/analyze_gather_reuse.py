"""CPU analysis (no GPU): how many input rows a 128-row output tile of a SubM 3x3x3 convolution gathers
  (a) today: once per (row, kernel offset) pair,
  (b) with rows in sorted (b, z, y, x) order and one staged run per (kz, ky) shared by the three kx offsets,
  (c) lower bound: every distinct input row of the tile once,
on the synthetic LiDAR scans of the bench (level 1 of the UNet).  Motivates the "gather less" item of DESIGN.md section 8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidarseg3d_b200 import synth
from oracle import sparse as osp
from oracle import voxelize as ov

for name in ("NUSC", "KITTI"):
    spec = getattr(synth, name)
    frames = [synth.lidar_scan(spec, s) for s in range(2)]
    vox = [ov.points_to_voxel(f, spec["voxel_size"], spec["pc_range"], 5, 300000) for f in frames]
    v, c, n, nv, pts = ov.collate_frames([(a, b, cc, f) for (a, b, cc), f in zip(vox, frames)])
    g = synth.grid_shape(spec)
    shape = (int(g[2]) + 1, int(g[1]), int(g[0]))
    for order_name in ("first-seen (reference order)", "sorted (b,z,y,x)"):
        idx = c.astype(np.int64)
        if order_name.startswith("sorted"):
            key = ((idx[:, 0] * shape[0] + idx[:, 1]) * shape[1] + idx[:, 2]) * shape[2] + idx[:, 3]
            idx = idx[np.argsort(key, kind="stable")]
        nbr = osp.subm_rulebook(idx.astype(np.int32), shape, 3)          # [27, M]
        M = nbr.shape[1]
        pairs = int((nbr >= 0).sum())
        per_run, per_tile, steps_now, steps_run = 0, 0, 0, 0
        for t0 in range(0, M, 128):
            blk = nbr[:, t0:t0 + 128]
            allrows = blk[blk >= 0]
            per_tile += np.unique(allrows).size
            steps_now += int((blk >= 0).any(axis=1).sum())
            for zy in range(9):
                sub = blk[3 * zy:3 * zy + 3]
                rows = sub[sub >= 0]
                if rows.size:
                    per_run += np.unique(rows).size
                    steps_run += 1
        print(f"{name:6s} {order_name:30s} rows {M:7d} pairs/row {pairs / M:5.2f} | gathered rows per output row: "
              f"per pair {pairs / M:5.2f}, per (kz,ky) run {per_run / M:5.2f}, distinct per tile {per_tile / M:5.2f} | "
              f"steps/tile now {steps_now / (M / 128):5.1f}, runs/tile {steps_run / (M / 128):4.1f}")


# ---- all four UNet levels: distinct input rows per 128-row tile (sizes the shared-memory stage of the next engine revision)
print()
for name in ("NUSC", "KITTI"):
    spec = getattr(synth, name)
    frames = [synth.lidar_scan(spec, s) for s in range(2)]
    vox = [ov.points_to_voxel(f, spec["voxel_size"], spec["pc_range"], 5, 300000) for f in frames]
    v, c, n, nv, pts = ov.collate_frames([(a, b, cc, f) for (a, b, cc), f in zip(vox, frames)])
    g = synth.grid_shape(spec)
    shape = (int(g[2]) + 1, int(g[1]), int(g[0]))
    idx = c.astype(np.int32)
    pads = {2: (1, 1, 1), 3: (1, 1, 1), 4: (0, 1, 1)}
    chans = {1: 32, 2: 64, 3: 128, 4: 128}
    for lv in (1, 2, 3, 4):
        if lv > 1:
            idx, shape, _ = osp.strided_rulebook(idx, shape, 3, 2, pads[lv])
        nbr = osp.subm_rulebook(idx, shape, 3)
        M = nbr.shape[1]
        d = []
        for t0 in range(0, M, 128):
            blk = nbr[:, t0:t0 + 128]
            d.append(np.unique(blk[blk >= 0]).size)
        d = np.asarray(d)
        print(f"{name:6s} level {lv}: rows {M:6d} tiles {d.size:4d} pairs/row {(nbr >= 0).sum() / M:5.2f} distinct rows per tile: mean "
              f"{d.mean():6.1f} p95 {np.percentile(d, 95):6.1f} max {d.max():4d} -> {d.max() * chans[lv] * 4 / 1024:6.1f} KB fp32 at C={chans[lv]}")


# ---- same statistic with rows grouped into spatially compact tiles (Morton order of (z, y, x) inside a frame)
def _part1by2(v):
    v = v.astype(np.uint64) & np.uint64(0x1fffff)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


print()
for name in ("NUSC", "KITTI"):
    spec = getattr(synth, name)
    frames = [synth.lidar_scan(spec, s) for s in range(2)]
    vox = [ov.points_to_voxel(f, spec["voxel_size"], spec["pc_range"], 5, 300000) for f in frames]
    v, c, n, nv, pts = ov.collate_frames([(a, b, cc, f) for (a, b, cc), f in zip(vox, frames)])
    g = synth.grid_shape(spec)
    shape = (int(g[2]) + 1, int(g[1]), int(g[0]))
    idx = c.astype(np.int32)
    pads = {2: (1, 1, 1), 3: (1, 1, 1), 4: (0, 1, 1)}
    for lv in (1, 2, 3, 4):
        if lv > 1:
            idx, shape, _ = osp.strided_rulebook(idx, shape, 3, 2, pads[lv])
        i64 = idx.astype(np.int64)
        morton = (_part1by2(i64[:, 3]) | (_part1by2(i64[:, 2]) << np.uint64(1)) | (_part1by2(i64[:, 1]) << np.uint64(2)))
        order = np.lexsort((morton, i64[:, 0]))
        nbr = osp.subm_rulebook(idx[order], shape, 3)
        M = nbr.shape[1]
        d = np.asarray([np.unique(nbr[:, t0:t0 + 128][nbr[:, t0:t0 + 128] >= 0]).size for t0 in range(0, M, 128)])
        print(f"{name:6s} level {lv} Morton tiles: distinct rows per tile mean {d.mean():6.1f} p95 {np.percentile(d, 95):6.1f} "
              f"max {d.max():4d} ({d.mean() / 128:4.2f} per output row)")
