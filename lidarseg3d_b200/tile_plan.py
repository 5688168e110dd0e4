"""Gather-once tile plan of a sparse convolution, built with device-side tensor ops (no host loop, no sync besides sizes).

Groundwork for the next gather-GEMM revision: NOT consumed by a kernel yet.  Today the gather-GEMM
fetches one input row per rulebook pair (5-15 per output row in LiDAR scenes); grouping the output rows into spatially
compact 128-row tiles (Morton order of (z, y, x) inside a frame) and listing each tile's DISTINCT input rows once lets a
kernel stage 1.3-2.2 rows per output row in shared memory and serve all kernel offsets from there
(scripts/analyze_gather_reuse.py).  A rulebook is cached per UNet level and reused by 6-9 convolutions, so the plan is built
once per level and frame batch.

The reference has no counterpart (spconv 1.x gathers per pair); the CPU ground truth of this structure
is oracle/sparse.py::tile_plan, compared in tests/test_tile_plan.py.
"""
import torch

TILE = 128
NONE = 0xFFFF


def _part1by2(v):
    v = v & 0x1FFFFF
    v = (v | (v << 32)) & 0x1F00000000FFFF
    v = (v | (v << 16)) & 0x1F0000FF0000FF
    v = (v | (v << 8)) & 0x100F00F00F00F00F
    v = (v | (v << 4)) & 0x10C30C30C30C30C3
    v = (v | (v << 2)) & 0x1249249249249249
    return v


def morton_order(coords):
    """coords [M, 4] int (b, z, y, x) -> permutation sorting by (b, Morton(z, y, x)); stable."""
    c = coords.long()
    code = _part1by2(c[:, 3]) | (_part1by2(c[:, 2]) << 1) | (_part1by2(c[:, 1]) << 2)
    # two stable passes = lexicographic (b, code)
    o1 = torch.sort(code, stable=True).indices
    o2 = torch.sort(c[o1, 0], stable=True).indices
    return o1[o2]


def build(nbr, out_order=None, n_in=None):
    """nbr [K, M_out] int32 (input row per kernel offset and output row, -1 = none), out_order [M_out] permutation or None.
    Returns dict(out_rows [T, 128] int32 (-1 = padding), stage_off [T + 1] int64, stage_rows [S] int32 ascending per tile,
    local [T, K, 128] int16 holding uint16 positions into the tile's stage, 0xFFFF = none)."""
    K, M = nbr.shape
    dev = nbr.device
    n_in = int(n_in) if n_in is not None else M
    T = (M + TILE - 1) // TILE
    order = torch.arange(M, device=dev) if out_order is None else out_order.long()
    out_rows = torch.full((T * TILE,), -1, dtype=torch.int64, device=dev)
    out_rows[:M] = order
    out_rows = out_rows.view(T, TILE)
    ok_slot = out_rows >= 0
    src = nbr.long()[:, out_rows.clamp(min=0)]                      # [K, T, TILE] input row of every (offset, tile slot)
    src = torch.where(ok_slot[None], src, torch.full_like(src, -1)).permute(1, 0, 2)      # [T, K, TILE]
    has = src >= 0
    tile_id = torch.arange(T, device=dev)[:, None, None].expand_as(src)
    key = tile_id[has] * n_in + src[has]                            # (tile, input row) of every pair
    ukey = torch.unique(key)                                        # sorted: by tile, then ascending row
    stage_rows = (ukey % n_in).to(torch.int32)
    counts = torch.bincount(ukey // n_in, minlength=T)
    stage_off = torch.zeros(T + 1, dtype=torch.int64, device=dev)
    stage_off[1:] = torch.cumsum(counts, 0)
    pos = torch.searchsorted(ukey, key) - stage_off[:-1][tile_id[has]]
    if pos.numel() and int(pos.max()) >= NONE:
        raise ValueError("a tile stages more than 65534 distinct rows")
    local = torch.full(src.shape, NONE, dtype=torch.int64, device=dev)
    local[has] = pos
    return dict(out_rows=out_rows.to(torch.int32), stage_off=stage_off, stage_rows=stage_rows,
                local=local.to(torch.int32).to(torch.int16))       # bit pattern of uint16 (0xFFFF -> -1)
