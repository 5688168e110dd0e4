"""lidarseg3d_b200/tile_plan.py (device-side builder of the gather-once tile plan, groundwork for the next gather-GEMM
revision) against the numpy ground truth oracle/sparse.py::tile_plan, on CPU tensors."""
import numpy as np
import pytest
import torch

from lidarseg3d_b200 import tile_plan as tp
from oracle import sparse as osp


def _sites(seed, B, shape, n):
    rng = np.random.default_rng(seed)
    cells = rng.choice(B * shape[0] * shape[1] * shape[2], size=n, replace=False)
    D, H, W = shape
    return np.stack([cells // (D * H * W), (cells // (H * W)) % D, (cells // W) % H, cells % W], 1).astype(np.int32)


@pytest.mark.parametrize("n,use_morton", [(1500, False), (1500, True), (100, True), (128, False), (1, True)])
def test_builder_matches_oracle_plan(n, use_morton):
    shape = (12, 24, 24)
    idx = _sites(n, 2, shape, n)
    nbr = osp.subm_rulebook(idx, shape, 3)
    order = osp.morton_order(idx) if use_morton else None
    want = osp.tile_plan(nbr, order, tile=128)
    if use_morton:
        assert np.array_equal(tp.morton_order(torch.from_numpy(idx)).numpy(), order)
    got = tp.build(torch.from_numpy(nbr), None if order is None else torch.from_numpy(order))
    np.testing.assert_array_equal(got["out_rows"].numpy(), want["out_rows"])
    np.testing.assert_array_equal(got["stage_off"].numpy(), want["stage_off"])
    np.testing.assert_array_equal(got["stage_rows"].numpy(), want["stage_rows"])
    np.testing.assert_array_equal(got["local"].numpy().view(np.uint16), want["local"])


def test_strided_rulebook_plan():
    """input and output row sets differ (strided conv): n_in is the number of input rows."""
    shape = (9, 20, 20)
    idx = _sites(5, 1, shape, 900)
    oi, osh, nb = osp.strided_rulebook(idx, shape, 3, 2, 1)
    want = osp.tile_plan(nb, None, 128)
    got = tp.build(torch.from_numpy(nb), None, n_in=idx.shape[0])
    np.testing.assert_array_equal(got["stage_rows"].numpy(), want["stage_rows"])
    np.testing.assert_array_equal(got["local"].numpy().view(np.uint16), want["local"])
