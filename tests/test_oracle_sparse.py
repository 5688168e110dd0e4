"""Sparse-conv oracle (spconv 1.x semantics, parity unpinned against spconv itself) pinned to the independent dense
ground truth F.conv3d / F.conv_transpose3d and to an O(N*K) dictionary enumeration of the rulebook."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sparse as osp


def random_sites(seed, B, shape, n):
    rng = np.random.default_rng(seed)
    cells = rng.choice(B * shape[0] * shape[1] * shape[2], size=n, replace=False)
    rng.shuffle(cells)
    D, H, W = shape
    return np.stack([cells // (D * H * W), (cells // (H * W)) % D, (cells // W) % H, cells % W], 1).astype(np.int32)


def densify(idx, feats, B, shape):
    d = torch.zeros(B, feats.shape[1], *shape, dtype=feats.dtype)
    d[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = feats
    return d


@pytest.mark.parametrize("seed", [0, 1])
def test_subm_matches_dense_conv(seed):
    B, shape, C, Co = 2, (9, 16, 16), 6, 5
    idx = random_sites(seed, B, shape, 600)
    feats = torch.randn(idx.shape[0], C, dtype=torch.float64)
    w = torch.randn(3, 3, 3, C, Co, dtype=torch.float64)
    nbr = osp.subm_rulebook(idx, shape, 3)
    out = osp.sparse_conv(feats, w, nbr)
    dense = F.conv3d(densify(idx, feats, B, shape), w.permute(4, 3, 0, 1, 2), padding=1)
    ref = dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
    torch.testing.assert_close(out, ref, rtol=1e-10, atol=1e-10)
    _, pairs = osp.rulebook_dict(idx, shape, 3, 1, 1, subm=True)
    assert osp.pairs_of(nbr) == pairs


@pytest.mark.parametrize("ks,st,pd", [((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)),
                                      ((3, 1, 1), (2, 1, 1), (0, 0, 0))])
def test_strided_and_inverse_match_dense(ks, st, pd):
    B, shape, C, Co = 2, (11, 16, 16), 4, 3
    idx = random_sites(3, B, shape, 500)
    feats = torch.randn(idx.shape[0], C, dtype=torch.float64)
    w = torch.randn(*ks, C, Co, dtype=torch.float64)
    oidx, oshape, nbr = osp.strided_rulebook(idx, shape, ks, st, pd)
    # ascending linear order of the output sites
    lin = ((oidx[:, 0].astype(np.int64) * oshape[0] + oidx[:, 1]) * oshape[1] + oidx[:, 2]) * oshape[2] + oidx[:, 3]
    assert np.all(np.diff(lin) > 0)
    out = osp.sparse_conv(feats, w, nbr)
    dense = F.conv3d(densify(idx, feats, B, shape), w.permute(4, 3, 0, 1, 2), stride=st, padding=pd)
    assert tuple(dense.shape[2:]) == tuple(oshape)
    ref = dense[oidx[:, 0], :, oidx[:, 1], oidx[:, 2], oidx[:, 3]]
    torch.testing.assert_close(out, ref, rtol=1e-10, atol=1e-10)
    # active output set = every cell with at least one active input in its receptive field
    occ = F.conv3d(densify(idx, torch.ones(idx.shape[0], 1, dtype=torch.float64), B, shape),
                   torch.ones(1, 1, *ks, dtype=torch.float64), stride=st, padding=pd)
    assert int((occ > 0).sum()) == oidx.shape[0]
    olist, pairs = osp.rulebook_dict(idx, shape, ks, st, pd, subm=False)
    assert [tuple(r) for r in oidx.tolist()] == olist and osp.pairs_of(nbr) == pairs
    # inverse conv == conv_transpose3d restricted to the original fine sites
    cf = torch.randn(oidx.shape[0], Co, dtype=torch.float64)
    wi = torch.randn(*ks, Co, C, dtype=torch.float64)
    nbr_up = osp.invert_rulebook(nbr, idx.shape[0])
    fine = osp.sparse_conv(cf, wi, nbr_up)
    opad = [shape[i] - ((oshape[i] - 1) * st[i] - 2 * pd[i] + ks[i]) for i in range(3)]
    dt = F.conv_transpose3d(densify(oidx, cf, B, oshape), wi.permute(3, 4, 0, 1, 2), stride=st, padding=pd,
                            output_padding=opad)
    ref = dt[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
    torch.testing.assert_close(fine, ref, rtol=1e-10, atol=1e-10)


def test_three_nn_ties_and_small_sets():
    from oracle import nets as on
    known = torch.tensor([[0., 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]])
    d2, idx = on.three_nn(torch.tensor([[0., 0, 0]]), known)
    assert idx.tolist() == [[0, 1, 2]] and d2.tolist() == [[0.0, 1.0, 1.0]]        # ties -> lowest index
    d2, idx = on.three_nn(torch.tensor([[0.5, 0, 0]]), known[:2])
    assert idx.tolist() == [[0, 1, 0]] and np.isinf(d2[0, 2].item())


# ---------------------------------------------------------------------------------------------- property tests
from hypothesis import given, settings, strategies as hst   # noqa: E402


@settings(max_examples=40, deadline=None)
@given(seed=hst.integers(0, 10 ** 6), n=hst.integers(0, 60), D=hst.integers(1, 6), H=hst.integers(1, 7), W=hst.integers(1, 7),
       ks=hst.tuples(hst.sampled_from([1, 3]), hst.sampled_from([1, 3]), hst.sampled_from([1, 3])),
       st=hst.tuples(hst.integers(1, 2), hst.integers(1, 2), hst.integers(1, 2)),
       pd=hst.tuples(hst.integers(0, 1), hst.integers(0, 1), hst.integers(0, 1)), B=hst.integers(1, 2))
def test_rulebooks_equal_dictionary_enumeration(seed, n, D, H, W, ks, st, pd, B):
    """Vectorised rulebook builders (numpy oracle and its torch twin) == the independent O(N*K) dictionary enumeration for
    random small grids, kernel shapes, strides and paddings - including empty inputs, 1-cell grids and fully dense grids."""
    from oracle import torch_backend as tb
    rng = np.random.default_rng(seed)
    cells = B * D * H * W
    n = min(n, cells)
    flat = rng.choice(cells, size=n, replace=False)
    idx = np.stack([flat // (D * H * W), (flat // (H * W)) % D, (flat // W) % H, flat % W], 1).astype(np.int32).reshape(-1, 4)
    shape = (D, H, W)
    osh = osp.out_shape(shape, ks, st, pd)
    if min(osh) < 1:
        return
    # SubM (only defined for odd kernels; spconv pads k // 2)
    want_out, want_pairs = osp.rulebook_dict(idx, shape, ks, (1, 1, 1), tuple(k // 2 for k in ks), subm=True)
    nbr = osp.subm_rulebook(idx, shape, ks)
    assert osp.pairs_of(nbr) == want_pairs
    if n:
        np.testing.assert_array_equal(tb.subm_rulebook(torch.from_numpy(idx), shape, ks).numpy(), nbr)
    if n == 0:
        return
    # strided conv: output sites in ascending linear order, pairs as enumerated
    want_out, want_pairs = osp.rulebook_dict(idx, shape, ks, st, pd, subm=False)
    oi, oshape, nb = osp.strided_rulebook(idx, shape, ks, st, pd)
    assert tuple(oshape) == tuple(osh)
    assert [tuple(int(v) for v in r) for r in oi] == [tuple(r) for r in want_out]
    assert osp.pairs_of(nb) == want_pairs
    oi_t, _, nb_t = tb.strided_rulebook(torch.from_numpy(idx), shape, ks, st, pd)
    np.testing.assert_array_equal(oi_t.numpy(), oi)
    np.testing.assert_array_equal(nb_t.numpy(), nb)
    # the inverse table is the exact swap
    up = osp.invert_rulebook(nb, n)
    assert {(k, j, i) for (k, i, j) in osp.pairs_of(nb)} == osp.pairs_of(up)


@pytest.mark.parametrize("use_morton", [False, True])
def test_tile_plan_gather_once_equals_per_pair_conv(use_morton):
    """The gather-once tile plan (design prototype of the next gather-GEMM revision) reproduces the per-pair sparse
    convolution exactly and lists every rulebook pair once."""
    rng = np.random.default_rng(7)
    shape, n, C, Co = (12, 24, 24), 1500, 8, 6
    cells = rng.choice(2 * shape[0] * shape[1] * shape[2], size=n, replace=False)
    D, H, W = shape
    idx = np.stack([cells // (D * H * W), (cells // (H * W)) % D, (cells // W) % H, cells % W], 1).astype(np.int32)
    nbr = osp.subm_rulebook(idx, shape, 3)
    order = osp.morton_order(idx) if use_morton else None
    plan = osp.tile_plan(nbr, order, tile=128)
    # every pair exactly once
    pairs = set()
    for t in range(plan["out_rows"].shape[0]):
        rows = plan["stage_rows"][plan["stage_off"][t]:plan["stage_off"][t + 1]]
        k, s = np.nonzero(plan["local"][t] != 0xFFFF)
        for kk, ss in zip(k.tolist(), s.tolist()):
            p = (kk, int(rows[plan["local"][t, kk, ss]]), int(plan["out_rows"][t, ss]))
            assert p not in pairs
            pairs.add(p)
    assert pairs == osp.pairs_of(nbr)
    g = torch.Generator().manual_seed(3)
    f = torch.randn(n, C, generator=g)
    w = torch.randn(27, C, Co, generator=g)
    torch.testing.assert_close(osp.sparse_conv_tiled(f, w, plan, n), osp.sparse_conv(f, w, nbr), rtol=1e-5, atol=1e-5)
    if use_morton:      # compact tiles stage fewer rows than row-order tiles
        assert plan["stage_rows"].size < osp.tile_plan(nbr, None, 128)["stage_rows"].size
