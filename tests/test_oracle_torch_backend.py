"""oracle/torch_backend.py (the torch-only restatement used for the spconv-style GPU baseline) against the numpy ground truth
oracle/sparse.py and oracle/nets.three_nn on the CPU."""
import numpy as np
import pytest
import torch

from oracle import nets as on
from oracle import sparse as osp
from oracle import torch_backend as tb


def _sites(seed, B, shape, n):
    rng = np.random.default_rng(seed)
    cells = rng.choice(B * shape[0] * shape[1] * shape[2], size=n, replace=False)
    rng.shuffle(cells)
    D, H, W = shape
    return np.stack([cells // (D * H * W), (cells // (H * W)) % D, (cells // W) % H, cells % W], 1).astype(np.int32)


@pytest.mark.parametrize("shape,n,ks,st,pd", [((9, 20, 20), 900, 3, 2, 1), ((11, 16, 12), 700, (3, 1, 1), (2, 1, 1), 0),
                                              ((8, 14, 14), 500, 3, 2, (0, 1, 1))])
def test_rulebooks_match_numpy_oracle(shape, n, ks, st, pd):
    idx = _sites(3, 2, shape, n)
    t = torch.from_numpy(idx)
    np.testing.assert_array_equal(tb.subm_rulebook(t, shape, 3).numpy(), osp.subm_rulebook(idx, shape, 3))
    oi, osh, nb = osp.strided_rulebook(idx, shape, ks, st, pd)
    oi_t, osh_t, nb_t = tb.strided_rulebook(t, shape, ks, st, pd)
    assert tuple(osh_t) == tuple(osh)
    np.testing.assert_array_equal(oi_t.numpy(), oi)
    np.testing.assert_array_equal(nb_t.numpy(), nb)
    np.testing.assert_array_equal(tb.invert_rulebook(nb_t, n).numpy(), osp.invert_rulebook(nb, n))


def test_sparse_conv_and_unet_match_numpy_oracle():
    shape, n, C, Co = (9, 16, 16), 600, 8, 12
    idx = _sites(5, 1, shape, n)
    g = torch.Generator().manual_seed(0)
    f = torch.randn(n, C, generator=g)
    w = torch.randn(3, 3, 3, C, Co, generator=g)
    nbr = osp.subm_rulebook(idx, shape, 3)
    ref = osp.sparse_conv(f, w, nbr)
    out = tb.sparse_conv(f, w, tb.subm_rulebook(torch.from_numpy(idx), shape, 3))
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)


def test_unet_forward_same_through_both_backends():
    """The full UNetSCN3D wiring of oracle/nets.py gives the same features with the numpy and the torch backend."""
    from lidarseg3d_b200.det3d import Config, build_detector
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, "configs", "sdseg3d_semantickitti.py"))
    torch.manual_seed(0)
    m = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    shape_xyz = (48, 40, 16)
    idx = _sites(9, 2, (16, 40, 48), 1500)
    g = torch.Generator().manual_seed(1)
    vf = torch.randn(idx.shape[0], 16, generator=g)
    vs, rg = [0.1, 0.1, 0.15], [-2.4, -2.0, -1.2, 2.4, 2.0, 1.2]
    a, ca = on.unet_scn3d(sd, "backbone.", vf, torch.from_numpy(idx), shape_xyz, vs, rg)
    b, cb = on.unet_scn3d(sd, "backbone.", vf, torch.from_numpy(idx), shape_xyz, vs, rg, backend=tb)
    torch.testing.assert_close(b, a, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(cb, ca)


def test_three_nn_matches_oracle_away_from_ties():
    g = torch.Generator().manual_seed(2)
    known = torch.rand(400, 3, generator=g) * 10
    unknown = torch.rand(900, 3, generator=g) * 10
    d_ref, i_ref = on.three_nn(unknown, known)
    d, i = tb.three_nn(unknown, known)
    torch.testing.assert_close(d, d_ref)
    assert torch.equal(i.int(), i_ref)


def test_segnet_forward_same_through_both_backends():
    """Whole SDSeg3D forward of the oracle (VFE -> UNet -> 3-NN devoxelize -> head) with the torch backend == numpy backend."""
    import os
    from lidarseg3d_b200 import synth
    from lidarseg3d_b200.det3d import Config, build_detector
    from oracle import voxelize as ov
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, "configs", "sdseg3d_semantickitti.py"))
    torch.manual_seed(0)
    m = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    spec = dict(synth.KITTI)
    spec.update(beams=8, azimuths=200)
    frame = synth.lidar_scan(spec, 0)
    v, c, n = ov.points_to_voxel(frame, spec["voxel_size"], spec["pc_range"], 5, 300000)
    vv, cc, nn_, nv, pts = ov.collate_frames([(v, c, n, frame)])
    ex = dict(voxels=torch.from_numpy(vv), coordinates=torch.from_numpy(cc), num_points=torch.from_numpy(nn_),
              num_voxels=torch.from_numpy(nv), shape=np.stack([synth.grid_shape(spec)]), points=torch.from_numpy(pts))
    ocfg = dict(voxel_size=spec["voxel_size"], pc_range=spec["pc_range"],
                reader=dict(type="TransformerVoxelFeatureExtractor", num_head=4, num_layers=3))
    with torch.no_grad():
        a = on.segnet_forward(sd, ex, ocfg)
        b = on.segnet_forward(sd, ex, ocfg, backend=tb)
    torch.testing.assert_close(b, a, rtol=1e-4, atol=1e-4)
