"""oracle/nets.py wiring (UNetSCN3D, SegNet, SegMSeg3DNet) and the TTA merge pinned to the REFERENCE's own forward code.

tests/golden/ref_backbones.pt is produced by oracle/make_golden.py::gen_backbones, which executes the reference files
det3d/models/backbones/scn_unet.py and det3d/models/detectors/{seg_net,seg_mseg3d_net}.py unmodified on top of the spconv
API shim (oracle/spconv_shim.py; spconv itself is not installable).  That pins UR_block_forward, channel_reduction,
indice_key reuse, conv_out and the state-dict names; the sparse arithmetic underneath is oracle/sparse.py in both."""
import os

import numpy as np
import pytest
import torch

from oracle import nets as on
from oracle.make_golden import SMALL_RANGE, SMALL_VOXEL, unet_fill


@pytest.fixture(scope="module")
def fx(golden_dir):
    return torch.load(os.path.join(golden_dir, "ref_backbones.pt"), weights_only=False)


def _sd_for(keys, module):
    sd = module.state_dict()
    assert sorted(sd.keys()) == keys, "state-dict key set differs from the reference module's"
    return unet_fill(sd)


@pytest.mark.parametrize("ratio", [2, 1])
def test_unet_wiring_vs_reference_forward(fx, ratio):
    from lidarseg3d_b200.det3d.backbones import UNetSCN3D
    g = fx[f"unet_r{ratio}"]
    net = UNetSCN3D(num_input_features=13, voxel_size=SMALL_VOXEL, point_cloud_range=SMALL_RANGE,
                    model_cfg=dict(SCALING_RATIO=ratio), ds_factor=8, us_factor=8)
    sd = _sd_for(g["keys"], net)
    feats, coords, ex = on.unet_scn3d(sd, "", g["voxel_features"], g["coordinates"], list(g["input_shape"]), SMALL_VOXEL,
                                      SMALL_RANGE, return_levels=True)
    scale = float(g["conv_point_features"].abs().max())
    assert scale > 1e-2
    assert float((feats - g["conv_point_features"]).abs().max()) <= 2e-5 * scale
    torch.testing.assert_close(coords, g["conv_point_coords"], rtol=0, atol=0)
    # encoder output (conv_out, k(3,1,1) s(2,1,1)): sites in ascending linear order + features
    assert np.array_equal(np.asarray(ex["encoded"].indices), g["encoded_indices"].numpy())
    assert list(ex["encoded"].shape) == g["encoded_shape"]
    e = g["encoded_features"]
    assert float((ex["encoded"].features - e).abs().max()) <= 2e-5 * float(e.abs().max())
    # decoder levels as the reference exposes them (multi_scale_3d_features)
    mine = {"x_conv1": ex["up2"], "x_conv2": ex["up3"], "x_conv3": ex["up4"], "x_conv4": ex["levels"][4]}
    for k, ref in g["multi_scale"].items():
        assert np.array_equal(np.asarray(mine[k].indices), ref["indices"].numpy()), k
        assert list(mine[k].shape) == ref["shape"], k
        assert float((mine[k].features - ref["features"]).abs().max()) <= 2e-5 * float(ref["features"].abs().max()), k


def test_segnet_vs_reference_forward(fx):
    from lidarseg3d_b200.det3d import Config, build_detector
    g = fx["segnet"]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, "configs", "sdseg3d_semantickitti.py"))
    m = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    sd = _sd_for(g["keys"], m)
    out = on.segnet_forward(sd, g["example"], dict(voxel_size=SMALL_VOXEL, pc_range=SMALL_RANGE,
                                                   reader=dict(type="TransformerVoxelFeatureExtractor", num_head=4, num_layers=3)))
    ref = g["out_logits"]
    assert float((out - ref).abs().max()) <= 5e-5 * float(ref.abs().max())
    lab = on.predict_labels(out, g["example"]["points"][:, :4], 2)
    agree = torch.cat([a == b for a, b in zip(lab, g["labels"])]).float().mean()
    assert agree >= 0.999


def test_mseg3d_vs_reference_forward(fx):
    from lidarseg3d_b200.det3d import Config, build_detector
    g = fx["mseg3d"]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, "configs", "mseg3d_nuscenes.py"))
    m = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    sd = _sd_for(g["keys"], m)
    r = on.mseg3d_forward(sd, g["example"], dict(voxel_size=SMALL_VOXEL, pc_range=SMALL_RANGE,
                                                 hrnet_extra=cfg.model.img_backbone.extra, nhead=4, nlayer=6, num_convs=2),
                          return_all=True)
    ref = g["out_logits"]
    assert float((r["out_logits"] - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    vl = g["voxel_logits"]
    assert float((r["voxel_logits"] - vl).abs().max()) <= 1e-4 * float(vl.abs().max())


def test_tta_merge_vs_reference_predict(fx):
    """_predict's TTA branch against the reference's own predict() (point_seg_mseg3d_head.py:398-453)."""
    from lidarseg3d_b200.det3d.point_heads import _predict
    g = fx["tta"]
    ntta = g["ntta"]
    n_frames = int(g["points"][:, 0].max()) + 1
    ex = dict(num_voxels=torch.zeros(n_frames), points=g["points"], metadata=[dict(token=i) for i in range(n_frames)],
              point_sem_labels=g["point_sem_labels"])
    ret = _predict(g["out_logits"], ex, dict(tta_flag=True, merge_type="ArithmeticMean", num_tta_tranforms=ntta))
    assert len(ret) == len(g["ret"]) == n_frames // ntta
    for a, b in zip(ret, g["ret"]):
        assert a["metadata"] == b["metadata"]
        assert torch.equal(a["pred_point_sem_labels"], b["pred"])
        assert torch.equal(a["point_sem_labels"], b["gt"])
    with pytest.raises(AssertionError):
        _predict(g["out_logits"], ex, dict(tta_flag=True, num_tta_tranforms=3))


def test_spmiddle_resnet_vs_reference_forward(fx):
    """SpMiddleResNetFHD (det3d/models/backbones/scn.py:84-177): oracle restatement vs the reference's own forward on the
    spconv shim; the product module carries the reference's state-dict key set."""
    from lidarseg3d_b200.det3d.backbones import SpMiddleResNetFHD
    g = fx["spmiddle"]
    net = SpMiddleResNetFHD(num_input_features=13)
    sd = _sd_for(g["keys"], net)
    dense, levels = on.sp_middle_resnet_fhd(sd, "", g["voxel_features"], g["coordinates"], 2, list(g["input_shape"]))
    ref = g["dense"]
    assert list(dense.shape) == list(ref.shape) and float(ref.abs().max()) > 1e-2
    assert float((dense - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    for k, r in g["multi_scale"].items():
        assert np.array_equal(np.asarray(levels[k].indices), r["indices"].numpy()), k
        assert list(levels[k].shape) == r["shape"], k
        assert float((levels[k].features - r["features"]).abs().max()) <= 2e-5 * float(r["features"].abs().max()), k
