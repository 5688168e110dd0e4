"""End-to-end parity: registry-built detectors on the CUDA kernels vs the CPU oracle on identical inputs and weights
(north-star gate: per-point logits within 1e-3 relative, >= 99.9 % argmax agreement)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nets as on
from oracle import voxelize as ov

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda"


def _randomize_bn(m, seed=0):
    g = torch.Generator().manual_seed(seed)
    for mod in m.modules():
        if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)


def _build(cfg_name, seed=0):
    from lidarseg3d_b200.det3d import Config, build_detector
    cfg = Config.fromfile(os.path.join(ROOT, "configs", cfg_name))
    torch.manual_seed(seed)
    m = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval()
    _randomize_bn(m, seed)
    return cfg, m


def _cpu_example(frames, spec, with_cam=False, img_hw=None, u8=False):
    from lidarseg3d_b200 import synth
    vox = [ov.points_to_voxel(f, spec["voxel_size"], spec["pc_range"], 5, 300000) for f in frames]
    v, c, n, nv, pts = ov.collate_frames([(a, b, cc, f) for (a, b, cc), f in zip(vox, frames)])
    B = len(frames)
    ex = dict(voxels=torch.from_numpy(v), coordinates=torch.from_numpy(c), num_points=torch.from_numpy(n),
              num_voxels=torch.from_numpy(nv), shape=np.stack([synth.grid_shape(spec)] * B), points=torch.from_numpy(pts))
    if with_cam:
        s2 = dict(spec)
        if img_hw:
            s2["net_hw"] = img_hw
        from oracle import camera as oc
        cbs = [synth.calibration(s2, b) for b in range(B)]
        ex["points_cuv"] = torch.from_numpy(np.concatenate([
            oc.project_points(f[:, :3], cb["ref_to_global"], cb["cams_from_global"], cb["intrinsics"], cb["img_hw"], s2["net_hw"])
            for f, cb in zip(frames, cbs)]))
        if u8:      # raw uint8 images; the oracle normalises them the reference's way (img_transforms.py:18-29) on the CPU
            ex["images_u8"] = torch.from_numpy(np.stack([synth.camera_images_u8(s2, b, s2["net_hw"]) for b in range(B)]))
            ex["images"] = torch.from_numpy(on.image_input_transform(ex["images_u8"].numpy(), synth.IMG_MEAN, synth.IMG_STD))
        else:
            ex["images"] = torch.from_numpy(np.stack([synth.camera_images(s2, b, s2["net_hw"]) for b in range(B)]))
    return ex


def _check_logits(out, ref, rel_tol, agree_tol):
    rel = float((out - ref).abs().max() / ref.abs().max())
    agree = float((out.argmax(1) == ref.argmax(1)).float().mean())
    assert rel <= rel_tol and agree >= agree_tol, (rel, agree)
    return rel, agree


def test_sdseg3d_forward_vs_oracle():
    from lidarseg3d_b200 import pipeline, synth
    cfg, m = _build("sdseg3d_semantickitti.py")
    spec = dict(synth.KITTI)
    spec.update(beams=16, azimuths=500)
    frames = [synth.lidar_scan(spec, s) for s in (0, 1)]
    ex_cpu = _cpu_example(frames, spec)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ref = on.segnet_forward(sd, ex_cpu, dict(voxel_size=spec["voxel_size"], pc_range=spec["pc_range"],
                                            reader=dict(type="TransformerVoxelFeatureExtractor", num_head=4, num_layers=3)))
    m = m.to(DEV)
    ex = pipeline.build_example(frames, spec["voxel_size"], spec["pc_range"])
    # the GPU-built example is bit-identical to the CPU (reference-semantics) one
    for k in ("voxels", "coordinates", "num_points", "points"):
        assert torch.equal(ex[k].cpu(), ex_cpu[k]), k
    assert torch.equal(ex["num_voxels"], ex_cpu["num_voxels"])
    preds = m(ex, return_loss=False)
    out = m.last_batch_dict["out_logits"].cpu()
    _check_logits(out, ref, 1e-3, 0.999)
    assert len(preds) == 2 and preds[0]["pred_point_sem_labels"].shape[0] == frames[0].shape[0]
    labels = on.predict_labels(ref, ex_cpu["points"], 2)
    assert float((preds[1]["pred_point_sem_labels"].cpu() == labels[1]).float().mean()) >= 0.999


@pytest.mark.parametrize("cfg_name,spec_name,image_dtype,u8", [("mseg3d_nuscenes.py", "NUSC", None, False),
                                                               ("mseg3d_waymo.py", "WAYMO", None, False),
                                                               ("mseg3d_nuscenes.py", "NUSC", torch.float16, False),
                                                               ("mseg3d_nuscenes.py", "NUSC", torch.float16, True),
                                                               ("mseg3d_nuscenes.py", "NUSC", "dual", False),
                                                               ("mseg3d_nuscenes.py", "NUSC", "dual", True),
                                                               ("mseg3d_waymo.py", "WAYMO", "dual", True),
                                                               ("mseg3d_waymo.py", "WAYMO", None, True)])
def test_mseg3d_forward_vs_oracle(cfg_name, spec_name, image_dtype, u8):
    """BASELINE.json configs[2] (nuScenes: 17 classes, 6 cameras) and configs[3] (Waymo: 23 classes, 5 cameras, z range
    [-2, 4]) at a reduced scan / image size the CPU oracle finishes in seconds."""
    from lidarseg3d_b200 import pipeline, synth
    cfg, m = _build(cfg_name)
    spec = dict(getattr(synth, spec_name))
    spec.update(beams=16, azimuths=400)
    hw = (128, 192)
    frames = [synth.lidar_scan(spec, s) for s in (0, 1)]
    ex_cpu = _cpu_example(frames, spec, with_cam=True, img_hw=hw, u8=u8)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ocfg = dict(voxel_size=spec["voxel_size"], pc_range=spec["pc_range"], hrnet_extra=cfg.model.img_backbone.extra,
                nhead=4, nlayer=6, num_convs=2)
    ref = on.mseg3d_forward(sd, ex_cpu, ocfg, return_all=True)
    m = m.to(DEV)
    m.image_dtype = image_dtype          # fp16: camera branch on fp16 maps (own tcgen05 3x3 kernel); same logits gate
    if u8:      # bench.py's path: uint8 images uploaded and normalised on the device in the camera branch's storage type
        ex = pipeline.build_example(frames, spec["voxel_size"], spec["pc_range"], images_u8=ex_cpu["images_u8"],
                                    img_mean=synth.IMG_MEAN, img_std=synth.IMG_STD,
                                    image_dtype=image_dtype if isinstance(image_dtype, torch.dtype) else torch.float32,
                                    points_cuv=ex_cpu["points_cuv"])
    else:
        ex = pipeline.build_example(frames, spec["voxel_size"], spec["pc_range"], images=ex_cpu["images"],
                                    points_cuv=ex_cpu["points_cuv"])
    preds = m(ex, return_loss=False)
    bd = m.last_batch_dict
    # stage-by-stage (helps localise a failure), then the north-star gate on the logits
    def rel(a, b):
        return float((a.cpu() - b).abs().max() / b.abs().max())
    assert rel(bd["image_features"].reshape(ref["image_features"].shape), ref["image_features"]) <= 2e-3
    assert rel(bd["conv_point_features"], ref["conv_point_features"]) <= 2e-3
    dbg = bd["_ls3d_debug"]
    assert rel(dbg["point_features_lidar_0"], ref["point_features_lidar_0"]) <= 2e-3
    assert rel(dbg["geo_fused"], ref["geo_fused"]) <= 2e-3
    _check_logits(bd["out_logits"].cpu(), ref["out_logits"], 1e-3, 0.999)
    assert len(preds) == 2


@pytest.mark.parametrize("image_dtype", [torch.float32, torch.float16, "dual"])
def test_mseg3d_full_size_parity_one_frame(image_dtype):
    """The BENCHMARKED configuration (32-beam ~30 k-point scan, 6 raw 900x1600 uint8 images resized to 640x960 on the
    device, GPU projection) through bench.py's own parity block, one frame: the gate the bench line carries
    (logits 1e-3 relative, >= 99.9 % argmax, bit-exact voxels)."""
    import sys
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        import bench
    finally:
        sys.argv = argv
    from lidarseg3d_b200 import synth
    wl = bench.WORKLOADS["mseg3d_nuscenes"]
    spec = synth.NUSC
    cfg, model = bench.build_model(wl)
    model = model.to(DEV)
    batch = bench.make_batches(wl, spec, 1, 1, 0, n_image_sets=1)[0]
    name = {torch.float32: "fp32", torch.float16: "fp16cam", "dual": "dual"}[image_dtype]
    parity, _, _ = bench.parity_block(wl, spec, cfg, model, batch, 1, lambda b, dt: bench.gpu_forward(wl, spec, model, b, dt, DEV),
                                      [(name, image_dtype)])
    m = parity["modes"][name]
    assert m["coords_bit_exact"]
    assert m["points_cuv_cam_valid_mismatches"] <= 2 and m["points_cuv_max_abs_diff"] <= 2e-6
    if image_dtype != torch.float16:
        assert m["resized_images_bit_exact"]
    assert m["rel_err"] <= 1e-3 and m["argmax_agreement"] >= 0.999, m
    assert parity["ok"]
