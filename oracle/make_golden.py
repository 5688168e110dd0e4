"""Generate tests/golden/* by running the REFERENCE's own code from /root/reference (build container only).

What is real reference code here:
  * det3d/ops/point_cloud/point_cloud_ops.py::points_to_voxel (numba) - imported by file path, unmodified;
  * the plain-PyTorch reference modules ImprovedMeanVoxelFeatureExtractor, TransformerVoxelFeatureExtractor,
    LiDARSemanticFeatureAggregationModule, PointSegMSeg3DHead (incl. SemanticFeatureFusionModule),
    PointSegBatchlossHead, HRNet, FCNMSeg3DHead (incl. CameraSemanticFeatureAggregationModule) - imported from
    the reference tree with their un-installable third-party imports (spconv, mmcv, torch_scatter, addict, the
    CUDA-only pointnet2 extension) replaced by the stubs below.  The mmcv stubs restate mmcv's ConvModule /
    build_conv_layer / build_norm_layer (conv -> norm -> act, conv bias off when a norm follows); the pointnet2
    stub is the CPU restatement of interpolate_gpu.cu (the kernel itself cannot run without a GPU).
Weights are not stored: every parameter/buffer is filled by ``seeded_fill`` (a pure function of its name), which the
tests re-apply.  Run:  python oracle/make_golden.py
"""
import importlib.util
import os
import sys
import types
import zlib

import numpy as np
import torch
from torch import nn

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def seeded_fill(module_or_sd, scale=1.0):
    """Deterministically fill every float tensor of a module / state dict from crc32(name)."""
    sd = module_or_sd.state_dict() if isinstance(module_or_sd, nn.Module) else module_or_sd
    out = {}
    for name, t in sd.items():
        if not t.is_floating_point():
            out[name] = t.clone()
            continue
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
        if name.endswith("running_var"):
            v = torch.rand(t.shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            v = torch.randn(t.shape, generator=g) * 0.1
        elif t.dim() <= 1:
            v = torch.randn(t.shape, generator=g) * 0.1 + (1.0 if name.endswith("weight") else 0.0)
        else:
            fan_in = max(1, int(np.prod(t.shape[1:]))) if not name.endswith(("conv1.weight", "conv2.weight")) or t.dim() != 5 \
                else int(np.prod(t.shape[:-1]))
            v = torch.randn(t.shape, generator=g) * scale / fan_in ** 0.5
        out[name] = v.to(t.dtype)
    if isinstance(module_or_sd, nn.Module):
        module_or_sd.load_state_dict(out)
    return out


# ----------------------------------------------------------------------------------------------- stubs
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    from oracle import nets as on

    _mod("spconv")
    _mod("torch_scatter")
    from lidarseg3d_b200.det3d.config import install_addict_shim
    install_addict_shim()

    # ---- mmcv (recalled semantics)
    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    def build_conv_layer(cfg, *args, **kwargs):
        assert cfg is None or cfg.get("type", "Conv2d") in ("Conv2d", "Conv")
        return nn.Conv2d(*args, **kwargs)

    def build_norm_layer(cfg, num_features, postfix=""):
        cfg = dict(cfg)
        t = cfg.pop("type")
        assert t in ("BN", "BN2d")
        rg = cfg.pop("requires_grad", True)
        cfg.setdefault("eps", 1e-5)
        layer = nn.BatchNorm2d(num_features, **cfg)
        for p in layer.parameters():
            p.requires_grad = rg
        return "bn" + str(postfix), layer

    class ConvModule(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                     bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True, **kw):
            super().__init__()
            with_norm = norm_cfg is not None
            if bias == "auto":
                bias = not with_norm
            self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
            self.with_norm, self.with_act = with_norm, act_cfg is not None
            if with_norm:
                self.bn = build_norm_layer(norm_cfg, out_channels)[1]
            if self.with_act:
                assert act_cfg["type"] == "ReLU"
                self.activate = nn.ReLU(inplace=inplace)

        def forward(self, x):
            x = self.conv(x)
            if self.with_norm:
                x = self.bn(x)
            if self.with_act:
                x = self.activate(x)
            return x

    _mod("mmcv")
    _mod("mmcv.cnn", build_conv_layer=build_conv_layer, build_norm_layer=build_norm_layer, ConvModule=ConvModule,
         build_plugin_layer=None, constant_init=None, kaiming_init=None, build_activation_layer=None)
    _mod("mmcv.runner", BaseModule=BaseModule, ModuleList=nn.ModuleList, Sequential=nn.Sequential,
         load_checkpoint=None, force_fp32=None)
    _mod("mmcv.utils")
    _mod("mmcv.utils.parrots_wrapper", _BatchNorm=nn.modules.batchnorm._BatchNorm)

    # ---- det3d package skeleton (only the files on the hot path are loaded for real)
    for pkg in ["det3d", "det3d.models", "det3d.models.readers", "det3d.models.point_heads", "det3d.models.img_heads",
                "det3d.models.img_backbones", "det3d.models.utils", "det3d.core", "det3d.core.utils", "det3d.ops",
                "det3d.ops.pointnet2_batch", "det3d.ops.mmseg_ops", "det3d.utils", "det3d.utils.dist"]:
        m = _mod(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split("."))]
    sys.modules["det3d"].torchie = _mod("det3d.torchie", is_str=lambda s: isinstance(s, str))
    _mod("det3d.core.utils.loss_utils", lovasz_softmax=None)
    _mod("det3d.core.utils.common_utils")
    sys.modules["det3d.core.utils"].common_utils = sys.modules["det3d.core.utils.common_utils"]
    _mod("det3d.utils.dist.dist_common")
    sys.modules["det3d.utils.dist"].dist_common = sys.modules["det3d.utils.dist.dist_common"]

    # pointnet2 extension: CPU restatement of interpolate_gpu.cu (the reference kernels are CUDA-only)
    def three_nn(unknown, known):
        d2, idx = on.three_nn(unknown[0], known[0])
        return torch.sqrt(d2).unsqueeze(0), idx.unsqueeze(0)

    def three_interpolate(features, idx, weight):      # features [1,C,M] -> [1,C,N]
        f = features[0].t()[idx[0].long()]              # [N,3,C]
        out = weight[0][:, 0:1] * f[:, 0] + weight[0][:, 1:2] * f[:, 1] + weight[0][:, 2:3] * f[:, 2]
        return out.t().unsqueeze(0)

    _mod("det3d.ops.pointnet2_batch.pointnet2_utils", three_nn=three_nn, three_interpolate=three_interpolate)


def load_ref(dotted, relpath):
    spec = importlib.util.spec_from_file_location(dotted, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = m
    spec.loader.exec_module(m)
    parent, _, leaf = dotted.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


def load_reference_modules():
    install_stubs()
    load_ref("det3d.utils.registry", "det3d/utils/registry.py")
    sys.modules["det3d.utils"].Registry = sys.modules["det3d.utils.registry"].Registry
    sys.modules["det3d.utils"].build_from_cfg = sys.modules["det3d.utils.registry"].build_from_cfg
    reg = load_ref("det3d.models.registry", "det3d/models/registry.py")
    _mod("det3d.models.builder", **{k: getattr(reg, k) for k in dir(reg) if k.isupper()})
    sys.modules["det3d.models"].builder = sys.modules["det3d.models.builder"]
    norm = load_ref("det3d.models.utils.norm", "det3d/models/utils/norm.py")
    sys.modules["det3d.models.utils"].build_norm_layer = norm.build_norm_layer
    wr = load_ref("det3d.ops.mmseg_ops.wrappers", "det3d/ops/mmseg_ops/wrappers.py")
    ms = sys.modules["det3d.ops.mmseg_ops"]
    ms.Upsample, ms.resize, ms.ResLayer = wr.Upsample, wr.resize, None
    R = {}
    R["vfe"] = load_ref("det3d.models.readers.voxel_encoder", "det3d/models/readers/voxel_encoder.py")
    R["ctx"] = load_ref("det3d.models.point_heads.context_module", "det3d/models/point_heads/context_module.py")
    load_ref("det3d.models.point_heads.point_utils", "det3d/models/point_heads/point_utils.py")
    R["mhead"] = load_ref("det3d.models.point_heads.point_seg_mseg3d_head", "det3d/models/point_heads/point_seg_mseg3d_head.py")
    R["bhead"] = load_ref("det3d.models.point_heads.point_seg_batchloss_head",
                          "det3d/models/point_heads/point_seg_batchloss_head.py")
    load_ref("det3d.models.img_backbones.resnet_mmcv", "det3d/models/img_backbones/resnet_mmcv.py")
    R["hrnet"] = load_ref("det3d.models.img_backbones.hrnet", "det3d/models/img_backbones/hrnet.py")
    load_ref("det3d.models.img_heads.sc_conv", "det3d/models/img_heads/sc_conv.py")
    load_ref("det3d.models.img_heads.decode_head", "det3d/models/img_heads/decode_head.py")
    R["fcn"] = load_ref("det3d.models.img_heads.fcn_mseg3d_head", "det3d/models/img_heads/fcn_mseg3d_head.py")
    return R


# ----------------------------------------------------------------------------------------------- fixtures
TINY_HRNET = dict(stage1=dict(num_modules=1, num_branches=1, block="BOTTLENECK", num_blocks=(2,), num_channels=(8,)),
                  stage2=dict(num_modules=1, num_branches=2, block="BASIC", num_blocks=(2, 2), num_channels=(4, 8)),
                  stage3=dict(num_modules=2, num_branches=3, block="BASIC", num_blocks=(2, 2, 2), num_channels=(4, 8, 16)),
                  stage4=dict(num_modules=2, num_branches=4, block="BASIC", num_blocks=(2, 2, 2, 2),
                              num_channels=(4, 8, 16, 32)))
HEAD_CFG = dict(VOXEL_IN_DIM=32, VOXEL_CLS_FC=[64], VOXEL_ALIGN_DIM=64, IMAGE_IN_DIM=48, IMAGE_ALIGN_DIM=64,
                GEO_FUSED_DIM=64, OUT_CLS_FC=[64, 64], IGNORED_LABEL=0, DP_RATIO=0.25, MIMIC_FC=[64, 64],
                SFPhase_CFG=dict(embeddings_proj_kernel_size=1, d_model=96, n_head=4, n_layer=2, n_ffn=192,
                                 drop_ratio=0, activation="relu", pre_norm=False))
BHEAD_CFG = dict(CONV_IN_DIM=32, CONV_CLS_FC=[64], CONV_ALIGN_DIM=64, OUT_CLS_FC=[64, 64], IGNORED_LABEL=0)


def voxel_scene(seed, n, feat, rng_xyz):
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.uniform(-rng_xyz[0], rng_xyz[0], (n, 2)), rng.uniform(rng_xyz[1], rng_xyz[2], (n, 1)),
                          rng.uniform(0, 1, (n, feat - 3))], 1).astype(np.float32)
    k = n // 3                                  # dense clusters -> voxels with > max_points points
    pts[:k, :3] = pts[k:2 * k, :3] + rng.normal(0, 0.03, (k, 3)).astype(np.float32)
    pts[2 * k:2 * k + 40, :3] = pts[0, :3]      # 40 identical points
    return pts


def gen_voxelize():
    spec = importlib.util.spec_from_file_location("ref_pco", os.path.join(REF, "det3d/ops/point_cloud/point_cloud_ops.py"))
    pco = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pco)
    cases = dict(nusc=dict(n=3000, feat=5, vs=[0.1, 0.1, 0.2], rg=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], ext=(54, -6, 4),
                           max_voxels=300000),
                 kitti=dict(n=3000, feat=4, vs=[0.1, 0.1, 0.15], rg=[-75.2, -75.2, -4, 75.2, 75.2, 2], ext=(78, -5, 3),
                            max_voxels=300000),
                 capped=dict(n=3000, feat=5, vs=[0.1, 0.1, 0.2], rg=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], ext=(54, -6, 4),
                             max_voxels=700))
    for name, c in cases.items():
        pts = voxel_scene(zlib.crc32(name.encode()), c["n"], c["feat"], c["ext"])
        v, co, nu = pco.points_to_voxel(pts, np.array(c["vs"], np.float32), np.array(c["rg"], np.float32), 5, True,
                                        c["max_voxels"])
        np.savez_compressed(os.path.join(OUT, f"voxelize_{name}.npz"), points=pts, voxel_size=np.array(c["vs"], np.float32),
                            pc_range=np.array(c["rg"], np.float32), max_points=5, max_voxels=c["max_voxels"],
                            voxels=v, coordinates=co, num_points=nu)
        print("voxelize", name, v.shape, int(nu.max()))


def gen_modules():
    R = load_reference_modules()
    g = torch.Generator().manual_seed(1234)
    fx = {}
    # ---- VFEs on voxels produced by the reference voxelizer fixture
    z = np.load(os.path.join(OUT, "voxelize_nusc.npz"))
    vox5 = torch.from_numpy(z["voxels"][:400]); num5 = torch.from_numpy(z["num_points"][:400])
    z4 = np.load(os.path.join(OUT, "voxelize_kitti.npz"))
    vox4 = torch.from_numpy(z4["voxels"][:400]); num4 = torch.from_numpy(z4["num_points"][:400])
    with torch.no_grad():
        m = R["vfe"].ImprovedMeanVoxelFeatureExtractor(num_input_features=5).eval()
        fx["improved_mean_vfe"] = dict(voxels=vox5, num=num5, out=m(vox5, num5))
        m = R["vfe"].MeanVoxelFeatureExtractor(num_input_features=5).eval()
        fx["mean_vfe"] = dict(voxels=vox5, num=num5, out=m(vox5, num5))
        m = R["vfe"].TransformerVoxelFeatureExtractor(num_input_features=4, num_compressed_features=16, num_embed=64,
                                                      num_head=4, num_layers=3).eval()
        seeded_fill(m)
        # torch>=2.0 passes is_causal to custom layers: drive the reference layer objects by hand
        class _Seq(nn.Module):
            def __init__(self, layers):
                super().__init__()
                self.layers = layers

            def forward(self, x):
                for l in self.layers:
                    x = l(x)
                return x
        m.chunck = _Seq(m.chunck.layers)
        fx["trans_vfe"] = dict(voxels=vox4, num=num4, out=m(vox4, num4))
        # ---- point heads
        B, M, N, ncls = 2, 300, 500, 17
        vcoord_idx = torch.stack([torch.arange(M) % 2 * 0 + (torch.arange(M) >= M // 2).long(),
                                  torch.randint(0, 40, (M,), generator=g), torch.randint(0, 200, (M,), generator=g),
                                  torch.randint(0, 200, (M,), generator=g)], 1)
        vs = torch.tensor([0.1, 0.1, 0.2]); lo = torch.tensor([-10.0, -10.0, -5.0])
        centers = (vcoord_idx[:, [3, 2, 1]].float() + 0.5) * vs + lo
        vcoords = torch.cat([vcoord_idx[:, :1].float(), centers], 1)
        pb = (torch.arange(N) >= N // 2).float()
        pts = torch.cat([pb[:, None], torch.rand(N, 3, generator=g) * torch.tensor([20.0, 20.0, 8.0]) + lo], 1)
        vfeat = torch.randn(M, 32, generator=g)
        cuv = torch.rand(N, 4, generator=g) * 2 - 1
        cuv[:, 0] = (torch.rand(N, generator=g) > 0.3).float()
        cam = torch.randint(0, 6, (N,), generator=g).float()
        cuv[:, 1] = cam / 5 * 2 - 1
        img_feat = torch.randn(B, 6, 48, 10, 15, generator=g)
        cam_emb = torch.randn(B, 48, ncls, 1, generator=g)
        head = R["mhead"].PointSegMSeg3DHead(class_agnostic=False, num_class=ncls, model_cfg=HEAD_CFG).eval()
        seeded_fill(head)
        bd = dict(batch_size=B, conv_point_features=vfeat, conv_point_coords=vcoords, points=pts, points_cuv=cuv,
                  image_features=img_feat, camera_semantic_embeddings=cam_emb)
        out = head(dict(bd), return_loss=False)
        lemb = head.lidar_sfam(feats=vfeat, probs=head.forward_ret_dict["voxel_logits"], batch_idx=vcoords[:, 0], batch_size=B)
        fx["mseg3d_head"] = dict(inputs=bd, out_logits=out["out_logits"], voxel_logits=head.forward_ret_dict["voxel_logits"],
                                 lidar_emb=lemb)
        bhead = R["bhead"].PointSegBatchlossHead(class_agnostic=False, num_class=20, model_cfg=BHEAD_CFG).eval()
        seeded_fill(bhead)
        out = bhead(dict(batch_size=B, conv_point_features=vfeat, conv_point_coords=vcoords, points=pts), return_loss=False)
        fx["batchloss_head"] = dict(out_logits=out["out_logits"], conv_logits=bhead.forward_ret_dict["conv_logits"])
        # ---- image branch (tiny HRNet, same class / code path as w18)
        hr = R["hrnet"].HRNet(extra=TINY_HRNET, norm_cfg=dict(type="BN", requires_grad=True))
        hr.eval()            # the reference's HRNet.train() returns None (hrnet.py:695-704)
        seeded_fill(hr)
        imgs = torch.randn(B * 2, 3, 64, 96, generator=g)
        ys = hr(imgs)
        fcn = R["fcn"].FCNMSeg3DHead(in_channels=[4, 8, 16, 32], in_index=(0, 1, 2, 3), channels=8, num_classes=5,
                                     input_transform="resize_concat", kernel_size=1, num_convs=2, concat_input=False,
                                     dropout_ratio=-1, norm_cfg=dict(type="BN", requires_grad=True), align_corners=False,
                                     ignore_index=0, loss_weight=0.5).eval()
        seeded_fill(fcn)
        r = fcn(dict(inputs=ys, batch_size=B), return_loss=False)
        fx["image_branch"] = dict(images=imgs, hrnet_out=[y.clone() for y in ys], image_features=r["image_features"],
                                  image_logits=r["image_logits"], camera_semantic_embeddings=r["camera_semantic_embeddings"],
                                  hrnet_keys=sorted(hr.state_dict().keys()), fcn_keys=sorted(fcn.state_dict().keys()))
        fx["state_dict_keys"] = dict(mseg3d_head=sorted(head.state_dict().keys()), batchloss_head=sorted(bhead.state_dict().keys()),
                                     trans_vfe=sorted(k.replace("chunck.layers.layers.", "chunck.layers.") for k in m.state_dict().keys()))
    torch.save(fx, os.path.join(OUT, "ref_modules.pt"))
    print("modules:", {k: (tuple(v["out"].shape) if isinstance(v, dict) and "out" in v else "...") for k, v in fx.items()})


def gen_losses():
    """Training losses from the reference's own det3d/core/utils/loss_utils.py + the loss assembly of
    the reference MSeg3D point head (forward_ret_dict filled by hand) -> tests/golden/ref_losses.pt."""
    for k in [k for k in sys.modules if k.startswith("det3d")]:
        del sys.modules[k]
    install_stubs()
    _mod("det3d.core.utils.box_utils")
    lu = load_ref("det3d.core.utils.loss_utils", "det3d/core/utils/loss_utils.py")
    g = torch.Generator().manual_seed(4321)
    fx = {"cases": []}
    for n, c, p_ignore in ((500, 17, 0.2), (64, 5, 0.5), (1000, 23, 0.0), (7, 4, 0.6)):
        logits = torch.randn(n, c, generator=g) * 2
        labels = torch.randint(0, c, (n,), generator=g)
        labels[torch.rand(n, generator=g) < p_ignore] = 0
        probas = torch.softmax(logits, -1)
        fx["cases"].append(dict(logits=logits, labels=labels,
                                lovasz_ignore0=lu.lovasz_softmax(probas, labels, ignore=0),
                                lovasz_noignore=lu.lovasz_softmax(probas, labels),
                                ce_ignore0=nn.CrossEntropyLoss(ignore_index=0)(logits, labels)))
    # dense (image) form [B, C, H, W]
    logits4 = torch.randn(2, 6, 9, 11, generator=g)
    labels4 = torch.randint(0, 6, (2, 9, 11), generator=g)
    fx["dense"] = dict(logits=logits4, labels=labels4, lovasz=lu.lovasz_softmax(torch.softmax(logits4, 1), labels4),
                       lovasz_per_image=lu.lovasz_softmax(torch.softmax(logits4, 1), labels4, per_image=True))
    # point-head assembly through the reference head object
    R = load_reference_modules()
    sys.modules["det3d.core.utils.loss_utils"].lovasz_softmax = lu.lovasz_softmax
    R["mhead"].lovasz_softmax = lu.lovasz_softmax
    head = R["mhead"].PointSegMSeg3DHead(class_agnostic=False, num_class=17, model_cfg=HEAD_CFG)
    vl, pl = torch.randn(300, 17, generator=g), torch.randn(800, 17, generator=g)
    vlab, plab = torch.randint(0, 17, (300,), generator=g), torch.randint(0, 17, (800,), generator=g)
    pc, cam = torch.randn(500, 64, generator=g), torch.randn(500, 64, generator=g)
    head.forward_ret_dict = dict(voxel_logits=vl, voxel_sem_labels=vlab, out_logits=pl, point_sem_labels=plab,
                                 point_features_pcamera=pc, point_features_camera=cam)
    loss, parts = head.get_loss()
    fx["point_head"] = dict(voxel_logits=vl, voxel_labels=vlab, out_logits=pl, point_labels=plab, pcamera=pc, camera=cam,
                            loss=loss.detach(), parts={k: v.clone() for k, v in parts.items()})
    # image-head loss (fcn_mseg3d_head.py:202-244) through the reference head object
    fcn = R["fcn"].FCNMSeg3DHead(in_channels=[4, 8, 16, 32], in_index=(0, 1, 2, 3), channels=8, num_classes=5,
                                 input_transform="resize_concat", kernel_size=1, num_convs=2, concat_input=False,
                                 dropout_ratio=-1, norm_cfg=dict(type="BN", requires_grad=True), align_corners=False,
                                 ignore_index=0, loss_weight=0.5, lovasz_loss_weight=-1.0)
    il = torch.randn(4, 5, 16, 24, generator=g)
    ilab = torch.randint(0, 5, (4, 1, 64, 96), generator=g).float()
    ilab[torch.rand(ilab.shape, generator=g) < 0.9] = 0         # sparse point-wise supervision
    fcn.forward_ret_dict = dict(image_logits=il, image_sem_labels=ilab)
    iloss, iparts = fcn.get_loss()
    fx["image_head"] = dict(image_logits=il, image_sem_labels=ilab, loss=iloss.detach(),
                            parts={k: v.clone() for k, v in iparts.items()})
    torch.save(fx, os.path.join(OUT, "ref_losses.pt"))
    print("losses:", [float(c["lovasz_ignore0"]) for c in fx["cases"]], float(loss), float(iloss))



# ----------------------------------------------------------------------------------------------- backbones / detectors
def unet_fill(sd):
    """seeded_fill, then every 5-D sparse-conv weight [kz,ky,kx,Cin,Cout] redrawn with fan_in = K*Cin (activations of O(1)
    through the 37 convs of the UNet; seeded_fill's generic fan-in is far too large for that layout)."""
    out = seeded_fill(sd)
    for name, t in out.items():
        if t.dim() == 5:
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
            fan_in = int(np.prod(t.shape[:4]))
            out[name] = (torch.randn(t.shape, generator=g) * (2.0 / fan_in) ** 0.5).to(t.dtype)
    return out


def surface_scene(seed, n, feat, half_xy=3.1, z_lo=-1.9, z_hi=1.9):
    """Points on a few noisy planes (dense local neighbourhoods, so SubM / strided rulebooks have many pairs)."""
    rng = np.random.default_rng(seed)
    parts = []
    for _ in range(4):
        m = n // 4
        xy = rng.uniform(-half_xy, half_xy, (m, 2))
        a, b, c = rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-1.0, 1.0)
        z = np.clip(a * xy[:, 0] + b * xy[:, 1] + c + rng.normal(0, 0.03, m), z_lo, z_hi)
        parts.append(np.concatenate([xy, z[:, None]], 1))
    xyz = np.concatenate(parts, 0)
    extra = rng.uniform(0, 1, (xyz.shape[0], feat - 3))
    return np.concatenate([xyz, extra], 1).astype(np.float32)


SMALL_RANGE = [-3.2, -3.2, -2.0, 3.2, 3.2, 2.0]
SMALL_VOXEL = [0.1, 0.1, 0.1]


def load_reference_detectors():
    """The reference's scn_unet.py / scn.py / detectors on the spconv shim (oracle/spconv_shim.py)."""
    from oracle import spconv_shim
    for k in [k for k in sys.modules if k.startswith("det3d")]:
        del sys.modules[k]
    R = load_reference_modules()
    spconv_shim.install()
    _mod("pycocotools")
    _mod("pycocotools.mask")
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    _mod("det3d.torchie.trainer", load_checkpoint=None)
    sys.modules["det3d.torchie"].trainer = sys.modules["det3d.torchie.trainer"]
    _mod("det3d.models.utils.finetune_utils", FrozenBatchNorm2d=None)
    cu = load_ref("det3d.core.utils.common_utils", "det3d/core/utils/common_utils.py")
    sys.modules["det3d.core.utils"].common_utils = cu
    for pkg in ["det3d.models.backbones", "det3d.models.detectors"]:
        m = _mod(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split("."))]
    load_ref("det3d.models.builder", "det3d/models/builder.py")
    R["scn_unet"] = load_ref("det3d.models.backbones.scn_unet", "det3d/models/backbones/scn_unet.py")
    R["scn"] = load_ref("det3d.models.backbones.scn", "det3d/models/backbones/scn.py")
    load_ref("det3d.models.detectors.base", "det3d/models/detectors/base.py")
    load_ref("det3d.models.detectors.single_stage", "det3d/models/detectors/single_stage.py")
    R["seg_net"] = load_ref("det3d.models.detectors.seg_net", "det3d/models/detectors/seg_net.py")
    R["seg_mseg3d_net"] = load_ref("det3d.models.detectors.seg_mseg3d_net", "det3d/models/detectors/seg_mseg3d_net.py")
    R["builder"] = sys.modules["det3d.models.builder"]
    return R


def _example(frames, feat, with_cam=None):
    """example dict through the reference's numba voxelizer + collate rules (oracle.voxelize.collate_frames)."""
    from oracle import voxelize as ov
    spec = importlib.util.spec_from_file_location("ref_pco", os.path.join(REF, "det3d/ops/point_cloud/point_cloud_ops.py"))
    pco = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pco)
    vox = [pco.points_to_voxel(f, np.array(SMALL_VOXEL, np.float32), np.array(SMALL_RANGE, np.float32), 5, True, 300000)
           for f in frames]
    v, c, n, nv, pts = ov.collate_frames([(a, b, cc, f) for (a, b, cc), f in zip(vox, frames)])
    grid = np.round((np.array(SMALL_RANGE[3:], np.float32) - np.array(SMALL_RANGE[:3], np.float32))
                    / np.array(SMALL_VOXEL, np.float32)).astype(np.int64)
    return dict(voxels=torch.from_numpy(v), coordinates=torch.from_numpy(c), num_points=torch.from_numpy(n),
                num_voxels=torch.from_numpy(nv), shape=np.stack([grid] * len(frames)), points=torch.from_numpy(pts),
                metadata=[dict(token=i) for i in range(len(frames))])


def gen_backbones():
    """Run the reference's OWN UNetSCN3D.forward / SegNet.forward / SegMSeg3DNet.forward / predict() on the spconv shim
    -> tests/golden/ref_backbones.pt (pins oracle/nets.py::unet_scn3d, segnet_forward, mseg3d_forward, the TTA merge and
    the state-dict key sets)."""
    R = load_reference_detectors()
    from lidarseg3d_b200.det3d import Config
    fx = {}
    g = torch.Generator().manual_seed(77)
    with torch.no_grad():
        # ---- UNetSCN3D alone, SCALING_RATIO 2 (shipped configs) and 1 (constructor default; red_c = 16)
        frames = [surface_scene(11, 1600, 5), surface_scene(12, 1200, 5)]
        ex = _example(frames, 5)
        vf = torch.randn(ex["voxels"].shape[0], 13, generator=g)
        for ratio in (2, 1):
            net = R["scn_unet"].UNetSCN3D(num_input_features=13, voxel_size=SMALL_VOXEL, point_cloud_range=SMALL_RANGE,
                                         model_cfg=dict(SCALING_RATIO=ratio), ds_factor=8, us_factor=8).eval()
            net.load_state_dict(unet_fill(net.state_dict()))
            bd = net(dict(voxel_features=vf.clone(), voxel_coords=ex["coordinates"], batch_size=2, input_shape=ex["shape"][0]))
            ms = bd["multi_scale_3d_features"]
            fx[f"unet_r{ratio}"] = dict(
                voxel_features=vf, coordinates=ex["coordinates"], input_shape=ex["shape"][0],
                conv_point_features=bd["conv_point_features"], conv_point_coords=bd["conv_point_coords"],
                encoded_features=bd["encoded_spconv_tensor"].features, encoded_indices=bd["encoded_spconv_tensor"].indices,
                encoded_shape=list(bd["encoded_spconv_tensor"].spatial_shape),
                multi_scale={k: dict(features=v.features, indices=v.indices, shape=list(v.spatial_shape)) for k, v in ms.items()},
                keys=sorted(net.state_dict().keys()))
        # ---- SpMiddleResNetFHD (scn.py:84-177): the reference's own forward on the shim
        net = R["scn"].SpMiddleResNetFHD(num_input_features=13).eval()
        net.load_state_dict(unet_fill(net.state_dict()))
        dense, ms = net(vf.clone(), ex["coordinates"], 2, ex["shape"][0])
        fx["spmiddle"] = dict(voxel_features=vf, coordinates=ex["coordinates"], input_shape=ex["shape"][0], dense=dense,
                              multi_scale={k: dict(features=v.features, indices=v.indices, shape=list(v.spatial_shape))
                                           for k, v in ms.items()},
                              keys=sorted(net.state_dict().keys()))
        # ---- SegNet (SDSeg3D SemanticKITTI model section of the reference config, small range)
        cfg = Config.fromfile(os.path.join(REF, "configs/semantickitti/SDSeg3D/semkitti_transVFE_unetscn3d_batchloss_e10.py"))
        mc = cfg.model
        mc["backbone"]["voxel_size"], mc["backbone"]["point_cloud_range"] = SMALL_VOXEL, SMALL_RANGE
        mc["pretrained"] = None
        det = R["builder"].build_detector(mc, train_cfg=None, test_cfg=cfg.test_cfg).eval()
        det.load_state_dict(unet_fill(det.state_dict()))

        class _Seq(nn.Module):
            def __init__(self, layers):
                super().__init__()
                self.layers = layers

            def forward(self, x):
                for l in self.layers:
                    x = l(x)
                return x
        det.reader.chunck = _Seq(det.reader.chunck.layers)          # torch>=2 is_causal drift
        frames4 = [surface_scene(21, 1400, 4), surface_scene(22, 1000, 4)]
        ex4 = _example(frames4, 4)
        preds = det(ex4, return_loss=False)
        fx["segnet"] = dict(example={k: v for k, v in ex4.items()}, out_logits=det.point_head.forward_ret_dict["out_logits"],
                            labels=[p["pred_point_sem_labels"] for p in preds],
                            keys=sorted(k.replace("chunck.layers.layers.", "chunck.layers.") for k in det.state_dict().keys()))
        # ---- SegMSeg3DNet (MSeg3D nuScenes model section of the reference config, small range, 2 cameras of 64x96)
        cfg = Config.fromfile(os.path.join(REF, "configs/semanticnusc/MSeg3D/semnusc_avgvfe_unetscn3d_hrnetw18_lr1en2_e12.py"))
        mc = cfg.model
        mc["backbone"]["voxel_size"], mc["backbone"]["point_cloud_range"] = SMALL_VOXEL, SMALL_RANGE
        mc["pretrained"] = None
        mc["img_backbone"]["pretrained"] = None
        mc["img_backbone"]["init_cfg"] = None
        det = R["builder"].build_detector(mc, train_cfg=None, test_cfg=cfg.test_cfg)
        det.eval()
        for m in det.modules():
            m.training = False                                      # the reference HRNet.train() returns None
        det.load_state_dict(unet_fill(det.state_dict()))
        B, ncam = 2, 2
        npts = ex["points"].shape[0]
        cuv = torch.rand(npts, 4, generator=g) * 2 - 1
        cuv[:, 0] = (torch.rand(npts, generator=g) > 0.3).float()
        cuv[:, 1] = torch.randint(0, ncam, (npts,), generator=g).float() / (ncam - 1) * 2 - 1
        ex5 = dict(ex)
        ex5["points_cuv"] = cuv
        ex5["images"] = torch.randn(B, ncam, 3, 64, 96, generator=g)
        preds = det(ex5, return_loss=False)
        fx["mseg3d"] = dict(example=ex5, out_logits=det.point_head.forward_ret_dict["out_logits"],
                            voxel_logits=det.point_head.forward_ret_dict["voxel_logits"],
                            labels=[p["pred_point_sem_labels"] for p in preds], keys=sorted(det.state_dict().keys()))
        # ---- TTA merge (point_seg_mseg3d_head.py:398-453): 4 variants of 2 frames, same point count inside a group
        ntta, nf, n_each, ncls = 4, 2, [37, 53], 17
        rows, logits = [], []
        for f in range(nf):
            for t in range(ntta):
                b = f * ntta + t
                rows.append(torch.cat([torch.full((n_each[f], 1), float(b)), torch.rand(n_each[f], 3, generator=g)], 1))
                logits.append(torch.randn(n_each[f], ncls, generator=g))
        pts_tta, log_tta = torch.cat(rows), torch.cat(logits)
        exT = dict(num_voxels=torch.zeros(nf * ntta), points=pts_tta, metadata=[dict(token=i) for i in range(nf * ntta)],
                   point_sem_labels=torch.randint(0, ncls, (pts_tta.shape[0],), generator=g))
        head = det.point_head
        head.forward_ret_dict = dict(out_logits=log_tta)
        rt = head.predict(example=exT, test_cfg=dict(tta_flag=True, merge_type="ArithmeticMean", num_tta_tranforms=ntta))
        fx["tta"] = dict(points=pts_tta, out_logits=log_tta, point_sem_labels=exT["point_sem_labels"], ntta=ntta,
                         ret=[dict(metadata=r["metadata"], pred=r["pred_point_sem_labels"], gt=r["point_sem_labels"]) for r in rt])
    torch.save(fx, os.path.join(OUT, "ref_backbones.pt"))
    print("backbones:", {k: tuple(v[next(n for n in ("conv_point_features", "out_logits", "dense") if n in v)].shape)
                         for k, v in fx.items()})



def gen_camera_inputs():
    """Drive the reference's OWN LoadPointCloudFromFile (SemanticNuscDataset branch, loading.py:361-416) and
    SegImagePreprocess (segpreprocess.py:388-676, val mode: cv2.resize + normalisation + points_cuv) on a synthetic frame
    -> tests/golden/ref_camera_inputs.npz (pins oracle/camera.py; raw images are regenerated from their seed by the tests)."""
    import hashlib
    import tempfile
    from lidarseg3d_b200 import synth
    for k in [k for k in sys.modules if k.startswith("det3d")]:
        del sys.modules[k]
    install_stubs()
    load_ref("det3d.utils.registry", "det3d/utils/registry.py")
    Registry = sys.modules["det3d.utils.registry"].Registry
    _mod("pycocotools")
    _mod("pycocotools.mask")
    if "turtle" not in sys.modules:          # loading.py:2 has a stray `from turtle import shape` (needs tkinter)
        _mod("turtle", shape=None)
    sys.modules["det3d.core"].box_np_ops = _mod("det3d.core.box_np_ops")
    for pkg in ["det3d.datasets", "det3d.datasets.pipelines", "det3d.core.sampler", "det3d.core.input", "det3d.ops.point_cloud"]:
        m = _mod(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split("."))]
    _mod("det3d.datasets.registry", PIPELINES=Registry("pipeline"))
    _mod("det3d.core.sampler.segpreprocess")
    sys.modules["det3d.core.sampler"].segpreprocess = sys.modules["det3d.core.sampler.segpreprocess"]
    load_ref("det3d.ops.point_cloud.point_cloud_ops", "det3d/ops/point_cloud/point_cloud_ops.py")
    load_ref("det3d.core.input.voxel_generator", "det3d/core/input/voxel_generator.py")
    load_ref("det3d.datasets.pipelines.img_transforms", "det3d/datasets/pipelines/img_transforms.py")
    ld = load_ref("det3d.datasets.pipelines.loading", "det3d/datasets/pipelines/loading.py")
    sp = load_ref("det3d.datasets.pipelines.segpreprocess", "det3d/datasets/pipelines/segpreprocess.py")

    spec = dict(synth.NUSC)
    spec.update(beams=16, azimuths=500)
    pts = synth.lidar_scan(spec, 4242)
    rig = synth.camera_rig(spec)
    # a non-trivial ego pose so that the lidar -> global -> camera chain of the loader is exercised
    a = 0.3
    G = np.eye(4)
    G[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    G[:3, 3] = [431.5, -1207.25, 3.5]
    Ginv = np.linalg.inv(G)
    chans = ["CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_BACK_RIGHT", "CAM_BACK", "CAM_BACK_LEFT", "CAM_FRONT_LEFT"]
    cams_from_global = {c: T @ Ginv for c, (T, K) in zip(chans, rig)}
    intr = {c: K for c, (T, K) in zip(chans, rig)}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "frame.bin")
        pts.astype(np.float32).tofile(path)
        loader = ld.LoadPointCloudFromFile(dataset="SemanticNuscDataset", use_img=True)
        res = dict(lidar=dict(nsweeps=1), cam=dict(chan=chans), mode="val")
        info = dict(lidar_path=path, cams_from_global=cams_from_global, cam_intrinsics=intr, ref_to_global=G)
        res, info = loader(res, info)
    points_cp = res["lidar"]["points_cp"].copy()
    img_seed = 31
    raw = synth.camera_images_u8(spec, img_seed, hw=spec["img_hw"])                       # [6, 900, 1600, 3] uint8
    names = [str(i + 1) for i in range(len(chans))]
    res["images"] = [raw[i] for i in range(len(chans))]
    res["cam"].update(names=names, annotations=None, resized_shape=(spec["net_hw"][1], spec["net_hw"][0]),
                      attributes={n: dict(mean=synth.IMG_MEAN, std=synth.IMG_STD) for n in names})
    from lidarseg3d_b200.det3d.config import ConfigDict
    pre = sp.SegImagePreprocess(cfg=ConfigDict(shuffle_points=False), save_img_for_tta=True)
    res, info = pre(res, info)
    resized = np.stack(res["images_for_tta"])                                              # [6, 640, 960, 3] uint8 (cv2.resize)
    np.savez_compressed(
        os.path.join(OUT, "ref_camera_inputs.npz"), points=pts, ref_to_global=G,
        cams_from_global=np.stack([cams_from_global[c] for c in chans]), intrinsics=np.stack([intr[c] for c in chans]),
        img_hw=np.array(spec["img_hw"]), net_hw=np.array(spec["net_hw"]), points_cp=points_cp,
        points_cuv=res["lidar"]["points_cuv"].astype(np.float32), img_seed=img_seed,
        resized_sha256=np.frombuffer(hashlib.sha256(resized.tobytes()).digest(), np.uint8),
        resized_rows=resized[:, ::40].copy(), raw_sha256=np.frombuffer(hashlib.sha256(raw.tobytes()).digest(), np.uint8),
        images_norm_rows=res["images"][:, :, ::80].astype(np.float32))
    v = res["lidar"]["points_cuv"][:, 0]
    print("camera_inputs: points", pts.shape, "valid", float(v.mean()), "resized", resized.shape)


def gen_image_norm():
    """image_input_transform of the reference (det3d/datasets/pipelines/img_transforms.py:18-29; cv2 is only needed by
    other functions of that file and is stubbed when absent) on random uint8 images -> tests/golden/ref_image_norm.npz."""
    import types
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules["cv2"] = types.SimpleNamespace(INTER_NEAREST=0, INTER_LINEAR=1, INTER_CUBIC=2, INTER_AREA=3,
                                                       INTER_LANCZOS4=4)
    spec = importlib.util.spec_from_file_location("ref_img_transforms",
                                                  os.path.join(REF, "det3d/datasets/pipelines/img_transforms.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.default_rng(99)
    img = rng.integers(0, 256, (3, 20, 32, 3), dtype=np.uint8)
    mean = [0.40789654, 0.44719302, 0.47026115]          # configs/semanticnusc/MSeg3D/...e12.py:19-20
    std = [0.28863828, 0.27408164, 0.27809835]
    out = np.stack([m.image_input_transform(img[i], mean=mean, std=std).astype(np.float32) for i in range(img.shape[0])])
    np.savez_compressed(os.path.join(OUT, "ref_image_norm.npz"), images_u8=img, mean=np.array(mean), std=np.array(std), out=out)
    print("image_norm", out.shape, out.dtype, float(out.mean()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--image-norm-only" in sys.argv:
        gen_image_norm()
        sys.exit(0)
    if "--camera-only" in sys.argv:
        gen_camera_inputs()
        sys.exit(0)
    if "--backbones-only" in sys.argv:
        gen_backbones()
        sys.exit(0)
    if "--losses-only" in sys.argv:
        gen_losses()
        sys.exit(0)
    gen_voxelize()
    gen_modules()
    gen_losses()
    gen_image_norm()
    gen_backbones()
    gen_camera_inputs()
    print(sorted((f, os.path.getsize(os.path.join(OUT, f))) for f in os.listdir(OUT)))
