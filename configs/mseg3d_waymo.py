# MSeg3D, Waymo-shaped (23 classes, 5 cameras, HRNet-w18).  Model section of the reference's
# configs/semanticwaymo/MSeg3D/semwaymo_avgvfe_unetscn3d_hrnetw18_lr1en2_e12.py restated for the GPU box
# (the reference tree is not present there); pretrained=None because no checkpoints travel either.
num_class = 23
point_cloud_range = [-75.2, -75.2, -2, 75.2, 75.2, 4]
voxel_size = [0.1, 0.1, 0.15]
norm_cfg = dict(type="BN", requires_grad=True)
widths = (18, 36, 72, 144)
model = dict(
    type="SegMSeg3DNet", pretrained=None,
    img_backbone=dict(type="HRNet", pretrained=None, norm_cfg=norm_cfg, norm_eval=False, frozen_stages=3, extra=dict(
        stage1=dict(num_modules=1, num_branches=1, block="BOTTLENECK", num_blocks=(4,), num_channels=(64,)),
        stage2=dict(num_modules=1, num_branches=2, block="BASIC", num_blocks=(4, 4), num_channels=widths[:2]),
        stage3=dict(num_modules=4, num_branches=3, block="BASIC", num_blocks=(4, 4, 4), num_channels=widths[:3]),
        stage4=dict(num_modules=3, num_branches=4, block="BASIC", num_blocks=(4, 4, 4, 4), num_channels=widths))),
    img_head=dict(type="FCNMSeg3DHead", in_channels=list(widths), in_index=(0, 1, 2, 3), channels=48,
                  input_transform="resize_concat", kernel_size=1, num_convs=2, concat_input=False, dropout_ratio=-1,
                  num_classes=num_class, norm_cfg=norm_cfg, align_corners=False, ignore_index=0, loss_weight=0.5),
    reader=dict(type="ImprovedMeanVoxelFeatureExtractor", num_input_features=5),
    backbone=dict(type="UNetSCN3D", num_input_features=13, ds_factor=8, us_factor=8, point_cloud_range=point_cloud_range,
                  voxel_size=voxel_size, model_cfg=dict(SCALING_RATIO=2)),
    point_head=dict(type="PointSegMSeg3DHead", class_agnostic=False, num_class=num_class, model_cfg=dict(
        VOXEL_IN_DIM=32, VOXEL_CLS_FC=[64], VOXEL_ALIGN_DIM=64, IMAGE_IN_DIM=48, IMAGE_ALIGN_DIM=64, GEO_FUSED_DIM=64,
        OUT_CLS_FC=[64, 64], IGNORED_LABEL=0, DP_RATIO=0.25, MIMIC_FC=[64, 64],
        SFPhase_CFG=dict(embeddings_proj_kernel_size=1, d_model=96, n_head=4, n_layer=6, n_ffn=192, drop_ratio=0,
                         activation="relu", pre_norm=False))))
train_cfg = dict()
test_cfg = dict()
voxel_generator = dict(range=point_cloud_range, voxel_size=voxel_size, max_points_in_voxel=5, max_voxel_num=300000)
