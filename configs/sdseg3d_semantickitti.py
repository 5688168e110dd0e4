# SDSeg3D, SemanticKITTI-shaped (20 classes, LiDAR only).  Model section of the reference's
# configs/semantickitti/SDSeg3D/semkitti_transVFE_unetscn3d_batchloss_e10.py restated for the GPU box.
num_class = 20
point_cloud_range = [-75.2, -75.2, -4, 75.2, 75.2, 2]
voxel_size = [0.1, 0.1, 0.15]
model = dict(
    type="SegNet", pretrained=None,
    reader=dict(type="TransformerVoxelFeatureExtractor", num_input_features=4, num_compressed_features=16, num_embed=64,
                num_head=4, num_layers=3),
    backbone=dict(type="UNetSCN3D", num_input_features=16, ds_factor=8, us_factor=8, point_cloud_range=point_cloud_range,
                  voxel_size=voxel_size, model_cfg=dict(SCALING_RATIO=2)),
    point_head=dict(type="PointSegBatchlossHead", class_agnostic=False, num_class=num_class,
                    model_cfg=dict(CONV_IN_DIM=32, CONV_CLS_FC=[64], CONV_ALIGN_DIM=64, OUT_CLS_FC=[64, 64], IGNORED_LABEL=0)))
train_cfg = dict()
test_cfg = dict()
voxel_generator = dict(range=point_cloud_range, voxel_size=voxel_size, max_points_in_voxel=5, max_voxel_num=300000)
