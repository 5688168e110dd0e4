"""Stage-by-stage error of the MSeg3D forward vs the oracle, and a kernel-time breakdown of one bench step."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_e2e import _build, _cpu_example
from oracle import nets as on
from lidarseg3d_b200 import pipeline, synth

def rel(a, b):
    return float((a.cpu() - b).abs().max() / b.abs().max())

def stage_errors(tag, image_dtype=None):
    cfg, m = _build("mseg3d_nuscenes.py")
    m.image_dtype = image_dtype
    spec = dict(synth.NUSC); spec.update(beams=16, azimuths=400)
    hw = (128, 192)
    frames = [synth.lidar_scan(spec, s) for s in (0, 1)]
    ex_cpu = _cpu_example(frames, spec, with_cam=True, img_hw=hw)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ocfg = dict(voxel_size=spec["voxel_size"], pc_range=spec["pc_range"], hrnet_extra=cfg.model.img_backbone.extra, nhead=4, nlayer=6, num_convs=2)
    ref = on.mseg3d_forward(sd, ex_cpu, ocfg, return_all=True)
    m = m.to("cuda")
    ex = pipeline.build_example(frames, spec["voxel_size"], spec["pc_range"], images=ex_cpu["images"], points_cuv=ex_cpu["points_cuv"])
    m(ex, return_loss=False)
    bd = m.last_batch_dict; dbg = bd["_ls3d_debug"]
    r = dict(tag=tag,
             image_features=rel(bd["image_features"].reshape(ref["image_features"].shape), ref["image_features"]),
             cam_emb=rel(bd["camera_semantic_embeddings"], ref["camera_semantic_embeddings"].squeeze(-1).permute(0, 2, 1)),
             conv_point_features=rel(bd["conv_point_features"], ref["conv_point_features"]),
             lidar0=rel(dbg["point_features_lidar_0"], ref["point_features_lidar_0"]),
             cam0=rel(dbg["point_features_camera_0"][ex_cpu["points_cuv"][:, 0] == 1], ref["point_features_camera_0"]),
             geo=rel(dbg["geo_fused"], ref["geo_fused"]), lidar_emb=rel(dbg["lidar_emb"], ref["lidar_emb"].squeeze(-1).permute(0, 2, 1)),
             sem_fused=rel(dbg["sem_fused"], ref["sem_fused"]), logits=rel(bd["out_logits"], ref["out_logits"]))
    print(json.dumps(r), flush=True)
    return r

stage_errors("cudnn_tf32")


# ---- kernel-time breakdown of the real bench step
sys.argv = ["bench.py"]
import bench
wl = bench.WORKLOADS["mseg3d_nuscenes"]; spec = synth.NUSC
cfg, model = bench.build_model(wl); model = model.cuda()
b = bench.make_batches(wl, spec, 1, 3, 0)[0]
db = dict(frames=[f.cuda() for f in b["frames"]], cuv=b["cuv"].cuda(), images=b["images"].cuda())
torch.backends.cudnn.benchmark = True
def step():
    ex = pipeline.build_example(db["frames"], spec["voxel_size"], spec["pc_range"], images=db["images"], points_cuv=db["cuv"])
    return model(ex, return_loss=False)
with torch.no_grad():
    for _ in range(3): step()
    torch.cuda.synchronize()
    import time
    t = time.perf_counter(); step(); t_launch = time.perf_counter() - t; torch.cuda.synchronize(); t_all = time.perf_counter() - t
    print("cpu launch s", t_launch, "wall s", t_all)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=70))
