"""Development: where one MSeg3D bench step spends its time - camera branch alone (fp32 / fp16, graph replay), LiDAR branch
alone, full step (overlapped and serial), plus a per-kernel table of the fp16 camera branch."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.argv = ["bench.py"]
import bench
from lidarseg3d_b200 import pipeline, synth

wl = bench.WORKLOADS["mseg3d_nuscenes"]; spec = synth.NUSC
cfg, model = bench.build_model(wl); model = model.cuda()
b = bench.make_batches(wl, spec, 1, 3, 0)[0]
from lidarseg3d_b200 import ops
db = dict(frames=[f.cuda() for f in b["frames"]], cuv=b["cuv"].cuda(), images_u8=b["images_u8"].cuda())
db["images"] = ops.normalize_images_u8(db["images_u8"], synth.IMG_MEAN, synth.IMG_STD, torch.float16)
torch.backends.cudnn.benchmark = True

def ev_time(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def step():
    ex = pipeline.build_example(db["frames"], spec["voxel_size"], spec["pc_range"], images=db["images"], points_cuv=db["cuv"])
    return model(ex, return_loss=False)

def lidar_only():
    ex = pipeline.build_example(db["frames"], spec["voxel_size"], spec["pc_range"], images=db["images"], points_cuv=db["cuv"])
    return model._lidar_branch(ex)

imgs = db["images"].view(-1, 3, db["images"].shape[3], db["images"].shape[4]).contiguous(memory_format=torch.channels_last)
def image_only():
    outs, side = model._image_branch_graphed(imgs, 3)
    torch.cuda.current_stream().wait_stream(side)

res = {}
with torch.no_grad():
    for name, dt in (("fp16", torch.float16),):
        model.image_dtype = dt
        model.__dict__.pop("_img_graphs", None)
        res[f"image_branch_{name}_ms"] = ev_time(image_only)
        model.use_image_graph = True
        res[f"step_{name}_overlapped_ms"] = ev_time(step)
        model.use_image_graph = False
        res[f"step_{name}_serial_ms"] = ev_time(step)
        model.use_image_graph = True
    res["lidar_branch_ms"] = ev_time(lidar_only)
    print(json.dumps(res), flush=True)
    model.image_dtype = torch.float16
    model.use_image_graph = False
    from torch.profiler import profile, ProfilerActivity
    step(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
