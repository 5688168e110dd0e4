"""ncu target: ONE eager MSeg3D forward of a bench batch inside a cudaProfilerStart/Stop range, plus the static description
(rows, channels, offsets, rulebook pairs, algorithmic bytes) of every gather-GEMM launch of that forward in launch order
(gpurun_out/unet_launches.json), so that scripts/ncu_traffic.py can put ncu's per-launch DRAM / L2 / tensor-pipe counters
next to the algorithmic bytes of the SAME launch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["bench.py"]
import torch  # noqa: E402

import bench  # noqa: E402
from lidarseg3d_b200 import gemm, synth  # noqa: E402

dev = torch.device("cuda")
wl = bench.WORKLOADS["mseg3d_nuscenes"]
spec = synth.NUSC
cfg, model = bench.build_model(wl)
model = model.to(dev)
model.use_image_graph = False
batch = bench.to_device(bench.make_batches(wl, spec, 1, 3, 0, n_image_sets=1)[0], dev)
with torch.no_grad():
    for _ in range(2):
        model(bench.build_gpu_example(spec, batch, torch.float32, dev), return_loss=False)
    inputs_in_range = os.environ.get("LS3D_PROFILE_INPUTS") == "1"      # also profile projection / resize / voxelization
    if not inputs_in_range:
        ex = bench.build_gpu_example(spec, batch, torch.float32, dev)
    gemm.PROFILE, gemm.COUNT = [], []
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if inputs_in_range:
        ex = bench.build_gpu_example(spec, batch, torch.float32, dev)
    model(ex, return_loss=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
rows = []
for p, pairs in zip(gemm.PROFILE, gemm.COUNT):
    byts = (p["rows_in"] * p["cin"] + p["m_out"] * p["cout"] + p["koff"] * p["cin"] * p["cout"]) * 4 + (pairs * 8 if p["sparse"] else 0)
    rows.append(dict(sparse=p["sparse"], rows_in=p["rows_in"], m_out=p["m_out"], cin=p["cin"], cout=p["cout"], koff=p["koff"],
                     pairs=pairs, algorithmic_bytes=byts, flops=2.0 * pairs * p["cin"] * p["cout"],
                     event_us=p["e0"].elapsed_time(p["e1"]) * 1e3))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(engine=gemm.ENGINE_NAME, launches=rows), open(os.path.join(ROOT, "gpurun_out", "unet_launches.json"), "w"), indent=1)
print("gather-GEMM launches in one forward:", len(rows), "sparse:", sum(r["sparse"] for r in rows))
