"""Development: per-shape time of every ls3d_conv_f16_ex launch of one eager camera-branch forward (dual mode) next to its
HBM floor (bytes the launch must move / measured HBM peak).  Writes gpurun_out/prof_camera.json."""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["bench.py"]
import torch  # noqa: E402

import bench  # noqa: E402
from lidarseg3d_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda")
mode = os.environ.get("CAM_MODE", "dual")
wl = bench.WORKLOADS["mseg3d_nuscenes"]
spec = synth.NUSC
cfg, model = bench.build_model(wl)
model = model.to(dev)
model.use_image_graph = False
model.image_dtype = {"dual": "dual", "fp16": torch.float16}[mode]
batch = bench.to_device(bench.make_batches(wl, spec, 1, 3, 0, n_image_sets=1)[0], dev)
with torch.no_grad():
    ex = bench.build_gpu_example(spec, batch, torch.float32 if mode == "dual" else torch.float16, dev)
    images = ex["images"]
    images = images.view(-1, 3, images.shape[3], images.shape[4]).contiguous(memory_format=torch.channels_last)
    for _ in range(2):
        model._image_branch(images, 3)
    torch.cuda.synchronize()
    ops.CONV_PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if os.environ.get("LS3D_PROFILE_RANGE") == "1":           # ncu --profile-from-start off: this forward only
        torch.cuda.profiler.start()
    e0.record()
    model._image_branch(images, 3)
    e1.record()
    torch.cuda.synchronize()
    if os.environ.get("LS3D_PROFILE_RANGE") == "1":
        torch.cuda.profiler.stop()
prof, ops.CONV_PROFILE = ops.CONV_PROFILE, None
peak = 6541.8e9
agg = collections.OrderedDict()
for p in prof:
    ho, wo = ((p["h"] + 1) // 2, (p["w"] + 1) // 2) if p["stride"] == 2 else (p["h"], p["w"])
    pin, pout = p["n"] * p["h"] * p["w"], p["n"] * ho * wo
    byts = pin * p["cin"] * 2 + pout * p["cout"] * (6 if p["out32"] else 2) + (pout * p["cout"] * 4 if p["res"] else 0)
    key = (p["h"], p["w"], p["cin"], p["cout"], p["k"], p["stride"], p["res"], p["out32"])
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += p["e0"].elapsed_time(p["e1"]) * 1e3
    a[2] += byts / peak * 1e6
rows = [dict(h=k[0], w=k[1], cin=k[2], cout=k[3], k=k[4], stride=k[5], res=k[6], out32=k[7], launches=v[0], us=v[1], floor_us=v[2])
        for k, v in agg.items()]
rows.sort(key=lambda r: -r["us"])
tot = sum(r["us"] for r in rows)
print(f"mode {mode}: camera branch {e0.elapsed_time(e1):.2f} ms eager; {len(prof)} conv launches, {tot / 1e3:.2f} ms in conv events, "
      f"HBM floor {sum(r['floor_us'] for r in rows) / 1e3:.2f} ms")
for r in rows[:40]:
    print(f"{r['h']:4d}x{r['w']:<4d} {r['cin']:3d}->{r['cout']:3d} k{r['k']} s{r['stride']} res={int(r['res'])} o32={int(r['out32'])} "
          f"n={r['launches']:3d} {r['us']:8.1f} us ({r['us'] / r['launches']:6.1f} each) floor {r['floor_us'] / r['launches']:6.1f} each")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(mode=mode, total_ms=e0.elapsed_time(e1), rows=rows), open(os.path.join(ROOT, "gpurun_out", f"prof_camera_{mode}.json"), "w"), indent=1)
