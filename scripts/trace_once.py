"""Development: CTA-0 event timeline of the gather-once kernel on the real level-1/2/3 SubM rulebooks of a bench batch
(LS3D_PROF build: python -c 'from lidarseg3d_b200 import build; build.build(prof=True)').  Writes gpurun_out/trace_once.json."""
import ctypes
import json
import os
import sys

os.environ["LS3D_PROF_SO"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["bench.py"]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from lidarseg3d_b200 import capi, gemm, synth  # noqa: E402

dev = torch.device("cuda")
wl = bench.WORKLOADS["mseg3d_nuscenes"]
spec = synth.NUSC
cfg, model = bench.build_model(wl)
model = model.to(dev)
model.use_image_graph = False
batch = bench.to_device(bench.make_batches(wl, spec, 1, 3, 0, n_image_sets=1)[0], dev)
with torch.no_grad():
    model(bench.build_gpu_example(spec, batch, torch.float32, dev), return_loss=False)
levels = model.last_batch_dict["_ls3d_levels"]
C = {1: 32, 2: 64, 3: 128}
CAP = 2048
out = {}
L = capi.lib()
L.ls3d_debug_once_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
for lv in (1, 2, 3):
    nbr = levels[lv].subm_table()
    m = nbr.shape[1]
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(m, C[lv], device=dev, generator=g)
    w = torch.randn(27, C[lv], C[lv], device=dev, generator=g) / 30
    pw = gemm.PackedWeight(w)
    o = torch.empty(m, C[lv], device=dev)
    for skip in (0, 31):
        gemm.DEBUG_SKIP = skip
        for _ in range(3):
            gemm.run(x, pw, nbr=nbr, out=o)
        torch.cuda.synchronize()
        L.ls3d_debug_once_trace_reset()
        gemm.run(x, pw, nbr=nbr, out=o)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * (8 * CAP))()
        cnt = (ctypes.c_uint * 8)()
        L.ls3d_debug_once_trace(buf, cnt)
        a = np.array(buf[:], dtype=np.uint64).reshape(8, CAP)
        roles = {}
        t0 = None
        for r in range(8):
            ev = [(int(v >> np.uint64(56)), int((v >> np.uint64(40)) & np.uint64(0xFFFF)), int(v & np.uint64(0xFFFFFFFFFF))) for v in a[r, :cnt[r]]]
            roles[r] = ev
            for e in ev:
                t0 = e[2] if t0 is None else min(t0, e[2])
        out[f"l{lv}_skip{skip}"] = {str(r): [(e[0], e[1], e[2] - t0) for e in ev] for r, ev in roles.items()}
    gemm.DEBUG_SKIP = 0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "trace_once.json"), "w"))
print("ok", {k: {r: len(v) for r, v in d.items()} for k, d in out.items()})
