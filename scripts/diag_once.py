"""GPU diagnostic: where does a sparse-convolution launch spend its time at FRAME sizes?  Real level-1..4 rulebooks of a
3-frame nuScenes-like batch; serial launches (no camera branch), CUDA events over 20 repetitions:
  * both engines (gather-once / per-pair), * m_out cut to the first n tiles (fixed cost vs per-tile cost),
  * the once kernel with parts switched off (debug_skip: 1 stager loads, 2 splitter copies, 4 MMA, 8 epilogue, 16 W copies).
Writes gpurun_out/diag_once.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["bench.py"]
import torch  # noqa: E402

import bench  # noqa: E402
from lidarseg3d_b200 import gemm, ops, synth  # noqa: E402

dev = torch.device("cuda")
wl = bench.WORKLOADS["mseg3d_nuscenes"]
spec = synth.NUSC
cfg, model = bench.build_model(wl)
model = model.to(dev)
model.use_image_graph = False
batch = bench.to_device(bench.make_batches(wl, spec, 1, 3, 0, n_image_sets=1)[0], dev)
with torch.no_grad():
    model(bench.build_gpu_example(spec, batch, torch.float32, dev), return_loss=False)
bd = model.last_batch_dict
levels = bd["_ls3d_levels"]
rb = bd["_ls3d_rulebooks"]
C = {1: 32, 2: 64, 3: 128, 4: 128}
out = []


def timeit(fn, n=20):
    """n back-to-back launches replayed from a CUDA graph (no host launch cost in the figure)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def case(name, nbr, cin, cout, rows_in):
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(rows_in, cin, device=dev, generator=g)
    w = torch.randn(nbr.shape[0], cin, cout, device=dev, generator=g) / (8 * cin) ** 0.5
    pw = gemm.PackedWeight(w)
    m = nbr.shape[1]
    pairs = int((nbr >= 0).sum())
    rec = dict(name=name, rows_in=rows_in, m_out=m, cin=cin, cout=cout, pairs=pairs, tiles=(m + 127) // 128)
    o = torch.empty(m, cout, device=dev)
    for eng in ("once", "pair"):
        gemm.USE_PLAN = eng == "once"
        rec[eng + "_us"] = timeit(lambda: gemm.run(x, pw, nbr=nbr, out=o))
    gemm.USE_PLAN = True
    fast = os.environ.get("DIAG_FAST") == "1"
    for skip in (() if fast else (1, 2, 4, 8, 16, 3, 7, 23, 31)):
        gemm.DEBUG_SKIP = skip
        rec[f"once_skip{skip}_us"] = timeit(lambda: gemm.run(x, pw, nbr=nbr, out=o))
    gemm.DEBUG_SKIP = 0
    for nt in (() if fast else (37, 148, 296, 592)):
        mm = min(nt * 128, m)
        sub = nbr[:, :mm].contiguous()
        rec[f"once_tiles{nt}_us"] = timeit(lambda: gemm.run(x, pw, nbr=sub, out=o[:mm]))
    out.append(rec)
    print(json.dumps(rec), flush=True)


with torch.no_grad():
    for lv in (1, 2, 3, 4):
        nbr = levels[lv].subm_table()
        case(f"subm_l{lv}", nbr, C[lv], C[lv], nbr.shape[1])
    for lv in (2, 3, 4):
        case(f"down_l{lv}", rb["down"][lv], C[lv - 1], C[lv], levels[lv - 1].coords.shape[0])
        case(f"up_l{lv}", rb["up"][lv], C[lv], C[lv - 1], levels[lv].coords.shape[0])
    # empty-kernel launch cost for scale
    t = torch.empty(1, device=dev)
    out.append(dict(name="torch_fill_launch", us=timeit(lambda: t.fill_(1.0))))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_once.json"), "w"), indent=1)
