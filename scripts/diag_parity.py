"""GPU diagnostic: full-size (3-frame) parity of every camera-branch mode against the CPU oracle, stage by stage, plus the
distribution of the oracle's top-2 logit margins (how sensitive the argmax gate is).  Writes gpurun_out/diag_parity.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["bench.py"]
import torch  # noqa: E402

import bench  # noqa: E402
from lidarseg3d_b200 import synth  # noqa: E402

dev = torch.device("cuda")
fpg = int(os.environ.get("FPG", 3))
wl = bench.WORKLOADS["mseg3d_nuscenes"]
spec = synth.NUSC
cfg, model = bench.build_model(wl)
model = model.to(dev)
batch = bench.make_batches(wl, spec, 1, fpg, 0, n_image_sets=1)[0]
bench.cpu_threads()
sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
with torch.no_grad():
    ex_cpu = bench.cpu_inputs(wl, spec, batch, fpg)
    ref = bench.cpu_forward(wl, spec, cfg, sd, ex_cpu, return_all=True)
rl = ref["out_logits"]
top2 = rl.topk(2, dim=1).values
margin = (top2[:, 0] - top2[:, 1]) / rl.abs().max()
res = dict(points=int(rl.shape[0]), margin_quantiles={str(q): float(torch.quantile(margin, q)) for q in (0.0005, 0.001, 0.002, 0.005, 0.01, 0.05, 0.5)},
           frac_margin_below={str(t): float((margin < t).float().mean()) for t in (1e-5, 3e-5, 1e-4, 3e-4, 1e-3)}, modes={})


def rel(a, b):
    return float((a.float().cpu() - b).abs().max() / b.abs().max())


modes = [("fp32_exact", torch.float32, False), ("fp32lib_tf32", torch.float32, True), ("dual", "dual", True), ("fp16cam", torch.float16, True)]
for name, dt, tf32 in modes:
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    model.use_image_graph = tf32          # the exact-fp32 run goes eagerly (its own cuDNN algorithms)
    try:
        ex, bd = bench.gpu_forward(wl, spec, model, batch, dt, dev)
        lg = bd["out_logits"].float().cpu()
        dbg = bd["_ls3d_debug"]
        m = dict(rel_err=rel(lg, rl), mismatches=int((lg.argmax(1) != rl.argmax(1)).sum()),
                 argmax_agreement=float((lg.argmax(1) == rl.argmax(1)).float().mean()),
                 image_features=rel(bd["image_features"].reshape(ref["image_features"].shape), ref["image_features"]),
                 conv_point_features=rel(bd["conv_point_features"], ref["conv_point_features"]),
                 geo=rel(dbg["geo_fused"], ref["geo_fused"]), sem_fused=rel(dbg["sem_fused"], ref["sem_fused"]))
        d = (lg - rl).abs().max(1).values / rl.abs().max()
        m["logit_err_quantiles"] = {str(q): float(torch.quantile(d, q)) for q in (0.5, 0.9, 0.99, 0.999)}
        bad = lg.argmax(1) != rl.argmax(1)
        m["margin_of_mismatches_max"] = float(margin[bad].max()) if bad.any() else 0.0
    except Exception as e:
        m = dict(error=repr(e)[:400])
    res["modes"][name] = m
    print(name, json.dumps(m), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "diag_parity.json"), "w"), indent=1)
print(json.dumps({k: res[k] for k in ("points", "margin_quantiles", "frac_margin_below")}))
