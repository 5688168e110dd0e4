"""CPU experiment (oracle only): exact weights everywhere, fp16 (RNE) rounding of the conv INPUTS (what the dual mode does), plus
optionally fp16 STORAGE of feature maps (every conv+BN(+residual)+ReLU output rounded) in (a) the stem + layer1 only, (b) the
whole camera branch.  Logit error / argmax agreement against the unrounded oracle.  Usage: python scripts/sim_storage_rounding.py [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import bench  # noqa: E402
from lidarseg3d_b200 import synth  # noqa: E402
from oracle import nets as on  # noqa: E402

fpg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
wl = bench.WORKLOADS["mseg3d_nuscenes"]
spec = synth.NUSC
cfg, model = bench.build_model(wl)
batches = bench.make_batches(wl, spec, 1, fpg, 0)
torch.set_num_threads(os.cpu_count())
sd = {k: v.detach() for k, v in model.state_dict().items()}
ex = bench.cpu_inputs(wl, spec, batches[0], fpg)
r11 = lambda x: x.half().float()
orig_conv, orig_bb, orig_bn = on._conv2d, on._basic_block, on._bottleneck
MODE = dict(store="none")


def conv(sd_, p, x, stride=1, padding=0):
    return F.conv2d(r11(x), sd_[p + ".weight"], sd_.get(p + ".bias"), stride=stride, padding=padding)


def bb(sd_, p, x):
    y = orig_bb(sd_, p, x)
    return r11(y) if MODE["store"] == "all" else y


def bn(sd_, p, x):
    y = orig_bn(sd_, p, x)
    return r11(y) if MODE["store"] in ("all", "layer1") else y


with torch.no_grad():
    ref = bench.cpu_forward(wl, spec, cfg, sd, ex, return_all=True)["out_logits"]
    on._conv2d, on._basic_block, on._bottleneck = conv, bb, bn
    for store in ("none", "layer1", "all"):
        MODE["store"] = store
        t0 = time.time()
        out = bench.cpu_forward(wl, spec, cfg, sd, ex, return_all=True)["out_logits"]
        rel = float((out - ref).abs().max() / ref.abs().max())
        d = (out - ref).abs().max(1).values / ref.abs().max()
        print(f"store={store:7s} rel_err {rel:.3e} median {float(d.median()):.3e} mismatches {int((out.argmax(1) != ref.argmax(1)).sum())} of {ref.shape[0]}"
              f"  ({time.time() - t0:.0f} s)", flush=True)
