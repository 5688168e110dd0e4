"""CPU experiment (oracle only, no product code): how much logit error / argmax disagreement does rounding the camera-branch
convolution OPERANDS to an 11-bit significand cause, and which side (activations, weights) carries it?  Guides the choice of
the own convolution's arithmetic.  Usage: python scripts/sim_operand_rounding.py [frames]"""
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


def r11(x):          # fp16 rounding = 11-bit significand (values are far from the fp16 range limits here)
    return x.half().float()


def make(round_x, round_w, split_w=False):
    def conv(sd_, p, x, stride=1, padding=0):
        w = sd_[p + ".weight"]
        if round_x:
            x = r11(x)
        if round_w:
            w = r11(w)
        return F.conv2d(x, w, sd_.get(p + ".bias"), stride=stride, padding=padding)
    return conv


orig = on._conv2d
t0 = time.time()
with torch.no_grad():
    ref = bench.cpu_forward(wl, spec, cfg, sd, ex, return_all=True)["out_logits"]
print(f"reference forward {time.time() - t0:.1f} s, {ref.shape[0]} points", flush=True)
for name, (rx, rw) in dict(both=(True, True), act_only=(True, False), w_only=(False, True)).items():
    on._conv2d = make(rx, rw)
    with torch.no_grad():
        out = bench.cpu_forward(wl, spec, cfg, sd, ex, return_all=True)["out_logits"]
    on._conv2d = orig
    rel = float((out - ref).abs().max() / ref.abs().max())
    agree = float((out.argmax(1) == ref.argmax(1)).float().mean())
    print(f"{name:9s} rel_err {rel:.3e}  argmax agreement {agree:.5f}  mismatches {int((out.argmax(1) != ref.argmax(1)).sum())}",
          flush=True)
