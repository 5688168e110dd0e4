import os, sys, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
print("torch ok", flush=True); faulthandler.dump_traceback_later(40, exit=True)
import torch.nn.functional as F
from lidarseg3d_b200 import ops, capi
capi.lib(); print("lib ok", flush=True)
dev = "cuda"
def case(n, h, w, cin, cout, res, relu):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, cin, h, w, device=dev, generator=g).half().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(cout, device=dev, generator=g)
    z = torch.randn(n, cout, h, w, device=dev, generator=g).half().contiguous(memory_format=torch.channels_last) if res else None
    pk = ops.pack_conv3x3_f16(wt)
    torch.cuda.synchronize(); print("packed", pk.shape, flush=True)
    y = ops.conv3x3_f16(x, pk, b, res=z, relu=relu)
    print("launched", flush=True)
    torch.cuda.synchronize(); print("synced", flush=True)
    ref = F.conv2d(x.double(), wt.half().double(), b.double(), padding=1)
    if z is not None: ref = ref + z.double()
    if relu: ref = ref.relu()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    print(f"case n={n} {h}x{w} {cin}->{cout} res={res} relu={relu}: rel err {err:.2e}", flush=True)
case(1, 16, 8, 16, 16, False, False)
case(1, 16, 8, 16, 16, True, True)
case(2, 20, 30, 24, 24, True, True)
case(3, 33, 17, 72, 72, False, True)
case(18, 160, 240, 24, 24, True, True)
