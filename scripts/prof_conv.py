"""Development: a few launches of the fused fp16 3x3 convolution at the HRNet branch shapes (for ncu -k regex:conv3x3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidarseg3d_b200 import ops
dev = "cuda"
for (n, h, w, c) in ((18, 160, 240, 24), (18, 80, 120, 40), (18, 40, 60, 72)):
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, c, h, w, device=dev, generator=g).half().contiguous(memory_format=torch.channels_last)
    z = torch.randn(n, c, h, w, device=dev, generator=g).half().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(c, c, 3, 3, device=dev, generator=g) / (c * 9) ** 0.5)
    b = torch.randn(c, device=dev, generator=g)
    pk = ops.pack_conv3x3_f16(wt)
    for _ in range(2):
        ops.conv3x3_f16(x, pk, b, res=z, relu=True)
    torch.cuda.synchronize()
