"""Development: fixed vs per-tile cost of the fused 3x3 convolution (time against the number of images)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidarseg3d_b200 import ops
from scripts.diag_conv import timeit
dev = "cuda"
for (h, w, c) in ((160, 240, 24), (80, 120, 40), (40, 60, 72)):
    for n in (1, 3, 9, 18, 36, 72):
        g = torch.Generator(device=dev).manual_seed(1)
        x = torch.randn(n, c, h, w, device=dev, generator=g).half().contiguous(memory_format=torch.channels_last)
        wt = (torch.randn(c, c, 3, 3, device=dev, generator=g) / (c * 9) ** 0.5)
        b = torch.randn(c, device=dev, generator=g)
        pk = ops.pack_conv3x3_f16(wt)
        t = timeit(lambda: ops.conv3x3_f16(x, pk, b, res=None, relu=True))
        tiles = n * ((h + 15) // 16) * ((w + 7) // 8)
        print(f"c={c} {h}x{w} n={n}: {t:.1f} us, {tiles} tiles, {tiles / 148:.1f} per CTA", flush=True)
