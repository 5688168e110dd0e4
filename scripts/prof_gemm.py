"""Isolated gather-GEMM launches for ncu: a LiDAR-like SubM conv (57k rows, C=32), a C=128 level and a large dense Linear."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidarseg3d_b200 import gemm
dev = "cuda"
gemm.PRECISE = os.environ.get("LS3D_PRECISE", "1") == "1"

def sparse_case(m, cin, cout, fill, koff=27, reps=2):
    g = torch.Generator(device=dev).manual_seed(9)
    x = torch.randn(m, cin, device=dev, generator=g)
    w = torch.randn(koff, cin, cout, device=dev, generator=g) / 10
    base = torch.arange(m, device=dev, dtype=torch.int32)
    nbr = (base[None, :] + torch.randint(-64, 64, (koff, m), device=dev, generator=g, dtype=torch.int32)).clamp_(0, m - 1)
    nbr[torch.rand(koff, m, device=dev, generator=g) > fill] = -1
    nbr[13] = base
    pw = gemm.PackedWeight(w); out = torch.empty(m, cout, device=dev)
    sc = torch.rand(cout, device=dev) + 0.5; sh = torch.randn(cout, device=dev)
    for i in range(reps):
        if i == reps - 1: torch.cuda.synchronize(); torch.cuda.profiler.start()
        gemm.run(x, pw, nbr=nbr, out=out, scale=sc, shift=sh, relu=True, res=x if cin == cout else None, res_mode=1 if cin == cout else 0)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()

def dense_case(m, cin, cout, reps=2):
    x = torch.randn(m, cin, device=dev); w = torch.randn(1, cin, cout, device=dev) / 8
    pw = gemm.PackedWeight(w); out = torch.empty(m, cout, device=dev)
    b = torch.randn(cout, device=dev)
    for i in range(reps):
        if i == reps - 1: torch.cuda.synchronize(); torch.cuda.profiler.start()
        gemm.run(x, pw, out=out, shift=b, relu=True)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()

sparse_case(57000, 32, 32, 0.2)
sparse_case(42000, 128, 128, 0.5)
dense_case(1900000, 64, 192)
