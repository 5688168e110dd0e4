"""Development experiment: time the gather-GEMM with pipeline pieces disabled (debug_skip bits) to locate the bottleneck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidarseg3d_b200 import gemm
dev = "cuda"

def setup(m, cin, cout, fill, koff=27, sparse=True):
    g = torch.Generator(device=dev).manual_seed(9)
    x = torch.randn(m, cin, device=dev, generator=g)
    w = torch.randn(koff, cin, cout, device=dev, generator=g) / 10
    nbr = None
    if sparse:
        base = torch.arange(m, device=dev, dtype=torch.int32)
        nbr = (base[None, :] + torch.randint(-64, 64, (koff, m), device=dev, generator=g, dtype=torch.int32)).clamp_(0, m - 1)
        nbr[torch.rand(koff, m, device=dev, generator=g) > fill] = -1
        nbr[13] = base
    return x, gemm.PackedWeight(w), nbr, torch.empty(m, cout, device=dev)

def t(x, pw, nbr, out, iters=10):
    for _ in range(3): gemm.run(x, pw, nbr=nbr, out=out)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): gemm.run(x, pw, nbr=nbr, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for precise in (True,):
    gemm.PRECISE = precise
    for name, args in [("sp_57k_c32", (57000, 32, 32, 0.2)), ("sp_42k_c128", (42000, 128, 128, 0.5)),
                       ("dense_1.9M_64_192", (1900000, 64, 192, 1.0, 1, False))]:
        x, pw, nbr, out = setup(*args)
        row = {}
        for skip in (0, 79, 79+16, 79+32, 79+16+32, 79+128, 79+16+32+128, 16, 32, 48):
            gemm.DEBUG_SKIP = skip
            row[skip] = round(t(x, pw, nbr, out), 1)
        gemm.DEBUG_SKIP = 0
        print(f"precise={precise} {name} us by skip mask (1=noA 2=noW 4=noMMA 8=noEpi 64=noCpAsyncIssue 256=noGlobalStore 512=noPanel 1024=pollWait):", row, flush=True)
