"""Development: per-role wait-cycle counters of the bf16x3 gather-GEMM (needs LS3D_PROF_SO=1 and build.build(prof=True))."""
import ctypes, os, sys
os.environ["LS3D_PROF_SO"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lidarseg3d_b200 import gemm, capi
dev = "cuda"
NAMES = ["prod.ISSUE", "prod.wait_rawempty", "prod.total", "split.CONVERT", "split.wait_land", "split.wait_aempty",
         "split.tmem_st", "split.total", "mma.ISSUE", "mma.wait_acce", "mma.wait_afull", "mma.wait_wfull", "mma.total",
         "epi.wait_accf", "epi.total", "-"]

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

for name, args in [("sp_57k_c32", (57000, 32, 32, 0.2)), ("sp_90k_c64", (90000, 64, 64, 0.45)), ("sp_42k_c128", (42000, 128, 128, 0.5)),
                   ("dense_100k_64_192", (100000, 64, 192, 1.0, 1, False))]:
    x, pw, nbr, out = setup(*args)
    for skip in (0,):
        gemm.DEBUG_SKIP = skip
        for _ in range(3):
            gemm.run(x, pw, nbr=nbr, out=out)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); gemm.run(x, pw, nbr=nbr, out=out); e1.record(); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * (148 * 16))()
        rc = capi.lib().ls3d_debug_gemm_prof(buf)
        a = np.array(buf[:], dtype=np.float64).reshape(148, 16)
        ntile = (args[0] + 127) // 128
        a = a[: min(148, ntile)]
        print(f"{name} skip={skip} us={e0.elapsed_time(e1) * 1e3:.1f} (single launch incl. launch gap)  mean kcycles per CTA:")
        print("   " + "  ".join(f"{n}={a[:, i].mean() / 1e3:.1f}" for i, n in enumerate(NAMES[:15])))
    gemm.DEBUG_SKIP = 0
