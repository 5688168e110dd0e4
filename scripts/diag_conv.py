"""Development: correctness + timing of the fused fp16 3x3 convolution (csrc/conv3x3_f16.cu) against cuDNN."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from lidarseg3d_b200 import ops
dev = "cuda"
torch.backends.cudnn.benchmark = True

def case(n, h, w, cin, cout, res, relu, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(n, cin, h, w, device=dev, generator=g).half().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(cout, device=dev, generator=g)
    z = torch.randn(n, cout, h, w, device=dev, generator=g).half().contiguous(memory_format=torch.channels_last) if res else None
    pk = ops.pack_conv3x3_f16(wt)
    y = ops.conv3x3_f16(x, pk, b, res=z, relu=relu)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), wt.half().double(), b.double(), padding=1)
    if z is not None: ref = ref + z.double()
    if relu: ref = ref.relu()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    print(f"case n={n} {h}x{w} {cin}->{cout} res={res} relu={relu}: rel err {err:.2e}", flush=True)
    return err

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(iters): fn()
    gr.replay(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 3 * 1e3

def timing(n, h, w, c):
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n, c, h, w, device=dev, generator=g).half().contiguous(memory_format=torch.channels_last)
    z = torch.randn(n, c, h, w, device=dev, generator=g).half().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(c, c, 3, 3, device=dev, generator=g) / (c * 9) ** 0.5)
    b = torch.randn(c, device=dev, generator=g)
    pk = ops.pack_conv3x3_f16(wt)
    wh = wt.half().contiguous(memory_format=torch.channels_last); bh = b.half()
    t_mine = timeit(lambda: ops.conv3x3_f16(x, pk, b, res=None, relu=True))
    t_mine_r = timeit(lambda: ops.conv3x3_f16(x, pk, b, res=z, relu=True))
    t_cud = timeit(lambda: torch.cudnn_convolution_relu(x, wh, bh, (1, 1), (1, 1), (1, 1), 1))
    t_cud_r = timeit(lambda: torch.cudnn_convolution_add_relu(x, wh, z, 1.0, bh, (1, 1), (1, 1), (1, 1), 1))
    x32, z32, w32 = x.float(), z.float(), wt.contiguous(memory_format=torch.channels_last)
    t_c32 = timeit(lambda: torch.cudnn_convolution_relu(x32, w32, b, (1, 1), (1, 1), (1, 1), 1))
    byts = n * h * w * c * 2
    print(f"timing n={n} {h}x{w} c={c}: ls3d {t_mine:.1f} us (+res {t_mine_r:.1f}) | cudnn fp16 {t_cud:.1f} (+res {t_cud_r:.1f}) | cudnn tf32 {t_c32:.1f}"
          f" | ideal HBM {2 * byts / 6.5e6:.1f} us (+res {3 * byts / 6.5e6:.1f})", flush=True)

if __name__ == "__main__":
    case(1, 16, 8, 16, 16, False, False)
    case(1, 16, 8, 32, 32, False, False)
    case(2, 20, 30, 24, 24, True, True)
    case(1, 40, 60, 40, 40, True, True)
    case(3, 33, 17, 72, 72, False, True)
    case(1, 160, 240, 24, 24, True, True)
    timing(18, 160, 240, 24)
    timing(18, 80, 120, 40)
    timing(18, 40, 60, 72)
