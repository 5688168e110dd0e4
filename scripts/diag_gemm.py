"""GPU diagnostic for the tcgen05 gather-GEMM: prints max errors for a ladder of cases."""
import sys, os, json, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidarseg3d_b200 import gemm

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
res = []

def prep(x):
    """single-pass mode expects tf32-representable activations (the tensor core truncates)"""
    return x if gemm.PRECISE else gemm.round_tf32(x)

def ref_gemm(x, w_kio, nbr):
    # fp64 reference: exact operands in precise (3xTF32) mode, tf32-rounded operands otherwise
    xr = x.double() if gemm.PRECISE else gemm.round_tf32(x).double()
    wr = w_kio.double() if gemm.PRECISE else gemm.round_tf32(w_kio).double()
    koff = w_kio.shape[0]
    m = nbr.shape[1] if nbr is not None else x.shape[0]
    out = torch.zeros(m, w_kio.shape[2], dtype=torch.float64, device=x.device)
    for k in range(koff):
        if nbr is None:
            out += xr @ wr[k]
        else:
            idx = nbr[k].long(); ok = idx >= 0
            out[ok] += xr[idx[ok]] @ wr[k]
    return out

def case(name, fn):
    try:
        err = fn()
        torch.cuda.synchronize()
        res.append((name, err)); print(name, err, flush=True)
    except Exception as e:
        traceback.print_exc(); res.append((name, "EXC " + repr(e))); print(name, "EXC", e, flush=True)

def dense(m, cin, cout, seed=0):
    def f():
        g = torch.Generator(device=dev).manual_seed(seed)
        x = prep(torch.randn(m, cin, device=dev, generator=g))
        w = torch.randn(1, cin, cout, device=dev, generator=g) / cin ** 0.5
        pw = gemm.PackedWeight(w)
        y = gemm.run(x, pw)
        r = ref_gemm(x, w, None)
        return float((y.double() - r).abs().max())
    return f

def ident():
    x = prep(torch.randn(128, 32, device=dev))
    w = torch.eye(32, device=dev).unsqueeze(0)
    y = gemm.run(x, gemm.PackedWeight(w))
    d = (y - x).abs()
    if d.max() > 0:
        bad = (d > 0).nonzero()[:8].tolist()
        print("ident mismatches at", bad, y[0, :8].tolist(), x[0, :8].tolist())
    return float(d.max())

def sparse(m_in, m_out, cin, cout, koff=27, fill=0.3, seed=1):
    def f():
        g = torch.Generator(device=dev).manual_seed(seed)
        x = prep(torch.randn(m_in, cin, device=dev, generator=g))
        w = torch.randn(koff, cin, cout, device=dev, generator=g) / (cin * koff * fill) ** 0.5
        nbr = torch.randint(0, m_in, (koff, m_out), device=dev, generator=g, dtype=torch.int32)
        drop = torch.rand(koff, m_out, device=dev, generator=g) > fill
        nbr[drop] = -1
        nbr[3:9, :256] = -1   # whole-tile-empty offsets -> exercises step skipping
        y = gemm.run(x, gemm.PackedWeight(w), nbr=nbr)
        r = ref_gemm(x, w, nbr)
        return float((y.double() - r).abs().max())
    return f

def epilogue():
    g = torch.Generator(device=dev).manual_seed(5)
    m, cin, cout = 777, 64, 96
    x = prep(torch.randn(m, cin, device=dev, generator=g))
    w = torch.randn(1, cin, cout, device=dev, generator=g) / 8
    sc = torch.rand(cout, device=dev, generator=g) + 0.5
    sh = torch.randn(cout, device=dev, generator=g)
    rs = torch.randn(m, cout, device=dev, generator=g)
    g0 = torch.rand(cout, device=dev, generator=g) + 0.5; b0 = torch.randn(cout, device=dev, generator=g)
    g1 = torch.rand(cout, device=dev, generator=g) + 0.5; b1 = torch.randn(cout, device=dev, generator=g)
    y = gemm.run(x, gemm.PackedWeight(w), scale=sc, shift=sh, relu=True, res=rs, res_mode=1,
                 ln=((g0, b0), (g1, b1)))
    r = ref_gemm(x, w, None) * sc.double() + sh.double() + rs.double()
    r = torch.relu(r)
    r = torch.nn.functional.layer_norm(r, (cout,), g0.double(), b0.double(), 1e-5)
    r = torch.nn.functional.layer_norm(r, (cout,), g1.double(), b1.double(), 1e-5)
    return float((y.double() - r).abs().max())

def concat_red():
    g = torch.Generator(device=dev).manual_seed(6)
    m, c = 500, 64
    a = prep(torch.randn(m, c, device=dev, generator=g)); b = prep(torch.randn(m, c, device=dev, generator=g))
    w = torch.randn(1, 2 * c, c, device=dev, generator=g) / 11
    y = gemm.run(a, gemm.PackedWeight(w), x1=b, relu=True, red=(a, b))
    cat = torch.cat([a, b], 1)
    r = torch.relu(ref_gemm(cat, w, None)) + cat.double().view(m, c, 2).sum(2)
    return float((y.double() - r).abs().max())

def attn():
    g = torch.Generator(device=dev).manual_seed(7)
    m, e, H, L, F = 1000, 96, 4, 34, 3
    x = prep(torch.randn(m, e, device=dev, generator=g))
    w = torch.randn(1, e, e, device=dev, generator=g) / e ** 0.5
    bias = torch.randn(e, device=dev, generator=g)
    K = torch.randn(F, H, L, 24, device=dev, generator=g); V = torch.randn(F, H, L, 24, device=dev, generator=g)
    fo = torch.tensor([0, 300, 650], dtype=torch.int32, device=dev)
    y = gemm.run(x, gemm.PackedWeight(w), shift=bias, attn=dict(k=K, v=V, frame_off=fo, scale=24 ** -0.5))
    q = (ref_gemm(x, w, None) + bias.double()).view(m, H, 24)
    fid = torch.bucketize(torch.arange(m, device=dev), fo[1:].long(), right=True)
    Kf = K.double()[fid]; Vf = V.double()[fid]           # [m,H,L,24]
    s = torch.einsum("mhd,mhld->mhl", q, Kf) * 24 ** -0.5
    o = torch.einsum("mhl,mhld->mhd", s.softmax(-1), Vf).reshape(m, e)
    return float((y.double() - o).abs().max())

def ladder(tag):
  global res
  print("=== mode", tag, flush=True)
  _ladder()

def _ladder():
  case("ident128x32", ident)
  case("dense_128_32_32", dense(128, 32, 32))
  case("dense_1000_64_64", dense(1000, 64, 64))
  case("dense_1000_16_32", dense(1000, 16, 32))
  case("dense_1000_48_64", dense(1000, 48, 64))
  case("dense_5000_128_128", dense(5000, 128, 128))
  case("dense_5000_256_128", dense(5000, 256, 128))
  case("dense_5000_96_192", dense(5000, 96, 192))
  case("dense_5000_192_96", dense(5000, 192, 96))
  case("dense_5000_96_17", dense(5000, 96, 17))
  case("dense_300_13p_32", dense(300, 16, 32))
  case("dense_3000_64_16", dense(3000, 64, 16))
  case("dense_3000_32_48", dense(3000, 32, 48))
  case("dense_3000_40_80", dense(3000, 40, 80))
  case("sparse_20000_32_32", sparse(20000, 20000, 32, 32))
  case("sparse_6000_128_128", sparse(6000, 5000, 128, 128))
  case("sparse_3000_256_128", sparse(3000, 3000, 256, 128, fill=0.5))
  case("epilogue_ln2", epilogue)
  case("concat_red", concat_red)
  case("attn", attn)

# quick timing of a realistic level-1 SubM conv: 57k sites, 32->32, ~5 nbrs/site
def timing(m, cin, cout, fill, koff=27, iters=20):
    g = torch.Generator(device=dev).manual_seed(9)
    x = torch.randn(m, cin, device=dev, generator=g)
    w = torch.randn(koff, cin, cout, device=dev, generator=g) / 10
    base = torch.arange(m, device=dev, dtype=torch.int32)
    nbr = (base[None, :] + torch.randint(-64, 64, (koff, m), device=dev, generator=g, dtype=torch.int32)).clamp_(0, m - 1)
    nbr[torch.rand(koff, m, device=dev, generator=g) > fill] = -1
    nbr[13] = base
    pw = gemm.PackedWeight(w); out = torch.empty(m, cout, device=dev)
    for _ in range(3): gemm.run(x, pw, nbr=nbr, out=out)
    torch.cuda.synchronize()
    # CUDA graph of `iters` launches: the python/ctypes launch path (~30 us) must not bound the measurement
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(iters): gemm.run(x, pw, nbr=nbr, out=out)
    gr.replay(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): gr.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters / 3
    pairs = int((nbr >= 0).sum())
    byts = m * cin * 4 + m * cout * 4 + koff * cin * cout * 4 + pairs * 8
    return dict(m=m, cin=cin, cout=cout, ms=ms, pairs=pairs, gbs=byts / ms / 1e6, gflops=2 * pairs * cin * cout / ms / 1e6)

def timings():
  try:
    for cfg in [(57000, 32, 32, 0.2), (90000, 64, 64, 0.45), (42000, 128, 128, 0.5), (17000, 128, 128, 0.5), (2000000, 32, 32, 0.2), (2000000, 64, 64, 0.3)]:
        t = timing(*cfg); t["precise"] = gemm.PRECISE; print("timing", t, flush=True); res.append(("timing", t))
  except Exception as e:
    traceback.print_exc()

MODES = [int(v) for v in os.environ.get("LS3D_MODES", "2,1,0").split(",")]
for mode in MODES:
    gemm.PRECISE = mode
    ladder({2: "bf16x3", 1: "precise(3xTF32)", 0: "single-pass TF32"}[mode])
    timings()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/diag_gemm.json", "w"), indent=1)
