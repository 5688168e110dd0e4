"""HRNet behind the IMG_BACKBONES registry (reference det3d/models/img_backbones/hrnet.py:229-704,
resnet_mmcv.py:20-313).  Same constructor kwargs and state-dict names (mmseg HRNet naming).

Channels-last, BatchNorm folded into the convolutions at inference.  fp16 maps (``image_dtype=torch.float16``): every 3x3
stride-1 conv+BN(+residual)+ReLU runs on csrc/conv3x3_f16.cu and every branch fusion on csrc/upsample_sum.cu; the stem,
stride-2, 1x1 and 144-channel convolutions stay on cuDNN (own kernels for those are a next row).
"""
import os
import warnings

import torch
import torch.nn.functional as F
from torch import nn

from .registry import IMG_BACKBONES


def _bn(c, norm_cfg):
    cfg = dict(norm_cfg or dict(type="BN"))
    cfg.pop("type", None)
    requires_grad = cfg.pop("requires_grad", True)
    cfg.setdefault("eps", 1e-5)
    m = nn.BatchNorm2d(c, **cfg)
    for p in m.parameters():
        p.requires_grad = requires_grad
    return m


FUSED_CUDNN = True      # eval mode: BatchNorm folded into the conv, cuDNN fused conv+bias(+residual)+ReLU


def _versions(*ts):
    return tuple(t._version for t in ts if t is not None)


FUSED_SUM = True        # eval/CUDA/fp32: branch fusion (resize + add + ReLU) in one hand-written kernel (csrc/upsample_sum.cu)
PAD_CHANNELS = True     # eval/CUDA: zero-pad channel counts to 16-byte multiples (18 -> 20 fp32 / 24 fp16) so cuDNN can use
                        # its aligned NHWC tensor-core kernels; padded channels stay exactly zero through conv/BN/ReLU/add


def _pad_to(c, dtype):
    q = 16 // torch.empty((), dtype=dtype).element_size()
    return (c + q - 1) // q * q if PAD_CHANNELS else c


class DualMap:
    """A camera feature map of the "fp32 residual stream" mode: the fp32 map and / or its fp16 tensor-core operand copy.
    ``f32`` is None for a map that only ever feeds an own convolution (conv1 of a BasicBlock); ``f16`` is produced by the
    writing kernel when it is an own one, else on first use (ls3d_cast_f16)."""
    __slots__ = ("f32", "_f16")

    def __init__(self, f32=None, f16=None):
        self.f32, self._f16 = f32, f16

    @property
    def f16(self):
        if self._f16 is None:
            from .. import ops
            self._f16 = ops.cast_f16(self.f32)
        return self._f16

    @property
    def shape(self):
        return (self.f32 if self.f32 is not None else self._f16).shape


def folded(conv, bn, x_channels, dtype, pad_dtype=None):
    """BatchNorm folded into the conv (eval): weight [Cout_p, Cin_p, kh, kw] channels-last and bias [Cout_p], zero padded;
    cached on the conv module, refreshed when a parameter / statistic changes."""
    # one entry per (dtype, channel padding): a captured CUDA graph of one camera-map mode keeps reading its own folded
    # tensors while another mode runs (replacing a single shared entry freed memory a graph still pointed to)
    ver = _versions(conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    store = conv.__dict__.setdefault("_ls3d_fold", {})
    if store.get("ver") != ver:
        store.clear()
        store["ver"] = ver
    pad_dtype = pad_dtype or dtype
    cache = store.get((dtype, x_channels, PAD_CHANNELS, pad_dtype))
    if cache is None:
        with torch.no_grad():
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = conv.weight * scale.view(-1, 1, 1, 1)
            b = bn.bias - bn.running_mean * scale
            cout, cin = w.shape[0], w.shape[1]
            cout_p = _pad_to(cout, pad_dtype)
            assert x_channels >= cin
            if cout_p != cout or x_channels != cin:
                wp = w.new_zeros(cout_p, x_channels, w.shape[2], w.shape[3])
                wp[:cout, :cin] = w
                bp = b.new_zeros(cout_p)
                bp[:cout] = b
                w, b = wp, bp
            w = w.to(dtype).contiguous(memory_format=torch.channels_last)
            b = b.to(dtype).contiguous()
        cache = store[(dtype, x_channels, PAD_CHANNELS, pad_dtype)] = (None, w, b)
    return cache[1], cache[2]


FUSED_CONV3X3 = True    # eval/CUDA/fp16: 3x3 and 1x1 stride-1 convs on the hand-written tcgen05 kernel (csrc/conv3x3_f16.cu)


def folded_packed(conv, bn, x_channels, dual=False):
    """BN-folded 3x3 / 1x1 weights packed for ls3d_conv_f16 (+ fp32 shift), cached like ``folded``; None when the weights do
    not fit the kernel's shared memory (the caller then uses cuDNN)."""
    ver = _versions(conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    store = conv.__dict__.setdefault("_ls3d_fold3", {})
    if store.get("ver") != ver:
        store.clear()
        store["ver"] = ver
    cache = store.get((x_channels, PAD_CHANNELS, dual))
    if cache is None:
        from .. import ops
        k = conv.kernel_size[0]
        cout_p = _pad_to(conv.out_channels, torch.float16)
        other = store.get((x_channels, PAD_CHANNELS, not dual))
        if not (ops.conv_f16_dual_supported if dual else ops.conv_f16_supported)(x_channels, cout_p, k):
            cache = (None, None, None, cout_p)
        elif other is not None and other[1] is not None:
            cache = other                                        # same packed block serves both kernels
        else:
            with torch.no_grad():
                scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                w = (conv.weight * scale.view(-1, 1, 1, 1)).float()
                b = (bn.bias - bn.running_mean * scale).float()
                wp = w.new_zeros(cout_p, x_channels, k, k)
                wp[:conv.out_channels, :conv.in_channels] = w
                bp = b.new_zeros(cout_p)
                bp[:conv.out_channels] = b
                cache = (None, ops.pack_conv_f16(wp), bp.contiguous(), cout_p)
        store[(x_channels, PAD_CHANNELS, dual)] = cache
    return cache[1], cache[2], cache[3]


def _is_plain(conv):
    """3x3 / stride 1 / pad 1 or 1x1 / stride 1 / pad 0: the shapes csrc/conv3x3_f16.cu serves."""
    k = conv.kernel_size
    return (k in ((3, 3), (1, 1)) and conv.stride == (1, 1) and conv.padding == (k[0] // 2, k[0] // 2)
            and conv.dilation == (1, 1) and conv.groups == 1)


# 1x1 convolutions: the kernel serves them (parity-tested), but inside the MSeg3D step routing them away from the library
# GEMM was measured slower on every map size (196.3 frames/s with the library against 191-193): each call is one more
# persistent 148-CTA launch competing with the concurrent LiDAR branch (its sparse-conv launches went from 99 to 128 us), for
# work the library already runs near its bandwidth bound.  Off by default; LS3D_OWN_1X1_MIN_PIXELS=<n> routes 1x1
# convolutions on maps of at least n pixels to the own kernel.
OWN_1X1_MIN_PIXELS = int(os.environ.get("LS3D_OWN_1X1_MIN_PIXELS", 1 << 62))
OWN_1X1_DUAL_MIN_PIXELS = int(os.environ.get("LS3D_OWN_1X1_DUAL_MIN_PIXELS", 1 << 62))      # same switch, fp32-map mode


def _own_conv_ok(conv, x):
    if not (FUSED_CONV3X3 and x.dtype == torch.float16 and _is_plain(conv) and x.shape[1] % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last)):
        return False
    return conv.kernel_size == (3, 3) or x.shape[0] * x.shape[2] * x.shape[3] >= OWN_1X1_MIN_PIXELS


def conv_deferred_bias(conv, bn, x, dual=False):
    if isinstance(x, DualMap):
        if own_dual_ok(conv, bn, x.shape[1]):
            plan = conv_plan(conv, bn, x.shape[1], pixels=x.shape[0] * x.shape[2] * x.shape[3])
            o32, _ = plan.run(x.f16, relu=False, want32=True, use_bias=False)
            return o32, plan.bias
        x = x.f32
    return _conv_deferred_bias(conv, bn, x, dual)


def _conv_deferred_bias(conv, bn, x, dual=False):
    if (F16_OWN_ALL and x.is_cuda and x.dtype == torch.float16 and x.shape[1] % 8 == 0 and not bn.training
            and x.is_contiguous(memory_format=torch.channels_last)):
        y = _cbr_f16_plan(conv, bn, x, False, None, use_bias=False)
        if y is not None:
            plan = conv_plan(conv, bn, x.shape[1], fp32out=False) or conv_plan(conv, bn, x.shape[1], fp32out=True)
            return y, plan.bias                                  # (the folded shift is the same vector in every plan)
    """conv -> BatchNorm (folded) WITHOUT the shift: returns (conv(x, w_folded), shift fp32 [Cout_p]).  For the linear terms
    of a branch fusion: the caller sums the shifts of all terms and ls3d_upsample_sum adds them once (a library convolution
    with a bias and no activation would run a separate elementwise add over every output map)."""
    if _own_conv_ok(conv, x):
        wp, bp, cout_p = folded_packed(conv, bn, x.shape[1])
        if wp is not None:
            from .. import ops
            return ops.conv_f16(x, wp, None, relu=False, cout=cout_p, ksize=conv.kernel_size[0]), bp
    w, b = folded(conv, bn, x.shape[1], x.dtype, pad_dtype=torch.float16 if dual else None)
    return F.conv2d(x, w, None, conv.stride, conv.padding, conv.dilation, conv.groups), b


DUAL_EXACT_WEIGHTS = os.environ.get("LS3D_DUAL_EXACT_WEIGHTS", "1") == "1"
DUAL_OWN_ALL = os.environ.get("LS3D_DUAL_OWN_ALL", "1") == "1"     # 0: only 3x3 stride-1 convs on the own kernel (development A/B)


# streamed-weight kernel (ls3d_conv_f16_kb) for the weight-heavy, pixel-light 3x3 / stride-1 convolutions: at least this many
# (padded) input channels and few enough output tiles (<= ~4 per SM) that streaming the weights once per tile group is cheaper
# than slicing the convolution into resident-weight passes
USE_KB = os.environ.get("LS3D_CONV_KB", "1") == "1"
KB_MIN_CIN = 64
KB_MIN_S2_WEIGHT = 24 * 72           # stride-2 fuse convolutions into the 72- / 144-channel branches
KB_MAX_PIXELS = 80000
KB_MIN_1X1_COUT = 128
KB_1X1 = os.environ.get("LS3D_CONV_KB_1X1", "0") == "1"


class ConvPlan:
    """One BN-folded convolution (3x3 stride 1 / 2, or 1x1) as launches of ls3d_conv_f16_ex (csrc/conv3x3_f16.cu).

    ``fp32out`` plans write an fp32 map + its fp16 operand copy (dual launches) and may cut the INPUT channels into slices whose
    launches accumulate through the fp32 residual input; operand-only plans (fp16 output, fp16 residual) are single-slice.
    Output channels are cut into slices when the accumulator (split weights: 2 n_pad <= 256 columns) or the staged output tile
    demands it.  With ``exact`` the fp32 weights enter exactly: split weights [W_hi ; W_lo] in one launch where the doubled
    weight block fits shared memory, else W_hi and W_lo as two accumulating launches ("hilo") when that needs fewer launches
    than slicing."""

    LAUNCH_US = 8.0          # cost model: fixed cost of a launch (prologue: resident weights, barriers, tensor memory)
    BYTES_PER_US = 4.5e6     # and sustained HBM bytes per microsecond

    @staticmethod
    def _cost(pixels, stride, cin_p, cout_p, ci, cs, hilo, fp32out, has_res):
        """Estimated microseconds of the launch list: per output slice every input-slice launch reads its input channels,
        re-reads the fp32 partial sums of the previous launch and writes the output slice again."""
        p_out = pixels / (4 if stride == 2 else 1)
        n_o, n_i = (cout_p + cs - 1) // cs, (cin_p + ci - 1) // ci
        per_in = n_i * (2 if hilo else 1)
        byts = n_o * per_in * pixels * ci * 2
        byts += n_o * per_in * p_out * cs * (6 if fp32out else 2)
        byts += n_o * (per_in - 1 + (1 if has_res else 0)) * p_out * cs * (4 if fp32out else 2)
        return n_o * per_in * ConvPlan.LAUNCH_US + byts / ConvPlan.BYTES_PER_US

    def __init__(self, w, bias, ksize, stride, exact=True, fp32out=True, pixels=100000):
        from .. import ops
        cout_p, cin_p = w.shape[:2]
        self.cout_p, self.cin_p, self.k, self.stride, self.exact, self.fp32out = cout_p, cin_p, ksize, stride, exact, fp32out
        self.bias = None if bias is None else bias.float().contiguous()
        self.kb = False
        # weight-heavy 3x3 / stride-1 convolutions (the 72- / 144-channel branches): streamed-weight kernel, whole input
        # channel range per launch, output slices only where the accumulator (2 n_pad <= 256 columns) demands it
        out_pixels = pixels // (4 if stride == 2 else 1)
        heavy = cin_p >= KB_MIN_CIN if stride == 1 else cin_p * cout_p >= KB_MIN_S2_WEIGHT
        use_kb = USE_KB and exact and ksize == 3 and heavy and out_pixels <= KB_MAX_PIXELS
        # (the kernel also serves 1x1 convolutions - 128-channel slices of Bottleneck conv3, 64 -> 256, fit thanks to its direct
        # epilogue - but at 160x240 its per-thread global stores lose to the staged TMA stores: measured 645 us against 554 us
        # for the three 88-channel passes of the resident-weight kernel, so the route stays off)
        use_kb = use_kb or (KB_1X1 and USE_KB and exact and ksize == 1 and stride == 1 and fp32out and cout_p >= KB_MIN_1X1_COUT)
        if use_kb:
            for n_out in (1, 2, 3, 4):
                cs = ((cout_p + n_out - 1) // n_out + 7) // 8 * 8
                if cs <= 128 and ops.conv_kb_supported(cin_p, cs, ksize, stride, fp32out, True, out_pixels):
                    self.kb, self.ok, self.hilo, self.split, self.cs, self.ci = True, True, False, True, cs, cin_p
                    self.passes = []
                    for o in range((cout_p + cs - 1) // cs):
                        o0 = min(o * cs, cout_p - cs)
                        self.passes.append((ops.pack_conv_ex(w[o0:o0 + cs].clone(), stride, True), 0, o0, True, True))
                    self.n_launch, self.multi, self._tables, self.est_us = 1, False, {}, 0.0
                    return
        best = None
        for hilo in ((False, True) if (exact and fp32out) else (False,)):
            split = exact and not hilo
            for n_out in (1, 2, 3, 4, 6, 8):
                cs = ((cout_p + n_out - 1) // n_out + 7) // 8 * 8
                if split and cs > 128:
                    continue
                for n_in in ((1, 2, 3, 4, 6, 8, 12, 16, 32) if fp32out else (1,)):
                    ci = ((cin_p + n_in - 1) // n_in + 7) // 8 * 8
                    if ops.conv_ex_supported(ci, cs, ksize, stride, dual=fp32out, split=split):
                        cand = (self._cost(pixels, stride, cin_p, cout_p, ci, cs, hilo, fp32out, True), hilo, cs, ci)
                        best = cand if best is None or cand < best else best
                        break
        self.ok = best is not None
        if not self.ok:
            return
        self.est_us, self.hilo, cs, ci = best
        self.split = exact and not self.hilo
        self.cs, self.ci = cs, ci
        # uniform slices: the last slice of a dimension is shifted back so that it ends with the tensor; input channels it
        # shares with the previous slice get zero weights, output channels it shares are simply computed twice
        self.passes = []            # (packed weights, in_off, out_off, first-of-out-slice, last-of-out-slice)
        n_o, n_i = (cout_p + cs - 1) // cs, (cin_p + ci - 1) // ci
        for o in range(n_o):
            o0 = min(o * cs, cout_p - cs)
            row = []
            for i in range(n_i):
                i0 = min(i * ci, cin_p - ci)
                ws = w[o0:o0 + cs, i0:i0 + ci].clone()
                if i0 < i * ci:
                    ws[:, :i * ci - i0] = 0
                if self.hilo:
                    hi = ws.half().float()
                    row.append((ops.pack_conv_ex(hi, stride, False), i0, o0))
                    row.append((ops.pack_conv_ex(ws - hi, stride, False), i0, o0))
                else:
                    row.append((ops.pack_conv_ex(ws, stride, self.split), i0, o0))
            for j, (wp, i0, oo) in enumerate(row):
                self.passes.append((wp, i0, oo, j == 0, j == len(row) - 1))
        self.n_launch = len(self.passes)
        self.multi = n_i * (2 if self.hilo else 1) > 1
        self._tables = {}

    def _table(self, has_res, relu, use_bias):
        from .. import capi
        key = (has_res, relu, use_bias)
        t = self._tables.get(key)
        if t is None:
            arr = (capi.ConvPass * len(self.passes))()
            for k, (wp, i0, o0, first, last) in enumerate(self.passes):
                f = 0
                if first:
                    f |= (1 if use_bias and self.bias is not None else 0) | (2 if has_res else 0)
                else:
                    f |= 4
                if last and relu:
                    f |= 8
                arr[k].w_packed, arr[k].in_c_off, arr[k].out_c_off, arr[k].flags = wp.data_ptr(), i0, o0, f
            t = self._tables[key] = arr
        return t

    def run(self, x16, res=None, relu=True, want32=True, use_bias=True):
        """x16 [N, cin_p, H, W] fp16 channels-last -> (out32 or None, out16): ONE launch, the slices are its passes.  ``res``:
        fp32 map for fp32out plans, fp16 map for operand-only plans."""
        from .. import ops
        N, C, H, W = x16.shape
        assert C == self.cin_p and x16.dtype == torch.float16 and x16.is_contiguous(memory_format=torch.channels_last)
        Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if self.stride == 2 else (H, W)
        out16 = torch.empty((N, self.cout_p, Ho, Wo), dtype=torch.float16, device=x16.device, memory_format=torch.channels_last)
        out32 = None
        if self.fp32out:
            assert res is None or res.dtype == torch.float32
            out32 = torch.empty((N, self.cout_p, Ho, Wo), dtype=torch.float32, device=x16.device,
                                memory_format=torch.channels_last)
        else:
            assert res is None or res.dtype == torch.float16
        tab = self._table(res is not None, bool(relu), bool(use_bias))
        if self.kb:
            ops.conv_kb(x16, tab, len(self.passes), self.bias if use_bias else None, cout=self.cs, out32=out32, out16=out16,
                        res32=res if self.fp32out else None, res16=None if self.fp32out else res, split=True, stride=self.stride,
                        ksize=self.k)
            return out32, out16
        ops.conv_multi(x16, tab, len(self.passes), self.bias if use_bias else None, cin=self.ci, cout=self.cs, out32=out32,
                       out16=out16, res32=res if self.fp32out else None, res16=None if self.fp32out else res, ksize=self.k,
                       stride=self.stride, split=self.split)
        return out32, out16


def _plannable(conv):
    k, st = conv.kernel_size, conv.stride
    if conv.groups != 1 or conv.dilation != (1, 1) or k[0] != k[1] or st[0] != st[1]:
        return False
    if k == (3, 3) and conv.padding == (1, 1) and st in ((1, 1), (2, 2)):
        return True
    return k == (1, 1) and conv.padding == (0, 0) and st == (1, 1)


def conv_plan(conv, bn, x_channels, fp32out=True, pixels=100000):
    """The cached ConvPlan of conv (+ folded eval BatchNorm ``bn``, may be None: the conv's own bias) for an input map with
    ``x_channels`` (zero-padded) channels; None when the own kernel does not serve the shape."""
    ts = (conv.weight, conv.bias) if bn is None else (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    ver = _versions(*ts)
    store = conv.__dict__.setdefault("_ls3d_plan", {})
    if store.get("ver") != ver:
        store.clear()
        store["ver"] = ver
    key = (x_channels, DUAL_EXACT_WEIGHTS, fp32out, int(pixels).bit_length())
    if key not in store:
        plan = None
        if _plannable(conv) and x_channels % 8 == 0 and x_channels >= conv.in_channels:
            with torch.no_grad():
                k = conv.kernel_size[0]
                cout_p = _pad_to(conv.out_channels, torch.float16)
                if bn is None:
                    w = conv.weight.float()
                    b = conv.bias.float() if conv.bias is not None else torch.zeros(conv.out_channels, device=w.device)
                else:
                    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                    w = (conv.weight * scale.view(-1, 1, 1, 1)).float()
                    b = (bn.bias - bn.running_mean * scale).float()
                wp = w.new_zeros(cout_p, x_channels, k, k)
                wp[:conv.out_channels, :conv.in_channels] = w
                bp = b.new_zeros(cout_p)
                bp[:conv.out_channels] = b
                plan = ConvPlan(wp, bp, k, conv.stride[0], DUAL_EXACT_WEIGHTS, fp32out, pixels)
                if not plan.ok:
                    plan = None
        store[key] = plan
    return store[key]


def own_dual_ok(conv, bn, channels):
    """Will _cbr_dual run ``conv`` on the own kernel (so that its input may be an operand-only map)?"""
    if not FUSED_CONV3X3:
        return False
    if not DUAL_OWN_ALL and not (conv.kernel_size == (3, 3) and conv.stride == (1, 1)):
        return False
    return conv_plan(conv, bn, channels) is not None


def _cbr_dual(conv, bn, x, relu, z, want):
    """cbr on DualMaps (fp32 maps, fp16 tensor-core operands).  want: "f16" = the result only feeds own convolutions (no fp32
    map is written), "both" / "f32" = fp32 map (+ operand copy)."""
    C = x.shape[1]
    if own_dual_ok(conv, bn, C):
        pix = x.shape[0] * x.shape[2] * x.shape[3]
        plan = conv_plan(conv, bn, C, pixels=pix)
        if want == "f16" and z is None:
            op = conv_plan(conv, bn, C, fp32out=False, pixels=pix)       # operand-only: no fp32 map is written
            if op is not None and op.est_us <= plan.est_us:
                return DualMap(None, op.run(x.f16, relu=relu)[1])
        if z is None or z.shape[1] == plan.cout_p:
            o32, o16 = plan.run(x.f16, res=None if z is None else z.f32, relu=relu)
            return DualMap(o32, o16)
    w, b = folded(conv, bn, C, torch.float32, pad_dtype=torch.float16)
    xf, zf = x.f32, (None if z is None else z.f32)
    if FUSED_CUDNN and relu and conv.groups == 1:
        if zf is None:
            return DualMap(torch.cudnn_convolution_relu(xf, w, b, conv.stride, conv.padding, conv.dilation, 1))
        return DualMap(torch.cudnn_convolution_add_relu(xf, w, zf, 1.0, b, conv.stride, conv.padding, conv.dilation, 1))
    y = F.conv2d(xf, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    if zf is not None:
        y = y + zf
    return DualMap(torch.relu_(y) if relu else y)


F16_OWN_ALL = os.environ.get("LS3D_F16_OWN_ALL", "1") == "1"     # fp16 maps: every convolution on the own kernel, exact weights


def _cbr_f16_plan(conv, bn, x, relu, z, use_bias=True):
    """fp16 maps: conv (+ folded BN) (+ z) (+ ReLU) on ls3d_conv_f16_ex with exact weights; None when not served."""
    from .. import ops
    C = x.shape[1]
    if not FUSED_CONV3X3:
        return None
    pix = x.shape[0] * x.shape[2] * x.shape[3]
    plan = conv_plan(conv, bn, C, fp32out=False, pixels=pix)
    alt = conv_plan(conv, bn, C, fp32out=True, pixels=pix)    # sliced shapes accumulate in an fp32 temporary
    if plan is not None and (alt is None or plan.est_us <= alt.est_us):
        if z is not None and (z.shape[1] != plan.cout_p or not z.is_contiguous(memory_format=torch.channels_last)):
            return None
        return plan.run(x, res=z, relu=relu, use_bias=use_bias)[1]
    plan = alt
    if plan is None or (z is not None and z.shape[1] != plan.cout_p):
        return None
    z32 = None if z is None else ops.cast_f32(z.contiguous(memory_format=torch.channels_last))
    return plan.run(x, res=z32, relu=relu, use_bias=use_bias)[1]


def cbr(conv, bn, x, relu, z=None, want="both"):
    """conv -> BatchNorm -> (+z) -> (ReLU).  Training / CPU: plain modules.  Eval on CUDA: BN folded into the conv weights
    (cached, refreshed when a parameter changes); fp16 3x3 / 1x1 stride-1 convs run on the hand-written tensor-core kernel, the rest
    as one cuDNN call."""
    if isinstance(x, DualMap):
        return _cbr_dual(conv, bn, x, relu, z, want)
    if bn.training or not x.is_cuda:
        y = bn(conv(x))
        if z is not None:
            y = y + z
        return torch.relu(y) if relu else y
    if F16_OWN_ALL and x.dtype == torch.float16 and x.shape[1] % 8 == 0 and x.is_contiguous(memory_format=torch.channels_last):
        y = _cbr_f16_plan(conv, bn, x, relu, z)
        if y is not None:
            return y
    if _own_conv_ok(conv, x):
        wp, bp, cout_p = folded_packed(conv, bn, x.shape[1])
        if wp is not None and (z is None or (z.shape[1] == cout_p and z.is_contiguous(memory_format=torch.channels_last))):
            from .. import ops
            return ops.conv_f16(x, wp, bp, res=z, relu=relu, cout=cout_p, ksize=conv.kernel_size[0])
    w, b = folded(conv, bn, x.shape[1], x.dtype)
    if FUSED_CUDNN and relu and conv.groups == 1:
        if z is None:
            return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, 1)
        return torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, conv.stride, conv.padding, conv.dilation, 1)
    y = F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    if z is not None:
        y = y + z
    return torch.relu_(y) if relu else y


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm_cfg=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = _bn(planes, norm_cfg)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = _bn(planes, norm_cfg)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else cbr(self.downsample[0], self.downsample[1], x, False)
        want = "both"
        if isinstance(x, DualMap) and own_dual_ok(self.conv2, self.bn2, _pad_to(self.conv1.out_channels, torch.float16)):
            want = "f16"                                          # conv1's result is only ever conv2's operand
        out = cbr(self.conv1, self.bn1, x, True, want=want)
        return cbr(self.conv2, self.bn2, out, True, z=idt)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm_cfg=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = _bn(planes, norm_cfg)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = _bn(planes, norm_cfg)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _bn(planes * 4, norm_cfg)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else cbr(self.downsample[0], self.downsample[1], x, False)
        w1 = w2 = "both"
        if isinstance(x, DualMap):              # conv1 / conv2 results are operands of the next convolution only
            c1 = _pad_to(self.conv1.out_channels, torch.float16)
            if own_dual_ok(self.conv2, self.bn2, c1):
                w1 = "f16"
            if own_dual_ok(self.conv3, self.bn3, _pad_to(self.conv2.out_channels, torch.float16)):
                w2 = "f16"
        out = cbr(self.conv1, self.bn1, x, True, want=w1)
        out = cbr(self.conv2, self.bn2, out, True, want=w2)
        return cbr(self.conv3, self.bn3, out, True, z=idt)


class Upsample(nn.Module):
    """det3d/ops/mmseg_ops/wrappers.py:30-51 (size = int(t * scale_factor))."""

    def __init__(self, scale_factor, mode="bilinear", align_corners=False):
        super().__init__()
        self.scale_factor, self.mode, self.align_corners = float(scale_factor), mode, align_corners

    def forward(self, x):
        size = [int(t * self.scale_factor) for t in x.shape[-2:]]
        return F.interpolate(x, size, None, self.mode, self.align_corners)


class HRModule(nn.Module):
    """hrnet.py:24-226."""

    def __init__(self, num_branches, block, num_blocks, in_channels, num_channels, multiscale_output, norm_cfg):
        super().__init__()
        self.in_channels = in_channels
        self.num_branches = num_branches
        self.multiscale_output = multiscale_output
        branches = []
        for i in range(num_branches):
            layers, down = [], None
            if in_channels[i] != num_channels[i] * block.expansion:
                down = nn.Sequential(nn.Conv2d(in_channels[i], num_channels[i] * block.expansion, 1, bias=False),
                                     _bn(num_channels[i] * block.expansion, norm_cfg))
            layers.append(block(in_channels[i], num_channels[i], 1, down, norm_cfg))
            in_channels[i] = num_channels[i] * block.expansion
            for _ in range(1, num_blocks[i]):
                layers.append(block(in_channels[i], num_channels[i], norm_cfg=norm_cfg))
            branches.append(nn.Sequential(*layers))
        self.branches = nn.ModuleList(branches)
        self.fuse_layers = self._make_fuse_layers(norm_cfg)
        self.relu = nn.ReLU(inplace=False)

    def _make_fuse_layers(self, norm_cfg):
        if self.num_branches == 1:
            return None
        nb, ch = self.num_branches, self.in_channels
        fuse = []
        for i in range(nb if self.multiscale_output else 1):
            row = []
            for j in range(nb):
                if j > i:
                    row.append(nn.Sequential(nn.Conv2d(ch[j], ch[i], 1, bias=False), _bn(ch[i], norm_cfg),
                                             Upsample(2 ** (j - i), "bilinear", False)))
                elif j == i:
                    row.append(None)
                else:
                    downs = []
                    for k in range(i - j):
                        if k == i - j - 1:
                            downs.append(nn.Sequential(nn.Conv2d(ch[j], ch[i], 3, 2, 1, bias=False), _bn(ch[i], norm_cfg)))
                        else:
                            downs.append(nn.Sequential(nn.Conv2d(ch[j], ch[j], 3, 2, 1, bias=False), _bn(ch[j], norm_cfg),
                                                       nn.ReLU(inplace=False)))
                    row.append(nn.Sequential(*downs))
            fuse.append(nn.ModuleList(row))
        return nn.ModuleList(fuse)

    def forward(self, x):
        if self.num_branches == 1:
            return [self.branches[0](x[0])]
        x = [self.branches[i](x[i]) for i in range(self.num_branches)]
        if isinstance(x[0], DualMap):
            return self._forward_fused(x)
        if FUSED_SUM and x[0].is_cuda and not self.training and x[0].dtype in (torch.float32, torch.float16) and all(
                t.shape[1] % (4 if t.dtype == torch.float32 else 8) == 0 and x[0].shape[2] == t.shape[2] << j
                and x[0].shape[3] == t.shape[3] << j for j, t in enumerate(x)):
            return self._forward_fused(x)
        outs = []
        for i in range(len(self.fuse_layers)):
            y = 0
            for j in range(self.num_branches):
                if i == j:
                    y = y + x[j]
                elif j > i:
                    fl = self.fuse_layers[i][j]                     # conv1x1 + BN + Upsample, then resize (hrnet.py:216-220)
                    t = fl[2](cbr(fl[0], fl[1], x[j], False))
                    if t.shape[2:] != x[i].shape[2:]:            # the reference's second resize; identity when sizes agree
                        t = F.interpolate(t, size=x[i].shape[2:], mode="bilinear", align_corners=False)
                    y = y + t
                else:
                    t = x[j]
                    for seq in self.fuse_layers[i][j]:              # conv3x3 s2 + BN (+ReLU except the last)
                        t = cbr(seq[0], seq[1], t, len(seq) == 3)
                    y = y + t
            outs.append(self.relu(y))
        return outs


def _forward_fused(self, x):
    """Eval / CUDA: every output branch = ONE ls3d_upsample_sum launch over its terms in the reference's j order (same-size
    terms as they are, coarser 1x1-conv outputs resized inside the kernel), the terms' folded-BN shifts summed into one bias
    vector (cached) that the kernel adds once, ReLU fused."""
    from .. import ops
    outs = []
    cache = self.__dict__.setdefault("_ls3d_fuse_bias", {})
    dual = isinstance(x[0], DualMap)
    for i in range(len(self.fuse_layers)):
        terms, shifts = [], []
        for j in range(self.num_branches):
            if i == j:
                terms.append(x[j].f32 if dual else x[j])
            elif j > i:
                fl = self.fuse_layers[i][j]
                t, b = conv_deferred_bias(fl[0], fl[1], x[j], dual)
                terms.append(t)
                shifts.append(b)
            else:
                t = x[j]
                for seq in self.fuse_layers[i][j]:
                    if len(seq) == 3:
                        t = cbr(seq[0], seq[1], t, True, want="f16")      # feeds the next stride-2 convolution only
                    else:
                        t, b = conv_deferred_bias(seq[0], seq[1], t, dual)
                        shifts.append(b)
                terms.append(t)
        key = (i,) + tuple((b.data_ptr(), b._version) for b in shifts)      # one entry per camera-map mode (see ``folded``)
        ent = cache.get(key)
        if ent is None:
            if len(cache) > 32:
                cache.clear()
            with torch.no_grad():
                ent = cache[key] = (key, torch.stack([b.float() for b in shifts]).sum(0).contiguous() if shifts else None)
        if dual:
            outs.append(DualMap(*ops.upsample_sum_dual(terms, relu=True, bias=ent[1])))
        else:
            outs.append(ops.upsample_sum(terms, relu=True, bias=ent[1]))
    return outs


HRModule._forward_fused = _forward_fused


@IMG_BACKBONES.register_module
class HRNet(nn.Module):
    blocks_dict = {"BASIC": BasicBlock, "BOTTLENECK": Bottleneck}

    def __init__(self, extra, in_channels=3, conv_cfg=None, norm_cfg=dict(type="BN", requires_grad=True),
                 norm_eval=False, with_cp=False, frozen_stages=-1, zero_init_residual=False, multiscale_output=True,
                 pretrained=None, init_cfg=None):
        super().__init__()
        self.pretrained, self.extra, self.norm_cfg = pretrained, extra, norm_cfg
        self.norm_eval, self.frozen_stages = norm_eval, frozen_stages
        self.conv1 = nn.Conv2d(in_channels, 64, 3, 2, 1, bias=False)
        self.bn1 = _bn(64, norm_cfg)
        self.conv2 = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
        self.bn2 = _bn(64, norm_cfg)
        self.relu = nn.ReLU(inplace=True)
        s1 = extra["stage1"]
        block = self.blocks_dict[s1["block"]]
        c = s1["num_channels"][0]
        down = None
        if 64 != c * block.expansion:
            down = nn.Sequential(nn.Conv2d(64, c * block.expansion, 1, bias=False), _bn(c * block.expansion, norm_cfg))
        layers = [block(64, c, 1, down, norm_cfg)] + [block(c * block.expansion, c, norm_cfg=norm_cfg)
                                                       for _ in range(1, s1["num_blocks"][0])]
        self.layer1 = nn.Sequential(*layers)
        pre = [c * block.expansion]
        for st in (2, 3, 4):
            cfg = extra[f"stage{st}"]
            blk = self.blocks_dict[cfg["block"]]
            chans = [ch * blk.expansion for ch in cfg["num_channels"]]
            setattr(self, f"transition{st - 1}", self._make_transition(pre, chans))
            mods, inch = [], list(chans)
            for i in range(cfg["num_modules"]):
                ms = True if (multiscale_output or st != 4 or i != cfg["num_modules"] - 1) else False
                mods.append(HRModule(cfg["num_branches"], blk, cfg["num_blocks"], inch, list(cfg["num_channels"]), ms,
                                     norm_cfg))
            setattr(self, f"stage{st}", nn.Sequential(*mods))
            pre = inch
        if isinstance(pretrained, str):
            self.load_pretrained_model()
        self._freeze_stages()

    def _make_transition(self, pre, cur):
        layers = []
        for i in range(len(cur)):
            if i < len(pre):
                if cur[i] != pre[i]:
                    layers.append(nn.Sequential(nn.Conv2d(pre[i], cur[i], 3, 1, 1, bias=False), _bn(cur[i], self.norm_cfg),
                                                nn.ReLU(inplace=True)))
                else:
                    layers.append(None)
            else:
                downs = []
                for j in range(i + 1 - len(pre)):
                    cin = pre[-1]
                    cout = cur[i] if j == i - len(pre) else cin
                    downs.append(nn.Sequential(nn.Conv2d(cin, cout, 3, 2, 1, bias=False), _bn(cout, self.norm_cfg),
                                               nn.ReLU(inplace=True)))
                layers.append(nn.Sequential(*downs))
        return nn.ModuleList(layers)

    def load_pretrained_model(self):
        """hrnet.py:435-483 loads ``pretrained`` unconditionally; the drop-in tolerates a missing file
         and reports it."""
        import os
        if not os.path.isfile(self.pretrained):
            warnings.warn(f"HRNet pretrained weights not found at {self.pretrained}; keeping random init")
            return
        sd = torch.load(self.pretrained, map_location="cpu")
        sd = sd.get("state_dict", sd)
        sd = {k[len("backbone."):] if k.startswith("backbone.") else k: v for k, v in sd.items()}
        self.load_state_dict(sd, strict=False)

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            for m in (self.conv1, self.bn1, self.conv2, self.bn2):
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            mods = [getattr(self, "layer1" if i == 1 else f"stage{i}")]
            if i < 4:
                mods.append(getattr(self, f"transition{i}"))
            for m in mods:
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False

    def forward(self, x):
        dual = getattr(self, "dual_maps", False) and x.is_cuda and not self.training and x.dtype == torch.float32
        if dual and x.shape[1] == 3 and x.is_contiguous(memory_format=torch.channels_last) and own_dual_ok(self.conv1, self.bn1, 8):
            from .. import ops
            x = DualMap(None, ops.pad3_f16(x))               # operand copy of the image, 8-channel pixel rows
            x = cbr(self.conv1, self.bn1, x, True, want="f16" if own_dual_ok(self.conv2, self.bn2, 64) else "both")
            l0 = self.layer1[0]                               # its conv1 and downsample conv are the only readers of the stem map
            only_ops = (l0.downsample is not None and own_dual_ok(l0.conv1, l0.bn1, 64)
                        and own_dual_ok(l0.downsample[0], l0.downsample[1], 64))
            x = cbr(self.conv2, self.bn2, x, True, want="f16" if only_ops else "both")
        else:
            if (F16_OWN_ALL and x.is_cuda and not self.training and x.dtype == torch.float16 and x.shape[1] == 3
                    and conv_plan(self.conv1, self.bn1, 8, fp32out=False) is not None):
                x = F.pad(x, (0, 0, 0, 0, 0, 5)).contiguous(memory_format=torch.channels_last)   # 16-byte pixel rows
            x = cbr(self.conv1, self.bn1, x, True)
            x = cbr(self.conv2, self.bn2, x, True)
            if dual:
                x = DualMap(x)
        x = self.layer1(x)
        ys = [x]
        for st in (2, 3, 4):
            tr = getattr(self, f"transition{st - 1}")
            xs = []
            for i in range(self.extra[f"stage{st}"]["num_branches"]):
                if tr[i] is not None:
                    t = ys[0] if st == 2 else ys[-1]
                    if isinstance(tr[i][0], nn.Conv2d):              # same resolution: conv3x3 + BN + ReLU
                        t = cbr(tr[i][0], tr[i][1], t, True)
                    else:                                            # new branch: chain of conv3x3 s2 + BN + ReLU
                        for seq in tr[i]:
                            t = cbr(seq[0], seq[1], t, True)
                    xs.append(t)
                else:
                    xs.append(ys[i])
            ys = getattr(self, f"stage{st}")(xs)
        if dual and not getattr(self, "keep_dual_maps", False):
            ys = [y.f32 for y in ys]
        if dual and getattr(self, "keep_dual_maps", False):
            return ys
        if not getattr(self, "keep_channel_padding", False):
            true_c = [c * self.blocks_dict[self.extra["stage4"]["block"]].expansion for c in self.extra["stage4"]["num_channels"]]
            ys = [y if y.shape[1] == c else y[:, :c] for y, c in zip(ys, true_c)]
        return ys

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
        return self
