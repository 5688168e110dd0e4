"""Tensor-level wrappers over the C ABI (one function per include/ls3d.h entry point).

PyTorch is used for device memory and the current CUDA stream only; all arithmetic happens in
``_ls3d.so``.  Every wrapper raises when the library is missing or returns a non-zero status.
"""
import ctypes

import numpy as np

import torch

from . import capi
from .capi import check, host_f32, host_i32, ptr, stream_ptr


def _i32(dev, *shape):
    return torch.empty(*shape, dtype=torch.int32, device=dev)


def _f32(dev, *shape):
    return torch.empty(*shape, dtype=torch.float32, device=dev)


# ------------------------------------------------------------------------------------------ voxelize
def voxelize(points, frame_offsets, voxel_size, pc_range, max_points=5, max_voxels=300000, want_point_map=False):
    """Hard-voxelize a batch of frames on the GPU (reference semantics: points_to_voxel,
    det3d/ops/point_cloud/point_cloud_ops.py:112-184, per frame, then collate_kitti's batch column).

    points [N, F] fp32 cuda (frames concatenated), frame_offsets: python ints [B+1].
    Returns dict(voxels [M,P,F], coordinates [M,4] int32 (b,z,y,x), num_points [M] int32,
                 num_voxels [B] int64 (host-synchronised), point_voxel [N] int32 or None).
    """
    assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous()
    n, f = points.shape
    B = len(frame_offsets) - 1
    dev = points.device
    nbytes = ctypes.c_int64()
    check(capi.lib().ls3d_voxelize_workspace_bytes(n, max_points, B, ctypes.byref(nbytes)), "voxelize_workspace")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    cap = max(n, 1)
    voxels = _f32(dev, cap, max_points, f)
    coords = _i32(dev, cap, 4)
    nump = _i32(dev, cap)
    counts = _i32(dev, B + 1)      # [B] per frame + total
    pmap = _i32(dev, n) if want_point_map else None
    check(capi.lib().ls3d_voxelize(ptr(points), n, f, host_i32(frame_offsets), B, host_f32(voxel_size),
                                   host_f32(pc_range), max_points, max_voxels, ptr(ws), nbytes.value, ptr(voxels),
                                   ptr(coords), ptr(nump), ptr(counts), counts.data_ptr() + 4 * B, ptr(pmap),
                                   stream_ptr()), "ls3d_voxelize")
    host = counts.cpu()            # the one host sync: exact output shapes, like the reference's return values
    m = int(host[B])
    return dict(voxels=voxels[:m], coordinates=coords[:m], num_points=nump[:m],
                num_voxels=host[:B].to(torch.int64), point_voxel=pmap)


# ------------------------------------------------------------------------------------------ readers
def vfe_descriptor(voxels, num_points, mode, ld_out=None, round_out=False):
    m, P, F = voxels.shape
    need = F if mode == 0 else (F + 8 if mode == 1 else 2 * F + 8)
    ld = ld_out or need
    rows = m * P if mode == 2 else m
    out = _f32(voxels.device, rows, ld)
    check(capi.lib().ls3d_vfe_descriptor(ptr(voxels.contiguous()), ptr(num_points.to(torch.int32).contiguous()), m, P, F,
                                         mode, ptr(out), ld, int(round_out), stream_ptr()), "ls3d_vfe_descriptor")
    return out


def vfe_token_attn(qkv, m, P, n_head, d_head, round_out=False):
    out = _f32(qkv.device, m * P, n_head * d_head)
    check(capi.lib().ls3d_vfe_token_attn(ptr(qkv), qkv.stride(0), m, P, n_head, d_head, ptr(out), out.stride(0),
                                         int(round_out), stream_ptr()), "ls3d_vfe_token_attn")
    return out


def vfe_token_max(x, m, P, round_out=False):
    E = x.shape[1]
    out = _f32(x.device, m, E)
    check(capi.lib().ls3d_vfe_token_max(ptr(x), x.stride(0), m, P, E, ptr(out), out.stride(0), int(round_out),
                                        stream_ptr()), "ls3d_vfe_token_max")
    return out


# ------------------------------------------------------------------------------------------ rulebooks
class Grid:
    """Occupancy bitmap of one sparse level: words {bits, rank prefix}, optional rank->row permutation."""

    def __init__(self, B, shape, dev):
        self.B, (self.D, self.H, self.W) = B, shape
        wb, sb = ctypes.c_int64(), ctypes.c_int64()
        check(capi.lib().ls3d_grid_bytes(B, self.D, self.H, self.W, ctypes.byref(wb), ctypes.byref(sb)), "grid_bytes")
        self.words = torch.empty(wb.value, dtype=torch.uint8, device=dev)
        self.scratch = torch.empty(max(sb.value, 4), dtype=torch.uint8, device=dev)
        self.perm = None
        self.total = _i32(dev, 1)

    @property
    def shape(self):
        return (self.D, self.H, self.W)


def grid_from_coords(coords, B, shape, need_perm):
    g = Grid(B, shape, coords.device)
    m = coords.shape[0]
    if need_perm:
        g.perm = _i32(coords.device, max(m, 1))
    check(capi.lib().ls3d_grid_build(ptr(coords), m, B, g.D, g.H, g.W, ptr(g.words), ptr(g.perm), ptr(g.scratch),
                                     ptr(g.total), stream_ptr()), "ls3d_grid_build")
    return g


def out_shape(shape, ksize, stride, pad):
    return tuple((shape[i] + 2 * pad[i] - (ksize[i] - 1) - 1) // stride[i] + 1 for i in range(3))


def grid_strided(in_coords, B, in_shape, ksize, stride, pad):
    """Output sites of SparseConv3d(ksize, stride, pad): returns (grid, out_coords [M',4]) in ascending linear order."""
    oshape = out_shape(in_shape, ksize, stride, pad)
    g = Grid(B, oshape, in_coords.device)
    check(capi.lib().ls3d_grid_build_strided(ptr(in_coords), in_coords.shape[0], B, host_i32(ksize), host_i32(stride),
                                             host_i32(pad), g.D, g.H, g.W, ptr(g.words), ptr(g.scratch), ptr(g.total),
                                             stream_ptr()), "ls3d_grid_build_strided")
    m_out = int(g.total.item())    # host sync: number of output sites sizes the next launches
    ocoords = _i32(in_coords.device, max(m_out, 1), 4)
    check(capi.lib().ls3d_grid_enumerate(ptr(g.words), B, g.D, g.H, g.W, ptr(ocoords), stream_ptr()),
          "ls3d_grid_enumerate")
    return g, ocoords[:m_out]


def rulebook_gather(in_grid, out_coords, ksize, stride, pad):
    """nbr[k][j] = input row feeding output row j at kernel offset k (output-stationary pair table)."""
    K = ksize[0] * ksize[1] * ksize[2]
    m = out_coords.shape[0]
    nbr = _i32(out_coords.device, K, max(m, 1))[:, :m].contiguous() if m == 0 else _i32(out_coords.device, K, m)
    check(capi.lib().ls3d_rulebook_gather(ptr(in_grid.words), ptr(in_grid.perm), in_grid.B, in_grid.D, in_grid.H,
                                          in_grid.W, ptr(out_coords), m, host_i32(ksize), host_i32(stride),
                                          host_i32(pad), ptr(nbr), stream_ptr()), "ls3d_rulebook_gather")
    return nbr


def rulebook_scatter(out_grid, in_coords, ksize, stride, pad):
    """Inverse-conv table: nbr[k][i] = coarse row o with o*stride - pad + k == fine row i."""
    K = ksize[0] * ksize[1] * ksize[2]
    m = in_coords.shape[0]
    nbr = _i32(in_coords.device, K, m)
    check(capi.lib().ls3d_rulebook_scatter(ptr(out_grid.words), out_grid.B, out_grid.D, out_grid.H, out_grid.W,
                                           ptr(in_coords), m, host_i32(ksize), host_i32(stride), host_i32(pad),
                                           ptr(nbr), stream_ptr()), "ls3d_rulebook_scatter")
    return nbr


class TilePlan:
    """Gather-once plan of a rulebook table (ls3d_tile_plan_build; csrc/gather_gemm_once.cu)."""

    def __init__(self, nbr):
        K, m = nbr.shape
        hb, lb, pb = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(capi.lib().ls3d_tile_plan_bytes(K, m, ctypes.byref(hb), ctypes.byref(lb), ctypes.byref(pb)), "ls3d_tile_plan_bytes")
        dev = nbr.device
        self.hdr = torch.empty(hb.value // 4, dtype=torch.int32, device=dev)
        self.local = torch.empty(lb.value // 2, dtype=torch.int16, device=dev)
        self.pool = torch.empty(pb.value // 4, dtype=torch.int32, device=dev)
        self.counter = torch.empty(1, dtype=torch.int32, device=dev)
        self.koff, self.m_out = K, m
        check(capi.lib().ls3d_tile_plan_build(ptr(nbr), K, m, ptr(self.hdr), ptr(self.local), ptr(self.pool), ptr(self.counter),
                                              stream_ptr()), "ls3d_tile_plan_build")


def tile_plan(nbr):
    """The plan of ``nbr`` [K, m] int32, built once and cached on the tensor (a rulebook is reused by 6-9 convolutions)."""
    plan = getattr(nbr, "_ls3d_plan", None)
    if plan is None:
        assert nbr.dtype == torch.int32 and nbr.is_contiguous()
        plan = TilePlan(nbr)
        nbr._ls3d_plan = plan
    return plan


# ------------------------------------------------------------------------------------------ devoxelize
def three_nn_grid(points, grid, voxel_size, range_min, point_off, voxel_off, voxel_coords):
    """Exact 3-NN of points [N, >=4] (b,x,y,z) among the voxel centres of ``grid`` (level 1).
    Returns (dist2 [N,3] fp32, idx [N,3] int32 global voxel rows)."""
    n = points.shape[0]
    dev = points.device
    d2, idx = _f32(dev, n, 3), _i32(dev, n, 3)
    todo, cnt = _i32(dev, 2 * max(n, 1)), _i32(dev, 2)          # two work lists + counters (pass 1 -> 2 -> brute force)
    check(capi.lib().ls3d_three_nn_grid(ptr(points), points.stride(0), n, ptr(grid.words), ptr(grid.perm), grid.B,
                                        grid.D, grid.H, grid.W, host_f32(voxel_size), host_f32(range_min),
                                        ptr(point_off), ptr(voxel_off), ptr(voxel_coords), ptr(todo), ptr(cnt),
                                        ptr(d2), ptr(idx), stream_ptr()), "ls3d_three_nn_grid")
    return d2, idx


def frame_offsets(batch_col, batch_size):
    """Row offsets [B+1] (int32, device) of the frames of a batch-sorted tensor; ``batch_col`` = its batch column (a strided
    1-D view, fp32 or int32)."""
    assert batch_col.dim() == 1 and batch_col.dtype in (torch.float32, torch.int32)
    off = _i32(batch_col.device, batch_size + 1)
    n = batch_col.shape[0]
    check(capi.lib().ls3d_frame_offsets(ptr(batch_col), int(batch_col.dtype == torch.float32), batch_col.stride(0) if n else 1, n,
                                        batch_size, ptr(off), stream_ptr()), "ls3d_frame_offsets")
    return off


def three_nn(unknown, known):
    """Reference-signature 3-NN (three_nn_wrapper_fast): unknown [B, N, 3], known [B, M, 3] -> (dist2 [B, N, 3] squared,
    idx [B, N, 3] int32 per-batch rows)."""
    unknown, known = unknown.contiguous(), known.contiguous()
    B, N, _ = unknown.shape
    d2, idx = _f32(unknown.device, B, N, 3), _i32(unknown.device, B, N, 3)
    check(capi.lib().ls3d_three_nn(B, N, known.shape[1], ptr(unknown), ptr(known), ptr(d2), ptr(idx), stream_ptr()),
          "ls3d_three_nn")
    return d2, idx


def three_interpolate(feat, d2, idx, C=None, round_out=False):
    C = C or feat.shape[1]
    n = idx.shape[0]
    out = _f32(feat.device, n, C)
    check(capi.lib().ls3d_three_interpolate(ptr(feat), feat.stride(0), C, ptr(d2), ptr(idx), n, ptr(out),
                                            out.stride(0), int(round_out), stream_ptr()), "ls3d_three_interpolate")
    return out


# ------------------------------------------------------------------------------------------ sampling / SF-Phase
def sample_image_features(feat_nhwc, points_cuv, point_off, round_out=False):
    """feat_nhwc [B, ncam, H, W, C] contiguous fp32 or fp16; points_cuv [N,4]; returns [N, C] fp32 (zeros on invalid rows)."""
    B, ncam, H, W, C = feat_nhwc.shape
    n = points_cuv.shape[0]
    assert feat_nhwc.dtype in (torch.float32, torch.float16) and feat_nhwc.is_contiguous()
    out = _f32(feat_nhwc.device, n, C)
    check(capi.lib().ls3d_sample_image_features(ptr(feat_nhwc), int(feat_nhwc.dtype == torch.float16), B, ncam, H, W, C,
                                                ptr(points_cuv.contiguous()), n,
                                                ptr(point_off), ptr(out), out.stride(0), int(round_out),
                                                stream_ptr()), "ls3d_sample_image_features")
    return out


def project_points(points, cam_from_lidar, intrinsics, img_hw, net_hw, xyz_off=0, ref_to_global=None):
    """points [N, >=3] fp32 cuda; cam_from_lidar [ncam,4,4], intrinsics [ncam,3,3] (host, float64).  Returns
    points_cuv [N,4] = (valid, cam, v, u) with the reference's projection rules (loading.py:373-416).
    ``ref_to_global`` [4,4]: the first argument then holds the loader's ``cams_from_global`` and the two-stage chain
    lidar -> global -> camera is evaluated like the numpy original (loading.py:386-395)."""
    import numpy as np
    T = np.ascontiguousarray(np.asarray(cam_from_lidar, dtype=np.float64))
    K = np.ascontiguousarray(np.asarray(intrinsics, dtype=np.float64))
    n = points.shape[0]
    out = _f32(points.device, n, 4)
    if ref_to_global is None:
        check(capi.lib().ls3d_project_points(ptr(points), points.stride(0), xyz_off, n, T.ctypes.data, K.ctypes.data, T.shape[0],
                                             img_hw[0], img_hw[1], net_hw[0], net_hw[1], ptr(out), stream_ptr()),
              "ls3d_project_points")
    else:
        G = np.ascontiguousarray(np.asarray(ref_to_global, dtype=np.float64))
        check(capi.lib().ls3d_project_points_global(ptr(points), points.stride(0), xyz_off, n, G.ctypes.data, T.ctypes.data,
                                                    K.ctypes.data, T.shape[0], int(img_hw[0]), int(img_hw[1]), int(net_hw[0]),
                                                    int(net_hw[1]), ptr(out), stream_ptr()), "ls3d_project_points_global")
    return out


def resize_images_u8(images_u8, net_hw, mean=None, std=None, dtype=torch.float32):
    """uint8 [..., H, W, 3] raw camera images -> cv2.resize(img, (net_w, net_h)) (bit-exact with OpenCV's uint8 INTER_LINEAR),
    then, unless ``dtype`` is torch.uint8, (x / 255 - mean) / std as [..., 3, net_h, net_w] ``dtype`` maps in channels-last
    memory (a permuted view, like normalize_images_u8).  Reference: img_transforms.py:78-99 + :18-29."""
    assert images_u8.dtype == torch.uint8 and images_u8.shape[-1] == 3 and images_u8.is_contiguous()
    H, W = images_u8.shape[-3], images_u8.shape[-2]
    n_img = images_u8.numel() // (H * W * 3)
    oh, ow = int(net_hw[0]), int(net_hw[1])
    kind = {torch.float32: 0, torch.float16: 1, torch.uint8: 2}[dtype]
    out = torch.empty(images_u8.shape[:-3] + (oh, ow, 3), dtype=dtype, device=images_u8.device)
    m = sd = None
    if kind != 2:
        m = (ctypes.c_float * 3)(*[float(v) for v in np.asarray(mean, np.float32).reshape(-1)[:3]])
        sd = (ctypes.c_float * 3)(*[float(v) for v in np.asarray(std, np.float32).reshape(-1)[:3]])
    check(capi.lib().ls3d_resize_images_u8(ptr(images_u8), n_img, H, W, oh, ow, m, sd, ptr(out), kind, stream_ptr()),
          "ls3d_resize_images_u8")
    if kind == 2:
        return out
    nd = out.dim()
    return out.permute(*range(nd - 3), nd - 1, nd - 3, nd - 2)


def token_attention(q, k, v, frame_off, scale, round_out=False):
    """q [N, H*dh] fp32; k, v [F, H, L, dh]; frame_off int32 [F] -> softmax(q K^T * scale) V per point and head, [N, H*dh]."""
    F_, H, L, dh = k.shape
    n = q.shape[0]
    out = _f32(q.device, n, H * dh)
    check(capi.lib().ls3d_token_attention(ptr(q), q.stride(0), n, ptr(k), ptr(v), ptr(frame_off), F_, L, H, dh, float(scale),
                                          ptr(out), out.stride(0), int(round_out), stream_ptr()), "ls3d_token_attention")
    return out


def sffm_decoder(tgt, w_image, vec, k, v, frame_off, scale, n_layer, n_head, d_ffn, final_norm=True, ln_eps=1e-5):
    """The point stream of the SF-Phase TransformerDecoder (all layers + norm_tgt) in one launch (csrc/sffm_decoder.cu).
    tgt [N, E] fp32; w_image uint8 / vec fp32 from det3d.point_heads (layout: include/ls3d.h); k, v [n_layer, F, H, L, dh]."""
    n, E = tgt.shape
    nl, F_, H, L, dh = k.shape
    assert nl == n_layer and H == n_head and H * dh == E and tgt.dtype == torch.float32 and tgt.stride(1) == 1
    out = _f32(tgt.device, n, E)
    check(capi.lib().ls3d_sffm_decoder(ptr(tgt), tgt.stride(0), n, ptr(w_image), ptr(vec), ptr(k), ptr(v), ptr(frame_off), F_, L,
                                       n_layer, n_head, E, d_ffn, int(final_norm), float(scale), float(ln_eps), ptr(out),
                                       out.stride(0), stream_ptr()), "ls3d_sffm_decoder")
    return out


def normalize_images_u8(images_u8, mean, std, dtype=torch.float32):
    """uint8 [..., H, W, 3] (HWC images as the loader decodes them) -> (x / 255 - mean) / std as [..., 3, H, W] ``dtype`` maps in
    channels-last memory (the tensor is a permuted view of the pixel-major result, so no transpose is ever materialised).
    Reference: image_input_transform (det3d/datasets/pipelines/img_transforms.py:18-29)."""
    assert images_u8.dtype == torch.uint8 and images_u8.shape[-1] == 3 and images_u8.is_contiguous()
    assert dtype in (torch.float32, torch.float16)
    out = torch.empty(images_u8.shape, dtype=dtype, device=images_u8.device)
    m = (ctypes.c_float * 3)(*[float(v) for v in np.asarray(mean, np.float32).reshape(-1)[:3]])
    sd = (ctypes.c_float * 3)(*[float(v) for v in np.asarray(std, np.float32).reshape(-1)[:3]])
    check(capi.lib().ls3d_normalize_images_u8(ptr(images_u8), images_u8.numel() // 3, m, sd, ptr(out),
                                              int(dtype == torch.float16), stream_ptr()), "ls3d_normalize_images_u8")
    nd = out.dim()
    return out.permute(*range(nd - 3), nd - 1, nd - 3, nd - 2)


def upsample_sum(terms, relu=True, bias=None):
    """terms: list (<= 4) of [N, C, h_k, w_k] fp32 channels-last CUDA tensors, the FIRST at the output resolution or any of
    them; output resolution = the largest term.  Returns act(sum_k resize(term_k)) [N, C, H, W] channels-last."""
    import ctypes
    N, C = terms[0].shape[:2]
    H = max(t.shape[2] for t in terms)
    W = max(t.shape[3] for t in terms)
    dt = terms[0].dtype
    assert dt in (torch.float32, torch.float16) and C % (4 if dt == torch.float32 else 8) == 0
    ts = []
    for t in terms:
        assert t.dtype == dt and t.shape[0] == N and t.shape[1] == C
        ts.append(t.contiguous(memory_format=torch.channels_last))
    out = torch.empty((N, C, H, W), dtype=dt, device=terms[0].device, memory_format=torch.channels_last)
    k = len(ts)
    ptrs = (ctypes.c_void_p * k)(*[ptr(t) for t in ts])
    hs = (ctypes.c_int32 * k)(*[t.shape[2] for t in ts])
    ws = (ctypes.c_int32 * k)(*[t.shape[3] for t in ts])
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == C and bias.is_contiguous() and bias.device == out.device
    if dt == torch.float32:
        check(capi.lib().ls3d_upsample_sum(ptrs, hs, ws, k, N, H, W, C, int(relu), ptr(bias), ptr(out), stream_ptr()),
              "ls3d_upsample_sum")
    else:
        check(capi.lib().ls3d_upsample_sum_f16(ptrs, hs, ws, k, N, H, W, C, int(relu), ptr(bias), ptr(out), stream_ptr()),
              "ls3d_upsample_sum_f16")
    return out


def conv_f16_supported(cin, cout, ksize=3):
    """The fused fp16 convolution keeps all taps of the weights in shared memory: true when they (and two stages) fit."""
    nb = ctypes.c_int64()
    if cin % 8 or cout % 8 or cin > 256 or cout > 256:
        return False
    check(capi.lib().ls3d_conv_f16_smem_bytes(cin, cout, ksize, ctypes.byref(nb)), "ls3d_conv_f16_smem_bytes")
    return nb.value <= 227 * 1024


def conv3x3_f16_supported(cin, cout):
    return conv_f16_supported(cin, cout, 3)


def pack_conv_f16(w_oihw):
    """[Cout_p, Cin_p, k, k] (k = 1 or 3; BatchNorm folded, zero-padded channels, both multiples of 8) -> the packed fp16
    weight block of ls3d_conv_f16 (packed on the device in the kernel's K order)."""
    cout, cin, k = w_oihw.shape[:3]
    w = w_oihw.detach().float().contiguous()
    nb = ctypes.c_int64()
    check(capi.lib().ls3d_conv_f16_packed_bytes(cin, cout, k, ctypes.byref(nb)), "ls3d_conv_f16_packed_bytes")
    out = torch.empty(nb.value // 2, dtype=torch.float16, device=w.device)
    check(capi.lib().ls3d_conv_f16_pack(ptr(w), cin, cout, k, ptr(out), stream_ptr()), "ls3d_conv_f16_pack")
    return out


def conv_f16(x, w_packed, bias, res=None, relu=True, cout=None, ksize=3):
    """x [N, Cin, H, W] fp16 channels-last, res [N, Cout, H, W] fp16 channels-last or None, bias fp32 [Cout] or None
    -> act(conv_kxk(x) + bias + res), k = ksize (stride 1, same size)."""
    N, cin, H, W = x.shape
    assert x.dtype == torch.float16 and x.is_contiguous(memory_format=torch.channels_last)
    cout = cout if cout is not None else bias.shape[0]
    if res is not None:
        assert res.dtype == torch.float16 and res.shape == (N, cout, H, W) and res.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((N, cout, H, W), dtype=torch.float16, device=x.device, memory_format=torch.channels_last)
    check(capi.lib().ls3d_conv_f16(ptr(x), ptr(w_packed), ptr(bias), ptr(res), ptr(out), N, H, W, cin, cout, ksize, int(relu),
                                   stream_ptr()), "ls3d_conv_f16")
    return out


def conv_f16_dual_supported(cin, cout, ksize=3):
    """ls3d_conv_f16_dual (fp32 maps, fp16 operands) has a shared-memory configuration for this shape."""
    nb = ctypes.c_int64()
    if cin % 8 or cout % 8 or cin > 256 or cout > 256:
        return False
    check(capi.lib().ls3d_conv_f16_dual_smem_bytes(cin, cout, ksize, ctypes.byref(nb)), "ls3d_conv_f16_dual_smem_bytes")
    return nb.value <= 227 * 1024


def conv_f16_split_supported(cin, cout, ksize=3, dual=True):
    """Split (exact) weights [W_hi ; W_lo] fit the kernel for this shape (2 n_pad <= 256 columns, twice the weight bytes)."""
    if cin % 8 or cout % 8 or cin > 256 or cout > 128:
        return False
    ok = ctypes.c_int32()
    check(capi.lib().ls3d_conv_f16_split_supported(cin, cout, ksize, int(dual), ctypes.byref(ok)), "ls3d_conv_f16_split_supported")
    return bool(ok.value)


def pack_conv_f16_split(w_oihw):
    """Like pack_conv_f16, split weights: fp16(w) and fp16(w - fp16(w)) stacked along N (ls3d_conv_f16_pack_split)."""
    cout, cin, k = w_oihw.shape[:3]
    w = w_oihw.detach().float().contiguous()
    nb = ctypes.c_int64()
    check(capi.lib().ls3d_conv_f16_packed_bytes(cin, cout, k, ctypes.byref(nb)), "ls3d_conv_f16_packed_bytes")
    out = torch.empty(nb.value, dtype=torch.float16, device=w.device)            # twice the unsplit block
    check(capi.lib().ls3d_conv_f16_pack_split(ptr(w), cin, cout, k, ptr(out), stream_ptr()), "ls3d_conv_f16_pack_split")
    return out


def conv_f16_dual(x16, w_packed, bias, res32=None, relu=True, cout=None, ksize=3, split=False, want32=True):
    """fp32 residual stream: x16 [N, Cin, H, W] fp16 channels-last (operand copy of the input map), res32 fp32 channels-last or
    None -> (out32, out16) = act(conv(x16) + bias + res32) as an fp32 map and its fp16 operand copy.  ``want32=False``: only the
    operand copy is produced (out32 is None).  ``split``: w_packed holds split weights (pack_conv_f16_split)."""
    N, cin, H, W = x16.shape
    assert x16.dtype == torch.float16 and x16.is_contiguous(memory_format=torch.channels_last)
    cout = cout if cout is not None else bias.shape[0]
    if res32 is not None:
        assert want32 and res32.dtype == torch.float32 and res32.shape == (N, cout, H, W) and \
            res32.is_contiguous(memory_format=torch.channels_last)
    out32 = torch.empty((N, cout, H, W), dtype=torch.float32, device=x16.device, memory_format=torch.channels_last) if want32 \
        else None
    out16 = torch.empty((N, cout, H, W), dtype=torch.float16, device=x16.device, memory_format=torch.channels_last)
    check(capi.lib().ls3d_conv_f16_dual(ptr(x16), ptr(w_packed), ptr(bias), ptr(res32), ptr(out32), ptr(out16), N, H, W, cin,
                                        cout, ksize, int(relu), int(split), stream_ptr()), "ls3d_conv_f16_dual")
    return out32, out16


def conv_ex_supported(cin, cout, ksize, stride, dual=True, split=True):
    """ls3d_conv_f16_ex has a shared / tensor-memory configuration for a (cin -> cout) launch of this kind."""
    if cin % 8 or cout % 8 or cin <= 0 or cout <= 0:
        return False
    ok = ctypes.c_int32()
    check(capi.lib().ls3d_conv_f16_ex_supported(cin, cout, ksize, stride, int(dual), int(split), ctypes.byref(ok)),
          "ls3d_conv_f16_ex_supported")
    return bool(ok.value)


def pack_conv_ex(w_oihw, stride=1, split=True):
    """[Cout, Cin, k, k] fp32 (both multiples of 8) -> packed block of ls3d_conv_f16_ex for this (stride, split)."""
    cout, cin, k = w_oihw.shape[:3]
    w = w_oihw.detach().float().contiguous()
    nb = ctypes.c_int64()
    check(capi.lib().ls3d_conv_f16_packed_bytes(cin, cout, k, ctypes.byref(nb)), "ls3d_conv_f16_packed_bytes")
    out = torch.empty(nb.value // 2 * (2 if split else 1), dtype=torch.float16, device=w.device)
    check(capi.lib().ls3d_conv_f16_pack_ex(ptr(w), cin, cout, k, stride, int(split), ptr(out), stream_ptr()),
          "ls3d_conv_f16_pack_ex")
    return out


CONV_PROFILE = None      # development: list -> (CUDA events, shape) per ls3d_conv_f16_ex launch (scripts/prof_camera.py)


def conv_ex(x16, w_packed, bias, *, cin, in_off, cout, out_off, out32, out16, res32=None, res16=None, ksize=3, stride=1,
            relu=False, split=True):
    """One ls3d_conv_f16_ex launch: input channels [in_off, in_off + cin) of the fp16 channels-last map ``x16`` -> output channels
    [out_off, out_off + cout) of ``out32`` (fp32, may be None: operand-only) / ``out16`` (fp16), both preallocated channels-last;
    ``res32`` (may alias out32) is added before the ReLU."""
    N, ct, H, W = x16.shape
    a = capi.ConvArgs()
    a.in16, a.w_packed, a.bias = ptr(x16), ptr(w_packed), ptr(bias)
    a.res32, a.out32, a.out16, a.res16 = ptr(res32), ptr(out32), ptr(out16), ptr(res16)
    a.in_c_total, a.in_c_off, a.cin = ct, in_off, cin
    a.out_c_total, a.out_c_off, a.cout = out16.shape[1], out_off, cout
    a.n_img, a.H_in, a.W_in, a.ksize, a.stride, a.relu, a.w_split = N, H, W, ksize, stride, int(relu), int(split)
    if CONV_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(capi.lib().ls3d_conv_f16_ex(ctypes.byref(a), stream_ptr()), "ls3d_conv_f16_ex")
        e1.record()
        CONV_PROFILE.append(dict(e0=e0, e1=e1, n=N, h=H, w=W, cin=cin, cout=cout, k=ksize, stride=stride,
                                 res=res32 is not None or res16 is not None,
                                 out32=out32 is not None, in_total=ct, out_total=out16.shape[1]))
        return
    check(capi.lib().ls3d_conv_f16_ex(ctypes.byref(a), stream_ptr()), "ls3d_conv_f16_ex")


def conv_multi(x16, passes, n_pass, bias, *, cin, cout, out32, out16, res32=None, res16=None, ksize=3, stride=1, split=True):
    """One ls3d_conv_f16_multi launch: ``passes`` = ctypes array of capi.ConvPass (weight block, channel offsets, flags)."""
    N, ct, H, W = x16.shape
    a = capi.ConvArgs()
    a.in16, a.bias = ptr(x16), ptr(bias)
    a.res32, a.out32, a.out16, a.res16 = ptr(res32), ptr(out32), ptr(out16), ptr(res16)
    a.in_c_total, a.cin, a.out_c_total, a.cout = ct, cin, out16.shape[1], cout
    a.n_img, a.H_in, a.W_in, a.ksize, a.stride, a.w_split = N, H, W, ksize, stride, int(split)
    if CONV_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(capi.lib().ls3d_conv_f16_multi(ctypes.byref(a), passes, n_pass, stream_ptr()), "ls3d_conv_f16_multi")
        e1.record()
        CONV_PROFILE.append(dict(e0=e0, e1=e1, n=N, h=H, w=W, cin=ct, cout=out16.shape[1], k=ksize, stride=stride,
                                 res=res32 is not None or res16 is not None, out32=out32 is not None, in_total=ct,
                                 out_total=out16.shape[1], passes=n_pass))
        return
    check(capi.lib().ls3d_conv_f16_multi(ctypes.byref(a), passes, n_pass, stream_ptr()), "ls3d_conv_f16_multi")


def conv_kb_supported(cin, cout, ksize, stride, dual, split, out_pixels):
    """ls3d_conv_f16_kb (streamed weights) has a configuration for a 3x3 (cin -> cout slice) launch with ``out_pixels`` outputs."""
    if cin % 8 or cout % 8 or cin <= 0 or cout <= 0:
        return False
    ok = ctypes.c_int32()
    check(capi.lib().ls3d_conv_f16_kb_supported(cin, cout, ksize, stride, int(dual), int(split), int(out_pixels), ctypes.byref(ok)),
          "ls3d_conv_f16_kb_supported")
    return bool(ok.value)


def conv_kb(x16, passes, n_slices, bias, *, cout, out32, out16, res32=None, res16=None, split=True, stride=1, ksize=3):
    """One ls3d_conv_f16_kb launch: ``passes`` = ctypes array of capi.ConvPass, one per output channel slice of ``cout``."""
    N, ct, H, W = x16.shape
    a = capi.ConvArgs()
    a.in16, a.bias = ptr(x16), ptr(bias)
    a.res32, a.out32, a.out16, a.res16 = ptr(res32), ptr(out32), ptr(out16), ptr(res16)
    a.in_c_total, a.cin, a.out_c_total, a.cout = ct, ct, out16.shape[1], cout
    a.n_img, a.H_in, a.W_in, a.ksize, a.stride, a.w_split = N, H, W, ksize, stride, int(split)
    e0 = e1 = None
    if CONV_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(capi.lib().ls3d_conv_f16_kb(ctypes.byref(a), passes, n_slices, stream_ptr()), "ls3d_conv_f16_kb")
    if CONV_PROFILE is not None:
        e1.record()
        CONV_PROFILE.append(dict(e0=e0, e1=e1, n=N, h=H, w=W, cin=ct, cout=out16.shape[1], k=ksize, stride=stride,
                                 res=res32 is not None or res16 is not None, out32=out32 is not None, in_total=ct,
                                 out_total=out16.shape[1], passes=n_slices, kb=True))


def pad3_f16(x):
    """fp32 channels-last image batch [N, 3, H, W] -> fp16 [N, 8, H, W] channels-last, channels 3..7 zero: the operand copy of
    the network input for the own stem convolution (16-byte pixel rows for the tensor-map copies)."""
    N, C, H, W = x.shape
    assert C == 3 and x.dtype == torch.float32 and x.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((N, 8, H, W), dtype=torch.float16, device=x.device, memory_format=torch.channels_last)
    check(capi.lib().ls3d_pad3_f16(ptr(x), N * H * W, ptr(out), stream_ptr()), "ls3d_pad3_f16")
    return out


def cast_f32(x):
    """fp32 copy of an fp16 tensor, same memory layout (ls3d_cast_f32)."""
    assert x.dtype == torch.float16 and x.numel() % 4 == 0
    assert x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty_like(x, dtype=torch.float32)
    check(capi.lib().ls3d_cast_f32(ptr(x), ptr(out), x.numel(), stream_ptr()), "ls3d_cast_f32")
    return out


def cast_f16(x):
    """fp16 (RNE) copy of an fp32 tensor, same memory layout (ls3d_cast_f16)."""
    assert x.dtype == torch.float32 and x.numel() % 4 == 0
    assert x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty_like(x, dtype=torch.float16)
    check(capi.lib().ls3d_cast_f16(ptr(x), ptr(out), x.numel(), stream_ptr()), "ls3d_cast_f16")
    return out


def upsample_sum_dual(terms, relu=True, bias=None):
    """upsample_sum on fp32 terms that also writes the fp16 operand copy of the result: returns (out32, out16)."""
    N, C = terms[0].shape[:2]
    H = max(t.shape[2] for t in terms)
    W = max(t.shape[3] for t in terms)
    assert C % 4 == 0
    ts = []
    for t in terms:
        assert t.dtype == torch.float32 and t.shape[0] == N and t.shape[1] == C
        ts.append(t.contiguous(memory_format=torch.channels_last))
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=terms[0].device, memory_format=torch.channels_last)
    out16 = torch.empty((N, C, H, W), dtype=torch.float16, device=terms[0].device, memory_format=torch.channels_last)
    k = len(ts)
    ptrs = (ctypes.c_void_p * k)(*[ptr(t) for t in ts])
    hs = (ctypes.c_int32 * k)(*[t.shape[2] for t in ts])
    ws = (ctypes.c_int32 * k)(*[t.shape[3] for t in ts])
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == C and bias.is_contiguous() and bias.device == out.device
    check(capi.lib().ls3d_upsample_sum_dual(ptrs, hs, ws, k, N, H, W, C, int(relu), ptr(bias), ptr(out), ptr(out16),
                                            stream_ptr()), "ls3d_upsample_sum_dual")
    return out, out16


def pack_conv3x3_f16(w_oihw):
    """[Cout_p, Cin_p, 3, 3] (BatchNorm folded, zero-padded channels, both multiples of 8) -> the packed fp16 weight block of
    ls3d_conv3x3_f16 (packed on the device in the kernel's K order)."""
    cout, cin = w_oihw.shape[:2]
    w = w_oihw.detach().float().contiguous()
    nb = ctypes.c_int64()
    check(capi.lib().ls3d_conv3x3_f16_packed_bytes(cin, cout, ctypes.byref(nb)), "ls3d_conv3x3_f16_packed_bytes")
    out = torch.empty(nb.value // 2, dtype=torch.float16, device=w.device)
    check(capi.lib().ls3d_conv3x3_f16_pack(ptr(w), cin, cout, ptr(out), stream_ptr()), "ls3d_conv3x3_f16_pack")
    return out


def conv3x3_f16(x, w_packed, bias, res=None, relu=True, cout=None):
    """x [N, Cin, H, W] fp16 channels-last, res [N, Cout, H, W] fp16 channels-last or None -> relu(conv3x3(x) + bias + res)."""
    N, cin, H, W = x.shape
    assert x.dtype == torch.float16 and x.is_contiguous(memory_format=torch.channels_last)
    cout = cout if cout is not None else bias.shape[0]
    if res is not None:
        assert res.dtype == torch.float16 and res.shape == (N, cout, H, W) and res.is_contiguous(memory_format=torch.channels_last)
    out = torch.empty((N, cout, H, W), dtype=torch.float16, device=x.device, memory_format=torch.channels_last)
    check(capi.lib().ls3d_conv3x3_f16(ptr(x), ptr(w_packed), ptr(bias), ptr(res), ptr(out), N, H, W, cin, cout, int(relu),
                                      stream_ptr()), "ls3d_conv3x3_f16")
    return out


def class_embed(logits, feats, seg_off, n_frames, max_rows, ncls=None, C=None):
    """softmax over rows per (frame, class) then probs^T @ feats -> [B, ncls, C]."""
    ncls = ncls or logits.shape[1]
    C = C or feats.shape[1]
    nb = ctypes.c_int64()
    check(capi.lib().ls3d_class_embed_workspace_bytes(n_frames, max_rows, ncls, C, ctypes.byref(nb)), "class_embed_ws")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=logits.device)
    emb = _f32(logits.device, n_frames, ncls, C)
    assert logits.dtype == feats.dtype and logits.dtype in (torch.float32, torch.float16)
    check(capi.lib().ls3d_class_embed(ptr(logits), logits.stride(0), ncls, ptr(feats), feats.stride(0), C,
                                      int(logits.dtype == torch.float16), ptr(seg_off),
                                      n_frames, max_rows, ptr(ws), ptr(emb), stream_ptr()), "ls3d_class_embed")
    return emb


def class_tokens(emb1, emb2, params, n_layer, n_head, d_model, want_memory=False):
    """Memory path of the SF-Phase decoder for all layers: returns K, V [n_layer, B, H, L, dh] (+ memory)."""
    B, ncls, C1 = emb1.shape
    C2 = emb2.shape[2]
    L, dh = 2 * ncls, d_model // n_head
    K = _f32(emb1.device, n_layer, B, n_head, L, dh)
    V = _f32(emb1.device, n_layer, B, n_head, L, dh)
    mem = _f32(emb1.device, n_layer, B, L, d_model) if want_memory else None
    check(capi.lib().ls3d_class_tokens(ptr(emb1), C1, ptr(emb2), C2, ncls, B, ptr(params), n_layer, n_head, d_model,
                                       ptr(K), ptr(V), ptr(mem), stream_ptr()), "ls3d_class_tokens")
    return (K, V, mem) if want_memory else (K, V)
