"""Functional CPU restatement (PyTorch fp32) of the reference's SegNet / SegMSeg3DNet forward.

TEST INFRASTRUCTURE ONLY: the checker for tests/, __graft_entry__.smoke() and bench.py's CPU baseline / reference arm;
nothing under lidarseg3d_b200/ imports it and the product path has no CPU fallback.

Each function takes the flat ``state_dict`` of the reference model (parameter names as in the reference,
SURVEY.md Appendix C) plus the tensors of the ``example`` dict and returns what the reference module
returns.  Eval mode only (BatchNorm running statistics, Dropout = identity).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import sparse as osp


# ------------------------------------------------------------------------------------------ helpers
def _bn(sd, p, x, eps):
    """BatchNorm in eval mode on [..., C] rows (nn.BatchNorm1d / BatchNorm2d with running stats)."""
    w, b, m, v = sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"]
    return (x - m) / torch.sqrt(v + eps) * w + b


def _bn2d(sd, p, x, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, eps)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


# ------------------------------------------------------------------------------------------ readers
def vfe_descriptor(features, num_voxels):
    """Shared descriptor of ImprovedMeanVFE / TransVFE (det3d/models/readers/voxel_encoder.py:80-121,
    214-243): [mean xyz, max xyz, min xyz, mean(other feats), density, std]."""
    P = features.shape[1]
    nv = num_voxels.type_as(features)
    points_mean = features.sum(dim=1) / nv.view(-1, 1)
    point_mask = (features.sum(dim=-1) != 0).float()                       # voxel_encoder.py:87
    xyz = features[:, :, :3]
    pmax = torch.stack([(xyz[:, :, a] - (1 - point_mask) * 1e5).max(dim=1)[0] for a in range(3)], -1)
    pmin = torch.stack([(xyz[:, :, a] + (1 - point_mask) * 1e5).min(dim=1)[0] for a in range(3)], -1)
    density = point_mask.sum(-1) / P
    norm = torch.norm((xyz - points_mean[:, None, 0:3]) * point_mask[:, :, None], p=2, dim=-1)
    std = norm.sum(1) / nv
    return torch.cat([points_mean[:, 0:3], pmax, pmin, points_mean[:, 3:], density[:, None], std[:, None]], -1)


def mean_vfe(features, num_voxels):
    """MeanVoxelFeatureExtractor.forward (voxel_encoder.py:51-58)."""
    return features.sum(dim=1) / num_voxels.type_as(features).view(-1, 1)


def improved_mean_vfe(features, num_voxels):
    """ImprovedMeanVoxelFeatureExtractor.forward (voxel_encoder.py:74-124)."""
    return vfe_descriptor(features, num_voxels).contiguous()


def _mha_self(sd, p, x, nhead):
    """nn.MultiheadAttention(q=k=v=x) with x [L, B, E], no masks, dropout 0."""
    L, B, E = x.shape
    qkv = F.linear(x, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    dh = E // nhead
    q = q.reshape(L, B * nhead, dh).transpose(0, 1) * dh ** -0.5
    k = k.reshape(L, B * nhead, dh).transpose(0, 1)
    v = v.reshape(L, B * nhead, dh).transpose(0, 1)
    a = torch.softmax(q @ k.transpose(1, 2), dim=-1)
    o = (a @ v).transpose(0, 1).reshape(L, B, E)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def trans_vfe(sd, p, features, num_voxels, num_head, num_layers):
    """TransformerVoxelFeatureExtractor.forward (voxel_encoder.py:202-270) with the custom pre-norm layer
    (voxel_encoder.py:149-163): src = LN1(src); src += MHA(src); src = LN2(src); src += FFN(src).
    Zero-padded slots are attended (no mask)."""
    P = features.shape[1]
    desc = vfe_descriptor(features, num_voxels)
    x = torch.cat([features, desc[:, None, :].expand(-1, P, -1)], dim=-1)          # [M, P, F+D]
    x = F.linear(x, sd[p + "feature_conv.0.weight"].squeeze(-1), sd[p + "feature_conv.0.bias"])  # Conv1d k=1
    x = x.permute(1, 0, 2)                                                         # [L=P, B=M, E]
    for i in range(num_layers):
        q = f"{p}chunck.layers.{i}"
        x = _ln(sd, q + ".norm1", x)
        x = x + _mha_self(sd, q + ".self_attn", x, num_head)
        x = _ln(sd, q + ".norm2", x)
        x = x + _lin(sd, q + ".linear2", torch.relu(_lin(sd, q + ".linear1", x)))
    x = x.max(dim=0)[0]                                                            # max over the P slots
    if (p + "compress_layer.0.weight") in sd:
        x = torch.relu(_lin(sd, p + "compress_layer.0", x))
    return x.contiguous()


# ------------------------------------------------------------------------------------------ UNetSCN3D
class _Sp:
    """Minimal SparseConvTensor: features + indices + spatial shape."""

    def __init__(self, features, indices, shape):
        self.features, self.indices, self.shape = features, indices, tuple(shape)


def sp_middle_resnet_fhd(sd, p, voxel_features, coors, batch_size, input_shape_xyz):
    """SpMiddleResNetFHD.forward (det3d/models/backbones/scn.py:146-177): conv_input (SubM, key res0) -> 2 residual blocks ->
    three times [SparseConv3d k3 s2 (padding 1 / 1 / [0,1,1]) + 2 residual blocks] -> extra_conv k(3,1,1) s(2,1,1) -> dense
    [N, C*D, H, W].  SparseBasicBlock (scn.py:37-81): its two SubM convs carry a BIAS (bias = norm_cfg is not None, :56), the
    sequence entries are (conv, BN, ReLU, block, block) so the blocks sit at indices 3 and 4.  Returns (dense, levels)."""
    EPS = 1e-3                                                                      # scn.py:50,96
    shape1 = tuple(int(v) for v in (np.array(input_shape_xyz[::-1]) + [1, 0, 0]))   # scn.py:149
    idx = coors.numpy().astype(np.int32)

    def block(x, q, nbr):
        out = osp.sparse_conv(x, sd[q + ".conv1.weight"], nbr) + sd[q + ".conv1.bias"]
        out = torch.relu(_bn(sd, q + ".bn1", out, EPS))
        out = osp.sparse_conv(out, sd[q + ".conv2.weight"], nbr) + sd[q + ".conv2.bias"]
        return torch.relu(_bn(sd, q + ".bn2", out, EPS) + x)

    n1 = osp.subm_rulebook(idx, shape1, 3)
    x = torch.relu(_bn(sd, p + "conv_input.1", osp.sparse_conv(voxel_features, sd[p + "conv_input.0.weight"], n1), EPS))
    x = block(block(x, p + "conv1.0", n1), p + "conv1.1", n1)
    levels = {"conv1": _Sp(x, idx, shape1)}
    cur = levels["conv1"]
    for lv, pad in ((2, (1, 1, 1)), (3, (1, 1, 1)), (4, (0, 1, 1))):                # scn.py:110-140
        oidx, oshape, nb = osp.strided_rulebook(cur.indices, cur.shape, 3, 2, pad)
        y = torch.relu(_bn(sd, f"{p}conv{lv}.1", osp.sparse_conv(cur.features, sd[f"{p}conv{lv}.0.weight"], nb), EPS))
        ns = osp.subm_rulebook(oidx, oshape, 3)
        y = block(block(y, f"{p}conv{lv}.3", ns), f"{p}conv{lv}.4", ns)
        cur = _Sp(y, oidx, oshape)
        levels[f"conv{lv}"] = cur
    oidx, oshape, nb = osp.strided_rulebook(cur.indices, cur.shape, (3, 1, 1), (2, 1, 1), (0, 0, 0))
    y = torch.relu(_bn(sd, p + "extra_conv.1", osp.sparse_conv(cur.features, sd[p + "extra_conv.0.weight"], nb), EPS))
    D, H, W = oshape
    dense = torch.zeros(batch_size, D, H, W, y.shape[1])
    oi = torch.as_tensor(oidx).long()
    dense[oi[:, 0], oi[:, 1], oi[:, 2], oi[:, 3]] = y
    dense = dense.permute(0, 4, 1, 2, 3).contiguous()                               # SparseConvTensor.dense(): [N, C, D, H, W]
    return dense.view(batch_size, y.shape[1] * D, H, W), levels


def unet_scn3d(sd, p, voxel_features, voxel_coords, input_shape_xyz, voxel_size, pc_range,
               with_conv_out=True, last_pad=0, return_levels=False, backend=None):
    """UNetSCN3D.forward (det3d/models/backbones/scn_unet.py:189-249) with spconv semantics from oracle.sparse.
    ``p`` is the key prefix (e.g. 'backbone.').  Returns (conv_point_features [M,C], conv_point_coords [M,4]).
    ``backend``: None = oracle.sparse (numpy rulebooks, the parity ground truth); oracle.torch_backend = the same algorithm in
    pure torch on the device of the inputs (the spconv-style GPU baseline)."""
    EPS = 1e-3                                                                      # scn_unet.py:86
    shape1 = tuple(int(v) for v in (np.array(input_shape_xyz[::-1]) + [1, 0, 0]))   # scn_unet.py:203
    osp = backend if backend is not None else globals()["osp"]
    idx1 = voxel_coords.numpy().astype(np.int32) if backend is None else backend.as_indices(voxel_coords)
    rb = {}                                                                         # indice_key -> table

    def subm_key(key, indices, shape):
        if key not in rb:
            rb[key] = osp.subm_rulebook(indices, shape, 3)
        return rb[key]

    def conv_bn_relu(x, q, nbr):           # SparseSequential(conv, BN, ReLU)
        y = osp.sparse_conv(x, sd[q + ".0.weight"], nbr)
        return torch.relu(_bn(sd, q + ".1", y, EPS))

    def basic_block(x, q, nbr):            # SparseBasicBlock.forward (scn_unet.py:51-69)
        out = torch.relu(_bn(sd, q + ".bn1", osp.sparse_conv(x, sd[q + ".conv1.weight"], nbr), EPS))
        out = _bn(sd, q + ".bn2", osp.sparse_conv(out, sd[q + ".conv2.weight"], nbr), EPS)
        return torch.relu(out + x)

    n1 = subm_key("subm1", idx1, shape1)
    x = conv_bn_relu(voxel_features, p + "conv_input", n1)
    x1 = basic_block(basic_block(x, p + "conv1.0", n1), p + "conv1.1", n1)

    levels = {1: _Sp(x1, idx1, shape1)}
    down = {}
    pads = {2: (1, 1, 1), 3: (1, 1, 1), 4: (0, 1, 1)}                               # scn_unet.py:106,113,120
    cur = levels[1]
    for lv in (2, 3, 4):
        oidx, oshape, nb_down = osp.strided_rulebook(cur.indices, cur.shape, 3, 2, pads[lv])
        down[lv] = (nb_down, cur.indices.shape[0])
        y = conv_bn_relu(cur.features, f"{p}conv{lv}.0", nb_down)
        ns = subm_key(f"subm{lv}", oidx, oshape)
        y = basic_block(basic_block(y, f"{p}conv{lv}.1", ns), f"{p}conv{lv}.2", ns)
        cur = _Sp(y, oidx, oshape)
        levels[lv] = cur

    extra = {}
    if with_conv_out and (p + "conv_out.0.weight") in sd:                          # scn_unet.py:125-134,218-222
        lp = osp._triple(last_pad)
        oidx, oshape, nbo = osp.strided_rulebook(levels[4].indices, levels[4].shape, (3, 1, 1), (2, 1, 1), lp)
        extra["encoded"] = _Sp(conv_bn_relu(levels[4].features, p + "conv_out", nbo), oidx, oshape)

    def ur_block(lat, bottom, lv, inv_key):  # UR_block_forward (scn_unet.py:163-171)
        ns = rb[f"subm{lv}"]
        t = basic_block(lat.features, f"{p}conv_up_t{lv}", ns)
        cat = torch.cat([bottom.features, t], dim=1)
        xm = conv_bn_relu(cat, f"{p}conv_up_m{lv}", ns)
        red = cat.view(cat.shape[0], xm.shape[1], -1).sum(dim=2)                    # channel_reduction :173-187
        y = xm + red
        if inv_key is not None:                                                    # SparseInverseConv3d
            nb_down, n_fine = down[inv_key]
            y = conv_bn_relu(y, f"{p}inv_conv{inv_key}", osp.invert_rulebook(nb_down, n_fine))
        else:                                                                      # conv5 = SubM on subm1
            y = conv_bn_relu(y, f"{p}conv5.0", ns)
        return y

    up4 = _Sp(ur_block(levels[4], levels[4], 4, 4), levels[3].indices, levels[3].shape)
    up3 = _Sp(ur_block(levels[3], up4, 3, 3), levels[2].indices, levels[2].shape)
    up2 = _Sp(ur_block(levels[2], up3, 2, 2), levels[1].indices, levels[1].shape)
    up1 = ur_block(levels[1], up2, 1, None)

    # get_voxel_centers (det3d/core/utils/common_utils.py:74-90)
    dev = voxel_features.device
    idx1_t = torch.as_tensor(idx1).to(dev)
    centers = (idx1_t[:, [3, 2, 1]].float() + 0.5) * torch.tensor(voxel_size, device=dev).float() \
        + torch.tensor(pc_range[0:3], device=dev).float()
    coords = torch.cat([idx1_t[:, 0:1].float(), centers], dim=1)
    if return_levels:
        return up1, coords, dict(levels=levels, up4=up4, up3=up3, up2=up2, rulebooks=rb, down=down, **extra)
    return up1, coords


# ------------------------------------------------------------------------------------------ devoxelize
def three_nn(unknown, known):
    """three_nn_kernel_fast (det3d/ops/pointnet2_batch/src/interpolate_gpu.cu:16-59): 3 nearest ``known`` for
    every ``unknown`` by fp32 squared distance ((ux-x)^2 + (uy-y)^2) + (uz-z)^2 evaluated without FMA
    contraction, strict '<' cascade in ascending index order (ties -> lowest index).
    Returns (dist2 [N,3] fp32, idx [N,3] int32).  Slots never filled keep (1e40 -> inf, 0) like the kernel."""
    u = unknown.numpy().astype(np.float32)
    k = known.numpy().astype(np.float32)
    N, M = u.shape[0], k.shape[0]
    d2o = np.full((N, 3), np.float32(np.inf), dtype=np.float32)
    ido = np.zeros((N, 3), dtype=np.int32)
    if M == 0:
        return torch.from_numpy(d2o), torch.from_numpy(ido)
    CH = max(1, (1 << 24) // max(M, 1))
    kk = min(3, M)
    C = min(8, M)                                                                  # candidate pool per row
    for s in range(0, N, CH):
        uu = u[s:s + CH]
        dx = uu[:, None, 0] - k[None, :, 0]
        dy = uu[:, None, 1] - k[None, :, 1]
        dz = uu[:, None, 2] - k[None, :, 2]
        d = (dx * dx + dy * dy) + dz * dz                                           # fp32, left to right
        # lexicographic (d, index) order == the sequential strict-'<' scan.  Exact top-3 without a full sort: take the C
        # smallest (unordered), order them by (d, index); rows whose 3rd distance ties with more than the pool holds fall
        # back to the full stable sort.
        cand = np.argpartition(d, C - 1, axis=1)[:, :C] if C < M else np.tile(np.arange(M), (d.shape[0], 1))
        cd = np.take_along_axis(d, cand, 1)
        order = np.lexsort((cand, cd), axis=1)[:, :kk]
        part = np.take_along_axis(cand, order, 1)
        pd = np.take_along_axis(cd, order, 1)
        thr = pd[:, kk - 1]
        bad = (d <= thr[:, None]).sum(1) > C - 1 if C < M else np.zeros(d.shape[0], bool)
        if bad.any():
            full = np.argsort(d[bad], axis=1, kind="stable")[:, :kk]
            part[bad] = full
            pd[bad] = np.take_along_axis(d[bad], full, 1)
        ido[s:s + CH, :kk] = part
        d2o[s:s + CH, :kk] = pd
    return torch.from_numpy(d2o), torch.from_numpy(ido)


def three_interpolate_wrap(new_coords, coords, features, batch_size, three_nn_fn=None):
    """three_interpolate_wrap (det3d/models/point_heads/point_utils.py:8-52): per frame 3-NN between raw points
    and voxel centres, weights 1/(sqrt(d2)+1e-8) normalised, weighted sum of the 3 feature rows
    (three_interpolate_kernel_fast, interpolate_gpu.cu:84-104)."""
    outs = []
    for i in range(batch_size):
        m = coords[:, 0] == i
        nm = new_coords[:, 0] == i
        d2, idx = (three_nn_fn or three_nn)(new_coords[nm][:, 1:4].contiguous(), coords[m][:, 1:4].contiguous())
        dist = torch.sqrt(d2)
        recip = 1.0 / (dist + 1e-8)
        w = recip / recip.sum(dim=1, keepdim=True)
        f = features[m]
        g = f[idx.long()]                                                           # [n, 3, C]
        outs.append(w[:, 0:1] * g[:, 0] + w[:, 1:2] * g[:, 1] + w[:, 2:3] * g[:, 2])
    return torch.cat(outs, 0)


# ------------------------------------------------------------------------------------------ point heads
def _convcls(sd, p, x, first):
    """make_convcls_head (point_seg_mseg3d_head.py:119-134): [Dropout] + (Linear(no bias), BN, ReLU)* + Linear."""
    i = first
    while (f"{p}.{i + 1}.running_mean") in sd:
        x = torch.relu(_bn(sd, f"{p}.{i + 1}", _lin(sd, f"{p}.{i}", x), 1e-5))
        i += 3
    return _lin(sd, f"{p}.{i}", x)


def batchloss_head(sd, p, conv_point_features, conv_point_coords, points, batch_size, three_nn_fn=None):
    """PointSegBatchlossHead.forward (det3d/models/point_heads/point_seg_batchloss_head.py:122-168)."""
    conv_logits = _convcls(sd, p + "conv_cls_layers", conv_point_features, 0)
    f = three_interpolate_wrap(points, conv_point_coords, conv_point_features, batch_size, three_nn_fn)
    f = torch.relu(_bn(sd, p + "conv_align_layers.1", _lin(sd, p + "conv_align_layers.0", f), 1e-6))
    return _convcls(sd, p + "out_cls_layers", f, 0), conv_logits


def lidar_sfam(feats, logits, batch_idx, batch_size):
    """LiDARSemanticFeatureAggregationModule.forward (context_module.py:25-53): softmax over the voxels of each
    class, probs[ncls, M_i] @ feats[M_i, C] -> [B, C, ncls, 1]."""
    outs = []
    for i in range(batch_size):
        m = batch_idx == i
        pr = F.softmax(logits[m].permute(1, 0).contiguous(), dim=1)
        outs.append(pr @ feats[m])
    return torch.stack(outs, 0).permute(0, 2, 1).contiguous().unsqueeze(3)


def point_cross_attention(sd, p, query, memory, batch_idx, batch_size, nhead):
    """SparsePointCorssAttention.forward (context_module.py:320-376)."""
    E = query.shape[1]
    dh = E // nhead
    q = _lin(sd, p + ".q_proj", query).reshape(-1, nhead, dh)
    mem = memory.permute(1, 2, 0)                                                   # [B, E, L]
    k = F.conv1d(mem, sd[p + ".k_proj.weight"], sd[p + ".k_proj.bias"]).reshape(batch_size, nhead, dh, -1)
    v = F.conv1d(mem, sd[p + ".v_proj.weight"], sd[p + ".v_proj.bias"]).reshape(batch_size, nhead, dh, -1)
    outs = []
    for i in range(batch_size):
        cq = q[batch_idx == i].permute(1, 0, 2)
        sim = F.softmax(dh ** -0.5 * torch.bmm(cq, k[i]), dim=-1)
        outs.append(torch.bmm(sim, v[i].permute(0, 2, 1)).permute(1, 0, 2))
    att = torch.cat(outs, 0).reshape(-1, E)
    return _lin(sd, p + ".out_proj", att)


def sffm(sd, p, point_feats, emb1, emb2, batch_idx, batch_size, nhead, nlayer, return_memory=False):
    """SemanticFeatureFusionModule.forward + TransformerDecoder + forward_post
    (context_module.py:89-117,147-171,211-250); normalize_before = False, dropout 0."""
    tgt = _lin(sd, p + "input_proj_point", point_feats)
    e1 = F.conv1d(emb1.squeeze(-1), sd[p + "input_proj_embeddings1.weight"], sd[p + "input_proj_embeddings1.bias"])
    e2 = F.conv1d(emb2.squeeze(-1), sd[p + "input_proj_embeddings2.weight"], sd[p + "input_proj_embeddings2.bias"])
    mem = torch.cat([e1.permute(2, 0, 1), e2.permute(2, 0, 1)], dim=0).contiguous()  # [L=2ncls, B, E]
    mems = []
    for i in range(nlayer):
        q = f"{p}decoder.layers.{i}"
        mem = _ln(sd, q + ".norm1", mem + _mha_self(sd, q + ".self_attn", mem, nhead))
        mems.append(mem)
        tgt = _ln(sd, q + ".norm2", tgt + point_cross_attention(sd, q + ".crossocr_attn", tgt, mem, batch_idx,
                                                               batch_size, nhead))
        tgt = _ln(sd, q + ".norm3", tgt + _lin(sd, q + ".linear2", torch.relu(_lin(sd, q + ".linear1", tgt))))
    tgt = _ln(sd, p + "decoder.norm_tgt", tgt)
    return (tgt, mems) if return_memory else tgt


def sample_image_features(image_features, points_cuv, batch_idx):
    """get_points_image_feature (point_seg_mseg3d_head.py:200-236): 3-D grid_sample over (cam, h, w),
    bilinear, zero padding, align_corners=True; grid order (u, v, cam)."""
    img = image_features.transpose(2, 1)                                            # [B, C, ncam, h, w]
    outs = []
    for i in range(img.shape[0]):
        cuv = points_cuv[batch_idx == i]
        grid = cuv.reshape(1, 1, 1, cuv.shape[0], cuv.shape[-1])[..., (3, 2, 1)]
        s = F.grid_sample(img[i].unsqueeze(0), grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        outs.append(s.flatten(0, 1).flatten(1, 3).transpose(1, 0))
    return torch.cat(outs, 0)


def mseg3d_head(sd, p, batch, nhead, nlayer, return_all=False, three_nn_fn=None):
    """PointSegMSeg3DHead.forward (point_seg_mseg3d_head.py:240-376), return_loss=False.
    batch: conv_point_features, conv_point_coords, points [N,4], points_cuv [N,4], image_features
    [B,ncam,C,h,w], camera_semantic_embeddings [B,C,ncls,1], batch_size."""
    B = batch["batch_size"]
    vf, vc, pts = batch["conv_point_features"], batch["conv_point_coords"], batch["points"]
    voxel_logits = _convcls(sd, p + "voxel_cls_layers", vf, 1)                      # index 0 = Dropout
    f0 = three_interpolate_wrap(pts, vc, vf, B, three_nn_fn)
    fl = torch.relu(_bn(sd, p + "gffm_lidar.1", _lin(sd, p + "gffm_lidar.0", f0), 1e-6))
    cuv = batch["points_cuv"]
    valid = cuv[:, 0] == 1
    fc0 = sample_image_features(batch["image_features"], cuv[valid], pts[:, 0][valid])
    fc = torch.relu(_bn(sd, p + "gffm_camera.1", _lin(sd, p + "gffm_camera.0", fc0), 1e-6))
    fpc = _convcls(sd, p + "lidar_camera_mimic_layer", fl[valid], 0)
    cam_pad = torch.zeros(valid.shape[0], fc.shape[1], device=fc.device)
    cam_pad[valid] = fc
    pcam_pad = torch.zeros(valid.shape[0], fpc.shape[1], device=fpc.device)
    pcam_pad[valid] = fpc
    # NOTE the reference computes the mimic layer on valid points only, so invalid points get ZERO pseudo
    # camera features (point_seg_mseg3d_head.py:305-334): where(valid, cam, pcam_pad0) -> 0 for invalid rows.
    ccam = torch.where(valid.unsqueeze(-1).expand_as(cam_pad), cam_pad, pcam_pad)
    geo = torch.relu(_bn(sd, p + "gffm_lc.1", _lin(sd, p + "gffm_lc.0", torch.cat([fl, ccam], 1)), 1e-5))
    lemb = lidar_sfam(vf, voxel_logits, vc[:, 0], B)
    fused = sffm(sd, p + "sffm.", geo, batch["camera_semantic_embeddings"], lemb, pts[:, 0], B, nhead, nlayer)
    out = _lin(sd, p + "out_cls_layers", fused)
    if return_all:
        return dict(out_logits=out, voxel_logits=voxel_logits, point_features_lidar_0=f0, point_features_lidar=fl,
                    point_features_camera_0=fc0, geo_fused=geo, lidar_emb=lemb, sem_fused=fused)
    return out


# ------------------------------------------------------------------------------------------ image branch
def _conv2d(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _basic_block(sd, p, x):
    """BasicBlock (det3d/models/img_backbones/resnet_mmcv.py:20-100)."""
    out = torch.relu(_bn2d(sd, p + ".bn1", _conv2d(sd, p + ".conv1", x, 1, 1)))
    out = _bn2d(sd, p + ".bn2", _conv2d(sd, p + ".conv2", out, 1, 1))
    idt = x
    if (p + ".downsample.0.weight") in sd:
        idt = _bn2d(sd, p + ".downsample.1", _conv2d(sd, p + ".downsample.0", x))
    return torch.relu(out + idt)


def _bottleneck(sd, p, x):
    """Bottleneck, style 'pytorch' (resnet_mmcv.py:105-313): 1x1 -> 3x3 -> 1x1 (x4), BN each, ReLU."""
    out = torch.relu(_bn2d(sd, p + ".bn1", _conv2d(sd, p + ".conv1", x)))
    out = torch.relu(_bn2d(sd, p + ".bn2", _conv2d(sd, p + ".conv2", out, 1, 1)))
    out = _bn2d(sd, p + ".bn3", _conv2d(sd, p + ".conv3", out))
    idt = x
    if (p + ".downsample.0.weight") in sd:
        idt = _bn2d(sd, p + ".downsample.1", _conv2d(sd, p + ".downsample.0", x))
    return torch.relu(out + idt)


def _hr_module(sd, p, xs, num_blocks, multiscale_output=True):
    """HRModule.forward (det3d/models/img_backbones/hrnet.py:199-226) with fuse layers (:134-197)."""
    nb = len(xs)
    xs = list(xs)
    for b in range(nb):
        for k in range(num_blocks[b]):
            xs[b] = _basic_block(sd, f"{p}.branches.{b}.{k}", xs[b])
    if nb == 1:
        return xs
    outs = []
    for i in range(nb if multiscale_output else 1):
        y = 0
        for j in range(nb):
            q = f"{p}.fuse_layers.{i}.{j}"
            if i == j:
                y = y + xs[j]
            elif j > i:
                t = _bn2d(sd, q + ".1", _conv2d(sd, q + ".0", xs[j]))
                size = [int(s * float(2 ** (j - i))) for s in t.shape[-2:]]                # Upsample module
                t = F.interpolate(t, size=size, mode="bilinear", align_corners=False)
                t = F.interpolate(t, size=xs[i].shape[2:], mode="bilinear", align_corners=False)  # resize
                y = y + t
            else:
                t = xs[j]
                for k in range(i - j):
                    t = _bn2d(sd, f"{q}.{k}.1", _conv2d(sd, f"{q}.{k}.0", t, 2, 1))
                    if k != i - j - 1:
                        t = torch.relu(t)
                y = y + t
        outs.append(torch.relu(y))
    return outs


def hrnet(sd, p, x, extra):
    """HRNet.forward (hrnet.py:658-693); ``extra`` = the config's stage dict (hrnet_cfg.py:13-44)."""
    x = torch.relu(_bn2d(sd, p + "bn1", _conv2d(sd, p + "conv1", x, 2, 1)))
    x = torch.relu(_bn2d(sd, p + "bn2", _conv2d(sd, p + "conv2", x, 2, 1)))
    for k in range(extra["stage1"]["num_blocks"][0]):
        x = (_bottleneck if extra["stage1"]["block"] == "BOTTLENECK" else _basic_block)(sd, f"{p}layer1.{k}", x)

    def transition(t, prev, nbranch, first):
        xs = []
        for i in range(nbranch):
            q = f"{p}transition{t}.{i}"
            if (q + ".0.weight") in sd:                       # same-resolution channel change
                src = prev[0] if first else prev[-1]
                xs.append(torch.relu(_bn2d(sd, q + ".1", _conv2d(sd, q + ".0", src, 1, 1))))
            elif (q + ".0.0.weight") in sd:                   # new lower-resolution branch(es)
                y = prev[-1]
                k = 0
                while (f"{q}.{k}.0.weight") in sd:
                    y = torch.relu(_bn2d(sd, f"{q}.{k}.1", _conv2d(sd, f"{q}.{k}.0", y, 2, 1)))
                    k += 1
                xs.append(y)
            else:
                xs.append(prev[i])
        return xs

    ys = [x]
    for st in (2, 3, 4):
        cfg = extra[f"stage{st}"]
        xs = transition(st - 1, ys, cfg["num_branches"], st == 2)
        for m in range(cfg["num_modules"]):
            xs = _hr_module(sd, f"{p}stage{st}.{m}", xs, cfg["num_blocks"])
        ys = xs
    return ys


def camera_sfam(feats, logits, batch_size):
    """CameraSemanticFeatureAggregationModule.forward (det3d/models/img_heads/fcn_mseg3d_head.py:24-51)."""
    _, ncls, h, w = logits.shape
    C = feats.shape[1]
    pr = logits.view(batch_size, -1, ncls, h, w).permute(0, 2, 1, 3, 4).contiguous().view(batch_size, ncls, -1)
    ft = feats.view(batch_size, -1, C, h, w).permute(0, 2, 1, 3, 4).contiguous().view(batch_size, C, -1)
    emb = torch.matmul(F.softmax(pr, dim=2), ft.permute(0, 2, 1))
    return emb.permute(0, 2, 1).contiguous().unsqueeze(3)


def fcn_head(sd, p, inputs, batch_size, num_convs=2):
    """FCNMSeg3DHead.forward (fcn_mseg3d_head.py:173-199) with resize_concat (_transform_inputs,
    decode_head.py:151-160), 1x1 ConvModules (conv, BN, ReLU; conv bias off), cls_seg (decode_head.py:213-218)."""
    ups = [F.interpolate(x, size=inputs[0].shape[2:], mode="bilinear", align_corners=False) for x in inputs]
    x = torch.cat(ups, dim=1)
    for i in range(num_convs):
        x = torch.relu(_bn2d(sd, f"{p}convs.{i}.bn", _conv2d(sd, f"{p}convs.{i}.conv", x)))
    logits = _conv2d(sd, p + "conv_seg", x)
    return x, logits, camera_sfam(x, logits, batch_size)


# ------------------------------------------------------------------------------------------ detectors
def segnet_forward(sd, example, cfg, backend=None):
    """SegNet.forward(return_loss=False) up to out_logits (det3d/models/detectors/seg_net.py:51-107).
    cfg: dict(voxel_size, pc_range, reader=dict(type, num_head, num_layers))."""
    B = len(example["num_voxels"])
    pts = example["points"][:, 0:4]
    rd = cfg["reader"]
    if rd["type"] == "TransformerVoxelFeatureExtractor":
        vf = trans_vfe(sd, "reader.", example["voxels"], example["num_points"], rd["num_head"], rd["num_layers"])
    elif rd["type"] == "ImprovedMeanVoxelFeatureExtractor":
        vf = improved_mean_vfe(example["voxels"], example["num_points"])
    else:
        vf = mean_vfe(example["voxels"], example["num_points"])
    feats, coords = unet_scn3d(sd, "backbone.", vf, example["coordinates"], list(example["shape"][0]),
                               cfg["voxel_size"], cfg["pc_range"], backend=backend)
    out, conv_logits = batchloss_head(sd, "point_head.", feats, coords, pts, B,
                                      three_nn_fn=backend.three_nn if backend is not None else None)
    return out


def mseg3d_forward(sd, example, cfg, return_all=False, backend=None):
    """SegMSeg3DNet.forward(return_loss=False) up to out_logits (det3d/models/detectors/seg_mseg3d_net.py:47-147).
    cfg: dict(voxel_size, pc_range, hrnet_extra, nhead, nlayer, num_convs)."""
    B = len(example["num_voxels"])
    pts = example["points"][:, 0:4]
    images = example["images"]
    ncam, hi, wi = images.shape[1], images.shape[3], images.shape[4]
    ys = hrnet(sd, "img_backbone.", images.view(-1, 3, hi, wi), cfg["hrnet_extra"])
    feat, logits, cam_emb = fcn_head(sd, "img_head.", ys, B, cfg.get("num_convs", 2))
    image_features = feat.view(B, ncam, feat.shape[1], feat.shape[2], feat.shape[3])
    vf = improved_mean_vfe(example["voxels"], example["num_points"])
    feats, coords = unet_scn3d(sd, "backbone.", vf, example["coordinates"], list(example["shape"][0]),
                               cfg["voxel_size"], cfg["pc_range"], backend=backend)
    batch = dict(batch_size=B, conv_point_features=feats, conv_point_coords=coords, points=pts,
                 points_cuv=example["points_cuv"], image_features=image_features,
                 camera_semantic_embeddings=cam_emb)
    r = mseg3d_head(sd, "point_head.", batch, cfg["nhead"], cfg["nlayer"], return_all=return_all,
                    three_nn_fn=backend.three_nn if backend is not None else None)
    if return_all:
        r.update(image_features=image_features, image_logits=logits, camera_semantic_embeddings=cam_emb,
                 conv_point_features=feats, conv_point_coords=coords, voxel_features=vf)
    return r


def predict_labels(out_logits, points, batch_size):
    """predict() non-TTA branch (point_seg_mseg3d_head.py:455-477): argmax, split per frame."""
    lab = torch.argmax(out_logits, dim=1)
    return [lab[points[:, 0] == i] for i in range(batch_size)]


def image_input_transform(images_u8, mean, std):
    """det3d/datasets/pipelines/img_transforms.py:18-29 applied per camera (segpreprocess.py:621-628), then the
    [.., H, W, 3] -> [.., 3, H, W] transpose of segpreprocess.py:637.  numpy fp32, same operation order; mean / std are
    rounded to fp32 first (the reference's in-place ops promote its python-float lists to fp64 and round the result back, so
    the two differ by at most 1 fp32 ulp per operation - tests/test_oracle_golden.py pins that bound)."""
    import numpy as np
    image = np.asarray(images_u8).astype(np.float32)
    image = image / 255.0
    image -= np.asarray(mean, np.float32).reshape(1, 1, 3)
    image /= np.asarray(std, np.float32).reshape(1, 1, 3)
    nd = image.ndim
    return np.ascontiguousarray(image.transpose(*range(nd - 3), nd - 1, nd - 3, nd - 2))
