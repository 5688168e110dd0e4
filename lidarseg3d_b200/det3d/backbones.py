"""UNetSCN3D behind the BACKBONES registry (reference det3d/models/backbones/scn_unet.py:73-249).

Same constructor kwargs, state-dict names and ``forward(batch_dict) -> batch_dict`` contract as the reference; the
spconv modules are replaced by parameter holders, and the forward runs on the bitmap rulebook kernels
(csrc/rulebook.cu) + the tcgen05 gather-GEMM (csrc/gather_gemm_once.cu) with BatchNorm / ReLU / residual / concat /
channel-reduction fused into the GEMM epilogue.
"""
import math
from functools import partial

import numpy as np
import torch
from torch import nn

from .. import gemm, ops
from .common import Prepared, fold_bn, pad_cols
from .registry import BACKBONES


class SparseConvWeight(nn.Module):
    """Holds ``weight [kz,ky,kx,Cin,Cout]`` (spconv 1.x layout, bias=False everywhere in UNetSCN3D)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=False):
        super().__init__()
        ks = tuple(kernel_size) if isinstance(kernel_size, (tuple, list)) else (kernel_size,) * 3
        self.kernel_size = ks
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(*ks, in_channels, out_channels))
        # spconv reset_parameters: kaiming_uniform(a=sqrt(5)) with fan_in = Cin * prod(k)
        bound = 1.0 / math.sqrt(in_channels * ks[0] * ks[1] * ks[2])
        nn.init.uniform_(self.weight, -bound, bound)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)

    def packed(self):
        w = self.weight.detach()
        return gemm.PackedWeight(w.reshape(-1, w.shape[-2], w.shape[-1]))


def _conv_bn_relu(cin, cout, ksize, norm_fn):
    """post_act_block (scn_unet.py:11-30): SparseSequential(conv, BN, ReLU) -> keys .0.weight, .1.*"""
    return nn.Sequential(SparseConvWeight(cin, cout, ksize), norm_fn(cout), nn.ReLU())


class SparseBasicBlock(nn.Module):
    """scn_unet.py:34-69; keys conv1.weight, bn1.*, conv2.weight, bn2.*"""

    def __init__(self, inplanes, planes, norm_fn):
        super().__init__()
        self.conv1 = SparseConvWeight(inplanes, planes, 3)
        self.bn1 = norm_fn(planes)
        self.conv2 = SparseConvWeight(planes, planes, 3)
        self.bn2 = norm_fn(planes)


class SparseLevel:
    """Active sites of one resolution level + its bitmap and cached SubM table (spconv's indice_dict entry)."""

    def __init__(self, coords, B, shape, grid):
        self.coords, self.B, self.shape, self.grid = coords, B, shape, grid
        self.subm = None

    def subm_table(self):
        if self.subm is None:
            self.subm = ops.rulebook_gather(self.grid, self.coords, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        return self.subm


class SparseTensorView:
    """What the reference exposes as a SparseConvTensor on ``batch_dict`` (features / indices / spatial_shape)."""

    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features, self.indices, self.spatial_shape, self.batch_size = features, indices, list(spatial_shape), batch_size


@BACKBONES.register_module
class UNetSCN3D(Prepared):
    def __init__(self, num_input_features=128, name="UNetSCN3D", voxel_size=[], point_cloud_range=[], model_cfg={},
                 **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        r = model_cfg.get("SCALING_RATIO", 1)
        c1, c2, c3, c4 = 16 * r, 32 * r, 64 * r, 64 * r
        self.conv_input = _conv_bn_relu(num_input_features, c1, 3, norm_fn)
        self.conv1 = nn.Sequential(SparseBasicBlock(c1, c1, norm_fn), SparseBasicBlock(c1, c1, norm_fn))
        self.conv2 = nn.Sequential(_conv_bn_relu(c1, c2, 3, norm_fn), SparseBasicBlock(c2, c2, norm_fn),
                                   SparseBasicBlock(c2, c2, norm_fn))
        self.conv3 = nn.Sequential(_conv_bn_relu(c2, c3, 3, norm_fn), SparseBasicBlock(c3, c3, norm_fn),
                                   SparseBasicBlock(c3, c3, norm_fn))
        self.conv4 = nn.Sequential(_conv_bn_relu(c3, c4, 3, norm_fn), SparseBasicBlock(c4, c4, norm_fn),
                                   SparseBasicBlock(c4, c4, norm_fn))
        if model_cfg.get("RETURN_ENCODED_TENSOR", True):
            self.last_pad = model_cfg.get("last_pad", 0)
            self.conv_out = _conv_bn_relu(c4, 128, (3, 1, 1), norm_fn)
        else:
            self.conv_out = None
        self.conv_up_t4 = SparseBasicBlock(c4, c4, norm_fn)
        self.conv_up_m4 = _conv_bn_relu(2 * c4, c4, 3, norm_fn)
        self.inv_conv4 = _conv_bn_relu(c4, c3, 3, norm_fn)
        self.conv_up_t3 = SparseBasicBlock(c3, c3, norm_fn)
        self.conv_up_m3 = _conv_bn_relu(2 * c3, c3, 3, norm_fn)
        self.inv_conv3 = _conv_bn_relu(c3, c2, 3, norm_fn)
        self.conv_up_t2 = SparseBasicBlock(c2, c2, norm_fn)
        self.conv_up_m2 = _conv_bn_relu(2 * c2, c2, 3, norm_fn)
        self.inv_conv2 = _conv_bn_relu(c2, c1, 3, norm_fn)
        self.conv_up_t1 = SparseBasicBlock(c1, c1, norm_fn)
        self.conv_up_m1 = _conv_bn_relu(2 * c1, c1, 3, norm_fn)
        self.conv5 = nn.Sequential(_conv_bn_relu(c1, c1, 3, norm_fn))
        self.num_point_features = c1
        # strided encoder geometry (scn_unet.py:106,113,120): kernel 3, stride 2, padding 1 / 1 / (0,1,1)
        self.down_geom = {2: ((3, 3, 3), (2, 2, 2), (1, 1, 1)), 3: ((3, 3, 3), (2, 2, 2), (1, 1, 1)),
                          4: ((3, 3, 3), (2, 2, 2), (0, 1, 1))}

    # ------------------------------------------------------------------ inference cache
    def _prepare(self):
        def cbr(seq):
            s, b = fold_bn(seq[1])
            return (seq[0].packed(), s, b)

        def blk(m):
            s1, b1 = fold_bn(m.bn1)
            s2, b2 = fold_bn(m.bn2)
            return ((m.conv1.packed(), s1, b1), (m.conv2.packed(), s2, b2))

        P = {"conv_input": cbr(self.conv_input), "conv1": [blk(m) for m in self.conv1], "conv5": cbr(self.conv5[0])}
        for lv in (2, 3, 4):
            seq = getattr(self, f"conv{lv}")
            P[f"conv{lv}"] = (cbr(seq[0]), [blk(seq[1]), blk(seq[2])])
            P[f"inv{lv}"] = cbr(getattr(self, f"inv_conv{lv}"))
        for lv in (1, 2, 3, 4):
            P[f"t{lv}"] = blk(getattr(self, f"conv_up_t{lv}"))
            P[f"m{lv}"] = cbr(getattr(self, f"conv_up_m{lv}"))
        if self.conv_out is not None:
            P["conv_out"] = cbr(self.conv_out)
        return P

    @staticmethod
    def channel_reduction(features, out_channels):
        """scn_unet.py:173-187 (host-side twin of the fused epilogue; kept for API parity)."""
        n, c = features.shape
        assert (c % out_channels == 0) and (c >= out_channels)
        return features.view(n, out_channels, -1).sum(dim=2)

    # ------------------------------------------------------------------ building blocks
    @staticmethod
    def _conv(x, pk, nbr, x1=None, relu=True, res=None, res_mode=0, red=None):
        pw, s, b = pk
        return gemm.run(x, pw, x1=x1, nbr=nbr, scale=s, shift=b, relu=relu, res=res, res_mode=res_mode, red=red)

    def _block(self, x, pk, nbr):
        out = self._conv(x, pk[0], nbr)                                   # relu(bn1(conv1 x))
        return self._conv(out, pk[1], nbr, res=x, res_mode=1)             # relu(bn2(conv2 .) + x)

    def forward(self, batch_dict):
        if self.training:
            raise NotImplementedError("lidarseg3d_b200 UNetSCN3D: inference path only (model.eval())")
        P = self.prep()
        voxel_features, voxel_coords = batch_dict["voxel_features"], batch_dict["voxel_coords"]
        B = batch_dict["batch_size"]
        shape1 = tuple(int(v) for v in (np.array(batch_dict["input_shape"][::-1]) + [1, 0, 0]))   # scn_unet.py:203
        coords1 = voxel_coords.int().contiguous()
        # ---- phase 1: geometry of every level (bitmaps, output sites, rulebooks).  It depends on the coordinates only, so
        # all of it - including the host reads of the site counts - is issued before any feature kernel; the GEMM chain
        # below then runs without a single host synchronisation.
        lv1 = SparseLevel(coords1, B, shape1, ops.grid_from_coords(coords1, B, shape1, need_perm=True))
        levels = {1: lv1}
        down, up = {}, {}
        for lv in (2, 3, 4):
            ks, st, pd = self.down_geom[lv]
            prev = levels[lv - 1]
            grid, ocoords = ops.grid_strided(prev.coords, B, prev.shape, ks, st, pd)
            levels[lv] = SparseLevel(ocoords, B, grid.shape, grid)
            down[lv] = ops.rulebook_gather(prev.grid, ocoords, ks, st, pd)
            up[lv] = ops.rulebook_scatter(grid, prev.coords, ks, st, pd)
        enc = None
        if self.conv_out is not None:
            lp = self.last_pad if isinstance(self.last_pad, (tuple, list)) else (self.last_pad,) * 3
            l4 = levels[4]
            g5, c5 = ops.grid_strided(l4.coords, B, l4.shape, (3, 1, 1), (2, 1, 1), tuple(lp))
            enc = (ops.rulebook_gather(l4.grid, c5, (3, 1, 1), (2, 1, 1), tuple(lp)), c5, g5.shape)
        for lv in (1, 2, 3, 4):
            levels[lv].subm_table()
        # ---- phase 2: features
        vf = pad_cols(voxel_features.float())
        x = self._conv(vf, P["conv_input"], lv1.subm_table())
        for pk in P["conv1"]:
            x = self._block(x, pk, lv1.subm_table())
        feats = {1: x}
        for lv in (2, 3, 4):
            cbr, blocks = P[f"conv{lv}"]
            y = self._conv(feats[lv - 1], cbr, down[lv])
            for pk in blocks:
                y = self._block(y, pk, levels[lv].subm_table())
            feats[lv] = y
        if enc is not None:
            batch_dict["encoded_spconv_tensor"] = SparseTensorView(self._conv(feats[4], P["conv_out"], enc[0]), enc[1], enc[2], B)
            batch_dict["encoded_spconv_tensor_stride"] = 8

        def ur_block(lat, bottom, lv):
            """UR_block_forward (scn_unet.py:163-171) with concat + channel_reduction fused into conv_m's GEMM."""
            ns = levels[lv].subm_table()
            t = self._block(lat, P[f"t{lv}"], ns)
            y = self._conv(bottom, P[f"m{lv}"], ns, x1=t, red=(bottom, t))          # relu(bn(conv_m cat)) + red
            if lv > 1:
                return self._conv(y, P[f"inv{lv}"], up[lv])                          # SparseInverseConv3d + BN + ReLU
            return self._conv(y, P["conv5"], ns)

        x_up4 = ur_block(feats[4], feats[4], 4)
        x_up3 = ur_block(feats[3], x_up4, 3)
        x_up2 = ur_block(feats[2], x_up3, 2)
        x_up1 = ur_block(feats[1], x_up2, 1)

        def view(f, lv):
            return SparseTensorView(f, levels[lv].coords, levels[lv].shape, B)

        batch_dict.update({"multi_scale_3d_features": {"x_conv1": view(x_up2, 1), "x_conv2": view(x_up3, 2),
                                                       "x_conv3": view(x_up4, 3), "x_conv4": view(feats[4], 4)}})
        batch_dict["conv_point_features"] = x_up1
        # get_voxel_centers (det3d/core/utils/common_utils.py:74-90)
        vs = torch.tensor(self.voxel_size, device=coords1.device).float()
        lo = torch.tensor(self.point_cloud_range[0:3], device=coords1.device).float()
        centers = (coords1[:, [3, 2, 1]].float() + 0.5) * vs + lo
        batch_dict["conv_point_coords"] = torch.cat((coords1[:, 0:1].float(), centers), dim=1)
        batch_dict["_ls3d_voxel_size"] = [float(v) for v in self.voxel_size]
        batch_dict["_ls3d_pc_range"] = [float(v) for v in self.point_cloud_range]
        batch_dict["_ls3d_level1"] = lv1           # bitmap of the voxel grid, reused by the devoxelization kernel
        batch_dict["_ls3d_levels"] = levels
        batch_dict["_ls3d_rulebooks"] = dict(down=down, up=up)
        return batch_dict


class SparseBasicBlockBias(nn.Module):
    """SparseBasicBlock of scn.py:37-81: like the UNet's, but its two SubM convolutions carry a bias (scn.py:56)."""

    def __init__(self, planes, norm_fn):
        super().__init__()
        self.conv1 = SparseConvWeight(planes, planes, 3, bias=True)
        self.bn1 = norm_fn(planes)
        self.conv2 = SparseConvWeight(planes, planes, 3, bias=True)
        self.bn2 = norm_fn(planes)


@BACKBONES.register_module
class SpMiddleResNetFHD(Prepared):
    """Reference det3d/models/backbones/scn.py:84-177 (the sparse ResNet encoder of the detection configs): same constructor,
    state-dict names and ``forward(voxel_features, coors, batch_size, input_shape) -> (dense [N, C*D, H, W], multi_scale)``
    contract, on the bitmap rulebooks + tcgen05 gather-GEMM (conv bias and eval BatchNorm folded into the epilogue)."""

    def __init__(self, num_input_features=128, norm_cfg=None, name="SpMiddleResNetFHD", **kwargs):
        super().__init__()
        self.name = name
        norm_cfg = norm_cfg or dict(type="BN1d", eps=1e-3, momentum=0.01)
        if norm_cfg.get("type", "BN1d") != "BN1d":
            raise NotImplementedError("SpMiddleResNetFHD: BN1d norm layers only (scn.py default)")
        norm_fn = partial(nn.BatchNorm1d, eps=norm_cfg.get("eps", 1e-3), momentum=norm_cfg.get("momentum", 0.01))
        self.conv_input = nn.Sequential(SparseConvWeight(num_input_features, 16, 3), norm_fn(16), nn.ReLU())
        self.conv1 = nn.Sequential(SparseBasicBlockBias(16, norm_fn), SparseBasicBlockBias(16, norm_fn))

        def stage(cin, cout):
            return nn.Sequential(SparseConvWeight(cin, cout, 3), norm_fn(cout), nn.ReLU(), SparseBasicBlockBias(cout, norm_fn),
                                 SparseBasicBlockBias(cout, norm_fn))

        self.conv2, self.conv3, self.conv4 = stage(16, 32), stage(32, 64), stage(64, 128)
        self.extra_conv = nn.Sequential(SparseConvWeight(128, 128, (3, 1, 1)), norm_fn(128), nn.ReLU())
        self.down_geom = {2: ((3, 3, 3), (2, 2, 2), (1, 1, 1)), 3: ((3, 3, 3), (2, 2, 2), (1, 1, 1)),
                          4: ((3, 3, 3), (2, 2, 2), (0, 1, 1))}

    def _prepare(self):
        def cbr(conv, bn):
            s, b = fold_bn(bn)
            if conv.bias is not None:
                b = b + conv.bias.detach().float() * s
            return (conv.packed(), s, b.contiguous())

        def blk(m):
            return (cbr(m.conv1, m.bn1), cbr(m.conv2, m.bn2))

        P = {"conv_input": cbr(self.conv_input[0], self.conv_input[1]), "conv1": [blk(m) for m in self.conv1],
             "extra": cbr(self.extra_conv[0], self.extra_conv[1])}
        for lv in (2, 3, 4):
            seq = getattr(self, f"conv{lv}")
            P[f"conv{lv}"] = (cbr(seq[0], seq[1]), [blk(seq[3]), blk(seq[4])])
        return P

    def forward(self, voxel_features, coors, batch_size, input_shape):
        if self.training:
            raise NotImplementedError("lidarseg3d_b200 SpMiddleResNetFHD: inference path only (model.eval())")
        P = self.prep()
        conv = UNetSCN3D._conv

        def block(_, x, pk, nbr):                     # relu(bn2(conv2(relu(bn1(conv1 x)))) + x), scn.py:63-81
            return conv(conv(x, pk[0], nbr), pk[1], nbr, res=x, res_mode=1)

        B = batch_size
        shape1 = tuple(int(v) for v in (np.array(input_shape[::-1]) + [1, 0, 0]))            # scn.py:149
        coords1 = coors.int().contiguous()
        lv1 = SparseLevel(coords1, B, shape1, ops.grid_from_coords(coords1, B, shape1, need_perm=True))
        levels, down = {1: lv1}, {}
        for lv in (2, 3, 4):
            ks, st, pd = self.down_geom[lv]
            prev = levels[lv - 1]
            grid, oc = ops.grid_strided(prev.coords, B, prev.shape, ks, st, pd)
            levels[lv] = SparseLevel(oc, B, grid.shape, grid)
            down[lv] = ops.rulebook_gather(prev.grid, oc, ks, st, pd)
        l4 = levels[4]
        g5, c5 = ops.grid_strided(l4.coords, B, l4.shape, (3, 1, 1), (2, 1, 1), (0, 0, 0))
        nb5 = ops.rulebook_gather(l4.grid, c5, (3, 1, 1), (2, 1, 1), (0, 0, 0))
        x = conv(pad_cols(voxel_features.float()), P["conv_input"], lv1.subm_table())
        for pk in P["conv1"]:
            x = block(self, x, pk, lv1.subm_table())
        feats = {1: x}
        for lv in (2, 3, 4):
            cbr, blocks = P[f"conv{lv}"]
            y = conv(feats[lv - 1], cbr, down[lv])
            for pk in blocks:
                y = block(self, y, pk, levels[lv].subm_table())
            feats[lv] = y
        y = conv(feats[4], P["extra"], nb5)
        D, H, W = g5.shape
        dense = torch.zeros(B, D, H, W, y.shape[1], dtype=y.dtype, device=y.device)           # SparseConvTensor.dense()
        ci = c5.long()
        dense[ci[:, 0], ci[:, 1], ci[:, 2], ci[:, 3]] = y
        ret = dense.permute(0, 4, 1, 2, 3).contiguous().view(B, y.shape[1] * D, H, W)         # scn.py:165-168
        ms = {f"conv{lv}": SparseTensorView(feats[lv], levels[lv].coords, levels[lv].shape, B) for lv in (1, 2, 3, 4)}
        return ret, ms
