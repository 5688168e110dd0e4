"""Voxel feature encoders behind the READERS registry (reference det3d/models/readers/voxel_encoder.py)."""
import torch
from torch import nn

from .. import gemm, ops
from .common import Prepared, linear_pack
from .registry import READERS


@READERS.register_module
class MeanVoxelFeatureExtractor(nn.Module):
    """voxel_encoder.py:40-58."""

    def __init__(self, num_input_features=4, name="MeanVoxelFeatureExtractor"):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features

    def forward(self, features, num_voxels, coors=None):
        assert self.num_input_features == features.shape[-1]
        return ops.vfe_descriptor(features, num_voxels, mode=0)


@READERS.register_module
class ImprovedMeanVoxelFeatureExtractor(nn.Module):
    """voxel_encoder.py:63-124: [mean xyz, max xyz, min xyz, mean feats, density, std]."""

    def __init__(self, num_input_features=4, norm_cfg=None, name="ImprovedMeanVoxelFeatureExtractor"):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features

    def forward(self, features, num_voxels, coors=None):
        assert self.num_input_features == features.shape[-1]
        F = features.shape[-1]
        ld = (F + 8 + 3) // 4 * 4
        out = ops.vfe_descriptor(features, num_voxels, mode=1, ld_out=ld)
        return out[:, :F + 8]       # view of a zero-padded, 16-byte aligned row buffer


class _EncoderLayerParams(nn.Module):
    """Parameter container with the names of TransformerEncoderLayerPreNorm (voxel_encoder.py:128-147)."""

    def __init__(self, d_model, nhead, dim_feedforward):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=0.0)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)


class _Chunk(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayerParams(d_model, nhead, dim_feedforward) for _ in range(num_layers)])


@READERS.register_module
class TransformerVoxelFeatureExtractor(Prepared):
    """TransVFE (voxel_encoder.py:167-270).  State-dict names: feature_conv.0, chunck.layers.{i}.*, compress_layer.0.

    Forward = descriptor kernel -> token GEMMs (tcgen05) with LayerNorm / residual epilogues -> 5-slot attention
    kernel -> slot max -> compress GEMM.  The layer adds its residual to the *normalised* tensor and attends over
    zero-padded slots, exactly like the reference (voxel_encoder.py:149-163).
    """

    def __init__(self, num_input_features=4, num_compressed_features=16, num_embed=64, num_head=4, num_layers=2,
                 norm_cfg=None, name="TransformerVoxelFeatureExtractor"):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features
        nd = num_input_features + 8
        self.num_embed, self.num_head, self.num_layers = num_embed, num_head, num_layers
        self.feature_conv = nn.Sequential(nn.Conv1d(num_input_features + nd, num_embed, 1, bias=True))
        self.chunck = _Chunk(num_embed, num_head, num_embed * 2, num_layers)
        if num_compressed_features > 0:
            self.compress_layer = nn.Sequential(nn.Linear(num_embed, num_compressed_features), nn.ReLU())
            self.num_out_features = num_compressed_features
        else:
            self.compress_layer = None
            self.num_out_features = num_embed

    def _prepare(self):
        P = dict(conv=linear_pack(self.feature_conv[0]), layers=[])
        for ly in self.chunck.layers:
            w_in = gemm.PackedWeight.from_linear(ly.self_attn.in_proj_weight.detach())
            P["layers"].append(dict(
                in_proj=(w_in, ly.self_attn.in_proj_bias.detach().float().contiguous()),
                out_proj=linear_pack(ly.self_attn.out_proj), lin1=linear_pack(ly.linear1), lin2=linear_pack(ly.linear2),
                n1=(ly.norm1.weight.detach().float().contiguous(), ly.norm1.bias.detach().float().contiguous()),
                n2=(ly.norm2.weight.detach().float().contiguous(), ly.norm2.bias.detach().float().contiguous())))
        if self.compress_layer is not None:
            P["compress"] = linear_pack(self.compress_layer[0])
        return P

    def forward(self, features, num_voxels, coors=None):
        assert self.num_input_features == features.shape[-1]
        M, Pn, F = features.shape
        E, H = self.num_embed, self.num_head
        Pk = self.prep()
        tok = ops.vfe_descriptor(features, num_voxels, mode=2, round_out=not gemm.PRECISE)                  # [M*P, 2F+8]
        L = Pk["layers"]
        # feature_conv (+bias) then LN1 of layer 0
        x = gemm.run(tok, Pk["conv"][0], shift=Pk["conv"][1], ln=(L[0]["n1"],))
        for i, ly in enumerate(L):
            qkv = gemm.run(x, ly["in_proj"][0], shift=ly["in_proj"][1])        # [M*P, 3E]
            ctx = ops.vfe_token_attn(qkv, M, Pn, H, E // H, round_out=not gemm.PRECISE)
            x = gemm.run(ctx, ly["out_proj"][0], shift=ly["out_proj"][1], res=x, res_mode=1, ln=(ly["n2"],))
            h = gemm.run(x, ly["lin1"][0], shift=ly["lin1"][1], relu=True)
            nxt = (L[i + 1]["n1"],) if i + 1 < len(L) else ()
            x = gemm.run(h, ly["lin2"][0], shift=ly["lin2"][1], res=x, res_mode=1, ln=nxt)
        v = ops.vfe_token_max(x, M, Pn, round_out=not gemm.PRECISE)
        if self.compress_layer is not None:
            v = gemm.run(v, Pk["compress"][0], shift=Pk["compress"][1], relu=True, round_out=False)
        return v
