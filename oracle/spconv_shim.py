"""spconv 1.x API shim over oracle/sparse.py, so that the REFERENCE's own backbone / detector code can be executed on the CPU.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py inside the build container to generate tests/golden/ref_backbones.pt);
nothing under lidarseg3d_b200/ imports it.

spconv 1.x @ fad3000249d27ca918f2655ff73c41f39b0f3127 is NOT vendored under /root/reference (docs/INSTALL.md:88-99), so this
is a restatement of its published Python surface: ``SparseConvTensor`` (mutable ``features``, shared
``indice_dict``), ``SparseModule``, ``SparseSequential`` (sparse modules get the tensor, plain modules the features),
``SubMConv3d`` / ``SparseConv3d`` / ``SparseInverseConv3d`` (weight ``[kz,ky,kx,Cin,Cout]``, rulebook cached under
``indice_key`` and reused whatever the later layer's own kernel size / padding says, inverse conv = cached pairs with in/out
swapped).  The arithmetic is oracle.sparse (pinned to dense F.conv3d / F.conv_transpose3d in tests/test_oracle_sparse.py).
With it installed as ``sys.modules['spconv']`` the reference files det3d/models/backbones/scn_unet.py, scn.py and
det3d/models/detectors/{seg_net,seg_mseg3d_net}.py run unmodified: their wiring (UR_block_forward, channel_reduction,
indice_key reuse, state-dict names) is what the goldens pin.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import sparse as osp


def _triple(v):
    return tuple(int(x) for x in v) if isinstance(v, (tuple, list, np.ndarray)) else (int(v),) * 3


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(v) for v in spatial_shape]
        self.batch_size = batch_size
        self.indice_dict = {}
        self.grid = grid

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key, None)

    def dense(self, channels_first=True):
        """[B, C, D, H, W] (scn.py:165)."""
        D, H, W = self.spatial_shape
        C = self.features.shape[1]
        out = torch.zeros(self.batch_size, D, H, W, C, dtype=self.features.dtype)
        idx = self.indices.long()
        out[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]] = self.features
        return out.permute(0, 4, 1, 2, 3).contiguous() if channels_first else out


class SparseModule(nn.Module):
    """Marker base class: SparseSequential hands these the SparseConvTensor itself."""


def is_spconv_module(m):
    return isinstance(m, SparseModule)


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            self.add_module(name, module)

    def __getitem__(self, idx):
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def forward(self, input):
        for module in self._modules.values():
            if is_spconv_module(module):
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input.features = module(input.features)
            else:
                input = module(input)
        return input


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None, fused_bn=False,
                 use_hash=False, algo=None):
        super().__init__()
        assert ndim == 3 and groups == 1 and _triple(dilation) == (1, 1, 1) and not transposed
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _triple(kernel_size), _triple(stride), _triple(padding)
        self.conv1x1 = int(np.prod(self.kernel_size)) == 1
        self.subm, self.inverse, self.indice_key = subm, inverse, indice_key
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        feats, indices, shape = input.features, input.indices, tuple(input.spatial_shape)
        if self.conv1x1:                                  # spconv short-circuits 1x1x1 kernels to a dense mm
            out = feats @ self.weight.view(self.in_channels, self.out_channels)
            if self.bias is not None:
                out = out + self.bias
            t = SparseConvTensor(out, indices, shape, input.batch_size)
            t.indice_dict, t.grid = input.indice_dict, input.grid
            return t
        datas = input.find_indice_pair(self.indice_key)
        if self.inverse:
            assert datas is not None and self.indice_key is not None, "SparseInverseConv3d needs the cached pairs of its key"
            nbr = osp.invert_rulebook(datas["nbr"], datas["in_indices"].shape[0])
            outids, oshape = datas["in_indices"], datas["in_shape"]
        elif self.indice_key is not None and datas is not None:
            nbr, outids, oshape = datas["nbr"], datas["outids"], datas["out_shape"]
        else:
            idx_np = indices.numpy().astype(np.int32)
            if self.subm:
                nbr = osp.subm_rulebook(idx_np, shape, self.kernel_size)
                outids, oshape = indices, shape
            else:
                oi, oshape, nbr = osp.strided_rulebook(idx_np, shape, self.kernel_size, self.stride, self.padding)
                outids = torch.from_numpy(oi)
            if self.indice_key is not None:
                input.indice_dict[self.indice_key] = dict(nbr=nbr, outids=outids, out_shape=tuple(oshape), in_indices=indices,
                                                          in_shape=shape)
        K = nbr.shape[0]
        assert K == int(np.prod(self.kernel_size)), \
            f"cached rulebook of key {self.indice_key!r} has {K} offsets, this layer's kernel {self.kernel_size}"
        out = osp.sparse_conv(feats, self.weight, nbr)
        if self.bias is not None:
            out = out + self.bias
        t = SparseConvTensor(out, outids, oshape, input.batch_size)
        t.indice_dict, t.grid = input.indice_dict, input.grid
        return t


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, algo=None):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key)


def install():
    """Publish this module as ``spconv`` (+ ``spconv.pytorch``-less 1.x layout) in sys.modules."""
    import sys
    import types
    m = types.ModuleType("spconv")
    for k in ("SparseConvTensor", "SparseModule", "SparseSequential", "SparseConv3d", "SubMConv3d", "SparseInverseConv3d"):
        setattr(m, k, globals()[k])
    sys.modules["spconv"] = m
    return m
