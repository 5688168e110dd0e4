"""Shared host helpers for the registry-built modules: BatchNorm folding, weight packing caches."""
import torch
from torch import nn

from .. import gemm


def fold_bn(bn: nn.modules.batchnorm._BatchNorm):
    """Eval-mode BatchNorm as y = x * scale + shift."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


def pad_cols(x, mult=4):
    """Zero-pad the channel dim to a multiple of ``mult`` (gather-GEMM reads 16-byte chunks)."""
    c = x.shape[1]
    p = (-c) % mult
    if p == 0 and x.stride(1) == 1:
        return x
    return torch.nn.functional.pad(x, (0, p)).contiguous()


class Prepared(nn.Module):
    """Mixin: lazily built inference cache (packed tf32 weights, folded BN), dropped when parameters change."""

    def __init__(self):
        super().__init__()
        self._prep = None

    def _prepare(self):
        raise NotImplementedError

    def prep(self):
        if self._prep is None:
            with torch.no_grad():
                self._prep = self._prepare()
        return self._prep

    def invalidate(self):
        self._prep = None

    def _apply(self, fn, *a, **k):
        self._prep = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._prep = None
        return super().load_state_dict(*a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._prep = None
        return super()._load_from_state_dict(*a, **k)

    def train(self, mode=True):
        self._prep = None
        return super().train(mode)


def linear_pack(lin):
    """nn.Linear / nn.Conv1d(k=1) -> (PackedWeight, bias or None)."""
    w = lin.weight.detach()
    if w.dim() == 3:
        w = w.squeeze(-1)
    b = lin.bias.detach().float().contiguous() if lin.bias is not None else None
    return gemm.PackedWeight.from_linear(w), b


def linear_bn_pack(lin, bn):
    """Linear (+bias) followed by eval BatchNorm -> (PackedWeight, scale, shift)."""
    pw, b = linear_pack(lin)
    scale, shift = fold_bn(bn)
    if b is not None:
        shift = shift + b * scale
    return pw, scale, shift.contiguous()
