"""Residual merge ``(a + b) * gain`` as one kernel (reference: two torch ops, models.py:178,227)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from ... import _lib
from ..._tensor import is_nhwc_dense, nhwc, ptr, require_cuda, stream_ptr


class AddScale(Function):
    @staticmethod
    def forward(ctx, a, b, gain):
        require_cuda(a, b)
        if a.shape != b.shape:
            raise RuntimeError(f"add_scale: shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
        if a.dim() == 4:
            a, b = nhwc(a), nhwc(b)
        else:
            a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        _lib.call("ideas_add_scale", ptr(out), ptr(a), ptr(b), float(gain), a.numel(), stream_ptr(a))
        ctx.gain = gain
        return out

    @staticmethod
    def backward(ctx, g):
        gg = g * ctx.gain          # linear: differentiable to any order through torch
        return gg, gg, None


def add_scale(a, b, gain):
    return AddScale.apply(a, b, gain)
