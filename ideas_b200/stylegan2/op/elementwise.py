"""Residual merge ``(a + b) * gain`` as one kernel (reference: two torch ops, models.py:178,227) and
NHWC reflection padding (reference: nn.ReflectionPad2d, models.py:102-108)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from ... import _lib
from ..._tensor import empty_nhwc, is_nhwc_dense, nhwc, ptr, require_cuda, stream_ptr


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


class ReflectPad(Function):
    """x (N,C,H,W) -> (N,C,H+2p,W+2p), reflection without repeating the edge (nn.ReflectionPad2d).  Stays NHWC:
    torch's CUDA kernel hands back NCHW-contiguous tensors (forward and gradient) for channels_last inputs, i.e.
    a layout-conversion copy on each side of every pad.  The op is linear, so backward and double backward are
    the adjoint gather and this op again."""

    @staticmethod
    def forward(ctx, x, pad):
        require_cuda(x)
        x = nhwc(x)
        n, c, h, w = x.shape
        out = empty_nhwc(n, c, h + 2 * pad, w + 2 * pad, x)
        _lib.call("ideas_reflect_pad2d", ptr(out), ptr(x), n, h, w, c, int(pad), 0, stream_ptr(x))
        ctx.pad = pad
        return out

    @staticmethod
    def backward(ctx, g):
        return ReflectPadBackward.apply(g, ctx.pad), None


class ReflectPadBackward(Function):
    @staticmethod
    def forward(ctx, g, pad):
        require_cuda(g)
        g = nhwc(g)
        n, c, oh, ow = g.shape
        h, w = oh - 2 * pad, ow - 2 * pad
        out = empty_nhwc(n, c, h, w, g)
        _lib.call("ideas_reflect_pad2d", ptr(out), ptr(g), n, h, w, c, int(pad), 1, stream_ptr(g))
        ctx.pad = pad
        return out

    @staticmethod
    def backward(ctx, gg):
        return ReflectPad.apply(gg, ctx.pad), None


def reflect_pad(x, pad):
    return ReflectPad.apply(x, int(pad))
