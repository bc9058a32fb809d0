"""FusedLeakyReLU on the B200 C ABI.

Public surface of the reference's stylegan2/op/fused_act.py: ``FusedLeakyReLU(channel,
negative_slope=0.2, scale=sqrt(2))`` (:75-83) and ``fused_leaky_relu(input, bias,
negative_slope=0.2, scale=sqrt(2))`` (:86-97), with first- and second-order autograd
(:20-72).  Differences by design: the backward computes the bias gradient in the same
pass as the input gradient (the reference launches a second reduction, :33-38), data is
NHWC so the channel index is the fastest one, and there is no CPU branch.
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.autograd import Function

from ... import _lib
from ..._tensor import is_nhwc_dense, ptr, require_cuda, stream_ptr


def _layout(x: torch.Tensor):
    """-> (tensor whose memory the kernel may index linearly, step_b, channels)."""
    if x.dim() < 2:
        raise RuntimeError("fused_leaky_relu expects at least 2 dimensions (batch, channel, ...)")
    c = x.shape[1]
    if x.dim() == 2:
        return x.contiguous(), 1, c
    if x.dim() == 4 and is_nhwc_dense(x):
        return x, 1, c
    if x.dim() == 4:                      # foreign layout: bring it into the package's NHWC
        return x.contiguous(memory_format=torch.channels_last), 1, c
    x = x.contiguous()                    # (B, C, *rest) row-major: bias varies every prod(rest) elements
    return x, int(math.prod(x.shape[2:])), c


def _bias_act(x, bias, ref, act, grad, alpha, scale):
    x, step_b, c = _layout(x)
    out = torch.empty_like(x)             # preserves NHWC strides
    if ref is not None and ref.stride() != x.stride():
        ref = ref.contiguous(memory_format=torch.channels_last) if x.dim() == 4 and step_b == 1 else ref.contiguous()
    _lib.call("ideas_fused_bias_act", ptr(out), ptr(x), ptr(bias), ptr(ref), act, grad, float(alpha), float(scale),
              x.numel(), step_b, c, stream_ptr(x))
    return out


class FusedLeakyReLUFunctionBackward(Function):
    """grad_output, out -> (grad_input, grad_bias); its own backward is the double-backward
    used by the R1 penalty (reference fused_act.py:20-49)."""

    @staticmethod
    def forward(ctx, grad_output, out, want_bias, negative_slope, scale):
        require_cuda(grad_output, out)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        out_l, step_b, c = _layout(out)
        gy = grad_output
        if gy.stride() != out_l.stride():
            gy = gy.contiguous(memory_format=torch.channels_last) if (gy.dim() == 4 and step_b == 1) else gy.contiguous()
        gx = torch.empty_like(out_l)
        gbias = torch.zeros(c, device=out.device, dtype=out.dtype) if want_bias else None
        _lib.call("ideas_bias_act_backward", ptr(gx), ptr(gbias), ptr(gy), ptr(out_l), float(negative_slope),
                  float(scale), out_l.numel(), step_b, c, stream_ptr(out))
        if gbias is None:
            gbias = gx.new_zeros(0)
            ctx.mark_non_differentiable(gbias)
        return gx, gbias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        (out,) = ctx.saved_tensors
        bias = gradgrad_bias if (gradgrad_bias is not None and gradgrad_bias.numel()) else None
        if gradgrad_input is None:
            gradgrad_input = torch.zeros_like(out)
        gg = _bias_act(gradgrad_input, bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gg, None, None, None, None


class FusedLeakyReLUFunction(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        require_cuda(input, bias)
        out = _bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        want_bias = ctx.has_bias and ctx.needs_input_grad[1]
        gx, gb = FusedLeakyReLUFunctionBackward.apply(grad_output, out, want_bias, ctx.negative_slope, ctx.scale)
        return gx, (gb if want_bias else None), None, None


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    """``scale * leaky_relu(input + bias[c], negative_slope)``; ``bias=None`` gives the
    reference's ScaledLeakyReLU (stylegan2/model.py:169-178) from the same kernel."""
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)

    def extra_repr(self):
        return f"{self.bias.shape[0]}, negative_slope={self.negative_slope}"
