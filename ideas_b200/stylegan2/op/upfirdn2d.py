"""upfirdn2d on the B200 C ABI.

Public surface of the reference's stylegan2/op/upfirdn2d.py: ``upfirdn2d(input, kernel,
up=1, down=1, pad=(0, 0))`` (:145-157) with first- and second-order autograd (:19-142).
The gradient of an upfirdn2d is another upfirdn2d with up and down swapped, the kernel
flipped and the padding ``g_pad`` derived below; the double-backward is the forward op.
Data is NHWC: the native view (major, H, W, minor) is (N, H, W, C).  ``blur_bias_act`` is
the fused Blur -> FusedLeakyReLU epilogue variant used by upsampling styled convs.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from ... import _lib
from ..._tensor import empty_nhwc, nhwc, ptr, require_cuda, stream_ptr
from .fused_act import FusedLeakyReLUFunctionBackward


def _out_size(n, up, down, p0, p1, k):
    return (n * up + p0 + p1 - k) // down + 1


def residual_ok(x, kernel, up, down) -> bool:
    """Whether ideas_upfirdn2d_res can merge a residual for these parameters (its fused up-sampling path)."""
    return tuple(up) == (2, 2) and tuple(down) == (1, 1) and kernel.shape[0] <= 4 and kernel.shape[1] <= 4 and \
        x.shape[1] % 4 == 0 and x.shape[0] <= 65535


def flipped(kernel):
    """torch.flip(kernel, [0, 1]) -- the taps of the adjoint -- once per training iteration and stream instead of one
    tiny launch per call (the per-step cache lives in op/conv.py, which imports this module)."""
    from .conv import cached
    return cached(kernel, "flip", lambda: torch.flip(kernel, [0, 1]), immutable=True)


def _run(x, kernel, up, down, pad, bias=None, alpha=0.2, gain=1.0, residual=None, res_scale=1.0):
    """x: NHWC-dense (N,C,H,W) tensor.  pad = (x0, x1, y0, y1)."""
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    oh = _out_size(h, up[1], down[1], pad[2], pad[3], kh)
    ow = _out_size(w, up[0], down[0], pad[0], pad[1], kw)
    if n > 0 and (oh < 1 or ow < 1):
        raise RuntimeError(f"upfirdn2d: non-positive output size {oh}x{ow}")
    out = empty_nhwc(n, c, max(oh, 0), max(ow, 0), x)
    if residual is not None and tuple(residual.shape) != tuple(out.shape):
        raise RuntimeError(f"upfirdn2d: residual shape {tuple(residual.shape)} != output shape {tuple(out.shape)}")
    _lib.call("ideas_upfirdn2d_res", ptr(out), ptr(x), ptr(kernel), n, h, w, c, kh, kw, up[0], up[1], down[0], down[1],
              pad[0], pad[1], pad[2], pad[3], ptr(bias), float(alpha), float(gain), ptr(residual), float(res_scale),
              stream_ptr(x))
    return out


def _grad_pad(in_hw, out_hw, k_hw, up, down, pad):
    """Padding of the adjoint operator (same algebra as reference upfirdn2d.py:111-116)."""
    (in_h, in_w), (out_h, out_w), (kh, kw) = in_hw, out_hw, k_hw
    gx0 = kw - pad[0] - 1
    gy0 = kh - pad[2] - 1
    gx1 = in_w * up[0] - out_w * down[0] + pad[0] - up[0] + 1
    gy1 = in_h * up[1] - out_h * down[1] + pad[2] - up[1] + 1
    return (gx0, gx1, gy0, gy1)


class UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        ctx.save_for_backward(kernel)
        ctx.cfg = (up, down, pad, in_size, out_size)
        gx = _run(nhwc(grad_output), grad_kernel, down, up, g_pad)
        assert gx.shape[2:] == tuple(in_size[2:]), (gx.shape, in_size)
        return gx

    @staticmethod
    def backward(ctx, gradgrad_input):
        (kernel,) = ctx.saved_tensors
        up, down, pad, in_size, out_size = ctx.cfg
        gg = _run(nhwc(gradgrad_input), kernel, up, down, pad)
        return gg, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    """out = upfirdn2d(input); with ``residual`` (out-shaped): out = (upfirdn2d(input) + residual) * res_scale."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad, residual=None, res_scale=1.0):
        require_cuda(input, kernel, residual)
        kernel = kernel.contiguous()
        x = nhwc(input)
        out = _run(x, kernel, up, down, pad, residual=nhwc(residual) if residual is not None else None, res_scale=res_scale)
        ctx.save_for_backward(kernel, flipped(kernel))
        ctx.cfg = (up, down, pad, tuple(input.shape), (out.shape[2], out.shape[3]))
        ctx.res_scale = res_scale if residual is not None else None
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        up, down, pad, in_size, out_size = ctx.cfg
        gres = None
        if ctx.res_scale is not None:
            if ctx.res_scale != 1.0:                   # folded residual blocks pass 1.0: no scaling pass
                grad_output = grad_output * ctx.res_scale
            gres = grad_output if ctx.needs_input_grad[5] else None
        g_pad = _grad_pad(in_size[2:], out_size, kernel.shape, up, down, pad)
        gx = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size)
        return gx, None, None, None, None, gres, None


class SplitDown(Function):
    """x -> (x, upfirdn2d(x, kernel, down=2, pad)): the two uses of a down-sampling residual block's input -- its
    main path and the Blur of its skip path (models.py:213-227) -- as ONE autograd node.  Forward is the fused
    down-sampling kernel it always was.  The point is the backward: the node receives both gradients at once, so
    the adjoint of the blur (the fused up-sampling kernel) adds the main path's gradient in its epilogue instead of
    autograd summing the two in a separate pass over the full-resolution tensor."""

    @staticmethod
    def forward(ctx, x, kernel, pad):
        require_cuda(x, kernel)
        kernel = kernel.contiguous()
        x = nhwc(x)
        low = _run(x, kernel, (1, 1), (2, 2), pad)
        ctx.save_for_backward(kernel, flipped(kernel))
        ctx.cfg = (pad, tuple(x.shape), (low.shape[2], low.shape[3]))
        ctx.set_materialize_grads(False)
        return x.view_as(x), low

    @staticmethod
    def backward(ctx, gx, glow):
        kernel, gkernel = ctx.saved_tensors
        pad, in_size, out_size = ctx.cfg
        if glow is None:
            return gx, None, None
        g_pad = _grad_pad(in_size[2:], out_size, kernel.shape, (1, 1), (2, 2), pad)
        if gx is None or torch.is_grad_enabled() or not residual_ok(glow, gkernel, (2, 2), (1, 1)):
            g = UpFirDn2dBackward.apply(glow, kernel, gkernel, (1, 1), (2, 2), pad, g_pad, in_size, out_size)
            return (g if gx is None else g + gx), None, None
        out = _run(nhwc(glow), gkernel, (2, 2), (1, 1), g_pad, residual=nhwc(gx), res_scale=1.0)
        assert tuple(out.shape) == tuple(in_size), (out.shape, in_size)
        return out, None, None


class BlurBiasAct(Function):
    """out = scale * lrelu(upfirdn2d(x, kernel, pad) + bias[c]) in one kernel (up = down = 1).
    Equivalent to Blur followed by FusedLeakyReLU (stylegan2/model.py:261,375)."""

    @staticmethod
    def forward(ctx, input, kernel, pad, bias, negative_slope, scale):
        require_cuda(input, kernel, bias)
        kernel = kernel.contiguous()
        x = nhwc(input)
        out = _run(x, kernel, (1, 1), (1, 1), pad, bias=bias, alpha=negative_slope, gain=scale)
        ctx.save_for_backward(kernel, flipped(kernel), out)
        ctx.cfg = (pad, tuple(input.shape), (out.shape[2], out.shape[3]), negative_slope, scale)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel, out = ctx.saved_tensors
        pad, in_size, out_size, slope, scale = ctx.cfg
        g1, gb = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.needs_input_grad[3], slope, scale)
        gx = None
        if ctx.needs_input_grad[0]:
            g_pad = _grad_pad(in_size[2:], out_size, kernel.shape, (1, 1), (1, 1), pad)
            gx = UpFirDn2dBackward.apply(g1, kernel, grad_kernel, (1, 1), (1, 1), pad, g_pad, in_size, out_size)
        return gx, None, None, (gb if ctx.needs_input_grad[3] else None), None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0), residual=None, res_scale=1.0):
    """Reference signature (upfirdn2d.py:145) plus an optional residual merged into the result (see UpFirDn2d)."""
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]), residual, res_scale)


def blur_bias_act(input, kernel, pad, bias, negative_slope=0.2, scale=2 ** 0.5):
    return BlurBiasAct.apply(input, kernel, (pad[0], pad[1], pad[0], pad[1]), bias, negative_slope, scale)
