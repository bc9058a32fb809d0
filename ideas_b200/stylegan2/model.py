"""StyleGAN2 building blocks of IDEAS on the B200 kernels.

Module surface of the reference's stylegan2/model.py:14-399 -- same class names,
constructor signatures, attribute names and state_dict keys, so
reference checkpoints load and ``models.py`` can be built on top unchanged in spirit.
Every forward runs on the hand-written CUDA ops of ``ideas_b200.stylegan2.op``:

  * EqualConv2d            -> ConvFwd on packed weights, bias / FusedLeakyReLU fused as epilogue
  * ModulatedConv2d        -> act(d * conv(s * x) + b): style ``s`` multiplies the activation,
                              demodulation ``d = rsqrt(scale^2 * sum_i s_i^2 * sum_k W_oik^2 + eps)``
                              is a small (B,Cin)x(Cin,Cout) product; no per-sample weights
                              (reference: stylegan2/model.py:239-275 builds B*Cout*Cin*k*k of them)
  * Blur/Upsample/Downsample -> upfirdn2d kernel on NHWC data
Activations are NHWC in memory (torch channels_last); logical shapes stay (N, C, H, W).
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn import functional as F

from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d
from .op import conv as _ops
from .op.conv import Geom, ModConv, ModConvUp, PackWeight, cached, packed_weight
from .op.linear import matmul_nt
from .._tensor import nhwc


import os as _os

_AB_CUBLAS_LINEAR = _os.environ.get("IDEAS_AB_CUBLAS_LINEAR", "0") == "1"


class PixelNorm(nn.Module):
    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    """1-D taps -> normalised 2-D FIR kernel (outer product), reference model.py:22-30."""
    k = torch.as_tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def _resample_pads(klen, factor, up):
    p = klen - factor
    return ((p + 1) // 2 + factor - 1, p // 2) if up else ((p + 1) // 2, p // 2)


class Upsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel) * (factor ** 2))
        self.pad = _resample_pads(self.kernel.shape[0], factor, True)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel))
        self.pad = _resample_pads(self.kernel.shape[0], factor, False)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        k = make_kernel(kernel)
        if upsample_factor > 1:
            k = k * (upsample_factor ** 2)
        self.register_buffer("kernel", k)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def derived_weights(self):
        return packed_weight(self.weight, False, self.scale)

    def forward(self, input, activation: FusedLeakyReLU | None = None, stride: int | None = None, residual=None,
                res_scale: float = 1.0, gain_mul: float = 1.0, weight_mul: float = 1.0):
        """``activation``: a FusedLeakyReLU module to fuse into the conv epilogue.  ``stride`` overrides the
        module's stride (ConvLayer folds the decimation of a 1x1 stride-2 conv into the preceding blur).
        ``residual``: merged in the epilogue, (conv(x) + bias + residual) * res_scale (activation-free only).
        ``gain_mul`` multiplies the fused activation's gain, ``weight_mul`` the (bias-free) convolution's weights:
        a residual block folds its 1/sqrt(2) into them (models.py), so the merge itself is a plain sum."""
        k = self.weight.shape[2]
        stride = self.stride if stride is None else stride
        if weight_mul != 1.0 and self.bias is not None:
            raise RuntimeError("EqualConv2d: weight_mul would not scale the bias")
        wp = packed_weight(self.weight, False, self.scale * weight_mul)
        if activation is not None:
            if self.bias is not None:
                raise RuntimeError("EqualConv2d: a fused activation brings its own bias")
            if residual is not None:
                raise RuntimeError("EqualConv2d: residual merge is available without a fused activation only")
            return _ops.conv2d(input, wp, activation.bias, K=self.weight.shape[0], kh=k, kw=k, stride=stride,
                               pad=self.padding, act=True, alpha=activation.negative_slope,
                               gain=activation.scale * gain_mul)
        return _ops.conv2d(input, wp, self.bias, K=self.weight.shape[0], kh=k, kw=k, stride=stride,
                           pad=self.padding, residual=residual, res_scale=res_scale)

    def forward_act_blur(self, input, activation: FusedLeakyReLU, blur: "Blur"):
        """Blur(activation(conv(input))) with the fused backward of ``ConvActBlur``."""
        if self.bias is not None:
            raise RuntimeError("EqualConv2d: a fused activation brings its own bias")
        k = self.weight.shape[2]
        wp = packed_weight(self.weight, False, self.scale)
        g = Geom.forward(input.shape, self.weight.shape[0], k, k, self.stride, self.padding)
        return _ops.ConvActBlur.apply(input, wp, activation.bias, g, activation.negative_slope, activation.scale,
                                      blur.kernel, tuple(blur.pad))

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]},"
                f" {self.weight.shape[2]}, stride={self.stride}, padding={self.padding})")


class EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def derived_weights(self):
        return self.effective_weight()

    def effective_weight(self):
        """weight * scale as a temporary (shared by the uses of one training iteration): the leaf itself is never
        saved for backward, which keeps train.py:209-216's second backward after optimizer.step() legal."""
        return cached(self.weight, "eqlin", lambda: self.weight * self.scale)

    def forward(self, input):
        lead = input.shape[:-1]
        if _AB_CUBLAS_LINEAR:           # measurement-only A/B switch (IDEAS_AB_CUBLAS_LINEAR=1): library GEMM instead of ours
            out = F.linear(input, self.effective_weight())
            if self.activation:
                return fused_leaky_relu(out, self.bias * self.lr_mul)
            return out if self.bias is None else out + self.bias * self.lr_mul
        out = matmul_nt(input.reshape(-1, input.shape[-1]), self.effective_weight())   # hand-written fp32 GEMM
        if self.activation:
            return fused_leaky_relu(out, self.bias * self.lr_mul).reshape(*lead, -1)
        if self.bias is not None:
            out = out + self.bias * self.lr_mul
        return out.reshape(*lead, -1)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class ScaledLeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return fused_leaky_relu(input, None, self.negative_slope, math.sqrt(2))


class ModulatedConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample}, downsample={self.downsample})")

    def derived_weights(self):
        """(packed scaled weight, sum_k (scale*W)^2): functions of the parameter alone, shared by every call of one
        training iteration (op/conv.py step cache)."""
        def build():
            ws = self.weight[0] * self.scale                         # temp, never the leaf (see op/conv.py)
            wsq = ws.pow(2).sum(dim=(2, 3)) if self.demodulate else None          # (Cout, Cin)
            # upsampling branch: (taps, Cin, Cout), the packed weight of the stride-2 conv it is the adjoint of
            return PackWeight.apply(ws, bool(self.upsample), 1.0), wsq

        return cached(self.weight, "modconv", build)

    def forward(self, input, style, activation: FusedLeakyReLU | None = None, modulation: torch.Tensor | None = None,
                premodulated: bool = False, post_modulation: torch.Tensor | None = None, gain_mul: float = 1.0):
        """Returns the modulated convolution; with ``activation`` the FusedLeakyReLU (bias,
        slope, gain) is applied inside the same kernels (StyledConv passes its own).  ``modulation``: the
        per-sample channel scales s = self.modulation(style), when the caller has already computed them (the
        Generator evaluates the modulation linears of all its layers in one GEMM).
        ``premodulated``: ``input`` is already s * x (written by the producing layer, below).
        ``post_modulation`` (B, Cout): also return out * post_modulation, the next layer's pre-modulated input --
        the result is then the pair (out, out_modulated)."""
        k = self.kernel_size
        if input.is_cuda:
            # layout conversion OUTSIDE the autograd nodes below: they save their inputs for the double backward
            # (path-length regularisation), which must stay connected to the caller's graph
            input = nhwc(input)
        s = (self.modulation(style) if modulation is None else modulation).contiguous()   # (B, Cin)
        wp, wsq = self.derived_weights()
        d = None
        if self.demodulate:
            # (B, Cout); the (s^2) . sum_k w^2 product on the repo's own fp32 GEMM like every other linear map here
            d = torch.rsqrt(matmul_nt(s * s, wsq) + self.eps)
        act = activation is not None
        bias = activation.bias if act else None
        alpha = activation.negative_slope if act else 0.2
        gain = activation.scale * gain_mul if act else 1.0          # gain_mul: see EqualConv2d.forward
        if gain_mul != 1.0 and not act:
            raise RuntimeError("ModulatedConv2d: gain_mul needs the fused activation")
        if self.upsample:
            if premodulated:
                raise RuntimeError("ModulatedConv2d: the up-sampling branch modulates its own input")
            g = Geom.transposed(input.shape, self.out_channel, k, k, 2, 0)
            pad = self.blur.pad
            return ModConvUp.apply(input, s, d, wp, self.blur.kernel, (pad[0], pad[1], pad[0], pad[1]), bias, g,
                                   act, alpha, gain, post_modulation)
        if self.downsample:
            if premodulated:
                raise RuntimeError("ModulatedConv2d: the down-sampling branch modulates its own input")
            input = self.blur(input)
            g = Geom.forward(input.shape, self.out_channel, k, k, 2, 0)
        else:
            g = Geom.forward(input.shape, self.out_channel, k, k, 1, self.padding)
        return ModConv.apply(input, s, d, wp, bias, g, act, alpha, gain, premodulated, post_modulation)


class NoiseInjection(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    """Modulated conv -> noise injection -> FusedLeakyReLU (reference model.py:307-341).
    The noise add sits between conv and activation, so only the activation kernel is shared."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None):
        return self.activate(self.noise(self.conv(input, style), noise=noise))


class StyledConv_without_noise(nn.Module):
    """The variant IDEAS uses (models.py:7): modulated conv + FusedLeakyReLU, ``noise`` accepted
    and ignored (reference model.py:343-377).  Conv, demodulation, bias and activation are fused."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None, modulation=None, premodulated=False, post_modulation=None, gain_mul=1.0):
        return self.conv(input, style, activation=self.activate, modulation=modulation, premodulated=premodulated,
                         post_modulation=post_modulation, gain_mul=gain_mul)


class ToRGB(nn.Module):
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None):
        out = self.conv(input, style) + self.bias
        if skip is not None:
            out = out + self.upsample(skip)
        return out
