"""The seven IDEAS networks on the B200 layer library.

Drop-in for the reference's models.py: ``init_model(name, args)`` (:468-513) and the classes
it builds keep their names, constructor arguments, sub-module names/indices and therefore
their state_dict keys (pinned by tests against tests/golden/contract.json),
and parameters are drawn in the same order so a seeded construction reproduces the
reference's initialisation bit for bit.

What differs is underneath: ConvLayer runs a peephole pass over its children so that
conv + FusedLeakyReLU execute as one kernel with a fused epilogue; residual merges are one
kernel; all activations are NHWC; every convolution, blur and activation is hand-written
CUDA behind the C ABI (no cuDNN, no per-sample weights).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .stylegan2.model import (Blur, EqualConv2d, EqualLinear, ScaledLeakyReLU,
                              StyledConv_without_noise as StyledConv)
from .stylegan2.op import FusedLeakyReLU, upfirdn2d
from .stylegan2.op.upfirdn2d import SplitDown, residual_ok
from .stylegan2.op import conv as _ops
from .stylegan2.op.conv import cached, packed_weight
from .stylegan2.op.linear import matmul_nt
from .stylegan2.op.elementwise import add_scale, reflect_pad

_INV_SQRT2 = 1.0 / math.sqrt(2.0)
_AB_NO_PREMOD = __import__("os").environ.get("IDEAS_AB_NO_PREMOD", "0") == "1"      # measurement-only A/B switch


class EqualConvTranspose2d(nn.Module):
    """Equalised-lr transposed convolution; weight is (in, out, k, k) as in models.py:17-19."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(in_channel, out_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def derived_weights(self):
        return packed_weight(self.weight, False, self.scale)

    def forward(self, input, stride=None, weight_mul: float = 1.0):
        cin, cout, k, _ = self.weight.shape
        stride = self.stride if stride is None else stride
        if weight_mul != 1.0 and self.bias is not None:
            raise RuntimeError("EqualConvTranspose2d: weight_mul would not scale the bias")
        # (in, out, k, k) is the OIHW weight of the stride-s conv whose adjoint this layer is
        wp = packed_weight(self.weight, False, self.scale * weight_mul)
        out = _ops.conv_transpose2d(input, wp, C_out=cout, kh=k, kw=k, stride=stride, pad=self.padding)
        if self.bias is not None:
            out = out + self.bias.view(1, -1, 1, 1)
        return out

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.weight.shape[0]}, {self.weight.shape[1]},"
                f" {self.weight.shape[2]}, stride={self.stride}, padding={self.padding})")


def _blur_pads(blur_kernel, kernel_size, up):
    """Blur padding around a stride-2 (transposed) conv, reference models.py:68-74,90-95."""
    factor = 2
    if up:
        p = (len(blur_kernel) - factor) - (kernel_size - 1)
        return ((p + 1) // 2 + factor - 1, p // 2 + 1)
    p = (len(blur_kernel) - factor) + (kernel_size - 1)
    return ((p + 1) // 2, p // 2)


class ReflectionPad2d(nn.Module):
    """nn.ReflectionPad2d on the NHWC kernel (parameter-free: no state_dict entry either way)."""

    def __init__(self, padding):
        super().__init__()
        self.padding = int(padding)

    def forward(self, input):
        return reflect_pad(input, self.padding)

    def extra_repr(self):
        return str(self.padding)


class ConvLayer(nn.Sequential):
    """[Blur] conv | convT Blur | [ReflectionPad] conv, then FusedLeakyReLU / ScaledLeakyReLU / Tanh.
    Child order and indices follow models.py:49-134 exactly (they are the state_dict keys)."""

    def __init__(self, in_channel, out_channel, kernel_size, upsample=False, downsample=False,
                 blur_kernel=(1, 3, 3, 1), bias=True, activate=True, padding="zero", tanh=False):
        stages = []
        conv_bias = bias and not activate
        if downsample:
            stages.append(Blur(blur_kernel, pad=_blur_pads(blur_kernel, kernel_size, up=False)))
        if upsample:
            stages.append(EqualConvTranspose2d(in_channel, out_channel, kernel_size, padding=0, stride=2,
                                               bias=conv_bias))
            stages.append(Blur(blur_kernel, pad=_blur_pads(blur_kernel, kernel_size, up=True)))
            conv_pad = 0
        else:
            conv_pad = 0
            if not downsample:
                half = (kernel_size - 1) // 2
                if padding == "zero":
                    conv_pad = half
                elif padding == "reflect":
                    if half > 0:
                        stages.append(ReflectionPad2d(half))
                elif padding != "valid":
                    raise ValueError('Padding should be "zero", "reflect", or "valid"')
            stages.append(EqualConv2d(in_channel, out_channel, kernel_size, padding=conv_pad,
                                      stride=2 if downsample else 1, bias=conv_bias))
        if activate:
            if tanh:
                stages.append(nn.Tanh())
            elif bias:
                stages.append(FusedLeakyReLU(out_channel))
            else:
                stages.append(ScaledLeakyReLU(0.2))
        super().__init__(*stages)
        self.padding = conv_pad

    def forward(self, input, start=0, residual=None, res_scale=1.0, gain_mul=1.0, weight_mul=1.0, pre_low=None):
        """``start``: index of the first child to run (ResBlock fuses conv1's tail with conv2's leading Blur).
        ``residual``: the layer returns (layer(input) + residual) * res_scale.  When the layer ends in an
        activation-free convolution or in the blur of an up-sampling skip -- every skip path of the residual blocks
        does -- the merge happens in that kernel's epilogue; otherwise it is one add_scale pass.
        ``gain_mul`` / ``weight_mul``: factor folded into the layer's fused activation gain / its bias-free
        convolution's weights (``scalable()`` tells whether the layer's structure allows it)."""
        if (gain_mul != 1.0 or weight_mul != 1.0) and not self.scalable(start):
            raise RuntimeError("ConvLayer: this layer cannot absorb an output scale")
        mods = list(self)
        i, out = start, input
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            last2, last1 = i + 2 >= len(mods), i + 1 >= len(mods)
            if isinstance(m, EqualConv2d) and isinstance(nxt, FusedLeakyReLU):
                out = m(out, activation=nxt, gain_mul=gain_mul)          # one kernel: conv + bias + leaky ReLU
                i += 2
            elif (isinstance(m, Blur) and isinstance(nxt, EqualConv2d) and nxt.weight.shape[2] == 1
                  and nxt.stride == 2 and nxt.padding == 0):
                # Blur -> 1x1 stride-2 conv (down-sampling skip): the conv reads every other blurred pixel, so
                # blur and decimate in one pass (upfirdn2d down=2) and run the 1x1 conv at the low resolution
                # ``pre_low``: the caller already ran this blur + decimation on the layer's input (ResBlock, SplitDown)
                low = pre_low if (pre_low is not None and i == start) else upfirdn2d(out, m.kernel, down=2, pad=m.pad)
                if residual is not None and last2:
                    out, residual = nxt(low, stride=1, residual=residual, res_scale=res_scale, weight_mul=weight_mul), None
                else:
                    out = nxt(low, stride=1, weight_mul=weight_mul)
                i += 2
            elif (isinstance(m, EqualConvTranspose2d) and isinstance(nxt, Blur) and m.weight.shape[2] == 1
                  and m.stride == 2 and m.padding == 0 and m.bias is None):
                # 1x1 stride-2 transposed conv -> Blur (up-sampling skip): a bias-free 1x1 conv commutes with zero
                # insertion, so convolve at the low resolution and let upfirdn2d(up=2) interleave the zeros on the
                # fly (its trailing zero replaces one unit of right padding)
                low = m(out, stride=1, weight_mul=weight_mul)
                pad = (nxt.pad[0], nxt.pad[1] - 1)
                if residual is not None and last2 and residual_ok(low, nxt.kernel, (2, 2), (1, 1)):
                    out, residual = upfirdn2d(low, nxt.kernel, up=2, pad=pad, residual=residual, res_scale=res_scale), None
                else:
                    out = upfirdn2d(low, nxt.kernel, up=2, pad=pad)
                i += 2
            elif isinstance(m, EqualConv2d) and last1 and (residual is not None or weight_mul != 1.0):
                out, residual = m(out, residual=residual, res_scale=res_scale, weight_mul=weight_mul), None
                i += 1
            else:
                out = m(out)
                i += 1
        if residual is not None:
            out = add_scale(out, residual, res_scale)
        return out


def _conv_layer_scalable(layer, start=0):
    """True when ``ConvLayer.forward(start=...)`` can fold an output factor into its kernels: the layer is
    [Blur] -> conv -> FusedLeakyReLU (factor goes into the activation gain) or a bias-free, activation-free
    [Blur ->] 1x1 conv / 1x1 transposed conv -> Blur (factor goes into the weights)."""
    mods = list(layer)[start:]
    convs = [m for m in mods if isinstance(m, (EqualConv2d, EqualConvTranspose2d))]
    if len(convs) != 1:
        return False
    rest = [m for m in mods if not isinstance(m, (EqualConv2d, EqualConvTranspose2d, Blur, ReflectionPad2d))]
    if not rest:
        return convs[0].bias is None
    return len(rest) == 1 and isinstance(rest[0], FusedLeakyReLU) and mods[-1] is rest[0] and convs[0].bias is None \
        and isinstance(convs[0], EqualConv2d)


ConvLayer.scalable = _conv_layer_scalable


class StyledResBlock(nn.Module):
    def __init__(self, in_channel, out_channel, style_dim, upsample, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.conv1 = StyledConv(in_channel, out_channel, 3, style_dim, upsample=upsample, blur_kernel=blur_kernel)
        self.conv2 = StyledConv(out_channel, out_channel, 3, style_dim)
        self.skip = None
        if upsample or in_channel != out_channel:
            self.skip = ConvLayer(in_channel, out_channel, 1, upsample=upsample, blur_kernel=blur_kernel,
                                  bias=False, activate=False)

    def forward(self, input, style, noise=None, modulation=None):
        """``modulation``: optional (s1, s2), the outputs of conv1 / conv2's modulation linears."""
        m1, m2 = modulation if modulation is not None else (None, None)
        if input.is_cuda and not _AB_NO_PREMOD:
            # conv1 hands conv2 its input already multiplied by conv2's style (s2 * a): conv2 runs no modulation
            # pass, and its style gradient comes out of conv1's activation backward for free (op/conv.py, ``post``)
            if m2 is None:
                m2 = self.conv2.conv.modulation(style)
            fold = self.skip is not None and self.skip.scalable()
            _, a_mod = self.conv1(input, style, noise, modulation=m1, post_modulation=m2)
            # (out + skip)/sqrt(2): the factor goes into conv2's activation gain (sqrt(2) * 1/sqrt(2)) and into the
            # skip convolution's weights, so the merge is a plain sum in the skip path's epilogue and the backward
            # pass has no scaling pass at all
            out = self.conv2(a_mod, style, noise, modulation=m2, premodulated=True, gain_mul=_INV_SQRT2 if fold else 1.0)
            if fold:
                return self.skip(input, residual=out, res_scale=1.0, weight_mul=_INV_SQRT2)
        else:
            out = self.conv2(self.conv1(input, style, noise, modulation=m1), style, noise, modulation=m2)
        if self.skip is None:
            return add_scale(out, input, _INV_SQRT2)
        return self.skip(input, residual=out, res_scale=_INV_SQRT2)      # (out + skip)/sqrt(2) in the skip's epilogue


class ResBlock(nn.Module):
    def __init__(self, in_channel, out_channel, downsample, padding="zero", blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, out_channel, 3, padding=padding)
        self.conv2 = ConvLayer(out_channel, out_channel, 3, downsample=downsample, padding=padding,
                               blur_kernel=blur_kernel)
        self.skip = None
        if downsample or in_channel != out_channel:
            self.skip = ConvLayer(in_channel, out_channel, 1, downsample=downsample, blur_kernel=blur_kernel,
                                  bias=False, activate=False)

    def _down_skip(self):
        """(blur, conv) when the skip path is Blur -> bias-free 1x1 stride-2 conv, else None."""
        sk = list(self.skip) if self.skip is not None else []
        if (len(sk) == 2 and isinstance(sk[0], Blur) and isinstance(sk[1], EqualConv2d) and sk[1].weight.shape[2] == 1
                and sk[1].stride == 2 and sk[1].padding == 0 and sk[1].bias is None):
            return sk[0], sk[1]
        return None

    def forward(self, input):
        c1, c2 = list(self.conv1), list(self.conv2)
        low = None
        block_input = input
        ds = self._down_skip() if input.is_cuda else None
        if ds is not None and input.requires_grad and input.shape[1] % 4 == 0:
            # the block input feeds the main path and the skip's blur: one node for both, whose backward adds the two
            # gradients inside the blur's adjoint kernel (op/upfirdn2d.py SplitDown)
            p = ds[0].pad
            input, low = SplitDown.apply(input, ds[0].kernel, (p[0], p[1], p[0], p[1]))
        if (input.is_cuda and isinstance(c2[0], Blur) and len(c1) >= 2 and isinstance(c1[-1], FusedLeakyReLU)
                and isinstance(c1[-2], EqualConv2d) and c1[-2].bias is None):
            # conv1 -> FusedLeakyReLU -> Blur as one autograd node: its backward fuses the blur's and the
            # activation's backward passes into one kernel (the blurred gradient is never written)
            h = input
            for m in c1[:-2]:                      # ReflectionPad2d of the encoder blocks
                h = m(h)
            fold = self.skip is not None and self.skip.scalable() and self.conv2.scalable(1)
            out = self.conv2(c1[-2].forward_act_blur(h, c1[-1], c2[0]), start=1, gain_mul=_INV_SQRT2 if fold else 1.0)
        else:
            fold = input.is_cuda and self.skip is not None and self.skip.scalable() and self.conv2.scalable()
            out = self.conv2(self.conv1(input), gain_mul=_INV_SQRT2 if fold else 1.0)
        if fold:        # 1/sqrt(2) lives in conv2's activation gain and the skip's weights: the merge is a plain sum
            return self.skip(block_input, residual=out, res_scale=1.0, weight_mul=_INV_SQRT2, pre_low=low)
        if self.skip is None:
            return add_scale(out, block_input, _INV_SQRT2)
        # (out + skip)/sqrt(2) in the skip's epilogue
        return self.skip(block_input, residual=out, res_scale=_INV_SQRT2, pre_low=low)


class DisentanglementEncoder(nn.Module):
    """image -> (structure code (B, structure_channel, H/16, W/16), texture vector (B, texture_channel))."""

    def __init__(self, channel, structure_channel=8, texture_channel=2048, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        widths = [channel * 2 ** i for i in range(5)]
        stem = [ConvLayer(3, widths[0], 1)]
        stem += [ResBlock(a, b, downsample=True, padding="reflect", blur_kernel=blur_kernel)
                 for a, b in zip(widths[:-1], widths[1:])]
        self.stem = nn.Sequential(*stem)
        top = widths[-1]
        self.structure = nn.Sequential(ConvLayer(top, top, 1, blur_kernel=blur_kernel),
                                       ConvLayer(top, structure_channel, 1, blur_kernel=blur_kernel))
        self.texture = nn.Sequential(
            ConvLayer(top, top * 2, 3, downsample=True, padding="valid", blur_kernel=blur_kernel),
            ConvLayer(top * 2, top * 4, 3, downsample=True, padding="valid", blur_kernel=blur_kernel),
            nn.AdaptiveAvgPool2d(1),
            ConvLayer(top * 4, texture_channel, 1, tanh=True, blur_kernel=blur_kernel))

    def forward(self, input):
        feat = self.stem(input)
        return self.structure(feat), torch.flatten(self.texture(feat), 1)


class Generator(nn.Module):
    """(structure code, texture vector) -> image through 8 styled residual blocks (4 at 1/16
    resolution, 4 upsampling) and a 1x1 to-RGB conv."""

    CH_MULT = (4, 8, 12, 16, 16, 16, 8, 4)
    UPSAMPLE = (False, False, False, False, True, True, True, True)

    def __init__(self, channel, structure_channel=8, texture_channel=2048, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.layers = nn.ModuleList()
        width = structure_channel
        for mult, up in zip(self.CH_MULT, self.UPSAMPLE):
            self.layers.append(StyledResBlock(width, channel * mult, texture_channel, up, blur_kernel))
            width = channel * mult
        self.to_rgb = ConvLayer(width, 3, 1, activate=False)

    def forward(self, structure, texture, noises=None):
        noises = noises if noises is not None else [None] * len(self.layers)
        out = structure
        mods = self.modulations(texture) if texture.is_cuda else [None] * len(self.layers)
        for block, noise, mod in zip(self.layers, noises, mods):
            out = block(out, texture, noise, modulation=mod)
        return self.to_rgb(out)

    def derived_weights(self):
        """Row-concatenated modulation weights and biases of all layers (cached per training iteration)."""
        lins = [conv.conv.modulation for block in self.layers for conv in (block.conv1, block.conv2)]

        def build():
            return (torch.cat([m.effective_weight() for m in lins], 0),
                    torch.cat([m.bias * m.lr_mul for m in lins], 0))

        return cached(lins[0].weight, "G.modulations", build)

    def modulations(self, texture):
        """The 16 modulation linears of one call (stylegan2/model.py:226,239: every ModulatedConv2d maps the SAME
        texture vector through its own EqualLinear) as ONE GEMM on the row-concatenated weights: s_all =
        texture @ cat(W_l * scale)^T + cat(bias_l), then split per layer."""
        lins = [conv.conv.modulation for block in self.layers for conv in (block.conv1, block.conv2)]
        w_all, b_all = self.derived_weights()
        s_all = matmul_nt(texture, w_all) + b_all
        parts = torch.split(s_all, [m.weight.shape[0] for m in lins], dim=1)
        return [(parts[2 * i], parts[2 * i + 1]) for i in range(len(self.layers))]


def _same_res_stack(c_in, widths, c_out, blur_kernel):
    layers = [ConvLayer(c_in, widths[0], 1, blur_kernel=blur_kernel)]
    layers += [ResBlock(a, b, downsample=False, padding="reflect", blur_kernel=blur_kernel)
               for a, b in zip(widths[:-1], widths[1:])]
    layers.append(ConvLayer(widths[-1], c_out, 1, blur_kernel=blur_kernel))
    return nn.Sequential(*layers)


class StructureGenerator(nn.Module):
    """secret tensor Z (B, N, h, w) -> structure code (B, structure_channel, h, w)."""

    def __init__(self, channel, N=1, structure_channel=8, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.structure = _same_res_stack(N, (channel, channel * 2, channel * 4, channel * 2), structure_channel,
                                         blur_kernel)

    def forward(self, noise):
        return self.structure(noise)


class ImageLevelDiscriminator(nn.Module):
    def __init__(self, size, channel_multiplier=1, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        m = channel_multiplier
        channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * m, 128: 128 * m, 256: 64 * m, 512: 32 * m,
                    1024: 16 * m}
        width = channels[size]
        convs = [ConvLayer(3, width, 1, blur_kernel=blur_kernel)]
        for level in range(int(math.log(size, 2)), 2, -1):
            nxt = channels[2 ** (level - 1)]
            convs.append(ResBlock(width, nxt, downsample=True, blur_kernel=blur_kernel))
            width = nxt
        self.convs = nn.Sequential(*convs)
        self.final_conv = ConvLayer(width, channels[4], 3, blur_kernel=blur_kernel)
        self.final_linear = nn.Sequential(EqualLinear(channels[4] * 4 * 4, channels[4], activation="fused_lrelu"),
                                          EqualLinear(channels[4], 1))

    def forward(self, input):
        out = self.final_conv(self.convs(input))
        # flatten in the logical (C, H, W) order the linear layer's weights expect
        return self.final_linear(out.reshape(out.shape[0], -1))


class CooccurenceDiscriminator(nn.Module):
    CH_MULT = (2, 4, 8, 12, 12, 24)
    DOWNSAMPLE = (True, True, True, True, True, False)

    def __init__(self, channel, size=256):
        super().__init__()
        encoder = [ConvLayer(3, channel, 1)]
        width = channel
        for mult, down in zip(self.CH_MULT, self.DOWNSAMPLE):
            encoder.append(ResBlock(width, channel * mult, down))
            width = channel * mult
        k_size, feat_size = (3, 2 * 2) if size > 511 else (2, 1 * 1)
        encoder.append(ConvLayer(width, channel * 12, k_size, padding="valid"))
        self.encoder = nn.Sequential(*encoder)
        self.linear = nn.Sequential(
            EqualLinear(channel * 12 * 2 * feat_size, channel * 32, activation="fused_lrelu"),
            EqualLinear(channel * 32, channel * 32, activation="fused_lrelu"),
            EqualLinear(channel * 32, channel * 16, activation="fused_lrelu"),
            EqualLinear(channel * 16, 1))

    def reference_code(self, reference, ref_batch):
        """The ``ref_input`` that forward() derives from the reference patches (models.py:417-420): their encoder
        features averaged over the ``ref_batch`` patches of each image.  Exposed so a caller can evaluate it early,
        on another stream, and pass it back as ``ref_input``."""
        ref = self.encoder(reference)
        _, c, h, w = ref.shape
        return ref.reshape(-1, ref_batch, c, h, w).mean(1)

    def forward(self, input, reference=None, ref_batch=None, ref_input=None):
        feat = self.encoder(input)
        if ref_input is None:
            ref_input = self.reference_code(reference, ref_batch)
        out = torch.flatten(torch.cat((feat, ref_input), 1), 1)
        return self.linear(out), ref_input


class DistributionDiscriminator(nn.Module):
    def __init__(self, texture_channel=2048):
        super().__init__()
        t = texture_channel
        dims = (t, t // 4, t // 16, t // 64, 1)
        self.model = nn.Sequential(*[EqualLinear(a, b, activation="fused_lrelu") for a, b in zip(dims[:-1], dims[1:])])

    def forward(self, input):
        return self.model(input)


class TensorExtractor(nn.Module):
    """recovered structure code -> secret tensor estimate (B, N, h, w)."""

    def __init__(self, channel, N=1, structure_channel=8, blur_kernel=(1, 3, 3, 1)):
        super().__init__()
        self.extract = _same_res_stack(structure_channel, (channel * 2, channel * 4, channel * 2, channel), N,
                                       blur_kernel)

    def forward(self, input):
        return self.extract(input)


_FACTORIES = {
    "DisentanglementEncoder": lambda a: DisentanglementEncoder(channel=a.channel, structure_channel=a.structure_channel,
                                                               texture_channel=a.texture_channel,
                                                               blur_kernel=a.blur_kernel),
    "Generator": lambda a: Generator(channel=a.channel, structure_channel=a.structure_channel,
                                     texture_channel=a.texture_channel, blur_kernel=a.blur_kernel),
    "StructureGenerator": lambda a: StructureGenerator(channel=a.channel, N=a.N, structure_channel=a.structure_channel,
                                                       blur_kernel=a.blur_kernel),
    "ImageLevelDiscriminator": lambda a: ImageLevelDiscriminator(size=a.image_size,
                                                                 channel_multiplier=a.channel_multiplier,
                                                                 blur_kernel=a.blur_kernel),
    "CooccurenceDiscriminator": lambda a: CooccurenceDiscriminator(channel=a.channel, size=a.image_size),
    "DistributionDiscriminator": lambda a: DistributionDiscriminator(texture_channel=a.texture_channel),
    "TensorExtractor": lambda a: TensorExtractor(channel=a.channel, N=a.N, structure_channel=a.structure_channel,
                                                 blur_kernel=a.blur_kernel),
}


def init_model(model, args):
    """Same contract as the reference's ``init_model`` (models.py:468-513): build one network by
    name from the fields of ``args`` (train.py:341-370)."""
    try:
        factory = _FACTORIES[model]
    except KeyError:
        raise NotImplementedError(model) from None
    return factory(args)
