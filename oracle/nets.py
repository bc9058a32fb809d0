"""The seven IDEAS networks as pure functions of a state_dict (TEST INFRASTRUCTURE).

Restates models.py:49-513 of the reference.  Nothing here is an nn.Module: a network
is ``fn(sd, inputs...)`` where ``sd`` maps the reference's state_dict keys (SURVEY.md
App. C) to tensors.  The same file holds the *parameter spec* of every network --
the ordered list of (key, shape, initialiser) -- which is at once the drop-in
checkpoint contract and a seeded initialiser that reproduces the reference's
``init_model`` draw for draw (validated by tests/golden: sha256 of a seeded init).

Layer-index bookkeeping follows how models.py:49-134 assembles a ``ConvLayer``
(an nn.Sequential): [Blur] (downsample) | ConvT, Blur (upsample) | [ReflectionPad2d] ,
then the conv, then the activation module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from . import functional as O

Spec = List[Tuple[str, Tuple[int, ...], str]]
BLUR_TAPS = (1, 3, 3, 1)


# ==========================================================================
# parameter specs (models.py / stylegan2/model.py constructors)
# ==========================================================================
def _conv_layer_spec(prefix, cin, cout, k, *, up=False, down=False, pad="zero", bias=True,
                     activate=True, tanh=False) -> Spec:
    """Keys of one ConvLayer (models.py:49-134).  Returns spec entries in creation order."""
    out: Spec = []
    i = 0
    if down:
        out.append((f"{prefix}.{i}.kernel", (4, 4), "blur1"))
        i += 1
    if up:
        out.append((f"{prefix}.{i}.weight", (cin, cout, k, k), "randn"))       # models.py:17-19
        if bias and not activate:
            out.append((f"{prefix}.{i}.bias", (cout,), "zeros"))
        i += 1
        out.append((f"{prefix}.{i}.kernel", (4, 4), "blur1"))   # models.py:95: no upsample_factor => no x4 gain
        i += 1
    else:
        if not down and pad == "reflect" and (k - 1) // 2 > 0:
            i += 1                                                             # ReflectionPad2d
        out.append((f"{prefix}.{i}.weight", (cout, cin, k, k), "randn"))
        if bias and not activate:
            out.append((f"{prefix}.{i}.bias", (cout,), "zeros"))
        i += 1
    if activate and not tanh and bias:
        out.append((f"{prefix}.{i}.bias", (cout,), "zeros"))                   # FusedLeakyReLU
    return out


def _res_block_spec(prefix, cin, cout, down, pad="zero") -> Spec:
    """models.py:181-227."""
    s = _conv_layer_spec(f"{prefix}.conv1", cin, cout, 3, pad=pad)
    s += _conv_layer_spec(f"{prefix}.conv2", cout, cout, 3, down=down, pad=pad)
    if down or cin != cout:
        s += _conv_layer_spec(f"{prefix}.skip", cin, cout, 1, down=down, bias=False, activate=False)
    return s


def _styled_conv_spec(prefix, cin, cout, style_dim, up) -> Spec:
    """StyledConv_without_noise (stylegan2/model.py:343-377) around ModulatedConv2d (:181-228)."""
    s: Spec = [(f"{prefix}.conv.weight", (1, cout, cin, 3, 3), "randn")]
    if up:
        s.append((f"{prefix}.conv.blur.kernel", (4, 4), "blur4"))
    s.append((f"{prefix}.conv.modulation.weight", (cin, style_dim), "randn"))
    s.append((f"{prefix}.conv.modulation.bias", (cin,), "ones"))
    s.append((f"{prefix}.activate.bias", (cout,), "zeros"))
    return s


def _linear_spec(prefix, cin, cout) -> Spec:
    return [(f"{prefix}.weight", (cout, cin), "randn"), (f"{prefix}.bias", (cout,), "zeros")]


G_CH_MUL = (4, 8, 12, 16, 16, 16, 8, 4)                       # models.py:281
G_UP = (False, False, False, False, True, True, True, True)   # models.py:282
D_CHANNELS = lambda m: {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * m, 128: 128 * m,  # noqa: E731
                        256: 64 * m, 512: 32 * m, 1024: 16 * m}  # models.py:336-346
DCO_CH_MUL = (2, 4, 8, 12, 12, 24)                            # models.py:388
DCO_DOWN = (True, True, True, True, True, False)              # models.py:389


def param_spec(name: str, *, channel=32, structure_channel=8, texture_channel=2048, N=1,
               image_size=256, channel_multiplier=1) -> Spec:
    """Ordered (key, shape, init) list of one network; keyword names follow train.py:341-359."""
    s: Spec = []
    if name == "DisentanglementEncoder":                       # models.py:230-260
        s += _conv_layer_spec("stem.0", 3, channel, 1)
        cin = channel
        for i in range(1, 5):
            ch = channel * 2 ** i
            s += _res_block_spec(f"stem.{i}", cin, ch, True, "reflect")
            cin = ch
        s += _conv_layer_spec("structure.0", cin, cin, 1)
        s += _conv_layer_spec("structure.1", cin, structure_channel, 1)
        s += _conv_layer_spec("texture.0", cin, cin * 2, 3, down=True, pad="valid")
        s += _conv_layer_spec("texture.1", cin * 2, cin * 4, 3, down=True, pad="valid")
        s += _conv_layer_spec("texture.3", cin * 4, texture_channel, 1, tanh=True)
    elif name == "Generator":                                  # models.py:271-294
        cin = structure_channel
        for i, (m, up) in enumerate(zip(G_CH_MUL, G_UP)):
            cout = channel * m
            s += _styled_conv_spec(f"layers.{i}.conv1", cin, cout, texture_channel, up)
            s += _styled_conv_spec(f"layers.{i}.conv2", cout, cout, texture_channel, False)
            if up or cin != cout:
                s += _conv_layer_spec(f"layers.{i}.skip", cin, cout, 1, up=up, bias=False, activate=False)
            cin = cout
        s += _conv_layer_spec("to_rgb", cin, 3, 1, activate=False)
    elif name in ("StructureGenerator", "TensorExtractor"):    # models.py:309-325, 444-460
        if name == "StructureGenerator":
            root, widths, cin, cout = "structure", (channel, channel * 2, channel * 4, channel * 2), N, structure_channel
        else:
            root, widths, cin, cout = "extract", (channel * 2, channel * 4, channel * 2, channel), structure_channel, N
        s += _conv_layer_spec(f"{root}.0", cin, widths[0], 1)
        for i in range(1, 4):
            s += _res_block_spec(f"{root}.{i}", widths[i - 1], widths[i], False, "reflect")
        s += _conv_layer_spec(f"{root}.4", widths[3], cout, 1)
    elif name == "ImageLevelDiscriminator":                    # models.py:332-367
        chans = D_CHANNELS(channel_multiplier)
        cin = chans[image_size]
        s += _conv_layer_spec("convs.0", 3, cin, 1)
        log_size = int(math.log(image_size, 2))
        for j, i in enumerate(range(log_size, 2, -1), start=1):
            cout = chans[2 ** (i - 1)]
            s += _res_block_spec(f"convs.{j}", cin, cout, True)
            cin = cout
        s += _conv_layer_spec("final_conv", cin, chans[4], 3)
        s += _linear_spec("final_linear.0", chans[4] * 16, chans[4])
        s += _linear_spec("final_linear.1", chans[4], 1)
    elif name == "CooccurenceDiscriminator":                   # models.py:379-411
        s += _conv_layer_spec("encoder.0", 3, channel, 1)
        cin = channel
        for j, (m, down) in enumerate(zip(DCO_CH_MUL, DCO_DOWN), start=1):
            s += _res_block_spec(f"encoder.{j}", cin, channel * m, down)
            cin = channel * m
        k_size, feat = (3, 4) if image_size > 511 else (2, 1)
        s += _conv_layer_spec("encoder.7", cin, channel * 12, k_size, pad="valid")
        s += _linear_spec("linear.0", channel * 12 * 2 * feat, channel * 32)
        s += _linear_spec("linear.1", channel * 32, channel * 32)
        s += _linear_spec("linear.2", channel * 32, channel * 16)
        s += _linear_spec("linear.3", channel * 16, 1)
    elif name == "DistributionDiscriminator":                  # models.py:429-437
        t = texture_channel
        for j, (a, b) in enumerate(((t, t // 4), (t // 4, t // 16), (t // 16, t // 64), (t // 64, 1))):
            s += _linear_spec(f"model.{j}", a, b)
    else:
        raise NotImplementedError(name)                        # models.py:512-513
    return s


def init_state(name: str, seed: int | None = None, **cfg) -> Dict[str, torch.Tensor]:
    """Seeded initialiser: walks the spec in creation order drawing ``randn`` exactly where
    the reference constructors do (stylegan2/model.py:100-103,138,222-223; models.py:17-19),
    so ``torch.manual_seed(s); init_state(name)`` equals ``torch.manual_seed(s);
    init_model(name, args).state_dict()`` of the reference."""
    if seed is not None:
        torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    base = O.make_kernel(BLUR_TAPS)
    for key, shape, init in param_spec(name, **cfg):
        if init == "randn":
            sd[key] = torch.randn(*shape)
        elif init == "zeros":
            sd[key] = torch.zeros(*shape)
        elif init == "ones":
            sd[key] = torch.ones(*shape)
        elif init == "blur1":
            sd[key] = base.clone()
        elif init == "blur4":
            sd[key] = base * 4.0
        else:
            raise ValueError(init)
    return sd


def is_buffer(key: str) -> bool:
    return key.endswith(".kernel")


# ==========================================================================
# functional forward passes
# ==========================================================================
def conv_layer(sd, prefix, x, k, *, up=False, down=False, pad="zero", bias=True,
               activate=True, tanh=False):
    """One ConvLayer (models.py:49-134) evaluated from its keys."""
    i = 0
    if down:
        p = (4 - 2) + (k - 1)                                  # models.py:68-74
        x = O.blur(x, sd[f"{prefix}.{i}.kernel"], ((p + 1) // 2, p // 2))
        i += 1
    if up:
        b = sd.get(f"{prefix}.{i}.bias") if (bias and not activate) else None
        x = O.equal_conv_transpose2d(x, sd[f"{prefix}.{i}.weight"], b, stride=2, padding=0)
        i += 1
        p = (4 - 2) - (k - 1)                                  # models.py:90-95
        x = O.blur(x, sd[f"{prefix}.{i}.kernel"], ((p + 1) // 2 + 1, p // 2 + 1))
        i += 1
    else:
        padding = 0
        if not down:
            if pad == "zero":
                padding = (k - 1) // 2
            elif pad == "reflect":
                r = (k - 1) // 2
                if r > 0:
                    x = F.pad(x, [r, r, r, r], mode="reflect")
                    i += 1
            elif pad != "valid":
                raise ValueError(pad)
        b = sd.get(f"{prefix}.{i}.bias") if (bias and not activate) else None
        x = O.equal_conv2d(x, sd[f"{prefix}.{i}.weight"], b, stride=2 if down else 1, padding=padding)
        i += 1
    if activate:
        if tanh:
            x = torch.tanh(x)
        elif bias:
            x = O.fused_leaky_relu(x, sd[f"{prefix}.{i}.bias"])
        else:
            x = O.scaled_leaky_relu(x)
    return x


def res_block(sd, prefix, x, down, pad="zero"):
    """models.py:181-227."""
    cout, cin = sd[f"{prefix}.conv1.{1 if pad == 'reflect' else 0}.weight"].shape[:2]
    y = conv_layer(sd, f"{prefix}.conv1", x, 3, pad=pad)
    y = conv_layer(sd, f"{prefix}.conv2", y, 3, down=down, pad=pad)
    if down or cin != cout:
        x = conv_layer(sd, f"{prefix}.skip", x, 1, down=down, bias=False, activate=False)
    return (y + x) / math.sqrt(2)


def styled_res_block(sd, prefix, x, style, up):
    """models.py:137-178 with StyledConv_without_noise (models.py:7)."""
    def sc(p, inp, upsample):
        return O.styled_conv(inp, style, sd[f"{p}.conv.weight"], sd[f"{p}.conv.modulation.weight"],
                             sd[f"{p}.conv.modulation.bias"], sd[f"{p}.activate.bias"],
                             upsample=upsample, blur_kernel=sd.get(f"{p}.conv.blur.kernel"))
    y = sc(f"{prefix}.conv1", x, up)
    y = sc(f"{prefix}.conv2", y, False)
    if any(kk.startswith(f"{prefix}.skip.") for kk in sd):
        x = conv_layer(sd, f"{prefix}.skip", x, 1, up=up, bias=False, activate=False)
    return (y + x) / math.sqrt(2)


def encoder(sd, x):
    """DisentanglementEncoder.forward (models.py:262-268) -> (structure, texture)."""
    h = conv_layer(sd, "stem.0", x, 1)
    for i in range(1, 5):
        h = res_block(sd, f"stem.{i}", h, True, "reflect")
    s = conv_layer(sd, "structure.0", h, 1)
    s = conv_layer(sd, "structure.1", s, 1)
    t = conv_layer(sd, "texture.0", h, 3, down=True, pad="valid")
    t = conv_layer(sd, "texture.1", t, 3, down=True, pad="valid")
    t = t.mean(dim=(2, 3), keepdim=True)                       # AdaptiveAvgPool2d(1)
    t = conv_layer(sd, "texture.3", t, 1, tanh=True)
    return s, torch.flatten(t, 1)


def generator(sd, structure, texture):
    """Generator.forward (models.py:296-306); noises are accepted and ignored upstream."""
    h = structure
    for i, up in enumerate(G_UP):
        h = styled_res_block(sd, f"layers.{i}", h, texture, up)
    return conv_layer(sd, "to_rgb", h, 1, activate=False)


def _five_stage(sd, root, x):
    h = conv_layer(sd, f"{root}.0", x, 1)
    for i in range(1, 4):
        h = res_block(sd, f"{root}.{i}", h, False, "reflect")
    return conv_layer(sd, f"{root}.4", h, 1)


def structure_generator(sd, z):
    """models.py:327-329."""
    return _five_stage(sd, "structure", z)


def extractor(sd, s):
    """models.py:462-465."""
    return _five_stage(sd, "extract", s)


def image_discriminator(sd, x):
    """ImageLevelDiscriminator.forward (models.py:369-376)."""
    h = conv_layer(sd, "convs.0", x, 1)
    j = 1
    while f"convs.{j}.conv1.0.weight" in sd:
        h = res_block(sd, f"convs.{j}", h, True)
        j += 1
    h = conv_layer(sd, "final_conv", h, 3)
    h = h.reshape(h.shape[0], -1)
    h = O.equal_linear(h, sd["final_linear.0.weight"], sd["final_linear.0.bias"], activation="fused_lrelu")
    return O.equal_linear(h, sd["final_linear.1.weight"], sd["final_linear.1.bias"])


def cooccur_encoder(sd, x):
    h = conv_layer(sd, "encoder.0", x, 1)
    for j, down in enumerate(DCO_DOWN, start=1):
        h = res_block(sd, f"encoder.{j}", h, down)
    k = sd["encoder.7.0.weight"].shape[-1]
    return conv_layer(sd, "encoder.7", h, k, pad="valid")


def cooccur_discriminator(sd, x, reference=None, ref_batch=None, ref_input=None):
    """CooccurenceDiscriminator.forward (models.py:413-426) -> (pred, ref_input)."""
    out_input = cooccur_encoder(sd, x)
    if ref_input is None:
        r = cooccur_encoder(sd, reference)
        _, c, h, w = r.shape
        ref_input = r.reshape(-1, ref_batch, c, h, w).mean(1)
    h = torch.flatten(torch.cat((out_input, ref_input), 1), 1)
    for j in range(3):
        h = O.equal_linear(h, sd[f"linear.{j}.weight"], sd[f"linear.{j}.bias"], activation="fused_lrelu")
    return O.equal_linear(h, sd["linear.3.weight"], sd["linear.3.bias"]), ref_input


def distribution_discriminator(sd, t):
    """models.py:439-441 (all four linears are activated, :432-436)."""
    h = t
    for j in range(4):
        h = O.equal_linear(h, sd[f"model.{j}.weight"], sd[f"model.{j}.bias"], activation="fused_lrelu")
    return h


FORWARD = {
    "DisentanglementEncoder": encoder,
    "Generator": generator,
    "StructureGenerator": structure_generator,
    "TensorExtractor": extractor,
    "ImageLevelDiscriminator": image_discriminator,
    "CooccurenceDiscriminator": cooccur_discriminator,
    "DistributionDiscriminator": distribution_discriminator,
}
