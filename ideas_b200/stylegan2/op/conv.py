"""Convolution family on the B200 C ABI (new native unit).

The reference computes every convolution with cuDNN through ``F.conv2d`` /
``F.conv_transpose2d`` (stylegan2/model.py:115,258,267,273; models.py:32) and, for the
modulated convolution, on per-sample weights materialised as a (B*Cout, Cin, k, k) tensor
(stylegan2/model.py:240-275).  Here the three GEMMs of a convolution are explicit autograd
Functions over NHWC activations and *packed* weights ``wp[tap][K][C]``:

    ConvFwd(x, wp)    y  = conv(x, w)                      (+ bias, leaky ReLU epilogue)
    ConvDgrad(dy, wp) dx = conv_transpose(dy, w)           (also the forward of a transposed conv)
    ConvWgrad(x, dy)  dwp

Each one's backward is written with the other two, so gradients of any order exist -- the
R1 penalty (reference utils.py:112-118) needs second order through the discriminators.
``ModConv`` / ``ModConvUp`` implement ModulatedConv2d + FusedLeakyReLU (and the upsampling
variant with its Blur) as  act(d * conv(s * x) + b)  -- algebraically identical to the
reference's weight modulation/demodulation, never building per-sample weights.

Nothing here saves a leaf parameter for backward: only packed copies are saved, so the
reference's second ``backward()`` over a retained graph after ``optimizer.step()``
(train.py:209-216) keeps working.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import _lib
from ..._tensor import empty_nhwc, nhwc, ptr, require_cuda, stream_ptr
from .fused_act import FusedLeakyReLUFunctionBackward
from .upfirdn2d import UpFirDn2dBackward, _grad_pad, flipped, _run as _upfirdn_run

import os

# tests flip this to compare SIMT and tcgen05 paths; IDEAS_B200_UMMA=0 forces the fp32 FFMA kernels
DEFAULT_IMPL = _lib.IMPL_SIMT if os.environ.get("IDEAS_B200_UMMA", "1") == "0" else _lib.IMPL_AUTO


def set_default_impl(impl: int) -> int:
    global DEFAULT_IMPL
    old, DEFAULT_IMPL = DEFAULT_IMPL, impl
    return old


# --------------------------------------------------------------------------- per-step cache of derived weights
_AB_NO_CACHE = os.environ.get("IDEAS_AB_NO_CACHE", "0") == "1"      # measurement / debugging switch


class _StepCache:
    """Packed weights (and the other tensors derived from a parameter alone) are functions of the parameter, which
    changes once per optimiser step -- yet a training iteration runs E twice, G and Dreal three to six times.  Inside
    ``step_scope()`` such tensors are built once per (parameter, version, grad mode) and shared by every use: one
    PackWeight node per weight and phase, whose gradient autograd accumulates before ONE unpack.  Entries derived
    from a parameter are keyed by the parameter's ``_version`` AND by an epoch that the owner of the scope advances
    after every optimiser step (``invalidate_step_cache``; fused Adam updates parameters without touching their
    version counters).  The cache is dropped when the scope exits, so nothing derived from stale weights -- or
    living in a CUDA graph's private memory pool -- can be seen outside the iteration that built it.  Entries are
    per CUDA stream (see ``cached``): sharing one packed weight between the side-stream branches of a captured
    iteration produced wrong results under graph replay (scripts/debug_graph_streams.py), and a few extra
    microsecond-sized pack kernels per stream are cheaper than any cross-stream ordering of cached tensors."""
    depth = 0
    epoch = 0
    store: dict = {}


def invalidate_step_cache():
    """Call after every optimiser step taken inside a step_scope (train_step.Trainer registers it as a step hook)."""
    _StepCache.epoch += 1


class step_scope:
    def __enter__(self):
        if _StepCache.depth == 0:
            _StepCache.store = {}
        _StepCache.depth += 1
        return self

    def __exit__(self, *exc):
        _StepCache.depth -= 1
        if _StepCache.depth == 0:
            _StepCache.store = {}
        return False


def cached(key: torch.Tensor, tag, build, immutable: bool = False):
    """``build()`` once per (key tensor, version, epoch, autograd mode, tag) inside a step_scope; plain call outside.
    ``immutable``: the key is a temporary that nothing updates in place (a packed copy), so the epoch is ignored."""
    if _StepCache.depth == 0 or _AB_NO_CACHE:
        return build()
    # the CUDA stream is part of the key: an entry is only ever read on the stream that wrote it, so branches of the
    # iteration that run on side streams (train_step.py) never depend on another stream's kernels through the cache
    sid = torch.cuda.current_stream(key.device).cuda_stream if key.is_cuda else 0
    k = (id(key), key._version, -1 if immutable else _StepCache.epoch,
         bool(key.requires_grad and torch.is_grad_enabled()), sid, tag)
    hit = _StepCache.store.get(k)
    if hit is None:
        hit = (key, build())                 # holding ``key`` keeps id() unique while the entry lives
        _StepCache.store[k] = hit
    return hit[1]


def packed_weight(w: torch.Tensor, transpose: bool, scale: float):
    """PackWeight of a parameter, shared across the uses of one training iteration."""
    return cached(w, ("pack", bool(transpose), float(scale)), lambda: PackWeight.apply(w, transpose, scale))


class Geom(NamedTuple):
    """Forward-convolution geometry: x (N,C,H,W) -> y (N,K,OH,OW)."""
    N: int
    H: int
    W: int
    C: int
    K: int
    kh: int
    kw: int
    stride: int
    pad: int
    OH: int
    OW: int

    @staticmethod
    def forward(x_shape, K, kh, kw, stride, pad) -> "Geom":
        n, c, h, w = x_shape
        return Geom(n, h, w, c, K, kh, kw, stride, pad, (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1)

    @staticmethod
    def transposed(x_shape, C_out, kh, kw, stride, pad) -> "Geom":
        """Geometry whose *dgrad* is conv_transpose2d(x): x plays dy (N,K,OH,OW)."""
        n, k, oh, ow = x_shape
        return Geom(n, (oh - 1) * stride + kh - 2 * pad, (ow - 1) * stride + kw - 2 * pad, C_out, k, kh, kw, stride, pad, oh, ow)


# --------------------------------------------------------------------------- raw launches
def _fwd(x, wp, g: Geom, in_scale=None, out_scale=None, bias=None, act=_lib.ACT_NONE, alpha=0.2, gain=1.0, impl=None,
         residual=None, res_scale=1.0):
    y = empty_nhwc(g.N, g.K, g.OH, g.OW, x)
    _lib.call("ideas_conv2d_forward_res", ptr(y), ptr(x), ptr(wp), ptr(in_scale), ptr(out_scale), ptr(bias), ptr(residual),
              float(res_scale), g.N, g.H, g.W, g.C, g.K, g.kh, g.kw, g.stride, g.pad, act, float(alpha), float(gain),
              DEFAULT_IMPL if impl is None else impl, stream_ptr(x))
    return y


def _dgrad(dy, wp, g: Geom, in_scale=None, out_scale=None, impl=None):
    """dx (N,C,H,W) = in_scale * conv_transpose(out_scale * dy, w)."""
    wpt = dgrad_weights(wp)
    dx = empty_nhwc(g.N, g.C, g.H, g.W, dy)
    _lib.call("ideas_conv2d_dgrad", ptr(dx), ptr(dy), ptr(wpt), ptr(in_scale), ptr(out_scale), ptr(None),
              g.N, g.H, g.W, g.C, g.K, g.kh, g.kw, g.stride, g.pad, g.OH, g.OW, _lib.ACT_NONE, 0.2, 1.0,
              DEFAULT_IMPL if impl is None else impl, stream_ptr(dy))
    return dx


def dgrad_weights(wp):
    """The data-gradient operand of a packed weight (taps, K, C) -> (taps reversed, C, K); built once per packed
    weight inside a step_scope (on the stream of the first caller: see model.warm_weight_cache)."""
    def repack():
        t = torch.empty_like(wp)
        _lib.call("ideas_repack_dgrad", ptr(t), ptr(wp), wp.shape[1], wp.shape[2], wp.shape[0], stream_ptr(wp))
        return t

    return cached(wp, "dgrad", repack, immutable=True)


def _wgrad(x, dy, g: Geom, in_scale=None, out_scale=None, impl=None):
    dwp = torch.zeros((g.kh * g.kw, g.K, g.C), device=x.device, dtype=x.dtype)
    _lib.call("ideas_conv2d_wgrad", ptr(dwp), ptr(x), ptr(dy), ptr(in_scale), ptr(out_scale),
              g.N, g.H, g.W, g.C, g.K, g.kh, g.kw, g.stride, g.pad, g.OH, g.OW,
              DEFAULT_IMPL if impl is None else impl, stream_ptr(x))
    return dwp


def _check(x, shape, what):
    if tuple(x.shape) != tuple(shape):
        raise RuntimeError(f"{what}: expected shape {tuple(shape)}, got {tuple(x.shape)}")


# --------------------------------------------------------------------------- weight packing
class PackWeight(Function):
    """(O, I, kh, kw) parameter * scale -> packed (taps, O, I)  [transpose=False]
                                        or (taps, I, O)        [transpose=True]."""

    @staticmethod
    def forward(ctx, w, transpose, scale):
        require_cuda(w)
        w = w.contiguous()
        o, i, kh, kw = w.shape
        ctx.cfg = (o, i, kh, kw, transpose, scale)
        dst = torch.empty((kh * kw, i, o) if transpose else (kh * kw, o, i), device=w.device, dtype=w.dtype)
        _lib.call("ideas_pack_weight", ptr(dst), ptr(w), o, i, kh, kw, int(transpose), 0, float(scale), stream_ptr(w))
        return dst

    @staticmethod
    def backward(ctx, g):
        o, i, kh, kw, transpose, scale = ctx.cfg
        return UnpackWeight.apply(g, o, i, kh, kw, transpose, scale), None, None


class UnpackWeight(Function):
    @staticmethod
    def forward(ctx, g, o, i, kh, kw, transpose, scale):
        g = g.contiguous()
        ctx.cfg = (transpose, scale)
        dst = torch.empty((o, i, kh, kw), device=g.device, dtype=g.dtype)
        _lib.call("ideas_unpack_weight_grad", ptr(dst), ptr(g), o, i, kh, kw, int(transpose), float(scale), 0,
                  stream_ptr(g))
        return dst

    @staticmethod
    def backward(ctx, gg):
        transpose, scale = ctx.cfg
        return PackWeight.apply(gg, transpose, scale), None, None, None, None, None, None


# --------------------------------------------------------------------------- the three GEMMs
class ConvFwd(Function):
    """y = act(conv(x, w) + bias); with ``residual`` (same shape as y, activation-free convs only):
    y = (conv(x, w) + bias + residual) * res_scale -- the residual merge of a ResBlock whose skip path ends in this
    convolution (models.py:178,227), done in the convolution's epilogue."""

    @staticmethod
    def forward(ctx, x, wp, bias, g: Geom, act, alpha, gain, residual=None, res_scale=1.0):
        require_cuda(x, wp, bias, residual)
        x = nhwc(x)
        _check(x, (g.N, g.C, g.H, g.W), "conv forward input")
        _check(wp, (g.kh * g.kw, g.K, g.C), "conv forward packed weight")
        if residual is not None:
            if act != _lib.ACT_NONE:
                raise RuntimeError("conv forward: a residual operand is merged after activation-free convolutions only")
            residual = nhwc(residual)
            _check(residual, (g.N, g.K, g.OH, g.OW), "conv forward residual")
        y = _fwd(x, wp, g, bias=bias, act=act, alpha=alpha, gain=gain, residual=residual, res_scale=res_scale)
        ctx.g, ctx.act, ctx.alpha, ctx.gain, ctx.has_bias = g, act, alpha, gain, bias is not None
        ctx.res_scale = res_scale if residual is not None else None
        ctx.save_for_backward(x, wp, y if act != _lib.ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, wp, out = ctx.saved_tensors
        g = ctx.g
        want_bias = ctx.has_bias and ctx.needs_input_grad[2]
        gb = gres = None
        if ctx.res_scale is not None:
            if ctx.res_scale != 1.0:                   # (the residual blocks fold their 1/sqrt(2) into weights: no pass)
                gy = gy * ctx.res_scale                # one pass serves the residual branch and this convolution
            gres = gy if ctx.needs_input_grad[7] else None
        if ctx.act != _lib.ACT_NONE:
            gy, gb = FusedLeakyReLUFunctionBackward.apply(gy, out, want_bias, ctx.alpha, ctx.gain)
        elif want_bias:
            gb = gy.sum(dim=(0, 2, 3))
        gx = ConvDgrad.apply(gy, wp, g) if ctx.needs_input_grad[0] else None
        gw = ConvWgrad.apply(x, gy, g) if ctx.needs_input_grad[1] else None
        return gx, gw, (gb if want_bias else None), None, None, None, None, gres, None


class ConvDgrad(Function):
    @staticmethod
    def forward(ctx, dy, wp, g: Geom):
        require_cuda(dy, wp)
        dy = nhwc(dy)
        _check(dy, (g.N, g.K, g.OH, g.OW), "conv dgrad input")
        _check(wp, (g.kh * g.kw, g.K, g.C), "conv dgrad packed weight")
        ctx.g = g
        ctx.save_for_backward(dy, wp)
        return _dgrad(dy, wp, g)

    @staticmethod
    def backward(ctx, gg):
        dy, wp = ctx.saved_tensors
        g = ctx.g
        g_dy = ConvFwd.apply(gg, wp, None, g, _lib.ACT_NONE, 0.2, 1.0) if ctx.needs_input_grad[0] else None
        g_wp = ConvWgrad.apply(gg, dy, g) if ctx.needs_input_grad[1] else None
        return g_dy, g_wp, None


class ConvWgrad(Function):
    @staticmethod
    def forward(ctx, x, dy, g: Geom):
        require_cuda(x, dy)
        x, dy = nhwc(x), nhwc(dy)
        _check(x, (g.N, g.C, g.H, g.W), "conv wgrad input")
        _check(dy, (g.N, g.K, g.OH, g.OW), "conv wgrad output-gradient")
        ctx.g = g
        ctx.save_for_backward(x, dy)
        return _wgrad(x, dy, g)

    @staticmethod
    def backward(ctx, G):
        x, dy = ctx.saved_tensors
        g = ctx.g
        G = G.contiguous()
        g_x = ConvDgrad.apply(dy, G, g) if ctx.needs_input_grad[0] else None
        g_dy = ConvFwd.apply(x, G, None, g, _lib.ACT_NONE, 0.2, 1.0) if ctx.needs_input_grad[1] else None
        return g_x, g_dy, None


def _blur_act_bwd_ok(channels: int, kernel=None, batch: int = 1) -> bool:
    """Host-side statement of the fast-path envelope of ideas_blur_act_backward / ideas_blur_scale_dot_backward."""
    c4 = channels // 4
    if kernel is not None and (kernel.shape[0] > 4 or kernel.shape[1] > 4):
        return False
    return channels % 4 == 0 and batch <= 65535 and ((c4 <= 128 and 128 % c4 == 0) or c4 % 128 == 0)


def _modconv_act_backward(gy, out, d, alpha, gain, n, p, c):
    """g1d, gsum, dotz of ideas_modconv_act_backward; composed of torch ops outside the kernel's envelope
    (channels not a multiple of 4, or more than 2048 of them)."""
    if c % 4 == 0 and c <= 2048:
        g1d = torch.empty_like(out)
        gsum = torch.zeros((n, c), device=out.device, dtype=out.dtype)
        dotz = torch.zeros_like(gsum)
        _lib.call("ideas_modconv_act_backward", ptr(g1d), ptr(gsum), ptr(dotz), ptr(gy), ptr(out), ptr(d), float(alpha),
                  float(gain), n, p, c, stream_ptr(out))
        return g1d, gsum, dotz
    slope = torch.where(out > 0, 1.0, float(alpha))
    g1 = gy * slope * gain
    z = out / (slope * gain)
    g1d = g1 if d is None else g1 * d.view(n, c, 1, 1)
    return g1d, g1.sum(dim=(2, 3)), (g1 * z).sum(dim=(2, 3))


class ConvActBlur(Function):
    """z = Blur(gain * lrelu(conv(x, w) + b)): conv1 -> FusedLeakyReLU -> Blur of a down-sampling ResBlock
    (reference models.py:213-227 with :49-134).  Forward is the two kernels it always was; the point is the
    backward: the blur's backward and the activation's backward run as ONE kernel (ideas_blur_act_backward), so the
    blurred gradient never makes an HBM round trip.  Under create_graph=True (the R1 penalty, utils.py:112-118)
    the backward is instead composed of the differentiable pieces, exactly as without the fusion."""

    @staticmethod
    def forward(ctx, x, wp, bias, g: Geom, alpha, gain, kernel, pad):
        require_cuda(x, wp, bias, kernel)
        x = nhwc(x)
        kernel = kernel.contiguous()
        _check(x, (g.N, g.C, g.H, g.W), "conv forward input")
        y = _fwd(x, wp, g, bias=bias, act=_lib.ACT_LRELU, alpha=alpha, gain=gain)
        pad4 = (pad[0], pad[1], pad[0], pad[1])
        z = _upfirdn_run(y, kernel, (1, 1), (1, 1), pad4)
        ctx.save_for_backward(x, wp, y, kernel, flipped(kernel))
        ctx.cfg = (g, alpha, gain, pad4, (z.shape[2], z.shape[3]))
        return z

    @staticmethod
    def backward(ctx, gz):
        x, wp, y, kernel, gkernel = ctx.saved_tensors
        g, alpha, gain, pad4, zhw = ctx.cfg
        want_bias = ctx.needs_input_grad[2]
        g_pad = _grad_pad((y.shape[2], y.shape[3]), zhw, kernel.shape, (1, 1), (1, 1), pad4)
        if torch.is_grad_enabled() or not _blur_act_bwd_ok(g.K, gkernel, g.N):
            gy = UpFirDn2dBackward.apply(gz, kernel, gkernel, (1, 1), (1, 1), pad4, g_pad, tuple(y.shape), zhw)
            g1, gb = FusedLeakyReLUFunctionBackward.apply(gy, y, want_bias, alpha, gain)
            gx = ConvDgrad.apply(g1, wp, g) if ctx.needs_input_grad[0] else None
            gw = ConvWgrad.apply(x, g1, g) if ctx.needs_input_grad[1] else None
            return gx, gw, (gb if want_bias else None), None, None, None, None, None
        gz = nhwc(gz)
        g1 = torch.empty_like(y)
        gb = torch.zeros(g.K, device=y.device, dtype=y.dtype) if want_bias else None
        kh, kw = gkernel.shape
        _lib.call("ideas_blur_act_backward", ptr(g1), ptr(gb), ptr(gz), ptr(y), ptr(gkernel), g.N, gz.shape[2], gz.shape[3],
                  g.K, kh, kw, g_pad[0], g_pad[1], g_pad[2], g_pad[3], float(alpha), float(gain), stream_ptr(gz))
        gx = _dgrad(g1, wp, g) if ctx.needs_input_grad[0] else None
        gw = _wgrad(x, g1, g) if ctx.needs_input_grad[1] else None
        return gx, gw, gb, None, None, None, None, None


def conv2d(x, wp, bias=None, *, K, kh, kw, stride=1, pad=0, act=False, alpha=0.2, gain=2 ** 0.5, residual=None,
           res_scale=1.0):
    """y = [lrelu](conv(x, w) + bias) with packed weights (see PackWeight); optionally (.. + residual) * res_scale."""
    g = Geom.forward(x.shape, K, kh, kw, stride, pad)
    return ConvFwd.apply(x, wp, bias, g, _lib.ACT_LRELU if act else _lib.ACT_NONE, alpha, gain if act else 1.0, residual,
                         res_scale)


def conv_transpose2d(x, wp, *, C_out, kh, kw, stride=1, pad=0):
    """y = conv_transpose2d(x, w); wp is the packed weight of the conv this is the adjoint of:
    (taps, Cin_of_x, C_out)."""
    g = Geom.transposed(x.shape, C_out, kh, kw, stride, pad)
    return ConvDgrad.apply(x, wp, g)


# --------------------------------------------------------------------------- modulated convolutions
def _umma_shape_ok(g: Geom) -> bool:
    return _lib.umma_enabled() and g.C % 32 == 0 and g.K % 32 == 0


def _act_or_bias(z, bias, act, alpha, gain):
    from .fused_act import fused_leaky_relu
    if act:
        return fused_leaky_relu(z, bias, alpha, gain)
    return z if bias is None else z + bias.view(1, -1, 1, 1)


def _modconv_composite(x, s, d, wp, bias, g: Geom, act, alpha, gain):
    """ModConv written with the differentiable building blocks (each has a double backward): used to differentiate
    the backward itself (create_graph=True: path-length regularisation, stylegan2/train.py:85-98)."""
    xm = x if s is None else x * s.view(g.N, g.C, 1, 1)          # s None: x is already modulated
    u = ConvFwd.apply(xm, wp, None, g, _lib.ACT_NONE, 0.2, 1.0)
    if d is not None:
        u = u * d.view(g.N, g.K, 1, 1)
    return _act_or_bias(u, bias, act, alpha, gain)


def _modconv_up_composite(x, s, d, wp, kernel, blur_pad, bias, g: Geom, act, alpha, gain):
    from .upfirdn2d import UpFirDn2d
    xm = x * s.view(g.N, g.K, 1, 1)
    u = ConvDgrad.apply(xm, wp, g)
    if d is not None:
        u = u * d.view(g.N, g.C, 1, 1)
    z = UpFirDn2d.apply(u, kernel, (1, 1), (1, 1), blur_pad)
    return _act_or_bias(z, bias, act, alpha, gain)


def _grad_of_composite(fn, tensors, needs, gy):
    """Gradients of ``fn(*tensors)`` w.r.t. the tensors flagged in ``needs``, themselves differentiable: the
    forward is recomputed from the (graph-connected) saved inputs under enable_grad."""
    with torch.enable_grad():
        # Aliases make the inputs independent variables of the composite: the saved tensors are interior nodes of
        # the caller's graph (the demodulation d is itself a function of the style s), and differentiating w.r.t.
        # them directly would fold the d -> s path into ds as well as returning dd (counted twice by the caller's
        # graph).  The views keep the results connected to the originals for the second differentiation.
        alias = [t.view_as(t) if (t is not None and t.requires_grad) else t for t in tensors]
        out = fn(*alias)
        wanted = [a for a, n in zip(alias, needs) if n and a is not None and a.requires_grad]
        outs, gys = (list(out), list(gy)) if isinstance(out, tuple) else ([out], [gy])
        grads = torch.autograd.grad(outs, wanted, gys, create_graph=True, allow_unused=True) if wanted else ()
    it = iter(grads)
    return [next(it) if (n and t is not None and t.requires_grad) else None for t, n in zip(tensors, needs)]


def _scale_channels(x, sc, n, p, c):
    out = torch.empty_like(x)
    _lib.call("ideas_scale_channels", ptr(out), ptr(x), ptr(sc), n, p, c, stream_ptr(x))
    return out


def _post_general(gy, gym, out, post, n, p, c):
    """Both outputs of a post-modulating layer were used: fold the gradient of out*post into the gradient of out.
    Returns (gy_total, g_post) with g_post[n,c] = sum_p gym * out."""
    g_post = torch.zeros((n, c), device=out.device, dtype=out.dtype)
    tmp = torch.empty_like(out)
    _lib.call("ideas_channel_dot", ptr(g_post), ptr(tmp), ptr(out), ptr(nhwc(gym)), ptr(post), n, p, c, stream_ptr(out))
    return (tmp if gy is None else nhwc(gy) + tmp), g_post


def _post_act_backward(gym, out, post, d, bias, alpha, gain, n, p, c):
    """Activation backward of a layer whose ONLY used output is out*post (the next layer's pre-modulated input), in
    the one pass modconv_act_backward always was:  with the kernel's scale operand set to post*d,
        g1d = gym * mask * gain * post * d          (gradient w.r.t. the un-demodulated conv / blur output)
        gsum' = sum_p gym*mask*gain,  dotz' = sum_p gym*mask*gain*z = sum_p gym*out
    so g_post = dotz' (the next layer's style gradient, formerly its own channel_dot pass over (x, dxm)),
    gsum = post*gsum', dotz = post*dotz' (bias and demodulation gradients of this layer)."""
    gym = nhwc(gym)
    eff = post if d is None else (post * d).contiguous()
    g1d, gsum, dotz = _modconv_act_backward(gym, out, eff, alpha, gain, n, p, c)
    return g1d, post * gsum, post * dotz, dotz


class ModConv(Function):
    """out = gain * lrelu(d[n,k] * conv(s[n,c] * x, w)[n,k] + bias[k])   (same resolution).

    Restates ModulatedConv2d.forward's non-resampling branch + FusedLeakyReLU
    (stylegan2/model.py:236-248,271-275,375).  ``d`` may be None (demodulate=False) and
    ``bias`` None with ``act=False`` (plain modulated conv, e.g. ToRGB).

    ``premod``: ``x`` is already s*x -- the producing layer wrote it (see ``post``) -- so no modulation pass runs
    here, and the gradient returned for ``x`` is the raw data gradient (``s`` gets none from this node).
    ``post`` (N, K): additionally returns out * post[n,k], the NEXT layer's pre-modulated input; the style gradient
    of that layer comes back as this node's gradient for ``post``."""

    @staticmethod
    def forward(ctx, x, s, d, wp, bias, g: Geom, act, alpha, gain, premod=False, post=None):
        require_cuda(x, s, d, wp, bias, post)
        x = nhwc(x)
        s = s.contiguous() if s is not None else None
        d = d.contiguous() if d is not None else None
        _check(x, (g.N, g.C, g.H, g.W), "modconv input")
        if post is not None and not act:
            raise RuntimeError("modconv: post-modulation needs the fused activation")
        xm = None
        code = _lib.ACT_LRELU if act else _lib.ACT_NONE
        if premod:
            xm = x
            out = _fwd(xm, wp, g, out_scale=d, bias=bias, act=code, alpha=alpha, gain=gain)
        elif _umma_shape_ok(g) and DEFAULT_IMPL != _lib.IMPL_SIMT:
            xm = _scale_channels(x, s, g.N, g.H * g.W, g.C)
            out = _fwd(xm, wp, g, out_scale=d, bias=bias, act=code, alpha=alpha, gain=gain)
        else:
            out = _fwd(x, wp, g, in_scale=s, out_scale=d, bias=bias, act=code, alpha=alpha, gain=gain,
                       impl=_lib.IMPL_SIMT)
        ctx.g, ctx.act, ctx.alpha, ctx.gain, ctx.premod = g, act, alpha, gain, bool(premod)
        ctx.set_materialize_grads(False)
        post = post.contiguous() if post is not None else None
        # the bias is a leaf parameter: save a private copy, never the leaf (train.py:209-216 steps the
        # optimiser between two backwards over this graph)
        ctx.save_for_backward(None if premod else x, s, d, wp, bias.clone() if bias is not None else None, out, xm, post)
        if post is None:
            return out
        return out, _scale_channels(out, post, g.N, g.OH * g.OW, g.K)

    @staticmethod
    def backward(ctx, gy, gym=None):
        x, s, d, wp, bias, out, xm, post = ctx.saved_tensors
        g = ctx.g
        if gy is None and gym is None:
            return (None,) * 11
        if torch.is_grad_enabled():
            # create_graph=True: differentiate a recomputed composite instead of the fused kernels.  The bias leaf
            # was saved as a detached copy (train.py:209-216), so second-order terms do not reach it; none exist
            # (the output is piecewise linear in the bias).
            act, alpha, gain, premod = ctx.act, ctx.alpha, ctx.gain, ctx.premod

            def fn(x_, s_, d_, w_, b_, p_):
                o = _modconv_composite(x_, None if premod else s_, d_, w_, b_, g, act, alpha, gain)
                return o if p_ is None else (o, o * p_.view(g.N, g.K, 1, 1))

            xin = xm if premod else x
            needs = [ctx.needs_input_grad[0], ctx.needs_input_grad[1] and not premod, ctx.needs_input_grad[2],
                     ctx.needs_input_grad[3], False, ctx.needs_input_grad[10]]
            gys = gy if post is None else (gy if gy is not None else torch.zeros_like(out),
                                           gym if gym is not None else torch.zeros_like(out))
            gx, gs, gd, gw, _, gp = _grad_of_composite(fn, [xin, s, d, wp, bias, post], needs, gys)
            gb = None
            if bias is not None and ctx.needs_input_grad[4]:
                gtot = gys if post is None else gys[0] + gys[1] * post.view(g.N, g.K, 1, 1)
                mask = torch.where(out > 0, 1.0, alpha) * gain if act else 1.0
                gb = (gtot * mask).sum(dim=(0, 2, 3))
            return gx, gs, gd, gw, gb, None, None, None, None, None, gp
        st = stream_ptr(out)
        P = g.OH * g.OW
        gb = gd = g_post = None
        fast_post = post is not None and gy is None
        if post is not None and not fast_post:
            gy, g_post = _post_general(gy, gym, out, post, g.N, P, g.K)
        if fast_post:
            g1d, gsum, dotz, g_post = _post_act_backward(gym, out, post, d, bias, ctx.alpha, ctx.gain, g.N, P, g.K)
            if bias is not None:
                gb = gsum.sum(0)
            if d is not None:
                gd = (dotz - (bias * gsum if bias is not None else 0.0)) / d
        elif ctx.act:
            gy = nhwc(gy)
            g1d, gsum, dotz = _modconv_act_backward(gy, out, d, ctx.alpha, ctx.gain, g.N, P, g.K)
            if bias is not None:
                gb = gsum.sum(0)
            if d is not None:
                gd = (dotz - (bias * gsum if bias is not None else 0.0)) / d
        else:
            gy = nhwc(gy)
            if d is not None:
                # out = d*u (+bias): dL/dd = sum_p gy*u = sum_p gy*(out-bias)/d ; g1d = gy*d
                g1d = torch.empty_like(out)
                dot = torch.zeros((g.N, g.K), device=gy.device, dtype=gy.dtype)
                _lib.call("ideas_channel_dot", ptr(dot), ptr(g1d), ptr(out), ptr(gy), ptr(d), g.N, P, g.K, st)
                gsum = gy.sum(dim=(2, 3))
                gd = (dot - (bias * gsum if bias is not None else 0.0)) / d
                if bias is not None:
                    gb = gsum.sum(0)
            else:
                g1d = gy
                if bias is not None:
                    gb = gy.sum(dim=(0, 2, 3))
        gx = gs = gw = None
        if ctx.needs_input_grad[0] or (ctx.needs_input_grad[1] and not ctx.premod):
            dxm = _dgrad(g1d, wp, g)
            if ctx.premod:
                gx = dxm                       # the producer turns it into its own gradient and the style gradient
            else:
                gx = torch.empty_like(x)
                gs = torch.zeros((g.N, g.C), device=out.device, dtype=out.dtype)
                _lib.call("ideas_channel_dot", ptr(gs), ptr(gx), ptr(x), ptr(dxm), ptr(s), g.N, g.H * g.W, g.C, st)
        if ctx.needs_input_grad[3]:
            gw = _wgrad(xm, g1d, g) if xm is not None else _wgrad(x, g1d, g, in_scale=s, impl=_lib.IMPL_SIMT)
        return gx, gs, gd, gw, gb, None, None, None, None, None, g_post


class ModConvUp(Function):
    """out = gain * lrelu(blur(d * conv_transpose(s * x, w, stride 2)) + bias)  -- the upsampling
    branch of ModulatedConv2d + its Blur + FusedLeakyReLU (stylegan2/model.py:250-261,375).
    ``wp`` is (taps, Cin, Cout): the packed weight of the stride-2 conv this is the adjoint of.
    ``g`` is that conv's geometry (x here plays its output-gradient).
    ``post`` (N, Cout): additionally returns out * post, the next layer's pre-modulated input (see ModConv)."""

    @staticmethod
    def forward(ctx, x, s, d, wp, blur_kernel, blur_pad, bias, g: Geom, act, alpha, gain, post=None):
        require_cuda(x, s, d, wp, blur_kernel, bias, post)
        if act and bias is None:
            raise RuntimeError("modconv-up: the fused activation needs a bias")
        if post is not None and not act:
            raise RuntimeError("modconv-up: post-modulation needs the fused activation")
        x = nhwc(x)
        s = s.contiguous()
        d = d.contiguous() if d is not None else None
        _check(x, (g.N, g.K, g.OH, g.OW), "modconv-up input")
        xm = None
        if _umma_shape_ok(g) and DEFAULT_IMPL != _lib.IMPL_SIMT:
            xm = _scale_channels(x, s, g.N, g.OH * g.OW, g.K)
            u = _dgrad(xm, wp, g, in_scale=d)
        else:
            u = _dgrad(x, wp, g, in_scale=d, out_scale=s, impl=_lib.IMPL_SIMT)
        kernel = blur_kernel.contiguous()
        post = post.contiguous() if post is not None else None
        out_mod = None
        if post is not None and kernel.shape[0] <= 4 and kernel.shape[1] <= 4 and g.C % 4 == 0 and g.N <= 65535:
            # Blur + bias + leaky ReLU writing BOTH the activation and its post-modulated copy (one kernel)
            kh, kw = kernel.shape
            oh = u.shape[2] + blur_pad[2] + blur_pad[3] - kh + 1
            ow = u.shape[3] + blur_pad[0] + blur_pad[1] - kw + 1
            out = empty_nhwc(g.N, g.C, oh, ow, u)
            out_mod = torch.empty_like(out)
            _lib.call("ideas_blur_bias_act_post", ptr(out), ptr(out_mod), ptr(u), ptr(kernel), ptr(bias), ptr(post), g.N,
                      u.shape[2], u.shape[3], g.C, kh, kw, blur_pad[0], blur_pad[1], blur_pad[2], blur_pad[3],
                      float(alpha), float(gain), stream_ptr(u))
        else:
            out = _upfirdn_run(u, kernel, (1, 1), (1, 1), blur_pad, bias=bias if act else None, alpha=alpha, gain=gain)
        ctx.g, ctx.act, ctx.alpha, ctx.gain, ctx.blur_pad = g, act, alpha, gain, blur_pad
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, s, d, wp, kernel, u, out if act else None, xm,
                              bias.clone() if (act and bias is not None) else None, post)
        if post is None:
            return out
        if out_mod is None:
            out_mod = _scale_channels(out, post, g.N, out.shape[2] * out.shape[3], g.C)
        return out, out_mod

    @staticmethod
    def backward(ctx, gy, gym=None):
        x, s, d, wp, kernel, u, out, xm, bias_copy, post = ctx.saved_tensors
        g = ctx.g
        if gy is None and gym is None:
            return (None,) * 12
        if torch.is_grad_enabled():
            act, alpha, gain, blur_pad = ctx.act, ctx.alpha, ctx.gain, ctx.blur_pad

            def fn(x_, s_, d_, w_, b_, p_):
                o = _modconv_up_composite(x_, s_, d_, w_, kernel, blur_pad, b_, g, act, alpha, gain)
                return o if p_ is None else (o, o * p_.view(g.N, g.C, 1, 1))

            gys = gy if post is None else (gy if gy is not None else torch.zeros_like(out),
                                           gym if gym is not None else torch.zeros_like(out))
            gx, gs, gd, gw, _, gp = _grad_of_composite(fn, [x, s, d, wp, bias_copy, post],
                                                       list(ctx.needs_input_grad[:4]) + [False, ctx.needs_input_grad[11]], gys)
            gb = None
            if act and ctx.needs_input_grad[6]:
                gtot = gys if post is None else gys[0] + gys[1] * post.view(g.N, g.C, 1, 1)
                gb = (gtot * torch.where(out > 0, 1.0, alpha) * gain).sum(dim=(0, 2, 3))
            return gx, gs, gd, gw, None, None, gb, None, None, None, None, gp
        ref = out if out is not None else u
        st = stream_ptr(ref)
        P_out = ref.shape[2] * ref.shape[3]
        gb = g_post = None
        if post is not None and gy is None:
            # the only consumer is the next layer's pre-modulated input: activation backward, the factor post[n,c] and
            # that layer's style gradient in one pass (see _post_act_backward)
            g1, gsum, _, g_post = _post_act_backward(gym, out, post, None, None, ctx.alpha, ctx.gain, g.N, P_out, g.C)
            if ctx.needs_input_grad[6]:
                gb = gsum.sum(0)
        else:
            if post is not None:
                gy, g_post = _post_general(gy, gym, out, post, g.N, P_out, g.C)
            gy = nhwc(gy)
            if ctx.act:
                g1, gb = FusedLeakyReLUFunctionBackward.apply(gy, out, ctx.needs_input_grad[6], ctx.alpha, ctx.gain)
            else:
                g1 = gy
        g_pad = _grad_pad((u.shape[2], u.shape[3]), (g1.shape[2], g1.shape[3]), kernel.shape, (1, 1), (1, 1),
                          ctx.blur_pad)
        gkernel = flipped(kernel)
        gd = None
        if d is not None and _blur_act_bwd_ok(g.C, kernel, g.N):
            # blur^T, the demodulation scale and dL/dd = sum_p gu * u / d in ONE kernel: the blurred gradient is
            # never written (was: blur backward, then a channel_dot pass over (u, gu))
            g1 = nhwc(g1)
            gud = torch.empty_like(u)
            dot = torch.zeros((g.N, g.C), device=u.device, dtype=u.dtype)
            _lib.call("ideas_blur_scale_dot_backward", ptr(gud), ptr(dot), ptr(g1), ptr(u), ptr(d), ptr(gkernel), g.N,
                      g1.shape[2], g1.shape[3], g.C, gkernel.shape[0], gkernel.shape[1], g_pad[0], g_pad[1], g_pad[2],
                      g_pad[3], st)
            gd = dot / d
        else:
            gu = _upfirdn_run(g1, gkernel, (1, 1), (1, 1), g_pad)
            if d is not None:
                gud = torch.empty_like(gu)
                dot = torch.zeros((g.N, g.C), device=u.device, dtype=u.dtype)
                _lib.call("ideas_channel_dot", ptr(dot), ptr(gud), ptr(u), ptr(gu), ptr(d), g.N, g.H * g.W, g.C, st)
                gd = dot / d
            else:
                gud = gu
        gx = gs = gw = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dxm = _fwd(gud, wp, g)                                  # adjoint of the transposed conv
            gx = torch.empty_like(x)
            gs = torch.zeros((g.N, g.K), device=u.device, dtype=u.dtype)
            _lib.call("ideas_channel_dot", ptr(gs), ptr(gx), ptr(x), ptr(dxm), ptr(s), g.N, g.OH * g.OW, g.K, st)
        if ctx.needs_input_grad[3]:
            gw = _wgrad(gud, xm, g) if xm is not None else _wgrad(gud, x, g, out_scale=s, impl=_lib.IMPL_SIMT)
        return gx, gs, gd, gw, None, None, (gb if (ctx.act and ctx.needs_input_grad[6]) else None), None, None, None, None, g_post
