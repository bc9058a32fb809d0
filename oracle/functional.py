"""Op-level CPU oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Each function restates, from its definition, one operator of the reference hot
path and cites the reference lines it follows (paths relative to
/root/reference).  Everything is torch-CPU fp32 and built only from
differentiable torch primitives, so first- and second-order gradients of the
oracle come from torch autograd and can be compared with the hand-written CUDA
backward / double-backward kernels.

The restatements deliberately do *not* share structure with the reference
implementation: the FIR is a sum of shifted slices (not a conv2d call), the
modulated convolution loops over samples (not a grouped conv on materialised
batch*Cout weights), and nothing here is an nn.Module.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# --------------------------------------------------------------------------
# A3  bias + leaky ReLU  (stylegan2/op/fused_act.py:52-97,
#                         stylegan2/op/fused_bias_act_kernel.cu:18-49)
# --------------------------------------------------------------------------
def bias_act(x: torch.Tensor, bias: Optional[torch.Tensor], ref: Optional[torch.Tensor],
             act: int, grad: int, alpha: float, scale: float) -> torch.Tensor:
    """The reference's native entry point ``fused_bias_act(input, bias, refer, act,
    grad, alpha, scale)`` (fused_bias_act.cpp:11-20) as one expression.

    act*10+grad selects (fused_bias_act_kernel.cu:36-47):
      10/11 linear, 12 zero, 30 ``x>0 ? x : alpha*x``, 31 ``ref>0 ? x : alpha*x``,
      32 zero; the result is multiplied by ``scale``.  ``bias`` is indexed by
      dimension 1 of ``x`` (kernel.cu:29, step_b = prod(shape[2:])).
    """
    v = x
    if bias is not None and bias.numel():
        v = v + bias.reshape(1, -1, *([1] * (x.ndim - 2)))
    code = act * 10 + grad
    if code in (12, 32):
        y = torch.zeros_like(v)
    elif code == 30:
        y = torch.where(v > 0, v, v * alpha)
    elif code == 31:
        assert ref is not None
        y = torch.where(ref > 0, v, v * alpha)
    else:  # 10, 11 and the kernel's default branch
        y = v
    return y * scale


def fused_leaky_relu(x: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.2,
                     scale: float = SQRT2) -> torch.Tensor:
    """``scale * leaky_relu(x + bias[c])`` -- fused_act.py:86-97 (CUDA branch semantics:
    honours ``negative_slope``; the reference's CPU branch hard-codes 0.2, fused_act.py:91)."""
    return bias_act(x, bias, None, 3, 0, negative_slope, scale)


def scaled_leaky_relu(x: torch.Tensor, negative_slope: float = 0.2) -> torch.Tensor:
    """stylegan2/model.py:169-178."""
    return torch.where(x > 0, x, x * negative_slope) * SQRT2


# --------------------------------------------------------------------------
# A4  upfirdn2d  (stylegan2/op/upfirdn2d.py:145-200, upfirdn2d_kernel.cu:107-207)
# --------------------------------------------------------------------------
def make_kernel(k: Sequence[float]) -> torch.Tensor:
    """Outer product of a 1-D tap list, normalised to sum 1 (stylegan2/model.py:22-30)."""
    k = torch.tensor(list(k), dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def upfirdn2d_out_size(n: int, up: int, down: int, p0: int, p1: int, k: int) -> int:
    """upfirdn2d.py:103-104."""
    return (n * up + p0 + p1 - k) // down + 1


def upfirdn2d(x: torch.Tensor, kernel: torch.Tensor, up: int = 1, down: int = 1,
              pad: Tuple[int, int] = (0, 0)) -> torch.Tensor:
    """Zero-insert upsample, pad (negative = crop), 2-D FIR with the *flipped* kernel,
    decimate.  x is (B, C, H, W); the same (pad0, pad1) applies to both axes, as in the
    public wrapper (upfirdn2d.py:145-157).  Definition: upfirdn2d_native, upfirdn2d.py:159-200.
    """
    B, C, H, W = x.shape
    kh, kw = kernel.shape
    p0, p1 = pad
    if up > 1:
        z = x.new_zeros(B, C, H * up, W * up)
        z[:, :, ::up, ::up] = x           # sample i lands on index i*up, (up-1) zeros follow
    else:
        z = x
    z = F.pad(z, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    hz, wz = z.shape[2], z.shape[3]
    z = z[:, :, max(-p0, 0): hz - max(-p1, 0), max(-p0, 0): wz - max(-p1, 0)]
    oh_full = z.shape[2] - kh + 1
    ow_full = z.shape[3] - kw + 1
    acc = None
    for i in range(kh):
        for j in range(kw):
            tap = kernel[kh - 1 - i, kw - 1 - j]            # true convolution = flipped taps
            term = z[:, :, i:i + oh_full, j:j + ow_full] * tap
            acc = term if acc is None else acc + term
    out = acc[:, :, ::down, ::down]
    assert out.shape[2] == upfirdn2d_out_size(H, up, down, p0, p1, kh)
    assert out.shape[3] == upfirdn2d_out_size(W, up, down, p0, p1, kw)
    return out


def blur(x: torch.Tensor, kernel: torch.Tensor, pad: Tuple[int, int]) -> torch.Tensor:
    """``Blur.forward`` (stylegan2/model.py:73-91): upfirdn2d with up=down=1; the kernel
    buffer already contains the ``upsample_factor**2`` gain when there is one."""
    return upfirdn2d(x, kernel, 1, 1, pad)


# --------------------------------------------------------------------------
# A5/A6  equalised-lr conv / transposed conv / linear
# --------------------------------------------------------------------------
def equal_conv2d(x, weight, bias=None, stride=1, padding=0):
    """stylegan2/model.py:94-123: conv2d with the weight scaled by 1/sqrt(Cin*k*k)."""
    cout, cin, k, _ = weight.shape
    scale = 1.0 / math.sqrt(cin * k * k)
    return F.conv2d(x, weight * scale, bias=bias, stride=stride, padding=padding)


def equal_conv_transpose2d(x, weight, bias=None, stride=1, padding=0):
    """models.py:11-40: conv_transpose2d, weight (Cin, Cout, k, k) scaled by 1/sqrt(Cin*k*k)."""
    cin, cout, k, _ = weight.shape
    scale = 1.0 / math.sqrt(cin * k * k)
    return F.conv_transpose2d(x, weight * scale, bias=bias, stride=stride, padding=padding)


def equal_linear(x, weight, bias=None, lr_mul=1.0, activation=None):
    """stylegan2/model.py:132-161."""
    out_dim, in_dim = weight.shape
    scale = (1.0 / math.sqrt(in_dim)) * lr_mul
    y = x @ (weight * scale).t()
    if activation:
        return fused_leaky_relu(y, bias * lr_mul)
    if bias is not None:
        y = y + bias * lr_mul
    return y


# --------------------------------------------------------------------------
# A1  modulated convolution  (stylegan2/model.py:181-277)
# --------------------------------------------------------------------------
def modulated_conv2d(x, style, weight, mod_weight, mod_bias, *, demodulate=True,
                     upsample=False, downsample=False, blur_kernel=None, eps=1e-8):
    """``ModulatedConv2d.forward`` restated sample by sample.

    x (B,Cin,H,W); style (B,style_dim); weight (1,Cout,Cin,k,k);
    mod_weight (Cin,style_dim) / mod_bias (Cin,) are the ``modulation`` EqualLinear
    (model.py:226, bias_init=1).  ``blur_kernel`` is the (4,4) buffer of the layer's
    ``Blur`` (already x4 for upsample, model.py:81-82,208).
    Per sample b (model.py:239-248):  w_b = scale*W*s_b ; w_b *= rsqrt(sum_{i,k} w_b^2 + eps).
    """
    B, cin, H, W = x.shape
    _, cout, _, k, _ = weight.shape
    scale = 1.0 / math.sqrt(cin * k * k)
    s = equal_linear(style, mod_weight, mod_bias)              # (B, Cin)
    outs = []
    for b in range(B):
        wb = scale * weight[0] * s[b].reshape(1, cin, 1, 1)    # (Cout, Cin, k, k)
        if demodulate:
            d = torch.rsqrt(wb.pow(2).sum(dim=(1, 2, 3)) + eps)
            wb = wb * d.reshape(cout, 1, 1, 1)
        xb = x[b:b + 1]
        if upsample:
            # model.py:250-261: transposed conv, stride 2, no padding, then Blur(pad0,pad1)
            yb = F.conv_transpose2d(xb, wb.transpose(0, 1), stride=2, padding=0)
            p = (blur_kernel.shape[0] - 2) - (k - 1)
            yb = blur(yb, blur_kernel, ((p + 1) // 2 + 1, p // 2 + 1))
        elif downsample:
            # model.py:263-269: Blur(pad0,pad1) then stride-2 conv without padding
            p = (blur_kernel.shape[0] - 2) + (k - 1)
            yb = F.conv2d(blur(xb, blur_kernel, ((p + 1) // 2, p // 2)), wb, stride=2, padding=0)
        else:
            yb = F.conv2d(xb, wb, padding=k // 2)
        outs.append(yb)
    return torch.cat(outs, 0)


def styled_conv(x, style, weight, mod_weight, mod_bias, act_bias, *, upsample=False,
                blur_kernel=None, demodulate=True, noise=None, noise_weight=None):
    """``StyledConv`` / ``StyledConv_without_noise`` (stylegan2/model.py:307-377):
    modulated conv -> [noise injection, model.py:280-291] -> FusedLeakyReLU."""
    y = modulated_conv2d(x, style, weight, mod_weight, mod_bias, demodulate=demodulate,
                         upsample=upsample, blur_kernel=blur_kernel)
    if noise_weight is not None:
        assert noise is not None, "oracle needs explicit noise to stay deterministic"
        y = y + noise_weight * noise
    return fused_leaky_relu(y, act_bias)
