"""Trainer-side helpers of IDEAS on the GPU.

Surface of the reference's utils.py:42-149 (``requires_grad``, ``accumulate``,
``d_logistic_loss``, ``d_r1_loss``, ``g_nonsaturating_loss``, ``patchify_image``,
``message_to_tensor``, ``tensor_to_message``) with the same call signatures.  The bit
mapping runs as integer CUDA kernels on bit-packed messages (the reference runs float torch
ops on CPU tensors, utils.py:74-97) and is bit-exact with it; ``bit_error_rate`` is the
XOR/popcount form of train.py:285.
"""
from __future__ import annotations

import random

import torch
from torch import autograd
from torch.nn import functional as F

from . import _lib
from ._tensor import ptr, stream_ptr


def time_change(seconds):
    """Elapsed-time formatter of the training log ("2h 3m 4s" / "3m 4s" / "4s"; reference utils.py:12-34:
    above one hour the seconds are shown after the minutes, at most two units otherwise)."""
    seconds = float(seconds)
    if seconds / 3600 > 1:
        h = int(seconds / 3600)
        m = int((seconds - h * 3600) / 60)
        return f"{h}h {m}m {int(seconds - h * 3600 - m * 60)}s"
    if seconds / 60 > 1:
        m = int(seconds / 60)
        return f"{m}m {int(seconds - m * 60)}s"
    return f"{int(seconds)}s"


def data_sampler(dataset, shuffle):
    from torch.utils import data
    return data.RandomSampler(dataset) if shuffle else data.SequentialSampler(dataset)


def requires_grad(model, flag=True):
    for p in model.parameters():
        p.requires_grad = flag


def accumulate(model1, model2, decay=0.999):
    """EMA of parameters (buffers untouched, reference utils.py:55-60) as two multi-tensor ops
    instead of ~140 tiny launches per network."""
    par1 = dict(model1.named_parameters())
    par2 = dict(model2.named_parameters())
    dst = [par1[k].data for k in par1]
    src = [par2[k].data for k in par1]
    if not dst:
        return
    torch._foreach_mul_(dst, decay)
    torch._foreach_add_(dst, src, alpha=1 - decay)


def sample_data(loader):
    while True:
        for batch in loader:
            yield batch


# ---- losses (reference utils.py:105-124)
def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    (grad_real,) = autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def g_nonsaturating_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


# ---- random crops (reference utils.py:127-149)
def draw_crops(n_crop, height, width, min_size=1 / 8, max_size=1 / 4):
    """Crop boxes [(y, x, h, w)], drawn exactly like the reference: sizes from ``torch.rand`` on the
    CPU generator, corners from ``random.randrange``."""
    size = torch.rand(n_crop) * (max_size - min_size) + min_size
    hs = (size * height).type(torch.int64).tolist()
    ws = (size * width).type(torch.int64).tolist()
    return [(random.randrange(0, height - h), random.randrange(0, width - w), h, w) for h, w in zip(hs, ws)]


class _Patchify(autograd.Function):
    """(B,C,H,W) image, device int32 boxes (n_crop,4)=(y,x,h,w) -> (B*n_crop, C, th, tw), one kernel."""

    @staticmethod
    def forward(ctx, img, boxes, th, tw):
        from ._tensor import empty_nhwc, nhwc, require_cuda
        require_cuda(img)
        img = nhwc(img)
        b, c, h, w = img.shape
        n_crop = boxes.shape[0]
        out = empty_nhwc(b * n_crop, c, th, tw, img)
        _lib.call("ideas_patchify_forward", ptr(out), ptr(img), ptr(boxes), b, h, w, c, n_crop, th, tw, stream_ptr(img))
        ctx.save_for_backward(boxes)
        ctx.cfg = (b, c, h, w, n_crop, th, tw)
        return out

    @staticmethod
    @autograd.function.once_differentiable
    def backward(ctx, g):
        from ._tensor import nhwc
        (boxes,) = ctx.saved_tensors
        b, c, h, w, n_crop, th, tw = ctx.cfg
        g = nhwc(g)
        gimg = torch.zeros((b, h, w, c), device=g.device, dtype=g.dtype).permute(0, 3, 1, 2)
        _lib.call("ideas_patchify_backward", ptr(gimg), ptr(g), ptr(boxes), b, h, w, c, n_crop, th, tw, stream_ptr(g))
        return gimg, None, None, None


def crops_to_device(crops, device):
    """[(y, x, h, w)] -> int32 (n_crop, 4) device tensor (one small async copy from pinned memory)."""
    host = torch.tensor(crops, dtype=torch.int32).reshape(-1, 4)
    if torch.device(device).type == "cuda":
        host = host.pin_memory()
    return host.to(device, non_blocking=True)


def patchify_image(img, n_crop, min_size=1 / 8, max_size=1 / 4, crops=None):
    """n_crop random crops per image, each resized (bilinear) to (H*max_size, W*max_size);
    returns (B*n_crop, C, th, tw), crops of one image adjacent (reference utils.py:127-149).
    ``crops``: None (drawn like the reference), a list of (y, x, h, w), or a device int32 (n_crop, 4)
    tensor.  One kernel launch whatever the crop sizes (the reference loops over F.interpolate)."""
    batch, channel, height, width = img.shape
    th, tw = int(height * max_size), int(width * max_size)
    if crops is None:
        crops = draw_crops(n_crop, height, width, min_size, max_size)
    boxes = crops if torch.is_tensor(crops) else crops_to_device(crops, img.device)
    return _Patchify.apply(img, boxes, th, tw)


# ---- bit path (reference utils.py:74-97, train.py:254-286)
def pack_message(message: torch.Tensor) -> torch.Tensor:
    """(B, nbits) 0/1 (any dtype) -> (B, ceil(nbits/32)) int32 words on the same device,
    bit j of word w = message[:, 32*w + j]."""
    b, n = message.shape
    words = (n + 31) // 32
    m = torch.zeros((b, words * 32), dtype=torch.int64, device=message.device)
    m[:, :n] = (message != 0).to(torch.int64)
    weights = (1 << torch.arange(32, dtype=torch.int64, device=message.device))
    packed = (m.view(b, words, 32) * weights).sum(-1)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed)      # two's complement into int32
    return packed.to(torch.int32)


def unpack_message(words: torch.Tensor, nbits: int) -> torch.Tensor:
    w = words.to(torch.int64) & 0xFFFFFFFF
    bits = (w.unsqueeze(-1) >> torch.arange(32, dtype=torch.int64, device=words.device)) & 1
    return bits.reshape(words.shape[0], -1)[:, :nbits].to(torch.float32)


def encode_packed(bits: torch.Tensor, n_groups: int, sigma: int, delta: float, u: torch.Tensor | None = None):
    """packed message -> secret tensor (B, n_groups) on the GPU (ideas_bits_encode)."""
    if not bits.is_cuda:
        raise RuntimeError("encode_packed needs CUDA tensors (no CPU fallback)")
    b = bits.shape[0]
    z = torch.empty((b, n_groups), dtype=torch.float32, device=bits.device)
    if u is not None:
        u = u.to(device=bits.device, dtype=torch.float32).contiguous()
    _lib.call("ideas_bits_encode", ptr(z), ptr(bits.contiguous()), ptr(u), b, n_groups, sigma, float(delta),
              stream_ptr(z))
    return z


def decode_packed(z: torch.Tensor, sigma: int) -> torch.Tensor:
    """secret tensor (B, L) -> packed message (B, ceil(sigma*L/32)) int32 (ideas_bits_decode)."""
    if not z.is_cuda:
        raise RuntimeError("decode_packed needs CUDA tensors (no CPU fallback)")
    z = z.to(torch.float32).contiguous()
    b, l = z.shape
    out = torch.empty((b, (l * sigma + 31) // 32), dtype=torch.int32, device=z.device)
    _lib.call("ideas_bits_decode", ptr(out), ptr(z), b, l, sigma, stream_ptr(z))
    return out


def bit_errors_packed(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """number of differing bits (device scalar, int64) via XOR + popcount (ideas_bits_count_errors)."""
    a, b = a.contiguous(), b.contiguous()
    cnt = torch.zeros(1, dtype=torch.int64, device=a.device)
    _lib.call("ideas_bits_count_errors", ptr(cnt), ptr(a), ptr(b), a.numel(), stream_ptr(a))
    return cnt[0]


def message_to_tensor(message, sigma, delta, device=None):
    """Reference signature (utils.py:74): (B, sigma*L) float 0/1 -> (B, L) secret tensor.  The
    jitter is drawn with ``torch.rand`` on the CPU generator exactly where the reference calls
    ``torch.rand_like``, so a seeded run produces the same tensor bit for bit."""
    device = torch.device(device if device is not None else ("cuda" if not message.is_cuda else message.device))
    n_groups = message.shape[1] // sigma
    u = torch.rand(message.shape[0], n_groups)
    bits = pack_message(message[:, :n_groups * sigma].to(device))
    return encode_packed(bits, n_groups, sigma, delta, u)


def tensor_to_message(secret_tensor, sigma):
    """Reference signature (utils.py:86): (B, L) tensor -> (B, sigma*L) float 0/1 message
    (returned on the tensor's device; the packed form is ``decode_packed``)."""
    words = decode_packed(secret_tensor, sigma)
    return unpack_message(words, secret_tensor.shape[1] * sigma)


def bit_error_rate(message_bits: torch.Tensor, decoded_bits: torch.Tensor, nbits_total: int) -> float:
    """BER = mean |M - M_hat| (train.py:285) from packed words."""
    return float(bit_errors_packed(message_bits, decoded_bits).item()) / float(nbits_total)
