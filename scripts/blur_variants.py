"""Blur kernel tuning variants (option blur_variant: 0 = 32 rows x 2 cols/thread (default), 1 = 32x4, 2 = 64x4, 3 = 64x2) and a
layout probe of torch's reflection_pad2d on channels_last input."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200._tensor import ptr, stream_ptr

dev = torch.device("cuda")
kk = torch.tensor([1., 3., 3., 1.], device=dev)
kk = torch.outer(kk, kk)
kk = (kk / kk.sum()).contiguous()


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


for (B, C, H, pd) in ((32, 128, 256, 2), (32, 64, 256, 1), (32, 256, 128, 2), (96, 512, 32, 2), (32, 32, 256, 2)):
    x = torch.randn(B, H, H, C, device=dev)
    Ho = H + 2 * pd - 3
    y = torch.empty(B, Ho, Ho, C, device=dev)
    res = []
    for v in range(4):
        _lib.call("ideas_set_option", b"blur_variant", v)
        t = timeit(lambda: _lib.call("ideas_upfirdn2d", ptr(y), ptr(x), ptr(kk), B, H, H, C, 4, 4, 1, 1, 1, 1, pd, pd, pd, pd, ptr(None),
                                     0.2, 1.0, stream_ptr(x)))
        res.append(4.0 * (x.numel() + y.numel()) / t / 1e9)
    _lib.call("ideas_set_option", b"blur_variant", 0)
    print(f"blur {B}x{C}x{H}^2 pad{pd}: " + "  ".join(f"v{v} {r:6.0f} GB/s" for v, r in enumerate(res)), flush=True)
    del x, y

# fused blur^T * activation-mask backward (ideas_blur_act_backward): bytes = gz read + y read + gx write
for (B, C, H) in ((32, 128, 256), (32, 256, 128), (32, 64, 256), (96, 512, 64)):
    gz = torch.randn(B, H + 1, H + 1, C, device=dev)
    yv = torch.randn(B, H, H, C, device=dev)
    gx = torch.empty_like(yv)
    gb = torch.zeros(C, device=dev)
    for name, gbp in (("no bias grad", None), ("bias grad", gb)):
        t = timeit(lambda: _lib.call("ideas_blur_act_backward", ptr(gx), ptr(gbp), ptr(gz), ptr(yv), ptr(kk), B, H + 1, H + 1, C, 4, 4,
                                     1, 1, 1, 1, 0.2, 2 ** 0.5, stream_ptr(gz)))
        print(f"blur_act_backward {B}x{C}x{H}^2 ({name}): {t * 1e3:.3f} ms  {4.0 * (gz.numel() + 2 * yv.numel()) / t / 1e9:6.0f} GB/s", flush=True)
    del gz, yv, gx

x = torch.randn(4, 32, 64, 64, device=dev).contiguous(memory_format=torch.channels_last)
p = torch.nn.functional.pad(x, (1, 1, 1, 1), mode="reflect")
print("reflection_pad2d(channels_last input): output channels_last-contiguous =", p.is_contiguous(memory_format=torch.channels_last),
      " NCHW-contiguous =", p.is_contiguous())
xg = x.clone().requires_grad_(True)
pg = torch.nn.functional.pad(xg, (1, 1, 1, 1), mode="reflect")
(g,) = torch.autograd.grad(pg, xg, torch.randn_like(pg))
print("its gradient: channels_last =", g.is_contiguous(memory_format=torch.channels_last), " NCHW =", g.is_contiguous())
