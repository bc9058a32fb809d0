"""Fused resampling kernels (upfirdn2d up=2 / down=2, 4x4 binomial taps) on the shapes of a 256x256 training step,
strip-shape variants (option resample_variant): GB/s of algorithmic traffic."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200._tensor import ptr, stream_ptr

dev = torch.device("cuda")
k = torch.tensor([1., 3., 3., 1.], device=dev)
k = torch.outer(k, k)
k = (k / k.sum()).contiguous()
# kind, N, C, H (input size)
CASES = [("down2", 32, 32, 256), ("down2", 32, 64, 128), ("down2", 32, 128, 64), ("down2", 96, 64, 256), ("down2", 96, 128, 128),
         ("up2", 32, 32, 128), ("up2", 32, 64, 64), ("up2", 32, 128, 128), ("up2", 96, 64, 128), ("up2res", 32, 32, 128),
         ("up2res", 96, 64, 128), ("up2res", 32, 128, 128)]
NV = 7
print(f"{'kind':7s} {'N':>3s} {'C':>4s} {'H':>4s} | " + " | ".join(f"v{v} GB/s" for v in range(NV)))
for kind, N, C, H in CASES:
    x = torch.randn(N, H, H, C, device=dev)
    if kind == "down2":
        OH = (H + 2 - 4) // 2 + 1              # pad (1, 1)
        out = torch.empty(N, OH, OH, C, device=dev)
        args = (1, 1, 2, 2, 1, 1, 1, 1)
        res = None
    else:
        OH = 2 * H                             # up = 2, pad (2, 1): the adjoint / the up-sampling skip
        out = torch.empty(N, OH, OH, C, device=dev)
        args = (2, 2, 1, 1, 2, 1, 2, 1)
        res = torch.randn_like(out) if kind == "up2res" else None
    st = stream_ptr(x)
    fn = lambda: _lib.call("ideas_upfirdn2d_res", ptr(out), ptr(x), ptr(k), N, H, H, C, 4, 4, *args, ptr(None), 0.2, 1.0,  # noqa: E731
                           ptr(res), 1.0, st)
    nbytes = 4.0 * (x.numel() + out.numel() * (2 if res is not None else 1))
    cells = []
    for v in range(NV):
        _lib.call("ideas_set_option", b"resample_variant", v)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        cells.append(nbytes / (a.elapsed_time(e) / 10 * 1e-3) / 1e9)
    print(f"{kind:7s} {N:3d} {C:4d} {H:4d} | " + " | ".join(f"{c:7.0f}" for c in cells), flush=True)
    del x, out, res
    torch.cuda.empty_cache()
_lib.call("ideas_set_option", b"resample_variant", 0)
