"""Pixel-major forward-form kernel, 256-channel tiles: one CTA per tile (pair=0) vs CTA pairs sharing the weight tile
through tcgen05 cta_group::2 (pair=1), on the wide layers of a 256x256 training step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200._tensor import ptr, stream_ptr

dev = torch.device("cuda")
MODES = (0, 1)
CASES128 = [("forward", 32, 128, 128, 128, 3, 1, 1), ("forward", 32, 256, 384, 16, 3, 1, 1), ("forward", 32, 384, 384, 16, 3, 1, 1),
            ("forward", 32, 64, 128, 129, 3, 2, 0), ("forward", 32, 256, 128, 256, 1, 1, 0), ("forward", 96, 128, 128, 64, 3, 1, 1),
            ("dgrad", 32, 128, 128, 128, 3, 1, 1), ("dgrad", 32, 384, 384, 16, 3, 1, 1), ("dgrad", 96, 128, 128, 64, 3, 1, 1)]
CASES = [("forward", 16, 512, 512, 64, 3, 1, 1), ("forward", 32, 512, 512, 64, 3, 1, 1), ("forward", 32, 256, 256, 128, 3, 1, 1),
         ("forward", 32, 512, 512, 32, 3, 1, 1), ("forward", 32, 512, 512, 16, 3, 1, 1), ("forward", 32, 256, 512, 65, 3, 2, 0),
         ("forward", 32, 128, 256, 129, 3, 2, 0), ("forward", 96, 256, 256, 32, 3, 1, 1),
         ("dgrad", 32, 512, 512, 64, 3, 1, 1), ("dgrad", 32, 256, 256, 128, 3, 1, 1), ("dgrad", 32, 512, 512, 32, 3, 1, 1)]
if len(sys.argv) > 1 and sys.argv[1] == "128":       # 128-channel tiles: pair mode 2 vs 0, halo / pmh heuristics off
    CASES, MODES = CASES128, (0, 2)
    _lib.call("ideas_set_option", b"halo", 0)
    _lib.call("ideas_set_option", b"pmh", 0)
print(f"{'kind':8s} {'N':>5s} {'C':>4s} {'K':>4s} {'H':>4s} k s p | pair=0 ms  TF/s | pair=1 ms  TF/s | speed-up")
for kind, N, C, K, H, k, s, pad in CASES:
    OH = (H + 2 * pad - k) // s + 1
    x = torch.randn(N, H, H, C, device=dev)
    y = torch.randn(N, OH, OH, K, device=dev)
    wp = torch.randn(k * k, K, C, device=dev) / (C * k * k) ** 0.5
    wpt = torch.randn(k * k, C, K, device=dev) / (K * k * k) ** 0.5
    b = torch.randn(K, device=dev)
    d = torch.rand(N, K, device=dev) + 0.5
    st = stream_ptr(x)
    if kind == "forward":
        fn = lambda: _lib.call("ideas_conv2d_forward", ptr(y), ptr(x), ptr(wp), ptr(None), ptr(d), ptr(b), N, H, H, C, K, k, k,  # noqa: E731
                               s, pad, 1, 0.2, 2 ** 0.5, 0, st)
    else:
        fn = lambda: _lib.call("ideas_conv2d_dgrad", ptr(x), ptr(y), ptr(wpt), ptr(None), ptr(None), ptr(None), N, H, H, C, K, k, k,  # noqa: E731
                               s, pad, OH, OH, 0, 0.2, 1.0, 0, st)
    flops = 2.0 * N * OH * OH * K * C * k * k
    res = []
    for mode in MODES:
        _lib.call("ideas_set_option", b"pair", mode)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(8):
            fn()
        e.record()
        torch.cuda.synchronize()
        res.append(a.elapsed_time(e) / 8)
    print(f"{kind:8s} {N:5d} {C:4d} {K:4d} {H:4d} {k} {s} {pad} | {res[0]:7.3f} {flops / res[0] / 1e9:6.1f} | "
          f"{res[1]:7.3f} {flops / res[1] / 1e9:6.1f} | {res[0] / res[1]:5.2f}x", flush=True)
    del x, y
    torch.cuda.empty_cache()
