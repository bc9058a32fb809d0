"""Per-shape comparison of the tcgen05 forward-form kernels through the C ABI: current dispatch without the
pixel-major halo kernel (pmh=0) vs pmh forced wherever eligible (pmh=2).  Shapes are the low-channel and strided
layers of one training step (scripts/profile_convs.py).  Usage: python scripts/bench_pmh.py > gpurun_out/bench_pmh.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200._tensor import ptr, stream_ptr

dev = torch.device("cuda")
if os.environ.get("PMH_PROBE"):
    # compute-dense, L2-resident shapes: what MMA rate does the pixel-major halo kernel sustain at N = 64 / 32 / 128 ?
    PROBE = [("forward", 32, 128, 64, 64, 3, 1, 1), ("forward", 32, 128, 32, 64, 3, 1, 1), ("forward", 32, 128, 128, 64, 3, 1, 1),
             ("forward", 32, 128, 64, 64, 1, 1, 0), ("forward", 32, 32, 64, 258, 3, 1, 0)]
CASES = [
    # kind, N, C, K, H, k, stride, pad
    ("forward", 32, 32, 64, 258, 3, 1, 0), ("dgrad", 32, 32, 64, 258, 3, 1, 0),
    ("forward", 1024, 32, 64, 64, 3, 1, 1), ("dgrad", 1024, 32, 64, 64, 3, 1, 1),
    ("forward", 96, 64, 128, 256, 3, 1, 1), ("dgrad", 96, 64, 128, 256, 3, 1, 1),
    ("forward", 32, 64, 128, 130, 3, 1, 0), ("dgrad", 32, 64, 128, 130, 3, 1, 0),
    ("forward", 32, 128, 128, 256, 3, 1, 1), ("dgrad", 32, 128, 128, 256, 3, 1, 1),
    ("forward", 32, 256, 256, 128, 3, 1, 1), ("forward", 32, 512, 512, 64, 3, 1, 1),
    ("dgrad", 32, 128, 256, 257, 3, 2, 0), ("dgrad", 96, 128, 128, 257, 3, 2, 0), ("dgrad", 32, 64, 64, 257, 3, 2, 0),
    ("dgrad", 32, 256, 512, 129, 3, 2, 0), ("dgrad", 32, 512, 512, 65, 3, 2, 0), ("dgrad", 1024, 64, 64, 65, 3, 2, 0),
    ("forward", 32, 128, 256, 16, 3, 1, 1), ("dgrad", 32, 128, 256, 16, 3, 1, 1), ("forward", 96, 512, 512, 16, 3, 1, 1),
]
iters = int(os.environ.get("ITERS", "5"))
if os.environ.get("PMH_PROBE"):
    CASES = PROBE
print(f"{'kind':8s} {'N':>5s} {'C':>4s} {'K':>4s} {'H':>4s} k s p | pmh=0 ms  TF/s | pmh=2 ms  TF/s | speed-up")
for kind, N, C, K, H, k, s, pad in CASES:
    OH = (H + 2 * pad - k) // s + 1
    x = torch.randn(N, H, H, C, device=dev)
    y = torch.randn(N, OH, OH, K, device=dev)
    wp = torch.randn(k * k, K, C, device=dev) / (C * k * k) ** 0.5
    wpt = torch.randn(k * k, C, K, device=dev) / (K * k * k) ** 0.5
    b = torch.randn(K, device=dev)
    st = stream_ptr(x)
    if kind == "forward":
        fn = lambda: _lib.call("ideas_conv2d_forward", ptr(y), ptr(x), ptr(wp), ptr(None), ptr(None), ptr(b), N, H, H, C, K, k, k,  # noqa: E731
                               s, pad, 1, 0.2, 2 ** 0.5, 0, st)
    else:
        fn = lambda: _lib.call("ideas_conv2d_dgrad", ptr(x), ptr(y), ptr(wpt), ptr(None), ptr(None), ptr(None), N, H, H, C, K, k, k,  # noqa: E731
                               s, pad, OH, OH, 0, 0.2, 1.0, 0, st)
    flops = 2.0 * N * OH * OH * K * C * k * k
    res = []
    for mode in (0, 2):
        _lib.call("ideas_set_option", b"pmh", mode)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        res.append(a.elapsed_time(e) / iters)
    _lib.call("ideas_set_option", b"pmh", 1)
    print(f"{kind:8s} {N:5d} {C:4d} {K:4d} {H:4d} {k} {s} {pad} | {res[0]:7.3f} {flops / res[0] / 1e9:6.1f} | "
          f"{res[1]:7.3f} {flops / res[1] / 1e9:6.1f} | {res[0] / res[1]:5.2f}x", flush=True)
    del x, y
    torch.cuda.empty_cache()
