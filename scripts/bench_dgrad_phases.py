"""Strided data gradients: one halo launch per output phase (dgrad_phases=0) vs all four phases in one launch with
phase-interleaved items (dgrad_phases=2), on the stride-2 shapes of a 256x256 training step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200._tensor import ptr, stream_ptr

dev = torch.device("cuda")
# N, C (dx channels), K (dy channels), H (dx size)
CASES = [(32, 128, 256, 129), (96, 128, 256, 129), (32, 256, 512, 65), (96, 256, 512, 65), (32, 512, 512, 33),
         (96, 512, 512, 33), (32, 64, 128, 257), (32, 128, 128, 257)]
print(f"{'N':>5s} {'C':>4s} {'K':>4s} {'H':>4s} | per-phase ms  TF/s | one launch ms  TF/s | speed-up")
for N, C, K, H in CASES:
    k, s, pad = 3, 2, 0
    OH = (H + 2 * pad - k) // s + 1
    x = torch.empty(N, H, H, C, device=dev)
    y = torch.randn(N, OH, OH, K, device=dev)
    wpt = torch.randn(k * k, C, K, device=dev) / (K * k * k) ** 0.5
    st = stream_ptr(x)
    fn = lambda: _lib.call("ideas_conv2d_dgrad", ptr(x), ptr(y), ptr(wpt), ptr(None), ptr(None), ptr(None), N, H, H, C, K, k, k,  # noqa: E731
                           s, pad, OH, OH, 0, 0.2, 1.0, 0, st)
    flops = 2.0 * N * OH * OH * K * C * k * k
    res = []
    for mode in (0, 2, -1):
        # -1: halo kernel off -> one launch per phase of the pixel-major kernels (CTA pairs for 256-channel tiles)
        _lib.call("ideas_set_option", b"halo", 0 if mode < 0 else 1)
        _lib.call("ideas_set_option", b"dgrad_phases", max(mode, 0))
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        e.record()
        torch.cuda.synchronize()
        res.append(a.elapsed_time(e) / 5)
    print(f"{N:5d} {C:4d} {K:4d} {H:4d} | {res[0]:7.3f} {flops / res[0] / 1e9:6.1f} | "
          f"{res[1]:7.3f} {flops / res[1] / 1e9:6.1f} | {res[0] / res[1]:5.2f}x | pixel-major per phase {res[2]:7.3f} "
          f"{flops / res[2] / 1e9:6.1f}", flush=True)
    del x, y
    torch.cuda.empty_cache()
_lib.call("ideas_set_option", b"dgrad_phases", 1)
_lib.call("ideas_set_option", b"halo", 1)
