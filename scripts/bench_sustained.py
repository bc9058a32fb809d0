"""Power-limited regime: the same wide conv launched back to back for ~0.6 s per variant (the burst numbers of
bench_pair.py last a few ms).  Variants: one-CTA pixel-major kernel, CTA pairs (cta_group::2), channel-major halo
kernel, pixel-major halo kernel.  Reports TFLOP/s over the last 0.4 s and the SM clock nvidia-smi saw meanwhile."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200._tensor import ptr, stream_ptr

dev = torch.device("cuda")
VARIANTS = [("pm one CTA", dict(pair=0, halo=0, pmh=0)), ("pm CTA pair", dict(pair=1, halo=0, pmh=0)),
            ("halo (channel-major)", dict(pair=0, halo=2, pmh=0)), ("pmh (pixel-major halo)", dict(pair=0, halo=0, pmh=2))]
if len(sys.argv) > 1 and sys.argv[1] == "pair":
    VARIANTS = VARIANTS[:2]
CASES = [("forward", 16, 512, 512, 64), ("forward", 32, 256, 256, 128), ("dgrad", 16, 512, 512, 64)]


def clock():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        return out[0].strip() + " MHz " + out[1].strip() + " W"
    except Exception:
        return "?"


for kind, N, C, K, H in CASES:
    x = torch.randn(N, H, H, C, device=dev)
    y = torch.randn(N, H, H, K, device=dev)
    wp = torch.randn(9, K, C, device=dev) / (C * 9) ** 0.5
    b = torch.randn(K, device=dev)
    d = torch.rand(N, K, device=dev) + 0.5
    st = stream_ptr(x)
    if kind == "forward":
        fn = lambda: _lib.call("ideas_conv2d_forward", ptr(y), ptr(x), ptr(wp), ptr(None), ptr(d), ptr(b), N, H, H, C, K, 3, 3,  # noqa: E731
                               1, 1, 1, 0.2, 2 ** 0.5, 0, st)
    else:
        fn = lambda: _lib.call("ideas_conv2d_dgrad", ptr(x), ptr(y), ptr(wp), ptr(None), ptr(None), ptr(None), N, H, H, C, K, 3, 3,  # noqa: E731
                               1, 1, H, H, 0, 0.2, 1.0, 0, st)
    flops = 2.0 * N * H * H * K * C * 9
    for name, opts in VARIANTS:
        for k, v in opts.items():
            _lib.call("ideas_set_option", k.encode(), v)
        fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        e.record()
        torch.cuda.synchronize()
        t1 = a.elapsed_time(e) / 5
        n_burn = max(8, int(200.0 / t1))
        for _ in range(n_burn):
            fn()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = max(8, int(400.0 / t1))
        for _ in range(n):
            fn()
        e.record()
        time.sleep(0.25)
        c = clock()
        torch.cuda.synchronize()
        t = a.elapsed_time(e) / n
        print(f"{kind:8s} {N:3d}x{C}x{H}x{H}->{K} {name:24s} burst {flops / t1 / 1e9:6.1f}  sustained {flops / t / 1e9:6.1f} TF/s  ({c})",
              flush=True)
        time.sleep(1.0)
    del x, y
    torch.cuda.empty_cache()
for k, v in dict(pair=1, halo=1, pmh=1).items():
    _lib.call("ideas_set_option", k.encode(), v)
