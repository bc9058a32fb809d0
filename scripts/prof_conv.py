"""Time one convolution geometry through the C ABI under the kernel-selection options.
Usage: prof_conv.py kind N C K H k stride pad [iters]     kind = forward | dgrad | wgrad
Prints ms and TFLOP/s for halo = 0/1/2 (forward, dgrad) or wgrad_reuse = 0/1/2 (wgrad)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200._tensor import ptr, stream_ptr

for kv in filter(None, os.environ.get("IDEAS_OPTS", "").split(",")):      # e.g. IDEAS_OPTS=halo_resident=0
    name, val = kv.split("=")
    _lib.call("ideas_set_option", name.encode(), int(val))
kind = sys.argv[1]
N, C, K, H, k, s, pad = (int(v) for v in sys.argv[2:9])
iters = int(sys.argv[9]) if len(sys.argv) > 9 else 5
dev = torch.device("cuda")
OH = (H + 2 * pad - k) // s + 1
x = torch.randn(N, H, H, C, device=dev)
y = torch.randn(N, OH, OH, K, device=dev)
wp = torch.randn(k * k, K, C, device=dev) / (C * k * k) ** 0.5
wpt = torch.randn(k * k, C, K, device=dev) / (K * k * k) ** 0.5
dwp = torch.zeros(k * k, K, C, device=dev)
st = stream_ptr(x)
fns = {
    "forward": lambda: _lib.call("ideas_conv2d_forward", ptr(y), ptr(x), ptr(wp), ptr(None), ptr(None), ptr(None), N, H, H, C, K, k, k,
                                 s, pad, 0, 0.2, 1.0, 0, st),
    "dgrad": lambda: _lib.call("ideas_conv2d_dgrad", ptr(x), ptr(y), ptr(wpt), ptr(None), ptr(None), ptr(None), N, H, H, C, K, k, k,
                               s, pad, OH, OH, 0, 0.2, 1.0, 0, st),
    "wgrad": lambda: _lib.call("ideas_conv2d_wgrad", ptr(dwp), ptr(x), ptr(y), ptr(None), ptr(None), N, H, H, C, K, k, k, s, pad,
                               OH, OH, 0, st),
}
fn = fns[kind]
flops = 2.0 * N * OH * OH * K * C * k * k
opt = b"wgrad_reuse" if kind == "wgrad" else b"halo"
for mode in (0, 1, 2):
    _lib.call("ideas_set_option", opt, mode)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / iters
    print(f"{kind} N={N} C={C} K={K} H={H} k={k} s={s} p={pad} {opt.decode()}={mode}: {t:.4f} ms {flops / t / 1e9:7.1f} TFLOP/s", flush=True)
_lib.call("ideas_set_option", opt, 1)
