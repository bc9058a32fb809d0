"""Micro-benchmarks of the hot kernels through the C ABI (BASELINE.json configs[1] and configs[2]):
cfg-3 modulated-conv shape fwd / dgrad / wgrad, blur, bias+lrelu fwd/bwd.  Prints one JSON line per
kernel with achieved TFLOP/s or GB/s.  Also the command ncu wraps for profiles/."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200._tensor import ptr, stream_ptr

iters = int(os.environ.get("ITERS", "10"))
dev = torch.device("cuda")
if "IDEAS_HALO" in os.environ:
    _lib.call("ideas_set_option", b"halo", int(os.environ["IDEAS_HALO"]))


def timeit(fn, n=iters, warm=2):
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


def conv_suite(N, C, K, H, k=3, stride=1, pad=1, tag=""):
    OH = (H + 2 * pad - k) // stride + 1
    x = torch.randn(N, H, H, C, device=dev)
    wp = torch.randn(k * k, K, C, device=dev) / (C * k * k) ** 0.5
    wpt = torch.randn(k * k, C, K, device=dev) / (K * k * k) ** 0.5
    d = torch.rand(N, K, device=dev) + 0.5
    bias = torch.randn(K, device=dev)
    y = torch.empty(N, OH, OH, K, device=dev)
    dx = torch.empty_like(x)
    dwp = torch.zeros(k * k, K, C, device=dev)
    st = stream_ptr(x)
    flops = 2.0 * N * OH * OH * K * C * k * k
    for name, fn in (
        ("fwd", lambda: _lib.call("ideas_conv2d_forward", ptr(y), ptr(x), ptr(wp), ptr(None), ptr(d), ptr(bias), N, H, H, C, K, k, k,
                                  stride, pad, 1, 0.2, 2 ** 0.5, 0, st)),
        ("dgrad", lambda: _lib.call("ideas_conv2d_dgrad", ptr(dx), ptr(y), ptr(wpt), ptr(None), ptr(None), ptr(None), N, H, H, C, K, k, k,
                                    stride, pad, OH, OH, 0, 0.2, 1.0, 0, st)),
        ("wgrad", lambda: _lib.call("ideas_conv2d_wgrad", ptr(dwp), ptr(x), ptr(y), ptr(None), ptr(None), N, H, H, C, K, k, k, stride, pad,
                                    OH, OH, 0, st)),
    ):
        t = timeit(fn)
        print(json.dumps({"kernel": f"conv_{name}{tag}", "shape": [N, C, K, H, k, stride, pad], "ms": t * 1e3, "tflops": flops / t / 1e12}))


conv_suite(16, 512, 512, 64, tag="_cfg3")
conv_suite(32, 128, 128, 256, tag="_g256")
conv_suite(32, 64, 128, 258, pad=0, tag="_d256")
conv_suite(32, 128, 256, 129, stride=2, pad=0, tag="_down")
conv_suite(32, 512, 512, 16, tag="_g16")
conv_suite(32, 256, 256, 128, tag="_g128")
conv_suite(32, 512, 512, 32, tag="_g32")
conv_suite(32, 256, 512, 64, tag="_d64")
if os.environ.get("CONV_ONLY"):
    sys.exit(0)

# blur (cfg 2)
B, Cb, Hb = 32, 128, 256
xb = torch.randn(B, Hb, Hb, Cb, device=dev)
yb = torch.empty(B, Hb + 1, Hb + 1, Cb, device=dev)
kk = torch.tensor([1., 3., 3., 1.], device=dev)
kk = torch.outer(kk, kk)
kk = (kk / kk.sum()).contiguous()
t = timeit(lambda: _lib.call("ideas_upfirdn2d", ptr(yb), ptr(xb), ptr(kk), B, Hb, Hb, Cb, 4, 4, 1, 1, 1, 1, 2, 2, 2, 2, ptr(None), 0.2, 1.0,
                             stream_ptr(xb)), n=20)
print(json.dumps({"kernel": "blur_pad22", "ms": t * 1e3, "gbs": 4.0 * (xb.numel() + yb.numel()) / t / 1e9}))
y2 = torch.empty(B, Hb - 1, Hb - 1, Cb, device=dev)
t = timeit(lambda: _lib.call("ideas_upfirdn2d", ptr(y2), ptr(xb), ptr(kk), B, Hb, Hb, Cb, 4, 4, 1, 1, 1, 1, 1, 1, 1, 1, ptr(None), 0.2, 1.0,
                             stream_ptr(xb)), n=20)
print(json.dumps({"kernel": "blur_pad11", "ms": t * 1e3, "gbs": 4.0 * (xb.numel() + y2.numel()) / t / 1e9}))
for (Bq, Cq, Hq, pd) in ((32, 64, 256, 1), (32, 256, 128, 2), (96, 512, 32, 2), (32, 32, 256, 2)):
    xq = torch.randn(Bq, Hq, Hq, Cq, device=dev)
    Ho = Hq + 2 * pd - 3
    yq = torch.empty(Bq, Ho, Ho, Cq, device=dev)
    t = timeit(lambda: _lib.call("ideas_upfirdn2d", ptr(yq), ptr(xq), ptr(kk), Bq, Hq, Hq, Cq, 4, 4, 1, 1, 1, 1, pd, pd, pd, pd, ptr(None),
                                 0.2, 1.0, stream_ptr(xq)), n=20)
    print(json.dumps({"kernel": f"blur_{Bq}x{Cq}x{Hq}_pad{pd}", "ms": t * 1e3, "gbs": 4.0 * (xq.numel() + yq.numel()) / t / 1e9}))
    del xq, yq
# bias + lrelu forward / backward (cfg 2 iii)
bias = torch.randn(Cb, device=dev)
out = torch.empty_like(xb)
n = xb.numel()
t = timeit(lambda: _lib.call("ideas_fused_bias_act", ptr(out), ptr(xb), ptr(bias), ptr(None), 3, 0, 0.2, 2 ** 0.5, n, 1, Cb, stream_ptr(xb)), n=20)
print(json.dumps({"kernel": "bias_act_fwd", "ms": t * 1e3, "gbs": 8.0 * n / t / 1e9}))
gx = torch.empty_like(xb)
gb = torch.zeros(Cb, device=dev)
gy = torch.randn_like(xb)
t = timeit(lambda: _lib.call("ideas_bias_act_backward", ptr(gx), ptr(gb), ptr(gy), ptr(out), 0.2, 2 ** 0.5, n, 1, Cb, stream_ptr(xb)), n=20)
print(json.dumps({"kernel": "bias_act_bwd", "ms": t * 1e3, "gbs": 12.0 * n / t / 1e9}))
# a plain device copy of the same size for reference
t = timeit(lambda: out.copy_(xb), n=20)
print(json.dumps({"kernel": "torch_copy", "ms": t * 1e3, "gbs": 8.0 * n / t / 1e9}))
