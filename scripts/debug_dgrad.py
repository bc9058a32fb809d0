"""Debug helper: compare the autograd-path dgrad (tcgen05) against the FFMA path and localise errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ideas_b200 import _lib
from ideas_b200.stylegan2.op import conv as C
from ideas_b200._tensor import ptr, stream_ptr

torch.manual_seed(0)


def run(N, Ci, K, H, k, stride, pad, impl):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, Ci, H, H, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(K, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).cuda().requires_grad_(True)
    old = C.set_default_impl(impl)
    wp = C.PackWeight.apply(w, False, 1.0)
    y = C.conv2d(x, wp, None, K=K, kh=k, kw=k, stride=stride, pad=pad)
    gy = torch.randn(y.shape, generator=g).cuda()
    gx, gw = torch.autograd.grad(y, [x, w], gy)
    C.set_default_impl(old)
    return y.detach(), gx, gw


for case in [(2, 32, 64, 18, 3, 1, 0), (2, 64, 64, 16, 3, 1, 1), (1, 128, 128, 33, 3, 2, 0)]:
    ref = run(*case, _lib.IMPL_SIMT)
    for rep in range(2):
        got = run(*case, _lib.IMPL_AUTO)
        for name, a, b in zip(("y", "gx", "gw"), got, ref):
            d = (a - b).abs()
            print(case, rep, name, "max", float(d.max()), "ref max", float(b.abs().max()))
            if name == "gx" and float(d.max()) > 1e-2 * float(b.abs().max()):
                bad = (d > 1e-2 * b.abs().max())
                idx = bad.nonzero()
                print("  bad count", int(bad.sum()), "of", bad.numel())
                for dim, nm in enumerate("nchw"):
                    print("   ", nm, sorted(set(idx[:, dim].tolist()))[:40])
# direct C-ABI dgrad with and without in_scale on the first case geometry
N, Cc, K, H, k, stride, pad = 2, 32, 64, 18, 3, 1, 0
OH = (H + 2 * pad - k) // stride + 1
dy = torch.randn(N, OH, OH, K, device="cuda")
wpt = torch.randn(k * k, Cc, K, device="cuda") / 24
s = torch.rand(N, Cc, device="cuda") + 0.5
for scale in (None, s):
    outs = []
    for impl in (_lib.IMPL_SIMT, _lib.IMPL_AUTO):
        dx = torch.full((N, H, H, Cc), float("nan"), device="cuda")
        _lib.call("ideas_conv2d_dgrad", ptr(dx), ptr(dy), ptr(wpt), ptr(scale), ptr(None), ptr(None), N, H, H, Cc, K, k, k, stride,
                  pad, OH, OH, 0, 0.2, 1.0, impl, stream_ptr(dy))
        outs.append(dx)
    d = (outs[0] - outs[1]).abs()
    print("direct dgrad in_scale", scale is not None, "max diff", float(d.max()), "nan", int(torch.isnan(outs[1]).sum()))
    bad = (d > 1e-2).nonzero()
    if len(bad):
        for dim, nm in enumerate("nhwc"):
            print("   ", nm, sorted(set(bad[:, dim].tolist()))[:40])
