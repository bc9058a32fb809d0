"""GPU-side comparators for bench.py (`library_baseline` key): the SAME GPU, the same inputs, library kernels instead
of ours.  BENCH INFRASTRUCTURE -- never imported by the product.

  cfg 2  (BASELINE.json configs[1])  Blur pad (2,2) and FusedLeakyReLU on (B,128,256,256), B in {1, 32, 96}:
         ours (C ABI, NHWC)  vs  the reference's own CUDA operators K1 / K2 recompiled for sm_100a
         (oracle/_ref/*.so, built by oracle/build_ref.py from /root/reference/stylegan2/op; NCHW as the reference
         calls them)  vs  the torch-native formulation (depthwise F.conv2d / leaky_relu).
  cfg 3  (configs[2])  ModulatedConv2d(512,512,3) fwd and fwd+bwd on x (16,512,64,64): ours (the nn.Module) vs the
         reference's formulation -- per-sample weights w = scale*W*s, demodulated, one grouped F.conv2d with
         groups = B (stylegan2/model.py:239-275) -- on cuDNN with cudnn.benchmark = True (train.py:327),
         cudnn.allow_tf32 on (torch default) and off.
  cfg 4  (configs[3])  the unmodified reference's train step on this GPU (scripts/reference_step.py) when a
         reference tree is reachable; otherwise the committed result of the staged run is quoted.
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def cfg2(batches=(1, 32, 96)):
    import torch
    import torch.nn.functional as F
    from ideas_b200 import _lib
    from ideas_b200._tensor import ptr, stream_ptr
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    ref = build_ref.load()
    dev = torch.device("cuda")
    C, H = 128, 256
    k1 = torch.tensor([1., 3., 3., 1.], device=dev)
    k = torch.outer(k1, k1)
    k = (k / k.sum()).contiguous()
    out = {"ref_ops": "oracle/_ref (reference K1/K2 recompiled for sm_100a)" if ref else "not built on this box"}
    for B in batches:
        row = {}
        n_in, n_out = B * C * H * H, B * C * (H + 1) * (H + 1)
        blur_bytes = 4.0 * (n_in + n_out)
        x = torch.randn(B, H, H, C, device=dev)                      # ours: NHWC
        y = torch.empty(B, H + 1, H + 1, C, device=dev)
        t = _time(lambda: _lib.call("ideas_upfirdn2d", ptr(y), ptr(x), ptr(k), B, H, H, C, 4, 4, 1, 1, 1, 1, 2, 2, 2, 2,
                                    ptr(None), 0.2, 1.0, stream_ptr(x)))
        row["blur_ours_gbs"] = blur_bytes / t / 1e9
        xn = torch.randn(B, C, H, H, device=dev)                     # reference layout: NCHW
        if ref:
            t = _time(lambda: ref[1].upfirdn2d(xn.reshape(-1, H, H, 1), k, 1, 1, 1, 1, 2, 2, 2, 2))
            row["blur_reference_op_gbs"] = blur_bytes / t / 1e9
        wdw = torch.flip(k, [0, 1]).view(1, 1, 4, 4)
        t = _time(lambda: F.conv2d(F.pad(xn.view(-1, 1, H, H), [2, 2, 2, 2]), wdw))
        row["blur_torch_native_gbs"] = blur_bytes / t / 1e9
        del y
        # FusedLeakyReLU forward / backward
        bias = torch.randn(C, device=dev)
        n = n_in
        ya = torch.empty_like(x)
        t = _time(lambda: _lib.call("ideas_fused_bias_act", ptr(ya), ptr(x), ptr(bias), ptr(None), 3, 0, 0.2, 2 ** 0.5, n, 1, C,
                                    stream_ptr(x)))
        row["lrelu_fwd_ours_gbs"] = (8.0 * n + 4 * C) / t / 1e9
        gb = torch.zeros(C, device=dev)
        gy = torch.randn_like(x)
        t = _time(lambda: _lib.call("ideas_bias_act_backward", ptr(ya), ptr(gb), ptr(gy), ptr(x), 0.2, 2 ** 0.5, n, 1, C,
                                    stream_ptr(x)))
        row["lrelu_bwd_ours_gbs"] = (12.0 * n + 4 * C) / t / 1e9
        if ref:
            empty = xn.new_empty(0)
            t = _time(lambda: ref[0].fused_bias_act(xn, bias, empty, 3, 0, 0.2, 2 ** 0.5))
            row["lrelu_fwd_reference_op_gbs"] = (8.0 * n + 4 * C) / t / 1e9
            # the reference's backward is two passes: the masked gradient (act=3, grad=1) and the bias-grad .sum
            # (fused_act.py:29-38); same algorithmic bytes credited as for ours
            gyn = torch.randn_like(xn)
            t = _time(lambda: (ref[0].fused_bias_act(gyn, empty, xn, 3, 1, 0.2, 2 ** 0.5).sum(dim=(0, 2, 3))))
            row["lrelu_bwd_reference_op_gbs"] = (12.0 * n + 4 * C) / t / 1e9
        t = _time(lambda: F.leaky_relu(xn + bias.view(1, -1, 1, 1), 0.2) * 2 ** 0.5)
        row["lrelu_fwd_torch_native_gbs"] = (8.0 * n + 4 * C) / t / 1e9
        out[f"B{B}"] = {kk: round(v, 1) for kk, v in row.items()}
        del x, xn, ya, gy
        torch.cuda.empty_cache()
    return out


def _ref_modconv(x, style, W, mod_w, mod_b):
    """The reference's formulation of ModulatedConv2d.forward (stylegan2/model.py:236-277, same-resolution branch)."""
    import torch
    import torch.nn.functional as F
    B, cin, H, Wd = x.shape
    cout, k = W.shape[1], W.shape[3]
    s = F.linear(style, mod_w * (1 / math.sqrt(style.shape[1])), mod_b)
    w = (1 / math.sqrt(cin * k * k)) * W * s.view(B, 1, cin, 1, 1)
    d = torch.rsqrt(w.pow(2).sum([2, 3, 4]) + 1e-8)
    w = (w * d.view(B, cout, 1, 1, 1)).view(B * cout, cin, k, k)
    return F.conv2d(x.reshape(1, B * cin, H, Wd), w, padding=k // 2, groups=B).view(B, cout, H, Wd)


def cfg3():
    import torch
    from ideas_b200.stylegan2.model import ModulatedConv2d
    dev = torch.device("cuda")
    B, C, H, SD = 16, 512, 64, 2048
    torch.manual_seed(0)
    flops = 2.0 * B * H * H * C * C * 9
    out = {"shape": "ModulatedConv2d(512,512,3,style_dim=2048), x (16,512,64,64), fwd 309.2 GFLOP, fwd+bwd 927.7"}
    m = ModulatedConv2d(C, C, 3, SD).to(dev)
    x = torch.randn(B, C, H, H, device=dev).to(memory_format=torch.channels_last).requires_grad_(True)
    style = (torch.rand(B, SD, device=dev) * 2 - 1).requires_grad_(True)
    gy = torch.randn(B, C, H, H, device=dev).to(memory_format=torch.channels_last)

    def ours_fwd():
        with torch.no_grad():
            m(x, style)

    def ours_fb():
        y = m(x, style)
        torch.autograd.grad(y, [x, style, m.weight, m.modulation.weight, m.modulation.bias], gy)

    t = _time(ours_fwd, iters=10)
    out["ours_fwd_ms"], out["ours_fwd_tflops"] = t * 1e3, flops / t / 1e12
    t = _time(ours_fb, iters=5)
    out["ours_fwd_bwd_ms"], out["ours_fwd_bwd_tflops"] = t * 1e3, 3 * flops / t / 1e12
    W = m.weight.detach().clone().requires_grad_(True)
    mw = m.modulation.weight.detach().clone().requires_grad_(True)
    mb = m.modulation.bias.detach().clone().requires_grad_(True)
    xr = x.detach().contiguous().requires_grad_(True)                       # NCHW, as the reference runs
    gyr = gy.contiguous()
    old_b, old_t = torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.benchmark = True
    try:
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            tag = "cudnn_tf32" if tf32 else "cudnn_fp32"

            def lib_fwd():
                with torch.no_grad():
                    _ref_modconv(xr, style, W, mw, mb)

            def lib_fb():
                y = _ref_modconv(xr, style, W, mw, mb)
                torch.autograd.grad(y, [xr, style, W, mw, mb], gyr)

            t = _time(lib_fwd, iters=5 if tf32 else 2, warm=2)
            out[f"{tag}_fwd_ms"], out[f"{tag}_fwd_tflops"] = t * 1e3, flops / t / 1e12
            t = _time(lib_fb, iters=3 if tf32 else 1, warm=2)
            out[f"{tag}_fwd_bwd_ms"], out[f"{tag}_fwd_bwd_tflops"] = t * 1e3, 3 * flops / t / 1e12
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = old_b, old_t
    out["speedup_fwd_vs_cudnn_tf32"] = out["cudnn_tf32_fwd_ms"] / out["ours_fwd_ms"]
    out["speedup_fwd_bwd_vs_cudnn_tf32"] = out["cudnn_tf32_fwd_bwd_ms"] / out["ours_fwd_bwd_ms"]
    return {k: (round(v, 3) if isinstance(v, float) else v) for k, v in out.items()}


def cfg4(line):
    """The unmodified reference's train step on this GPU, batch 32 (falling back to 16 / 8 if it does not fit)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import bench as _b  # noqa: F401  (find_reference_tree lives in bench.py)
    ref = _b.find_reference_tree()
    if ref is None:
        try:
            committed = json.load(open(os.path.join(ROOT, "profiles", "reference_gpu_r2.json")))
        except Exception:
            committed = None
        return {"available": False, "why": "no reference tree on this box (the reference is a Python project and does not "
                "travel with the repository); result of the staged run on a B200 quoted from profiles/reference_gpu_r2.json",
                "staged_run": committed}
    import reference_step
    import torch
    torch.cuda.empty_cache()
    for batch in (32, 16, 8):
        try:
            s, n = reference_step.time_reference(ref, device="cuda", batch=batch, steps=4, warmup=2, budget_s=120.0)
            out = {"available": True, "kind": "reference", "batch": batch, "ms_per_step": s * 1e3, "images_per_s": batch / s,
                   "timed_steps": n, "settings": "cudnn.benchmark=True, cudnn.allow_tf32=True (torch defaults, train.py:327)"}
            if line.get("value"):
                out["ours_over_reference"] = line["value"] / (batch / s)
            return out
        except Exception as e:
            last = f"{type(e).__name__}: {str(e)[-300:]}"
    return {"available": False, "why": "reference step failed at batch 32/16/8: " + last}


def run(line):
    out = {}
    for name, fn in (("cfg2", cfg2), ("cfg3", cfg3), ("cfg4", lambda: cfg4(line))):
        try:
            out[name] = fn()
        except Exception as e:           # a comparator must never take the bench line down
            out[name] = {"error": f"{type(e).__name__}: {str(e)[-400:]}"}
    return out


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    print(json.dumps(run({}), indent=1))
