"""Tolerances of the GPU parity tests.

fp32 mode (FFMA kernels): everything is held to a few ulp-level multiples -- same arithmetic as the
reference up to summation order.

tf32 mode (tcgen05 kind::tf32, operands rounded to a 10-bit mantissa by the TMA unit, fp32
accumulation): forward outputs of ONE op are held to the north-star 1e-3 (max |a-b| / max |b|).
Gradients that pass through a leaky ReLU cannot be compared element-wise at that level: the backward
mask is the sign of the forward output, and a pre-activation within ~1e-3 of zero may legitimately
land on the other side (the reference itself shows this between its cuDNN-TF32 and fp32 runs).  Such a
flip changes one upstream gradient element by 0.8*sqrt(2)*|g| and is spread by the next dgrad over a
3x3xC neighbourhood.  So those gradients are checked with a flip-robust pair of metrics: the 95th
percentile of |a-b| / max|b| must meet the per-op tolerance, and the relative L2 error must stay small.
Whole networks (16+ stacked ops) get a proportionally wider output tolerance.
"""
import torch

OUT = {"fp32": 3e-5, "tf32": 1e-3}          # one op, forward
GRAD = {"fp32": 1e-4, "tf32": 3e-3}         # one op, backward (no activation mask in between)
NET = {"fp32": 1e-4, "tf32": 5e-3}          # whole network forward


def _auto(kind, v):
    import os
    t = os.environ.get("PYTEST_CURRENT_TEST", "")
    if t:
        record(t.split(" ")[0].split("::", 1)[-1] + ":" + kind, v)
    return v


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return _auto("rel", float((a - b).abs().max() / b.abs().max().clamp_min(1e-9)))


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return _auto("rel_l2", float((a - b).norm() / b.norm().clamp_min(1e-12)))


def rel_q(a, b, q=0.95):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = (a - b).abs().flatten()
    if d.numel() > 4_000_000:
        d = d[:: d.numel() // 4_000_000 + 1]
    return _auto("q95", float(torch.quantile(d, q) / b.abs().max().clamp_min(1e-9)))


def assert_grad_through_act(a, b, mode, what=""):
    """gradient comparison that tolerates leaky-ReLU mask flips in tf32 mode (see module docstring)."""
    if mode == "fp32":
        assert rel(a, b) <= GRAD["fp32"], (what, rel(a, b))
    elif a.numel() < 10000:
        # small reductions over all pixels (style, bias, per-sample scale gradients): every flip lands in them
        assert rel_l2(a, b) <= 5e-2, (what, "l2", rel_l2(a, b))
    else:
        assert rel_q(a, b) <= GRAD["tf32"], (what, "q95", rel_q(a, b))
        assert rel_l2(a, b) <= 5e-2, (what, "l2", rel_l2(a, b))


def record(name, value, tol=None):
    """Append a measured error to gpurun_out/parity_errors.jsonl (when that directory exists): the tolerances in
    this file are set from these measurements (<= ~2x the worst value seen), see DESIGN.md §4."""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_errors.jsonl"), "a") as f:
            f.write(json.dumps({"name": name, "err": float(value), "tol": tol}) + "\n")
    return float(value)
