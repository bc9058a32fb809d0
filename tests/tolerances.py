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

# Set from the errors measured on a B200 (gpurun_out/parity_errors.jsonl, summarised in DESIGN.md §4): each bound is
# about twice the worst value observed, and never looser than the north-star 1e-3 for a single op.
OUT = {"fp32": 3e-5, "tf32": 7e-4}          # one op, forward                       (measured <= 3.3e-4)
GRAD = {"fp32": 1e-4, "tf32": 1e-3}         # one op, backward, no activation mask  (measured <= 3.3e-4)
NET = {"fp32": 1e-4, "tf32": 2.5e-3}        # whole network forward at reduced size (measured <= 1.3e-3)
NET_CFG1 = {"fp32": 1e-4, "tf32": 1e-3}     # BASELINE cfg 1, SURVEY §8(d): <= 1e-3 (measured <= 9.0e-4, E -> G -> E -> Ex)
NET_256 = {"fp32": 2e-4, "tf32": 7e-3}      # G at full depth and width on 256x256  (measured 3.5e-3)
GRAD_ACT_Q95 = 3e-3                         # gradients through a leaky-ReLU mask: 95th percentile (measured <= 1.7e-3)
GRAD_ACT_L2 = 3e-2                          #   and relative L2 (measured <= 1.5e-2; mask flips, see above)


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


def assert_grad_through_act(a, b, mode, what="", reduced=False):
    """gradient comparison that tolerates leaky-ReLU mask flips in tf32 mode (see module docstring).
    ``reduced``: the tensor is a reduction over all pixels (style, bias, modulation-weight gradients): every flip
    lands in every element, so only the relative L2 error is meaningful."""
    if mode == "fp32":
        assert rel(a, b) <= GRAD["fp32"], (what, rel(a, b))
    elif a.numel() < 10000 or reduced:
        assert rel_l2(a, b) <= GRAD_ACT_L2, (what, "l2", rel_l2(a, b))
    else:
        assert rel_q(a, b) <= GRAD_ACT_Q95, (what, "q95", rel_q(a, b))
        assert rel_l2(a, b) <= GRAD_ACT_L2, (what, "l2", rel_l2(a, b))


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
